// TEST INFRASTRUCTURE (oracle) -- plain C API over the C++ oracle for ctypes (tests/, smoke(),
// bench.py cpu_baseline / --impl reference only).  See orc.hpp header.
#include <algorithm>
#include <map>

#include "orc.hpp"
using namespace orc;

namespace orc {
void CalcScalarSourcesForK(Model& M, int q_ix, Counters& cnt, int& err, std::string& errmsg);
}

extern "C" {

void* orc_new() { return new Model(); }
void orc_free(void* h) { delete (Model*)h; }

// Set a parameter by name.  Returns 0 if known.
int orc_set(void* h, const char* name, double v) {
  Model& M = *(Model*)h;
  Params& P = M.P;
  std::string n(name);
#define D(f) if (n == #f) { P.f = v; return 0; }
#define I(f) if (n == #f) { P.f = (int)v; return 0; }
  D(H0) D(omegab) D(omegac) D(omegav) D(omegan) D(TCMB) D(YHe) D(Num_Nu_massless)
  I(Num_Nu_massive) I(Nu_mass_eigenstates) I(share_delta_neff)
  I(use_galileon) D(c2) D(c3) D(c4) D(cG)
  D(ScalarPowerAmp) D(an) D(n_run) D(n_runrun) D(k_0_scalar)
  I(Reionization) I(use_optical_depth) D(optical_depth) D(redshift) D(delta_redshift) D(fraction)
  D(helium_redshift) D(helium_delta_redshift) D(helium_redshiftstart)
  D(RECFAST_fudge) D(RECFAST_fudge_He) I(RECFAST_Heswitch) I(RECFAST_Hswitch)
  I(Max_l) D(Max_eta_k) I(DoLensing) I(AccuratePolarization) I(AccurateReionization) I(MassiveNuMethod)
  I(DoLateRadTruncation) I(HighAccuracyDefault) D(AccuracyBoost) D(lSampleBoost) I(use_spline_template) I(nthreads)
  I(WantTransfer) I(transfer_high_precision) D(transfer_kmax) I(transfer_k_per_logint) I(transfer_num_redshifts)
  D(ALens) D(ALens_Fiducial)
#undef D
#undef I
  if (n == "lAccuracyBoost") { P.lAccuracyBoost = (float)v; return 0; }
  // per-eigenstate inputs: name + 1-based eigenstate index (camb/inidriver.F90:143-165)
  if (!n.empty() && n.back() >= '1' && n.back() <= '0' + max_nu) {
    const int i = n.back() - '1';
    const std::string b = n.substr(0, n.size() - 1);
    if (b == "Nu_mass_numbers") { P.Nu_mass_numbers[i] = (int)v; return 0; }
    if (b == "Nu_mass_fractions") { P.Nu_mass_fractions[i] = v; return 0; }
    if (b == "Nu_mass_degeneracies") { P.Nu_mass_degeneracies[i] = v; return 0; }
  }
  return 1;
}

// transfer redshifts, descending like CP%Transfer%redshifts (camb/modules.f90:2538-2570)
int orc_set_transfer_redshifts(void* h, const double* z, int n) {
  Params& P = ((Model*)h)->P;
  if (n < 0 || n > max_transfer_redshifts) return 1;
  P.transfer_num_redshifts = n;
  for (int i = 0; i < n; i++) P.transfer_redshifts[i] = z[i];
  return 0;
}
// matter-transfer results: dims = {Transfer_max, num_q_trans, nz}; any pointer may be NULL
int orc_get_transfer(void* h, int* dims, double* TransferData, double* q_trans, double* tautf, double* sigma8) {
  Model& M = *(Model*)h;
  const int nz = M.P.WantTransfer ? M.P.transfer_num_redshifts : 0;
  if (dims) {
    dims[0] = Transfer_max;
    dims[1] = M.num_q_trans;
    dims[2] = nz;
  }
  if (TransferData) std::copy(M.TransferData.begin(), M.TransferData.end(), TransferData);
  if (q_trans)
    for (int i = 1; i <= M.num_q_trans; i++) q_trans[i - 1] = M.q_trans[i];
  if (tautf)
    for (int i = 1; i <= nz; i++) tautf[i - 1] = M.tautf[i];
  if (sigma8)
    for (int i = 0; i < nz && i < (int)M.sigma_8.size(); i++) sigma8[i] = M.sigma_8[i];
  return 0;
}
int orc_stage_transfer(void* h) {  // TransferOut + Transfer_Get_sigmas (after orc_stage_sources)
  Model& M = *(Model*)h;
  if (!M.P.WantTransfer) return 0;
  TransferOut(M);
  if (M.error_flag) return M.error_flag;
  Transfer_Get_sigmas(M);
  return 0;
}

// "physical" convenience of camb/inidriver.F90:107-118 (use_physical = T, omk = 0)
void orc_set_physical(void* h, double ombh2, double omch2, double omnuh2, double H0) {
  Params& P = ((Model*)h)->P;
  const double h2 = (H0 / 100) * (H0 / 100);
  P.H0 = H0;
  P.omegab = ombh2 / h2;
  P.omegac = omch2 / h2;
  P.omegan = omnuh2 / h2;
  P.omegav = 1 - 0.0 - P.omegab - P.omegac - P.omegan;
}

int orc_load_template(void* h, const char* path) { return load_highL_template(*(Model*)h, path); }
int orc_set_template(void* h, const double* tmpl /* [4][7999] TT,EE,TE,PP */) {
  Model& M = *(Model*)h;
  M.highL_template.assign(tmpl, tmpl + (size_t)4 * (lmax_extrap_highl - lmin + 1));
  return 0;
}

int orc_run(void* h, double* stage_times, int reuse_bessels) {
  return run_model(*(Model*)h, stage_times, reuse_bessels != 0);
}

// Staged execution (same order as run_model) so tests can feed the CUDA stages with oracle inputs.
int orc_stage_background(void* h) { return CAMBParams_Set(*(Model*)h); }
int orc_stage_initvars(void* h) {
  Model& M = *(Model*)h;
  M.WantLateTime = M.P.DoLensing != 0;
  M.maximum_l = M.P.Max_l;
  M.maximum_qeta = M.Max_eta_k;
  initlval(M, M.maximum_l);
  cmb_InitVars(M);
  if (M.error_flag) return M.error_flag;
  SetkValuesForSources(M);
  if (M.P.WantTransfer) InitTransfer(M);
  M.SourceNum = M.WantLateTime ? 3 : 2;
  M.C_last = M.WantLateTime ? 6 : 3;
  return 0;
}
int orc_stage_sources(void* h) {
  Model& M = *(Model*)h;
  CalcAllSources(M);
  return M.error_flag;
}
// evolve a single source k (1-based index); Src must have been allocated by orc_alloc_sources
int orc_alloc_sources(void* h) {
  Model& M = *(Model*)h;
  size_t n = (size_t)M.Evolve_q.npoints * M.SourceNum * M.TimeSteps.npoints;
  M.Src.assign(n, 0.0);
  M.ddSrc.assign(n, 0.0);
  return 0;
}
int orc_source_k(void* h, int q_ix, double* n_derivs) {
  Model& M = *(Model*)h;
  Counters c;
  int err = 0;
  std::string msg;
  CalcScalarSourcesForK(M, q_ix, c, err, msg);
  if (n_derivs) *n_derivs = c.n_derivs;
  return err;
}
int orc_stage_spline_k(void* h) {
  InitSourceInterpolation(*(Model*)h);
  return 0;
}
int orc_stage_bessels(void* h) {
  Model& M = *(Model*)h;
  M.max_bessels_l_index = M.lSamp.l0;
  M.max_bessels_etak = M.maximum_qeta;
  GenerateBessels(M);
  return 0;
}
int orc_stage_qgrid(void* h) {  // the integration k grid alone (host pre-stage of the GPU path)
  SetkValuesForInt(*(Model*)h);
  return 0;
}
int orc_stage_los(void* h) {
  Model& M = *(Model*)h;
  SetkValuesForInt(M);
  IntegrateAll(M);
  return 0;
}
int orc_stage_cls(void* h) {
  Model& M = *(Model*)h;
  CalcScalCls(M);
  InterpolateCls(M);
  return 0;
}

int orc_stage_lensing(void* h) { return lens_Cls(*(Model*)h); }

const char* orc_error(void* h) { return ((Model*)h)->error_msg.c_str(); }

double orc_get(void* h, const char* name) {
  Model& M = *(Model*)h;
  std::string n(name);
#define G(f) if (n == #f) return (double)M.f;
  G(tau0) G(chi0) G(taurst) G(taurend) G(tau_maxvis) G(dtaurec) G(adotrad) G(akthom) G(grhom) G(grhog) G(grhor) G(grhob)
  G(grhoc) G(grhov) G(grhornomass) G(grhok) G(qmin) G(qmax) G(max_etak_scalar) G(SourceNum) G(C_last) G(num_xx)
  G(kmaxfile) G(epsw) G(Max_eta_k) G(Nu_mass_eigenstates) G(Num_Nu_massive) G(reion_redshift) G(reion_tau_start)
  G(reion_tau_complete) G(error_flag) G(maximum_l) G(scale) G(omegak)
#undef G
  if (n == "lmax_lensed") return M.lmax_lensed;
  if (n == "nk") return M.Evolve_q.npoints;
  if (n == "nt") return M.TimeSteps.npoints;
  if (n == "nq") return M.q_int.npoints;
  if (n == "nl") return M.lSamp.l0;
  if (n == "tight_tau") return M.th.tight_tau;
  if (n == "matter_verydom_tau") return M.th.matter_verydom_tau;
  if (n == "tauminn") return M.th.tauminn;
  if (n == "dlntau") return M.th.dlntau;
  if (n == "actual_opt_depth") return M.th.actual_opt_depth;
  if (n == "nu_dlnam") return M.nu.dlnam;
  if (n == "nu_nqmax") return M.nu.nqmax;
  if (n == "gal_c2") return M.gal.c2;
  if (n == "gal_c3") return M.gal.c3;
  if (n == "gal_c4") return M.gal.c4;
  if (n == "gal_c5") return M.gal.c5;
  if (n == "gal_cG") return M.gal.cG;
  if (n == "gal_h0") return M.gal.h0;
  if (n == "gal_orad") return M.gal.orad;
  if (n == "gal_om") return M.gal.om;
  if (n == "gal_n") return M.gal.n;
  if (n == "n_trial") return M.cnt.n_trial;
  if (n == "n_derivs") return M.cnt.n_derivs;
  if (n == "n_out") return M.cnt.n_out;
  if (n == "sum_n_trial") return M.cnt.sum_n_trial;
  if (n == "flop_ode") return M.cnt.flop_ode;
  if (n == "n_iter") return M.cnt.n_iter;
  if (!n.empty() && n.back() >= '1' && n.back() <= '0' + max_nu) {  // per-eigenstate values, 1-based suffix
    const int i = n.back() - '0';
    const std::string b = n.substr(0, n.size() - 1);
    if (b == "grhormass") return M.grhormass[i];
    if (b == "nu_masses") return M.nu_masses[i];
    if (b == "nu_tau_nonrelativistic") return M.nu_tau_nonrelativistic[i];
    if (b == "nu_tau_massive") return M.nu_tau_massive[i];
    if (b == "Nu_mass_degeneracies") return M.Nu_mass_degeneracies[i];
  }
  return NAN;
}

// Copy an array out.  Returns the array length (call with out = NULL to query).  1-based oracle
// arrays are returned 0-based (the unused slot 0 is dropped).
long orc_get_array(void* h, const char* name, double* out, long cap) {
  Model& M = *(Model*)h;
  std::string n(name);
  const std::vector<double>* v = nullptr;
  long skip = 0;
  std::vector<double> tmp;
  auto one = [&](const std::vector<double>& a) { v = &a; skip = 1; };
  auto zero = [&](const std::vector<double>& a) { v = &a; skip = 0; };
  if (n == "tau") one(M.TimeSteps.points);
  else if (n == "dtau") one(M.TimeSteps.dpoints);
  else if (n == "k") one(M.Evolve_q.points);
  else if (n == "q") one(M.q_int.points);
  else if (n == "dq") one(M.q_int.dpoints);
  else if (n == "bess_x") one(M.BessRanges.points);
  else if (n == "Src") zero(M.Src);
  else if (n == "ddSrc") zero(M.ddSrc);
  else if (n == "ajl") zero(M.ajl);
  else if (n == "ajlpr") zero(M.ajlpr);
  else if (n == "Delta_p_l_k") zero(M.Delta_p_l_k);
  else if (n == "iCl") zero(M.iCl_scalar);
  else if (n == "Cl") zero(M.Cl_scalar);
  else if (n == "template") zero(M.highL_template);
  else if (n == "Cl_lensed") zero(M.Cl_lensed);
  else if (n == "cs2") one(M.th.cs2);
  else if (n == "dcs2") one(M.th.dcs2);
  else if (n == "dotmu") one(M.th.dotmu);
  else if (n == "ddotmu") one(M.th.ddotmu);
  else if (n == "dddotmu") one(M.th.dddotmu);
  else if (n == "xe") one(M.th.xe);
  else if (n == "emmu") one(M.th.emmu);
  else if (n == "vis") one(M.th.vis);
  else if (n == "dvis") one(M.th.dvis);
  else if (n == "ddvis") one(M.th.ddvis);
  else if (n == "expmmu") one(M.th.expmmu);
  else if (n == "opac") one(M.th.opac);
  else if (n == "dopac") one(M.th.dopac);
  else if (n == "gal_a") zero(M.gal.a);
  else if (n == "gal_h") zero(M.gal.h);
  else if (n == "gal_x") zero(M.gal.x);
  else if (n == "gal_ch") zero(M.gal.ch);
  else if (n == "gal_cx") zero(M.gal.cx);
  else if (n == "nu_r1") one(M.nu.r1);
  else if (n == "nu_p1") one(M.nu.p1);
  else if (n == "nu_dr1") one(M.nu.dr1);
  else if (n == "nu_dp1") one(M.nu.dp1);
  else if (n == "nu_ddr1") one(M.nu.ddr1);
  else if (n == "nu_ddp1") one(M.nu.ddp1);
  else if (n == "xrec") one(M.rec.xrec);
  else if (n == "per_k_trials") one(M.per_k_trials);
  else if (n == "ls") {
    for (int i = 1; i <= M.lSamp.l0; i++) tmp.push_back(M.lSamp.l[i]);
    v = &tmp;
  } else if (n == "nu_q") {
    for (int i = 1; i <= M.nu.nqmax; i++) tmp.push_back(M.nu.nu_q[i]);
    v = &tmp;
  } else if (n == "nu_int_kernel") {
    for (int i = 1; i <= M.nu.nqmax; i++) tmp.push_back(M.nu.nu_int_kernel[i]);
    v = &tmp;
  } else if (n.rfind("nu_tau_notmassless", 0) == 0 && n.size() == 19 && n.back() >= '1' && n.back() <= '0' + max_nu) {
    for (int i = 1; i <= M.nu.nqmax; i++) tmp.push_back(M.nu_tau_notmassless[i][n.back() - '0']);
    v = &tmp;
  } else if (n == "regions_tau" || n == "regions_bess") {
    // flattened region table: per region {Low, High, delta, start_index, IsLog}
    const Regions& R = n == "regions_tau" ? M.TimeSteps : M.BessRanges;
    for (int i = 1; i <= R.count; i++) {
      tmp.push_back(R.R[i].Low);
      tmp.push_back(R.R[i].High);
      tmp.push_back(R.R[i].delta);
      tmp.push_back(R.R[i].start_index);
      tmp.push_back(R.R[i].IsLog ? 1 : 0);
    }
    v = &tmp;
  } else
    return -1;
  long len = (long)v->size() - skip;
  if (len < 0) len = 0;
  if (out) {
    long m = len < cap ? len : cap;
    for (long i = 0; i < m; i++) out[i] = (*v)[i + skip];
  }
  return len;
}

// Overwrite an oracle array (used to feed a stage with externally produced data, e.g. GPU Src).
long orc_put_array(void* h, const char* name, const double* in, long len) {
  Model& M = *(Model*)h;
  std::string n(name);
  std::vector<double>* v = nullptr;
  if (n == "Src") v = &M.Src;
  else if (n == "ddSrc") v = &M.ddSrc;
  else if (n == "Delta_p_l_k") v = &M.Delta_p_l_k;
  else if (n == "Cl") {  // feed the lensing stage alone (tests): (lmin..Max_l, C_last)
    v = &M.Cl_scalar;
    if (len == (long)(M.P.Max_l - lmin + 1) * M.C_last) v->resize(len);
  } else return -1;
  if ((long)v->size() != len) return -2;
  for (long i = 0; i < len; i++) (*v)[i] = in[i];
  return len;
}

// ---- known-answer-test hooks for the numerical primitives
double orc_bjl(int l, double x) {
  double j;
  bjl(l, x, j);
  return j;
}
void orc_spline(const double* x, const double* y, int n, double* d2) { spline(x - 1, y - 1, n, 1e40, 1e40, d2 - 1); }
void orc_splder(const double* y, double* dy, int n) {
  std::vector<double> g(n + 1);
  splini(g.data(), n);
  splder(y - 1, dy - 1, n, g.data());
}
static double kat_f(void*, double x) { return std::exp(-x) * std::cos(3 * x); }
double orc_rombint_kat(double a, double b, double tol) { return rombint(kat_f, nullptr, a, b, tol); }
// dverk on y1' = y2, y2' = -y1 (and a third, stiff-ish decaying component); returns #derivs
static int kat_n;
static void kat_osc(void*, int n, double, const double* y, double* yp) {
  kat_n++;
  yp[0] = y[1];
  yp[1] = -y[0];
  if (n > 2) yp[2] = -5 * y[2];
}
int orc_dverk_kat(double x0, double x1, int nout, double tol, double* y /*[3] in/out*/) {
  double c[25] = {0}, w[27] = {0};
  int ind = 1;
  kat_n = 0;
  double x = x0;
  for (int i = 1; i <= nout; i++) {
    double xe = x0 + (x1 - x0) * i / nout;
    dverk(nullptr, 3, kat_osc, x, y, xe, tol, ind, c, 3, w);
    if (ind != 3) return -ind;
  }
  return kat_n;
}
int orc_ranges_kat(const double* spec /*[n][4]: lo,hi,delta,islog*/, int n, double* pts, int cap) {
  Regions R;
  Ranges_Init(R);
  for (int i = 0; i < n; i++) Ranges_Add_delta(R, spec[4 * i], spec[4 * i + 1], spec[4 * i + 2], spec[4 * i + 3] != 0);
  Ranges_GetArray(R, true);
  for (int i = 1; i <= R.npoints && i <= cap; i++) pts[i - 1] = R.points[i];
  return R.npoints;
}
int orc_ranges_indexof(void* h, int which, double v) {
  Model& M = *(Model*)h;
  return Ranges_IndexOf(which == 0 ? M.TimeSteps : M.BessRanges, v);
}
double orc_dtauda(void* h, double a) { return dtauda(*(Model*)h, a); }
double orc_gal_H(void* h, double a) { return gal_GetH(*(Model*)h, a); }
double orc_gal_X(void* h, double a) { return gal_GetX(*(Model*)h, a); }
void orc_gal_tracker(void* h, double a, double* hh, double* xx) { gal_tracker(*(Model*)h, a, *hh, *xx); }
void orc_thermo(void* h, double tau, double* cs2, double* opac, double* dopac) {
  thermo(*(Model*)h, tau, *cs2, *opac, dopac);
}
void orc_nu_background(void* h, double am, double* rho, double* p) { Nu_background(*(Model*)h, am, *rho, *p); }
// all Galileon perturbation terms at one point: out[0..8] = dh, dx, grhogal, gpresgal, Chigal, qgal, Pigal,
// dphisecond, pigalprime
double orc_nu_dp(void* h, double am, double adotoa, double pnu) { return Nu_dp(*(Model*)h, am, adotoa, pnu); }
void orc_gal_terms(void* h, const double* in /*a,dgrho,dgq,dgpi,eta,dphi,dphip,k,deltafprime,pidot,grho,gpres*/,
                   double* out) {
  Model& M = *(Model*)h;
  const double a = in[0], dgrho = in[1], dgq = in[2], dgpi = in[3], eta = in[4], dphi = in[5], dphip = in[6], k = in[7],
               dfp = in[8], pidot = in[9], grho = in[10], gpres = in[11];
  const double x = gal_GetX(M, a), hc = a * gal_GetH(M, a);
  double dh, dx;
  gal_GetdHdX(M, a, hc, x, dh, dx);
  out[0] = dh;
  out[1] = dx;
  out[2] = gal_grhogal(M, a, hc, x);
  out[3] = gal_gpresgal(M, a, hc, x, dh, dx);
  out[4] = gal_Chigal(M, a, hc, x, dgrho, eta, dphi, dphip, k);
  out[5] = gal_qgal(M, a, hc, x, dgq, eta, dphi, dphip, k);
  out[6] = gal_Pigal(M, a, hc, x, dh, dx, dgrho, dgq, dgpi, eta, dphi, k);
  out[7] = gal_dphisecond(M, a, hc, x, dh, dx, dgrho, dgq, eta, dphi, dphip, k, dfp);
  out[8] = gal_pigalprime(M, a, hc, x, dh, dx, dgrho, dgq, dgpi, pidot, eta, dphi, dphip, k, grho, gpres);
}

}  // extern "C"
