// TEST INFRASTRUCTURE (oracle) -- see orc.hpp header.
// Scalar perturbation equations in the synchronous gauge with the Galileon stress-energy
// (module GaugeInterface of camb/equations_galileon.f90) and the per-k driver
// CalcScalarSources of camb/cmbmain.f90:917-1071.  Flat models, scalar modes, adiabatic
// initial conditions, w_lam = -1, no transfer-function outputs.
#include "orc.hpp"

namespace orc {

#define SP(x) ((double)(x##f))

// camb/equations_galileon.f90:200-265 (scalar subset).  Arrays are 1-based.
struct EvolutionVars {
  double q = 0, q2 = 0, k_buf = 0, k2_buf = 0;
  int w_ix = 0, r_ix = 0, g_ix = 0, q_ix = 0;
  int nvar = 0;
  int lmaxg = 0, lmaxnr = 0, lmaxnu = 0, lmaxgpol = 0, MaxlNeeded = 0;
  int polind = 0;
  int nu_ix[max_nu + 1] = {0}, nu_pert_ix = 0;
  int nq[max_nu + 1] = {0}, lmaxnu_pert = 0;
  bool has_nu_relativistic = false;
  double G11[max_nu + 1] = {0}, G30[max_nu + 1] = {0};
  bool MassiveNuApprox[max_nu + 1] = {false};
  double MassiveNuApproxTime[max_nu + 1] = {0};
  bool high_ktau_neutrino_approx = false;
  int NuMethod = 1;
  bool TightCoupling = true;
  double TightSwitchoffTime = 0;
  int ScalEqsToPropagate = 0;
  double pig = 0, pigdot = 0;
  double poltruncfac = 0;
  bool no_nu_multpoles = false, no_phot_multpoles = false;
  int lmaxnu_tau[max_nu + 1] = {0};
  bool nu_nonrelativistic[max_nu + 1] = {false};
  double denlk[max_l_evolve + 1], denlk2[max_l_evolve + 1], polfack[max_l_evolve + 1], Kf[max_l_evolve + 1];
  bool TransferOnly = false;         // camb/equations_galileon.f90:209
  double* OutputTransfer = nullptr;  // 1-based row of TransferData while outtransf runs (:264)
  // oracle bookkeeping
  const Model* M = nullptr;
  Counters* cnt = nullptr;
};

static const double ep0 = 1.0e-2;
static const int lmaxnu_high_ktau = 4;
static const int Nu_trunc = 1, Nu_best = 3;

static inline long nint(double x) { return std::lround(x); }

// camb/equations_galileon.f90:471-481
static double DeltaTimeMaxed(const Model& M, double a1, double a2, double tol = -1) {
  if (a1 > 1.0) return 0;
  if (a2 > 1.0) return DeltaTime(M, a1, 1.01, tol);
  return DeltaTime(M, a1, a2, tol);
}

// camb/equations_galileon.f90:510-553
void GaugeInterface_Init(Model& M) {
  M.epsw = 100 / M.tau0;
  for (int j = 2; j <= max_l_evolve; j++) M.polfac[j] = (double)((j + 3) * (j - 1)) / (j + 1);
  for (int j = 1; j <= max_l_evolve; j++) M.denl[j] = 1.0 / (2 * j + 1);
  const MassiveNuTables& T = M.nu;
  for (int nu_i = 1; nu_i <= M.Nu_mass_eigenstates; nu_i++) {
    double nu_mass = std::fmax(0.1, M.nu_masses[nu_i]);
    double a_mass = 1.e-1 / nu_mass / (double)M.P.lAccuracyBoost;
    double time = DeltaTime(M, 0.0, T.nu_q[1] * a_mass);
    M.nu_tau_notmassless[1][nu_i] = time;
    for (int j = 2; j <= T.nqmax; j++) {
      time = time + DeltaTimeMaxed(M, T.nu_q[j - 1] * a_mass, T.nu_q[j] * a_mass, 0.01);
      M.nu_tau_notmassless[j][nu_i] = time;
    }
    double a_nonrel = 2.5 / nu_mass * M.P.AccuracyBoost;
    M.nu_tau_nonrelativistic[nu_i] = DeltaTimeMaxed(M, 0.0, a_nonrel);
    double a_massive = 17.0 / nu_mass * M.P.AccuracyBoost;
    M.nu_tau_massive[nu_i] = M.nu_tau_nonrelativistic[nu_i] + DeltaTimeMaxed(M, a_nonrel, a_massive);
  }
}

// camb/equations_galileon.f90:298-312
static int next_nu_nq(const Model& M, int nq) {
  if (nq == 0) return 1;
  int q = (int)M.nu.nu_q[nq];  // integer q = nu_q(nq)
  // Fortran real->integer assignment truncates toward zero
  if (q >= 10) return M.nu.nqmax;
  return nq + 1;
}

// camb/equations_galileon.f90:555-644
static void SetupScalarArrayIndices(EvolutionVars& EV, int* max_num_eqns = nullptr) {
  const Model& M = *EV.M;
  const int basic_num_eqns = 5;
  int neq = basic_num_eqns, maxeq = neq;
  if (!EV.no_phot_multpoles) {
    EV.g_ix = basic_num_eqns + 1;
    if (EV.TightCoupling)
      neq = neq + 2;
    else {
      neq = neq + (EV.lmaxg + 1);
      EV.polind = neq - 1;
      neq = neq + EV.lmaxgpol - 1;
    }
  }
  if (!EV.no_nu_multpoles) {
    EV.r_ix = neq + 1;
    if (EV.high_ktau_neutrino_approx)
      neq = neq + 3;
    else
      neq = neq + (EV.lmaxnr + 1);
  }
  maxeq = maxeq + (EV.lmaxg + 1) + (EV.lmaxnr + 1) + EV.lmaxgpol - 1;
  if (M.P.use_galileon) {
    EV.w_ix = neq + 1;
    neq = neq + 2;
    maxeq = maxeq + 2;
  } else {
    EV.w_ix = 0;
  }
  if (M.Num_Nu_massive != 0) {
    const int nqmax = M.nu.nqmax;
    EV.has_nu_relativistic = false;
    for (int i = 1; i <= M.Nu_mass_eigenstates; i++)
      if (EV.nq[i] != nqmax) EV.has_nu_relativistic = true;
    if (EV.has_nu_relativistic) {
      EV.lmaxnu_pert = EV.lmaxnu;
      EV.nu_pert_ix = neq + 1;
      neq = neq + EV.lmaxnu_pert + 1;
      maxeq = maxeq + EV.lmaxnu_pert + 1;
    } else {
      EV.lmaxnu_pert = 0;
    }
    const double lAB = (double)M.P.lAccuracyBoost;
    for (int nu_i = 1; nu_i <= M.Nu_mass_eigenstates; nu_i++) {
      if (EV.high_ktau_neutrino_approx) {
        EV.lmaxnu_tau[nu_i] = (int)(lmaxnu_high_ktau * M.P.lAccuracyBoost);
      } else {
        EV.lmaxnu_tau[nu_i] =
            std::max((int)std::min(nint(0.8 * EV.q * M.nu_tau_nonrelativistic[nu_i] * lAB), (long)EV.lmaxnu), 3);
        if (EV.nu_nonrelativistic[nu_i])
          EV.lmaxnu_tau[nu_i] = std::min(EV.lmaxnu_tau[nu_i], (int)nint(4 * lAB));
      }
      EV.lmaxnu_tau[nu_i] = std::min(EV.lmaxnu, EV.lmaxnu_tau[nu_i]);
      EV.nu_ix[nu_i] = neq + 1;
      if (EV.MassiveNuApprox[nu_i])
        neq = neq + 4;
      else
        neq = neq + EV.nq[nu_i] * (EV.lmaxnu_tau[nu_i] + 1);
      maxeq = maxeq + nqmax * (EV.lmaxnu + 1);
    }
  } else {
    EV.has_nu_relativistic = false;
  }
  EV.ScalEqsToPropagate = neq;
  if (max_num_eqns) *max_num_eqns = maxeq;
}

// camb/equations_galileon.f90:646-724   (y, yout 1-based)
static void CopyScalarVariableArray(const double* y, double* yout, const EvolutionVars& EV,
                                    const EvolutionVars& EVout) {
  const Model& M = *EV.M;
  for (int i = 1; i <= EVout.nvar; i++) yout[i] = 0;
  for (int i = 1; i <= 5; i++) yout[i] = y[i];
  if (M.P.use_galileon) {
    yout[EVout.w_ix] = y[EV.w_ix];
    yout[EVout.w_ix + 1] = y[EV.w_ix + 1];
  }
  int lmax;
  if (!EV.no_phot_multpoles && !EVout.no_phot_multpoles) {
    if (EV.TightCoupling || EVout.TightCoupling)
      lmax = 1;
    else
      lmax = std::min(EV.lmaxg, EVout.lmaxg);
    for (int l = 0; l <= lmax; l++) yout[EVout.g_ix + l] = y[EV.g_ix + l];
    if (!EV.TightCoupling && !EVout.TightCoupling) {
      lmax = std::min(EV.lmaxgpol, EVout.lmaxgpol);
      for (int l = 2; l <= lmax; l++) yout[EVout.polind + l] = y[EV.polind + l];
    }
  }
  if (!EV.no_nu_multpoles && !EVout.no_nu_multpoles) {
    if (EV.high_ktau_neutrino_approx || EVout.high_ktau_neutrino_approx)
      lmax = 2;
    else
      lmax = std::min(EV.lmaxnr, EVout.lmaxnr);
    for (int l = 0; l <= lmax; l++) yout[EVout.r_ix + l] = y[EV.r_ix + l];
  }
  if (M.Num_Nu_massive != 0) {
    for (int nu_i = 1; nu_i <= M.Nu_mass_eigenstates; nu_i++) {
      int ix_off = EV.nu_ix[nu_i], ix_off2 = EVout.nu_ix[nu_i];
      if (EV.MassiveNuApprox[nu_i] && EVout.MassiveNuApprox[nu_i]) {
        for (int l = 0; l < 4; l++) yout[ix_off2 + l] = y[ix_off + l];
      } else if (!EV.MassiveNuApprox[nu_i] && !EVout.MassiveNuApprox[nu_i]) {
        lmax = std::min(EV.lmaxnu_tau[nu_i], EVout.lmaxnu_tau[nu_i]);
        int nq = std::min(EV.nq[nu_i], EVout.nq[nu_i]);
        for (int i = 1; i <= nq; i++) {
          int ind = ix_off + (i - 1) * (EV.lmaxnu_tau[nu_i] + 1);
          int ind2 = ix_off2 + (i - 1) * (EVout.lmaxnu_tau[nu_i] + 1);
          for (int l = 0; l <= lmax; l++) yout[ind2 + l] = y[ind + l];
        }
        for (int i = nq + 1; i <= EVout.nq[nu_i]; i++) {
          lmax = std::min(EVout.lmaxnu_tau[nu_i], EV.lmaxnr);
          int ind2 = ix_off2 + (i - 1) * (EVout.lmaxnu_tau[nu_i] + 1);
          for (int l = 0; l <= lmax; l++) yout[ind2 + l] = y[EV.r_ix + l];
          double q = M.nu.nu_q[i];
          double r = M.nu_masses[nu_i] / q;
          double pert_scale = (r * r) / 2;
          lmax = std::min(lmax, EV.lmaxnu_pert);
          for (int l = 0; l <= lmax; l++) yout[ind2 + l] = yout[ind2 + l] + y[EV.nu_pert_ix + l] * pert_scale;
        }
      }
    }
    if (EVout.has_nu_relativistic && EV.has_nu_relativistic) {
      lmax = std::min(EVout.lmaxnu_pert, EV.lmaxnu_pert);
      for (int l = 0; l <= lmax; l++) yout[EVout.nu_pert_ix + l] = y[EV.nu_pert_ix + l];
    }
  }
}

// camb/equations_galileon.f90:793-946 (scalar, flat, no transfer)
static void GetNumEqns(EvolutionVars& EV) {
  const Model& M = *EV.M;
  const Params& P = M.P;
  const double lAB = (double)P.lAccuracyBoost;
  double max_nu_mass;
  if (M.Num_Nu_massive == 0) {
    EV.lmaxnu = 0;
    max_nu_mass = 0;
  } else {
    const int nqmax = M.nu.nqmax;
    max_nu_mass = 0;
    for (int i = 1; i <= M.Nu_mass_eigenstates; i++) max_nu_mass = std::fmax(max_nu_mass, M.nu_masses[i]);
    for (int nu_i = 1; nu_i <= M.Nu_mass_eigenstates; nu_i++) {
      int q_rel = 0;
      for (int j = 1; j <= nqmax; j++) {
        if (M.nu.nu_q[j] > M.nu_masses[nu_i] * M.adotrad / EV.q) break;
        q_rel = q_rel + 1;
      }
      if (q_rel >= nqmax - 2)
        EV.nq[nu_i] = nqmax;
      else
        EV.nq[nu_i] = q_rel;
      EV.nu_nonrelativistic[nu_i] = false;
    }
    EV.NuMethod = P.MassiveNuMethod;
    if (EV.NuMethod == Nu_best) EV.NuMethod = Nu_trunc;
    EV.lmaxnu = std::max(3, (int)nint(10 * lAB));
    if (max_nu_mass > 700) EV.lmaxnu = std::max(3, (int)nint(15 * lAB));
  }
  EV.high_ktau_neutrino_approx = false;
  EV.TightCoupling = true;
  EV.no_phot_multpoles = false;
  EV.no_nu_multpoles = false;
  for (int i = 0; i <= max_nu; i++) EV.MassiveNuApprox[i] = false;
  if (P.HighAccuracyDefault && P.AccuratePolarization)
    EV.lmaxg = std::max((int)nint(11 * lAB), 3);
  else
    EV.lmaxg = std::max((int)nint(8 * lAB), 3);
  EV.lmaxnr = std::max((int)nint(14 * lAB), 3);
  if (max_nu_mass > 700 && P.HighAccuracyDefault) EV.lmaxnr = std::max((int)nint(32 * lAB), 3);
  EV.lmaxgpol = EV.lmaxg;
  if (!P.AccuratePolarization) EV.lmaxgpol = std::max((int)nint(4 * lAB), 3);
  if (EV.q < SP(0.05)) {
    double scal = 1;
    if (P.AccuratePolarization) scal = 4;
    EV.lmaxgpol = std::max(3, (int)nint(std::min(8L, nint(scal * 150 * EV.q)) * lAB));
    EV.lmaxnr = std::max(3, (int)nint(std::min(7L, nint(std::sqrt(scal) * 150 * EV.q)) * lAB));
    EV.lmaxg = std::max(3, (int)nint(std::min(8L, nint(std::sqrt(scal) * 300 * EV.q)) * lAB));
    if (P.AccurateReionization) {
      EV.lmaxg = EV.lmaxg * 4;
      EV.lmaxgpol = EV.lmaxgpol * 2;
    }
  }
  if (EV.TransferOnly) {  // :872-875
    EV.lmaxgpol = std::min(EV.lmaxgpol, (int)nint(5 * lAB));
    EV.lmaxg = std::min(EV.lmaxg, (int)nint(6 * lAB));
  }
  EV.poltruncfac = (double)EV.lmaxgpol / std::max(1, (EV.lmaxgpol - 2));
  EV.MaxlNeeded = std::max(std::max(EV.lmaxg, EV.lmaxnr), std::max(EV.lmaxgpol, EV.lmaxnu));
  if (EV.MaxlNeeded > max_l_evolve) {
    fprintf(stderr, "Need to increase max_l_evolve\n");
    std::abort();
  }
  SetupScalarArrayIndices(EV, &EV.nvar);
}

// camb/equations_galileon.f90:1058-1107 (y 1-based)
static void Nu_Integrate_L012(const EvolutionVars& EV, const double* y, double a, int nu_i, double& drhonu,
                              double& fnu, double* dpnu, double* pinu) {
  const Model& M = *EV.M;
  const MassiveNuTables& T = M.nu;
  drhonu = 0;
  fnu = 0;
  if (dpnu) {
    *dpnu = 0;
    *pinu = 0;
  }
  double am = a * M.nu_masses[nu_i];
  int ind = EV.nu_ix[nu_i];
  for (int iq = 1; iq <= EV.nq[nu_i]; iq++) {
    double aq = am / T.nu_q[iq];
    double v = 1.0 / std::sqrt(1.0 + aq * aq);
    drhonu = drhonu + T.nu_int_kernel[iq] * y[ind] / v;
    fnu = fnu + T.nu_int_kernel[iq] * y[ind + 1];
    if (dpnu) {
      *dpnu = *dpnu + T.nu_int_kernel[iq] * y[ind] * v;
      *pinu = *pinu + T.nu_int_kernel[iq] * y[ind + 2] * v;
    }
    ind = ind + EV.lmaxnu_tau[nu_i] + 1;
  }
  ind = EV.nu_pert_ix;
  for (int iq = EV.nq[nu_i] + 1; iq <= T.nqmax; iq++) {
    double aq = am / T.nu_q[iq];
    double v = 1.0 / std::sqrt(1.0 + aq * aq);
    double r = M.nu_masses[nu_i] / T.nu_q[iq];
    double pert_scale = (r * r) / 2;
    double tmp = T.nu_int_kernel[iq] * (y[EV.r_ix] + pert_scale * y[ind]);
    drhonu = drhonu + tmp / v;
    fnu = fnu + T.nu_int_kernel[iq] * (y[EV.r_ix + 1] + pert_scale * y[ind + 1]);
    if (dpnu) {
      *dpnu = *dpnu + tmp * v;
      *pinu = *pinu + T.nu_int_kernel[iq] * (y[EV.r_ix + 2] + pert_scale * y[ind + 2]) * v;
    }
  }
  if (dpnu) *dpnu = *dpnu / 3;
}

// camb/equations_galileon.f90:1109-1151
static void Nu_pinudot(const EvolutionVars& EV, const double* y, const double* ydot, double a, double adotoa, int nu_i,
                       double& pinudot) {
  const Model& M = *EV.M;
  const MassiveNuTables& T = M.nu;
  pinudot = 0.0;
  int ind = EV.nu_ix[nu_i] + 2;
  double am = a * M.nu_masses[nu_i];
  for (int iq = 1; iq <= EV.nq[nu_i]; iq++) {
    double q = T.nu_q[iq];
    double aq = am / q;
    double aqdot = aq * adotoa;
    double v = 1.0 / std::sqrt(1.0 + aq * aq);
    double vdot = -aq * aqdot / std::pow(1.0 + aq * aq, 1.5);
    pinudot = pinudot + T.nu_int_kernel[iq] * (ydot[ind] * v + y[ind] * vdot);
    ind = ind + EV.lmaxnu_tau[nu_i] + 1;
  }
  ind = EV.nu_pert_ix + 2;
  for (int iq = EV.nq[nu_i] + 1; iq <= T.nqmax; iq++) {
    double q = T.nu_q[iq];
    double aq = am / q;
    double aqdot = aq * adotoa;
    double r = M.nu_masses[nu_i] / q;
    double pert_scale = (r * r) / 2;
    double v = 1.0 / std::sqrt(1.0 + aq * aq);
    double vdot = -aq * aqdot / std::pow(1.0 + aq * aq, 1.5);
    double psi2dot = ydot[EV.r_ix + 2] + pert_scale * ydot[ind];
    double psi2 = y[EV.r_ix + 2] + pert_scale * y[ind];
    pinudot = pinudot + T.nu_int_kernel[iq] * (psi2dot * v + psi2 * vdot);
  }
}

// camb/equations_galileon.f90:1179-1207
static void Nu_Intvsq(const EvolutionVars& EV, const double* y, double a, int nu_i, double& G11, double& G30) {
  const Model& M = *EV.M;
  const MassiveNuTables& T = M.nu;
  double am = a * M.nu_masses[nu_i];
  int ind = EV.nu_ix[nu_i];
  G11 = 0.0;
  G30 = 0.0;
  if (EV.nq[nu_i] != T.nqmax) {
    fprintf(stderr, "Nu_Intvsq nq/=nqmax0\n");
    std::abort();
  }
  for (int iq = 1; iq <= EV.nq[nu_i]; iq++) {
    double q = T.nu_q[iq];
    double aq = am / q;
    double v = 1.0 / std::sqrt(1.0 + aq * aq);
    G11 = G11 + T.nu_int_kernel[iq] * y[ind + 1] * (v * v);
    if (EV.lmaxnu_tau[nu_i] > 2) G30 = G30 + T.nu_int_kernel[iq] * y[ind + 3] * (v * v);
    ind = ind + EV.lmaxnu_tau[nu_i] + 1;
  }
}

// camb/equations_galileon.f90:1209-1252
static void MassiveNuVars(const EvolutionVars& EV, const double* y, double a, double& grho, double& gpres,
                          double& dgrho, double& dgq, double* wnu_arr) {
  const Model& M = *EV.M;
  for (int nu_i = 1; nu_i <= M.Nu_mass_eigenstates; nu_i++) {
    double grhormass_t = M.grhormass[nu_i] / (a * a);
    double rhonu, pnu, clxnu, qnu;
    Nu_background(M, a * M.nu_masses[nu_i], rhonu, pnu);
    if (EV.MassiveNuApprox[nu_i]) {
      clxnu = y[EV.nu_ix[nu_i]];
      qnu = y[EV.nu_ix[nu_i] + 2];
    } else {
      Nu_Integrate_L012(EV, y, a, nu_i, clxnu, qnu, nullptr, nullptr);
      qnu = qnu / rhonu;
      clxnu = clxnu / rhonu;
    }
    double grhonu_t = grhormass_t * rhonu;
    double gpnu_t = grhormass_t * pnu;
    grho = grho + grhonu_t;
    gpres = gpres + gpnu_t;
    dgrho = dgrho + grhonu_t * clxnu;
    dgq = dgq + grhonu_t * qnu;
    if (wnu_arr) wnu_arr[nu_i] = pnu / rhonu;
  }
}

// camb/equations_galileon.f90:992-1056 with all optional arguments present (the only call, :1353)
static void MassiveNuVarsOut(const EvolutionVars& EV, const double* y, const double* yprime, double a, double& grho,
                             double& gpres, double& dgrho, double& dgq, double& dgpi, double& gdpi_diff,
                             double& pidot_sum) {
  const Model& M = *EV.M;
  double grhonu = 0, dgrhonu = 0;
  for (int nu_i = 1; nu_i <= M.Nu_mass_eigenstates; nu_i++) {
    double grhormass_t = M.grhormass[nu_i] / (a * a);
    double rhonu, pnu, clxnu, qnu, pinu, dpnu, pinudot;
    Nu_background(M, a * M.nu_masses[nu_i], rhonu, pnu);
    if (EV.MassiveNuApprox[nu_i]) {
      clxnu = y[EV.nu_ix[nu_i]];
      qnu = y[EV.nu_ix[nu_i] + 2];
      pinu = y[EV.nu_ix[nu_i] + 3];
      pinudot = yprime[EV.nu_ix[nu_i] + 3];
    } else {
      Nu_Integrate_L012(EV, y, a, nu_i, clxnu, qnu, &dpnu, &pinu);
      qnu = qnu / rhonu;
      clxnu = clxnu / rhonu;
      pinu = pinu / rhonu;
      double adotoa = 1 / (a * dtauda(M, a));
      double rhonudot = Nu_drho(M, a * M.nu_masses[nu_i], adotoa, rhonu);
      Nu_pinudot(EV, y, yprime, a, adotoa, nu_i, pinudot);
      pinudot = pinudot / rhonu - rhonudot / rhonu * pinu;
    }
    double grhonu_t = grhormass_t * rhonu;
    double gpnu_t = grhormass_t * pnu;
    grhonu = grhonu + grhonu_t;
    gpres = gpres + gpnu_t;
    dgrhonu = dgrhonu + grhonu_t * clxnu;
    dgq = dgq + grhonu_t * qnu;
    dgpi = dgpi + grhonu_t * pinu;
    gdpi_diff = gdpi_diff + pinu * (3 * gpnu_t - grhonu_t);
    pidot_sum = pidot_sum + grhonu_t * pinudot;
  }
  grho = grho + grhonu;
  dgrho = dgrho + dgrhonu;
}

// MassiveNuVarsOut as called from the transfer tap of derivs (camb/equations_galileon.f90:2335: only clxnu_all and
// dgpi present; the pinudot work of :1033-1037 feeds no requested output and is left out)
static void MassiveNuVarsOut_transfer(const EvolutionVars& EV, const double* y, double a, double& clxnu_all, double& dgpi) {
  const Model& M = *EV.M;
  double grhonu = 0, dgrhonu = 0;
  for (int nu_i = 1; nu_i <= M.Nu_mass_eigenstates; nu_i++) {
    const double grhormass_t = M.grhormass[nu_i] / (a * a);
    double rhonu, pnu, clxnu, qnu, pinu, dpnu;
    Nu_background(M, a * M.nu_masses[nu_i], rhonu, pnu);
    if (EV.MassiveNuApprox[nu_i]) {
      clxnu = y[EV.nu_ix[nu_i]];
      pinu = y[EV.nu_ix[nu_i] + 3];
    } else {
      Nu_Integrate_L012(EV, y, a, nu_i, clxnu, qnu, &dpnu, &pinu);
      clxnu = clxnu / rhonu;
      pinu = pinu / rhonu;
    }
    const double grhonu_t = grhormass_t * rhonu;
    grhonu = grhonu + grhonu_t;
    dgrhonu = dgrhonu + grhonu_t * clxnu;
    dgpi = dgpi + grhonu_t * pinu;
  }
  clxnu_all = dgrhonu / grhonu;
}

// camb/equations_galileon.f90:2098-2589.  ay, ayprime 1-based.  EV%pig / EV%pigdot are written
// as a side effect during tight coupling (SURVEY.md section 9 trap 1).
static void derivs(EvolutionVars& EV, int n, double tau, const double* ay, double* ayprime) {
  const Model& M = *EV.M;
  const bool gal = M.P.use_galileon != 0;
  const double k = EV.k_buf, k2 = EV.k2_buf;
  const double a = ay[1], a2 = a * a;
  const double etak = ay[2], clxc = ay[3], clxb = ay[4], vb = ay[5];
  double dphi = 0, dphiprime = 0;
  if (gal) {
    dphi = ay[EV.w_ix];
    dphiprime = ay[EV.w_ix + 1];
  }
  const double grhob_t = M.grhob / a, grhoc_t = M.grhoc / a, grhor_t = M.grhornomass / a2, grhog_t = M.grhog / a2;
  double grhov_t = 0, xgal = 0, hub = 0, dh = 0, dx = 0, grhogal_t = 0;
  if (gal) {
    xgal = gal_GetX(M, a);
    hub = a * gal_GetH(M, a);
    gal_GetdHdX(M, a, hub, xgal, dh, dx);
    grhogal_t = gal_grhogal(M, a, hub, xgal);
  } else {
    grhov_t = M.grhov * a2;
  }
  double cs2, opacity, dopacity = 0;
  if (EV.TightCoupling)
    thermo(M, tau, cs2, opacity, &dopacity);
  else
    thermo(M, tau, cs2, opacity, nullptr);
  double gpres = 0;
  gpres = gpres + (grhog_t + grhor_t) / 3;
  if (gal) {
    double gpresgal_t = gal_gpresgal(M, a, hub, xgal, dh, dx);
    gpres = gpres + gpresgal_t;
  } else {
    gpres = gpres + grhov_t * (-1.0);
  }
  double grho_matter = grhob_t + grhoc_t;
  double dgrho_matter = grhob_t * clxb + grhoc_t * clxc;
  double dgq = grhob_t * vb;
  double wnu_arr[max_nu + 1];
  if (M.Num_Nu_massive > 0) MassiveNuVars(EV, ay, a, grho_matter, gpres, dgrho_matter, dgq, wnu_arr);
  double grho = grho_matter + grhor_t + grhog_t;
  if (gal)
    grho = grho + grhogal_t;
  else
    grho = grho + grhov_t;
  (void)grho;
  const double adotoa = 1. / (a * dtauda(M, a));
  const double cothxor = 1.0 / tau;
  double dgrho = dgrho_matter;
  double z = 0, dz, clxr, qr, pir, clxg, qg, pig = 0, dgrhogal_t;
  if (EV.no_nu_multpoles) {
    if (gal) {
      dgrhogal_t = gal_Chigal(M, a, hub, xgal, dgrho, etak, dphi, dphiprime, k);
      z = (0.5 * (dgrho + dgrhogal_t) / k + etak) / adotoa;
      dz = -adotoa * z - 0.5 * (dgrho + dgrhogal_t) / k;
    } else {
      z = (0.5 * dgrho / k + etak) / adotoa;
      dz = -adotoa * z - 0.5 * dgrho / k;
    }
    clxr = -4 * dz / k;
    qr = -4.0 / 3 * z;
    pir = 0;
  } else {
    clxr = ay[EV.r_ix];
    qr = ay[EV.r_ix + 1];
    pir = ay[EV.r_ix + 2];
  }
  if (EV.no_phot_multpoles) {
    if (!EV.no_nu_multpoles) {
      if (gal) {
        dgrhogal_t = gal_Chigal(M, a, hub, xgal, dgrho, etak, dphi, dphiprime, k);
        z = (0.5 * (dgrho + dgrhogal_t) / k + etak) / adotoa;
        dz = -adotoa * z - 0.5 * (dgrho + dgrhogal_t) / k;
      } else {
        z = (0.5 * dgrho / k + etak) / adotoa;
        dz = -adotoa * z - 0.5 * dgrho / k;
      }
      clxg = -4 * dz / k - 4 / k * opacity * (vb + z);
      qg = -4.0 / 3 * z;
    } else {
      clxg = clxr - 4 / k * opacity * (vb + z);
      qg = qr;
    }
    pig = 0;
  } else {
    clxg = ay[EV.g_ix];
    qg = ay[EV.g_ix + 1];
    if (!EV.TightCoupling) pig = ay[EV.g_ix + 2];
  }
  dgrho = dgrho + grhog_t * clxg + grhor_t * clxr;
  dgq = dgq + grhog_t * qg + grhor_t * qr;
  const double photbar = grhog_t / grhob_t;
  const double pb43 = 4.0 / 3 * photbar;
  ayprime[1] = adotoa * a;
  if (gal) {
    dgrhogal_t = gal_Chigal(M, a, hub, xgal, dgrho, etak, dphi, dphiprime, k);
    dgrho = dgrho + dgrhogal_t;
    double dgqgal_t = gal_qgal(M, a, hub, xgal, dgq, etak, dphi, dphiprime, k);
    dgq = dgq + dgqgal_t;
  }
  z = (0.5 * dgrho / k + etak) / adotoa;
  const double sigma = (z + 1.5 * dgq / k2);
  ayprime[2] = 0.5 * dgq;
  if (EV.OutputTransfer) {  // :2325-2353 (pig is not yet assigned during tight coupling in the reference: 0 here)
    double* T = EV.OutputTransfer;
    T[Transfer_kh] = k / (M.P.H0 / 100.0);
    T[Transfer_cdm] = clxc;
    T[Transfer_b] = clxb;
    T[Transfer_g] = clxg;
    T[Transfer_r] = clxr;
    double clxnu_all = 0;
    double dgpi = grhor_t * pir + grhog_t * pig;
    if (M.Num_Nu_massive != 0) MassiveNuVarsOut_transfer(EV, ay, a, clxnu_all, dgpi);
    if (gal) {
      const double dgpigal_t = gal_Pigal(M, a, hub, xgal, dh, dx, dgrho, dgq, dgpi, etak, dphi, k);
      dgpi = dgpi + dgpigal_t;
    }
    T[Transfer_nu] = clxnu_all;
    T[Transfer_tot] = dgrho_matter / grho_matter;
    T[Transfer_nonu] = (grhob_t * clxb + grhoc_t * clxc) / (grhob_t + grhoc_t);
    T[Transfer_tot_de] = dgrho / grho_matter;
    T[Transfer_Weyl] = -(dgrho + 3 * dgq * adotoa / k) / (1.0 * 2) - dgpi / 2;
    T[Transfer_Newt_vel_cdm] = -k * sigma / adotoa;
    T[Transfer_Newt_vel_baryon] = -k * (vb + sigma) / adotoa;
    T[Transfer_vel_baryon_cdm] = vb;
  }
  const double clxcdot = -k * z;
  ayprime[3] = clxcdot;
  const double clxbdot = -k * (z + vb);
  ayprime[4] = clxbdot;
  const double clxgdot = -k * (4.0 / 3.0 * z + qg);
  double vbdot, qgdot, pigdot;
  if (EV.TightCoupling) {
    double adotdota = (adotoa * adotoa - gpres) / 2;
    pig = 32.0 / 45 / opacity * k * (sigma + vb);
    double slip = -(2 * adotoa / (1 + pb43) + dopacity / opacity) * (vb - 3.0 / 4 * qg) +
                  (-adotdota * vb - k / 2 * adotoa * clxg + k * (cs2 * clxbdot - clxgdot / 4)) / (opacity * (1 + pb43));
    {  // second_order_tightcoupling = .true.
      double dgs = grhog_t * pig + grhor_t * pir;
      if (gal) {
        double dgpigal_t = gal_Pigal(M, a, hub, xgal, dh, dx, dgrho, dgq, dgs, etak, dphi, k);
        dgs = dgs + dgpigal_t;
      }
      double sigmadot = -2 * adotoa * sigma - dgs / k + etak;
      qgdot = k * (clxg / 4.0 - pig / 2.0) + opacity * slip;
      const double op2 = opacity * opacity;
      pig = 32.0 / 45 / opacity * k * (sigma + 3.0 * qg / 4.0) * (1 + (dopacity * 11.0 / 6.0 / op2)) +
            (32.0 / 45.0 / op2) * k * (sigmadot + 3.0 * qgdot / 4.0) * (-11.0 / 6.0);
      pigdot = -(32.0 / 45.0) * (dopacity / op2) * k * (sigma + 3.0 * qg / 4.0) * (1 + dopacity * 11.0 / 6.0 / op2) +
               (32.0 / 45.0 / opacity) * k * (sigmadot + 3.0 * qgdot / 4.0) * (1 + (11.0 / 6.0) * (dopacity / op2));
      EV.pigdot = pigdot;
    }
    vbdot = (-adotoa * vb + cs2 * k * clxb + k / 4 * pb43 * (clxg - 2 * EV.Kf[1] * pig)) / (1 + pb43);
    vbdot = vbdot + pb43 / (1 + pb43) * slip;
    EV.pig = pig;
  } else {
    vbdot = -adotoa * vb + cs2 * k * clxb - photbar * opacity * (4.0 / 3 * vb - qg);
  }
  ayprime[5] = vbdot;
  if (!EV.no_phot_multpoles) {
    ayprime[EV.g_ix] = clxgdot;
    qgdot = 4.0 / 3 * (-vbdot - adotoa * vb + cs2 * k * clxb) / pb43 + EV.denlk[1] * clxg - EV.denlk2[1] * pig;
    ayprime[EV.g_ix + 1] = qgdot;
    if (!EV.TightCoupling) {
      double E2 = ay[EV.polind + 2];
      double polter = pig / 10 + 9.0 / 15 * E2;
      int ix = EV.g_ix + 2;
      if (EV.lmaxg > 2) {
        pigdot = EV.denlk[2] * qg - EV.denlk2[2] * ay[ix + 1] - opacity * (pig - polter) + 8.0 / 15.0 * k * sigma;
        ayprime[ix] = pigdot;
        for (int l = 3; l <= EV.lmaxg - 1; l++) {
          ix = ix + 1;
          ayprime[ix] = (EV.denlk[l] * ay[ix - 1] - EV.denlk2[l] * ay[ix + 1]) - opacity * ay[ix];
        }
        ix = ix + 1;
        ayprime[ix] = k * ay[ix - 1] - (EV.lmaxg + 1) * cothxor * ay[ix] - opacity * ay[ix];
      } else {
        pigdot = EV.denlk[2] * qg - opacity * (pig - polter) + 8.0 / 15.0 * k * sigma;
        ayprime[ix] = pigdot;
      }
      ix = EV.polind + 2;
      if (EV.lmaxgpol > 2) {
        ayprime[ix] = -opacity * (ay[ix] - polter) - k / 3.0 * ay[ix + 1];
        for (int l = 3; l <= EV.lmaxgpol - 1; l++) {
          ix = ix + 1;
          ayprime[ix] = -opacity * ay[ix] + (EV.denlk[l] * ay[ix - 1] - EV.polfack[l] * ay[ix + 1]);
        }
        ix = ix + 1;
        ayprime[ix] = -opacity * ay[ix] + k * EV.poltruncfac * ay[ix - 1] - (EV.lmaxgpol + 3) * cothxor * ay[ix];
      } else {
        ayprime[ix] = -opacity * (ay[ix] - polter);
      }
    }
  }
  // clxrdot is an uninitialised local in the reference when no_nu_multpoles (it is only assigned
  // inside the block below, :2482, but read at :2517).  The oracle pins it to 0, the value the
  // assignment would give under the approximation qr = -4z/3 used in that regime.
  double clxrdot = 0;
  if (!EV.no_nu_multpoles) {
    clxrdot = -k * (4.0 / 3.0 * z + qr);
    ayprime[EV.r_ix] = clxrdot;
    double qrdot = EV.denlk[1] * clxr - EV.denlk2[1] * pir;
    ayprime[EV.r_ix + 1] = qrdot;
    if (EV.high_ktau_neutrino_approx) {
      double pirdot = -3 * pir * cothxor - clxrdot;
      ayprime[EV.r_ix + 2] = pirdot;
    } else {
      int ix = EV.r_ix + 2;
      if (EV.lmaxnr > 2) {
        double pirdot = EV.denlk[2] * qr - EV.denlk2[2] * ay[ix + 1] + 8.0 / 15.0 * k * sigma;
        ayprime[ix] = pirdot;
        for (int l = 3; l <= EV.lmaxnr - 1; l++) {
          ix = ix + 1;
          ayprime[ix] = (EV.denlk[l] * ay[ix - 1] - EV.denlk2[l] * ay[ix + 1]);
        }
        ix = ix + 1;
        ayprime[ix] = k * ay[ix - 1] - (EV.lmaxnr + 1) * cothxor * ay[ix];
      } else {
        double pirdot = EV.denlk[2] * qr + 8.0 / 15.0 * k * sigma;
        ayprime[ix] = pirdot;
      }
    }
  }
  if (gal) {
    double dotdeltaf = grhob_t * (clxbdot - 3 * adotoa * clxb) + grhoc_t * (clxcdot - 3 * adotoa * clxc) +
                       grhor_t * (clxrdot - 4 * adotoa * clxr) + grhog_t * (clxgdot - 4 * adotoa * clxg);
    ayprime[EV.w_ix] = dphiprime;
    ayprime[EV.w_ix + 1] = gal_dphisecond(M, a, hub, xgal, dh, dx, dgrho, dgq, etak, dphi, dphiprime, k, dotdeltaf);
  }
  double flop_nu = 0;
  if (M.Num_Nu_massive != 0) {
    const MassiveNuTables& T = M.nu;
    for (int nu_i = 1; nu_i <= M.Nu_mass_eigenstates; nu_i++) {
      if (EV.MassiveNuApprox[nu_i]) {
        double G11_t = EV.G11[nu_i] / a / a2;
        double G30_t = EV.G30[nu_i] / a / a2;
        int off_ix = EV.nu_ix[nu_i];
        double w = wnu_arr[nu_i];
        ayprime[off_ix] = -k * z * (w + 1) + 3 * adotoa * (w * ay[off_ix] - ay[off_ix + 1]) - k * ay[off_ix + 2];
        ayprime[off_ix + 1] = (3 * w - 2) * adotoa * ay[off_ix + 1] - 5.0 / 3 * k * z * w - k / 3 * G11_t;
        ayprime[off_ix + 2] =
            (3 * w - 1) * adotoa * ay[off_ix + 2] - k * (2.0 / 3 * EV.Kf[1] * ay[off_ix + 3] - ay[off_ix + 1]);
        ayprime[off_ix + 3] =
            (3 * w - 2) * adotoa * ay[off_ix + 3] + 2 * w * k * sigma - k / 5 * (3 * EV.Kf[2] * G30_t - 2 * G11_t);
        flop_nu += 90 + 50;
      } else {
        int ind = EV.nu_ix[nu_i];
        for (int i = 1; i <= EV.nq[nu_i]; i++) {
          double q = T.nu_q[i];
          double aq = a * M.nu_masses[nu_i] / q;
          double v = 1.0 / std::sqrt(1.0 + aq * aq);
          ayprime[ind] = -k * (4.0 / 3.0 * z + v * ay[ind + 1]);
          ind = ind + 1;
          ayprime[ind] = v * (EV.denlk[1] * ay[ind - 1] - EV.denlk2[1] * ay[ind + 1]);
          ind = ind + 1;
          if (EV.lmaxnu_tau[nu_i] == 2) {
            ayprime[ind] = -ayprime[ind - 2] - 3 * cothxor * ay[ind];
          } else {
            ayprime[ind] = v * (EV.denlk[2] * ay[ind - 1] - EV.denlk2[2] * ay[ind + 1]) + k * 8.0 / 15.0 * sigma;
            for (int l = 3; l <= EV.lmaxnu_tau[nu_i] - 1; l++) {
              ind = ind + 1;
              ayprime[ind] = v * (EV.denlk[l] * ay[ind - 1] - EV.denlk2[l] * ay[ind + 1]);
            }
            ind = ind + 1;
            ayprime[ind] = k * v * ay[ind - 1] - (EV.lmaxnu_tau[nu_i] + 1) * cothxor * ay[ind];
          }
          ind = ind + 1;
        }
        flop_nu += 90 + 14 * T.nqmax + 6 * EV.nq[nu_i] * (EV.lmaxnu_tau[nu_i] + 1);
      }
    }
    if (EV.has_nu_relativistic) {
      int ind = EV.nu_pert_ix;
      ayprime[ind] = +k * a2 * qr - k * ay[ind + 1];
      int ind2 = EV.r_ix;
      for (int l = 1; l <= EV.lmaxnu_pert - 1; l++) {
        ind = ind + 1;
        ind2 = ind2 + 1;
        ayprime[ind] = -a2 * (EV.denlk[l] * ay[ind2 - 1] - EV.denlk2[l] * ay[ind2 + 1]) +
                       (EV.denlk[l] * ay[ind - 1] - EV.denlk2[l] * ay[ind + 1]);
      }
      ind = ind + 1;
      ind2 = ind2 + 1;
      ayprime[ind] = k * (ay[ind - 1] - a2 * ay[ind2 - 1]) - (EV.lmaxnu_pert + 1) * cothxor * ay[ind];
      flop_nu += 5 * (EV.lmaxnu_pert + 1);
    }
  }
  if (EV.cnt) {
    // canonical per-call FLOP table, SURVEY.md section 8(d)
    double F = 220 + 45;
    if (EV.TightCoupling) F += 25 + 140;
    if (!EV.no_phot_multpoles && !EV.TightCoupling) F += 7 * (EV.lmaxg - 2) + 7 * (EV.lmaxgpol - 2);
    if (!EV.no_nu_multpoles) F += EV.high_ktau_neutrino_approx ? 8 : 5 * (EV.lmaxnr - 2) + 8;
    F += gal ? (1150 + (EV.TightCoupling ? 160 : 0)) : 10;
    F += flop_nu;
    EV.cnt->n_derivs += 1;
    EV.cnt->sum_n_trial += n;
    EV.cnt->flop_ode += F + 90.0 * n / 8.0;
  }
}

static void derivs_cb(void* ctx, int n, double tau, const double* y, double* yp) {
  derivs(*(EvolutionVars*)ctx, n, tau, y - 1, yp - 1);
}

// camb/equations_galileon.f90:1256-1576.  sources 1-based [1..S]; j = time-step index.
static void output(EvolutionVars& EV, const double* y, int j, double tau, double* sources, int SourceNum,
                   std::vector<double>& yprime_buf) {
  const Model& M = *EV.M;
  const Thermo& T = M.th;
  const bool gal = M.P.use_galileon != 0;
  double* yprime = yprime_buf.data();
  for (int i = 0; i <= EV.nvar; i++) yprime[i] = 0;
  derivs(EV, EV.ScalEqsToPropagate, tau, y, yprime);
  double pol[4] = {0, 0, 0, 0}, polprime[4] = {0, 0, 0, 0};
  const bool localpol = EV.TightCoupling || EV.no_phot_multpoles;
  auto ypol = [&](int l) -> double { return localpol ? pol[l] : y[EV.polind + l]; };
  auto ypolprime = [&](int l) -> double { return localpol ? polprime[l] : yprime[EV.polind + l]; };
  const double k = EV.k_buf, k2 = EV.k2_buf;
  const double a = y[1], a2 = a * a, etak = y[2], clxc = y[3], clxb = y[4], vb = y[5], vbdot = yprime[5];
  double dphi = 0, dphiprime = 0;
  if (gal) {
    dphi = y[EV.w_ix];
    dphiprime = y[EV.w_ix + 1];
  }
  const double grhob_t = M.grhob / a, grhoc_t = M.grhoc / a, grhor_t = M.grhornomass / a2, grhog_t = M.grhog / a2;
  const double grhov_t = M.grhov * std::pow(a, (double)(-1 - 3 * (-1)));
  double grho = grhob_t + grhoc_t + grhor_t + grhog_t;
  double gpres = (grhog_t + grhor_t) / 3;
  double xgal = 0, hub = 0, dh = 0, dx = 0;
  if (gal) {
    xgal = gal_GetX(M, a);
    hub = a * gal_GetH(M, a);
    gal_GetdHdX(M, a, hub, xgal, dh, dx);
    double grhogal_t = gal_grhogal(M, a, hub, xgal);
    double gpresgal_t = gal_gpresgal(M, a, hub, xgal, dh, dx);
    grho = grho + grhogal_t;
    gpres = gpres + gpresgal_t;
  } else {
    grho = grho + grhov_t;
    gpres = gpres + grhov_t * (-1.0);
  }
  double dgrho = grhob_t * clxb + grhoc_t * clxc;
  double dgq = grhob_t * vb;
  double dgpi = 0, dgpi_diff = 0, pidot_sum = 0;
  if (M.Num_Nu_massive != 0) MassiveNuVarsOut(EV, y, yprime, a, grho, gpres, dgrho, dgq, dgpi, dgpi_diff, pidot_sum);
  const double adotoa = 1. / (a * dtauda(M, a));
  double z, dz, clxr, qr, pir, pirdot, clxg, qg, pig, pigdot = 0, octg, octgprime, qgdot, dgrhogal_t;
  if (EV.no_nu_multpoles) {
    if (gal) {
      dgrhogal_t = gal_Chigal(M, a, hub, xgal, dgrho, etak, dphi, dphiprime, k);
      z = (0.5 * (dgrho + dgrhogal_t) / k + etak) / adotoa;
      dz = -adotoa * z - 0.5 * (dgrho + dgrhogal_t) / k;
    } else {
      z = (0.5 * dgrho / k + etak) / adotoa;
      dz = -adotoa * z - 0.5 * dgrho / k;
    }
    clxr = -4 * dz / k;
    qr = -4.0 / 3 * z;
    pir = 0;
    pirdot = 0;
  } else {
    clxr = y[EV.r_ix];
    qr = y[EV.r_ix + 1];
    pir = y[EV.r_ix + 2];
    pirdot = yprime[EV.r_ix + 2];
  }
  if (EV.no_phot_multpoles) {
    if (gal) {
      dgrhogal_t = gal_Chigal(M, a, hub, xgal, dgrho, etak, dphi, dphiprime, k);
      z = (0.5 * (dgrho + dgrhogal_t) / k + etak) / adotoa;
      dz = -adotoa * z - 0.5 * (dgrho + dgrhogal_t) / k;
    } else {
      z = (0.5 * dgrho / k + etak) / adotoa;
      dz = -adotoa * z - 0.5 * dgrho / k;
    }
    clxg = -4 * dz / k - 4 / k * T.opac[j] * (vb + z);
    qg = -4.0 / 3 * z;
    pig = 0;
    pigdot = 0;
    octg = 0;
    octgprime = 0;
    qgdot = -4 * dz / 3;
  } else {
    if (EV.TightCoupling) {
      pig = EV.pig;
      octg = (3.0 / 7.0) * pig * (EV.k_buf / T.opac[j]);
      pol[2] = EV.pig / 4 + EV.pigdot * (1.0 / T.opac[j]) * (-5.0 / 8.0);
      pol[3] = (3.0 / 7.0) * (EV.k_buf / T.opac[j]) * pol[2];
      octgprime = 0;
    } else {
      pig = y[EV.g_ix + 2];
      pigdot = yprime[EV.g_ix + 2];
      octg = y[EV.g_ix + 3];
      octgprime = yprime[EV.g_ix + 3];
    }
    clxg = y[EV.g_ix];
    qg = y[EV.g_ix + 1];
    qgdot = yprime[EV.g_ix + 1];
  }
  dgrho = dgrho + grhog_t * clxg + grhor_t * clxr;
  dgq = dgq + grhog_t * qg + grhor_t * qr;
  dgpi = dgpi + grhor_t * pir + grhog_t * pig;
  double dgpigal_t = 0;
  if (gal) {
    dgrhogal_t = gal_Chigal(M, a, hub, xgal, dgrho, etak, dphi, dphiprime, k);
    dgrho = dgrho + dgrhogal_t;
    double dgqgal_t = gal_qgal(M, a, hub, xgal, dgq, etak, dphi, dphiprime, k);
    dgq = dgq + dgqgal_t;
    dgpigal_t = gal_Pigal(M, a, hub, xgal, dh, dx, dgrho, dgq, dgpi, etak, dphi, k);
    dgpi = dgpi + dgpigal_t;
  }
  z = (0.5 * dgrho / k + etak) / adotoa;
  const double sigma = (z + 1.5 * dgq / k2) / EV.Kf[1];
  const double polter = 0.1 * pig + 9.0 / 15.0 * ypol(2);
  const double x = k * (M.tau0 - tau);
  const double divfac = x * x;
  if (EV.TightCoupling) {
    pigdot = EV.pigdot;
    polprime[2] = (pigdot / 4.0) * (1 + (5.0 / 2.0) * (T.dopac[j] / (T.opac[j] * T.opac[j])));
  }
  pidot_sum = pidot_sum + grhog_t * pigdot + grhor_t * pirdot;
  double diff_rhopi;
  if (gal) {
    diff_rhopi = pidot_sum - (4 * (dgpi - dgpigal_t) + dgpi_diff) * adotoa;
    double pigaldot = gal_pigalprime(M, a, hub, xgal, dh, dx, dgrho, dgq, dgpi, diff_rhopi, etak, dphi, dphiprime, k,
                                     grho, gpres);
    diff_rhopi = diff_rhopi + pigaldot;
  } else {
    diff_rhopi = pidot_sum - (4 * dgpi + dgpi_diff) * adotoa;
  }
  const double Kf1 = EV.Kf[1], Kf2 = EV.Kf[2];
  const double kk = k * k;
  const double ISW = (4.0 / 3.0 * k * Kf1 * sigma + (-2.0 / 3.0 * sigma - 2.0 / 3.0 * etak / adotoa) * k -
                      diff_rhopi / kk - 1.0 / adotoa * dgrho / 3.0 + (3.0 * gpres + 5.0 * grho) * sigma / k / 3.0 -
                      2.0 / k * adotoa / Kf1 * etak) *
                     T.expmmu[j];
  sources[1] =
      ISW +
      ((-9.0 / 160.0 * pig - 27.0 / 80.0 * ypol(2)) / kk * T.opac[j] +
       (11.0 / 10.0 * sigma - 3.0 / 8.0 * Kf2 * ypol(3) + vb - 9.0 / 80.0 * Kf2 * octg + 3.0 / 40.0 * qg) / k -
       (-180.0 * ypolprime(2) - 30.0 * pigdot) / kk / 160.0) *
          T.dvis[j] +
      (-(9.0 * pigdot + 54.0 * ypolprime(2)) / kk * T.opac[j] / 160.0 + pig / 16.0 + clxg / 4.0 + 3.0 / 8.0 * ypol(2) +
       (-21.0 / 5.0 * adotoa * sigma - 3.0 / 8.0 * Kf2 * ypolprime(3) + vbdot + 3.0 / 40.0 * qgdot -
        9.0 / 80.0 * Kf2 * octgprime) /
           k +
       (-9.0 / 160.0 * T.dopac[j] * pig - 21.0 / 10.0 * dgpi - 27.0 / 80.0 * T.dopac[j] * ypol(2)) / kk) *
          T.vis[j] +
      (3.0 / 16.0 * T.ddvis[j] * pig + 9.0 / 8.0 * T.ddvis[j] * ypol(2)) / kk + 21.0 / 10.0 / k / Kf1 * T.vis[j] * etak;
  if (x > 0.0)
    sources[2] = T.vis[j] * polter * (15.0 / 8.0) / divfac;
  else
    sources[2] = 0;
  if (SourceNum > 2) {
    if (tau > M.tau_maxvis && M.tau0 - tau > 0.1) {
      double phi = -(dgrho + 3 * dgq * adotoa / k) / (k2 * Kf1 * 2) - dgpi / k2 / 2;
      sources[3] = -2 * phi * (tau - M.tau_maxvis) / ((M.tau0 - M.tau_maxvis) * (M.tau0 - tau));
    } else {
      sources[3] = 0;
    }
  }
  if (EV.cnt) {
    EV.cnt->n_out += 1;
    EV.cnt->flop_ode += 260 + (gal ? 1700 : 0);
  }
}

// camb/equations_galileon.f90:1714-1928 (adiabatic mode; y 1-based, zeroed here)
static void initial(EvolutionVars& EV, double* y, double tau) {
  const Model& M = *EV.M;
  const Params& P = M.P;
  EV.k_buf = EV.q;
  EV.k2_buf = EV.q2;
  for (int l = 1; l <= EV.MaxlNeeded; l++) EV.Kf[l] = 1.0;
  const double k = EV.k_buf;
  for (int j = 1; j <= EV.MaxlNeeded; j++) {
    EV.denlk[j] = M.denl[j] * k * j;
    EV.denlk2[j] = M.denl[j] * k * EV.Kf[j] * (j + 1);
    EV.polfack[j] = M.polfac[j] * k * EV.Kf[j] * M.denl[j];
  }
  double ep;
  if (EV.k_buf > M.epsw) {
    if (EV.k_buf > M.epsw * 5) {
      ep = ep0 * 5 / P.AccuracyBoost;
      if (P.HighAccuracyDefault) ep = ep * SP(0.65);
    } else {
      ep = ep0;
    }
  } else {
    ep = ep0;
  }
  ep = ep * 2;  // second_order_tightcoupling
  EV.TightSwitchoffTime = std::fmin(M.th.tight_tau, Thermo_OpacityToTime(M, EV.k_buf / ep));
  for (int i = 0; i <= EV.nvar; i++) y[i] = 0;
  const double x = k * tau, x2 = x * x, x3 = x2 * x;
  double rhomass = 0;
  for (int i = 1; i <= M.Nu_mass_eigenstates; i++) rhomass += M.grhormass[i];
  const double grhonu = rhomass + M.grhornomass;
  const double om = (M.grhob + M.grhoc) / std::sqrt(3 * (M.grhog + grhonu));
  const double omtau = om * tau;
  const double Rv = grhonu / (grhonu + M.grhog);
  const double Rp15 = 4 * Rv + 15;
  const double a = tau * M.adotrad * (1 + omtau / 4);
  const double chi = 1, Kf1 = EV.Kf[1];
  const double i_clxg = -chi * Kf1 / 3 * x2 * (1 - omtau / 5);
  const double i_clxr = i_clxg;
  const double i_clxb = 0.75 * i_clxg;
  const double i_clxc = i_clxb;
  const double i_qg = i_clxg * x / 9.0;
  const double i_qr = -chi * Kf1 * (4 * Rv + 23) / Rp15 * x3 / 27;
  const double i_vb = 0.75 * i_qg;
  const double i_pir = chi * 4.0 / 3 * x2 / Rp15 * (1 + omtau / 4 * (4 * Rv - 5) / (2 * Rv + 15));
  const double i_aj3r = chi * 4 / 21.0 / Rp15 * x3;
  const double i_eta = -chi * 2 * Kf1 * (1 - x2 / 12 * (-10.0 / Rp15 + Kf1));
  // InitVec = -initv(1,:)
  y[1] = a;
  y[2] = -(-i_eta) * k / 2;
  y[3] = -i_clxc;
  y[4] = -i_clxb;
  y[5] = -i_vb;
  y[EV.g_ix] = -i_clxg;
  y[EV.g_ix + 1] = -i_qg;
  y[EV.r_ix] = -i_clxr;
  y[EV.r_ix + 1] = -i_qr;
  y[EV.r_ix + 2] = -i_pir;
  if (EV.lmaxnr > 2) y[EV.r_ix + 3] = -i_aj3r;
  if (M.Num_Nu_massive == 0) return;
  for (int nu_i = 1; nu_i <= M.Nu_mass_eigenstates; nu_i++) {
    EV.MassiveNuApproxTime[nu_i] = M.nu_tau_massive[nu_i];
    double a_massive = 20000 * k / M.nu_masses[nu_i] * P.AccuracyBoost * (double)P.lAccuracyBoost;
    if (a_massive >= SP(0.99)) {
      EV.MassiveNuApproxTime[nu_i] = M.tau0 + 1;
    } else if (a_massive > 17.0 / M.nu_masses[nu_i] * P.AccuracyBoost) {
      EV.MassiveNuApproxTime[nu_i] = std::fmax(EV.MassiveNuApproxTime[nu_i], DeltaTime(M, 0.0, a_massive, 0.01));
    }
    int ind = EV.nu_ix[nu_i];
    for (int i = 1; i <= EV.nq[nu_i]; i++) {
      y[ind] = y[EV.r_ix];
      y[ind + 1] = y[EV.r_ix + 1];
      y[ind + 2] = y[EV.r_ix + 2];
      if (EV.lmaxnu_tau[nu_i] > 2) y[ind + 3] = -i_aj3r;
      ind = ind + EV.lmaxnu_tau[nu_i] + 1;
    }
  }
}

// camb/equations_galileon.f90:949-990
static void SwitchToMassiveNuApprox(EvolutionVars& EV, double* y, int nu_i, std::vector<double>& ybuf) {
  const Model& M = *EV.M;
  double a = y[1], a2 = a * a;
  EvolutionVars EVout = EV;
  EVout.MassiveNuApprox[nu_i] = true;
  SetupScalarArrayIndices(EVout);
  double* yout = ybuf.data();
  CopyScalarVariableArray(y, yout, EV, EVout);
  double rhonu, pnu, clxnu, qnu, dpnu, pinu;
  Nu_background(M, a * M.nu_masses[nu_i], rhonu, pnu);
  Nu_Integrate_L012(EV, y, a, nu_i, clxnu, qnu, &dpnu, &pinu);
  dpnu = dpnu / rhonu;
  qnu = qnu / rhonu;
  clxnu = clxnu / rhonu;
  pinu = pinu / rhonu;
  yout[EVout.nu_ix[nu_i]] = clxnu;
  yout[EVout.nu_ix[nu_i] + 1] = dpnu;
  yout[EVout.nu_ix[nu_i] + 2] = qnu;
  yout[EVout.nu_ix[nu_i] + 3] = pinu;
  Nu_Intvsq(EV, y, a, nu_i, EVout.G11[nu_i], EVout.G30[nu_i]);
  EVout.G11[nu_i] = EVout.G11[nu_i] * a2 * a / rhonu;
  EVout.G30[nu_i] = EVout.G30[nu_i] * a2 * a / rhonu;
  EV = EVout;
  for (int i = 1; i <= EV.nvar; i++) y[i] = yout[i];
}

// camb/equations_galileon.f90:278-296
static void GaugeInterface_ScalEv(EvolutionVars& EV, double* y, double& tau, double tauend, double tol1, int& ind,
                                  double* c, double* w, Model& Mw) {
  dverk(&EV, EV.ScalEqsToPropagate, derivs_cb, tau, y + 1, tauend, tol1, ind, c, EV.nvar, w);
  if (ind == -3) GlobalError(Mw, "Dverk error -3", error_evolution);
}

// camb/equations_galileon.f90:314-469 (tail recursion written as a loop)
static void GaugeInterface_EvolveScal(EvolutionVars& EV, double& tau, double* y, double tauend, double tol1, int& ind,
                                      double* c, double* w, Model& Mw, std::vector<double>& ybuf) {
  const Model& M = *EV.M;
  const Params& P = M.P;
  const int nqmax = M.nu.nqmax;
  for (;;) {
    const double noSwitch = M.tau0 + 1;
    const double smallTime = std::fmin(tau, 1 / EV.k_buf) / 100;
    double tau_switch_ktau = noSwitch, tau_switch_no_nu_multpoles = noSwitch, tau_switch_no_phot_multpoles = noSwitch,
           tau_switch_nu_massless = noSwitch, tau_switch_nu_nonrel = noSwitch, tau_switch_nu_massive = noSwitch;
    if (!EV.high_ktau_neutrino_approx && !EV.no_nu_multpoles)
      tau_switch_ktau = std::max(20, EV.lmaxnr - 4) / EV.k_buf;
    if (M.Num_Nu_massive != 0) {
      for (int nu_i = 1; nu_i <= M.Nu_mass_eigenstates; nu_i++) {
        if (EV.nq[nu_i] != nqmax)
          tau_switch_nu_massless =
              std::fmin(tau_switch_nu_massless, M.nu_tau_notmassless[next_nu_nq(M, EV.nq[nu_i])][nu_i]);
        else if (!EV.nu_nonrelativistic[nu_i])
          tau_switch_nu_nonrel = std::fmin(M.nu_tau_nonrelativistic[nu_i], tau_switch_nu_nonrel);
        else if (EV.NuMethod == Nu_trunc && !EV.MassiveNuApprox[nu_i])
          tau_switch_nu_massive = std::fmin(tau_switch_nu_massive, EV.MassiveNuApproxTime[nu_i]);
      }
    }
    if (P.DoLateRadTruncation) {
      if (!EV.no_nu_multpoles)
        tau_switch_no_nu_multpoles =
            std::fmax(15 / EV.k_buf * P.AccuracyBoost, std::fmin(M.taurend, M.th.matter_verydom_tau));
      if (!EV.no_phot_multpoles && (EV.k_buf > SP(0.03) * P.AccuracyBoost))
        tau_switch_no_phot_multpoles = std::fmax(15 / EV.k_buf, M.taurend) * P.AccuracyBoost;
    }
    double next_switch = std::fmin(tau_switch_ktau, tau_switch_nu_massless);
    next_switch = std::fmin(next_switch, EV.TightSwitchoffTime);
    next_switch = std::fmin(next_switch, tau_switch_nu_massive);
    next_switch = std::fmin(next_switch, tau_switch_no_nu_multpoles);
    next_switch = std::fmin(next_switch, tau_switch_no_phot_multpoles);
    next_switch = std::fmin(next_switch, tau_switch_nu_nonrel);
    next_switch = std::fmin(next_switch, noSwitch);
    if (!(next_switch < tauend)) break;
    if (next_switch > tau + smallTime) {
      GaugeInterface_ScalEv(EV, y, tau, next_switch, tol1, ind, c, w, Mw);
      if (Mw.error_flag) return;
    }
    EvolutionVars EVout = EV;
    double* yout = ybuf.data();
    auto relayout = [&]() {
      SetupScalarArrayIndices(EVout);
      CopyScalarVariableArray(y, yout, EV, EVout);
      EV = EVout;
      for (int i = 1; i <= EV.nvar; i++) y[i] = yout[i];
    };
    if (next_switch == EV.TightSwitchoffTime) {
      EVout.TightCoupling = false;
      EVout.TightSwitchoffTime = noSwitch;
      relayout();
      ind = 1;
      y[EV.g_ix + 2] = EV.pig;
      double cs2, opacity, dopacity;
      thermo(M, tau, cs2, opacity, &dopacity);
      const double op2 = opacity * opacity;
      y[EV.g_ix + 3] = (3.0 / 7.0) * y[EV.g_ix + 2] * (EV.k_buf / opacity) * (1.0 + dopacity / op2) +
                       (3.0 / 7.0) * EV.pigdot * (EV.k_buf / op2) * (-1.0);
      y[EV.polind + 2] = EV.pig / 4 + EV.pigdot * (1.0 / opacity) * (-5.0 / 8.0 - (25.0 / 16.0) * dopacity / op2) +
                         EV.pig * ((EV.k_buf / opacity) * (EV.k_buf / opacity)) * (-5.0 / 56.0);
      y[EV.polind + 3] = (3.0 / 7.0) * (EV.k_buf / opacity) * y[EV.polind + 2] * (1.0 + dopacity / op2) +
                         (3.0 / 7.0) * (EV.k_buf / op2) * ((EV.pigdot / 4.0) * (1.0 + (5.0 / 2.0) * dopacity / op2)) *
                             (-1.0);
    } else if (next_switch == tau_switch_ktau) {
      EVout.high_ktau_neutrino_approx = true;
      // (the reference also sets EV%nq = nqmax on the source descriptor, then overwrites it: no effect)
      relayout();
    } else if (next_switch == tau_switch_nu_massless) {
      for (int nu_i = 1; nu_i <= M.Nu_mass_eigenstates; nu_i++) {
        if (EV.nq[nu_i] != nqmax && next_switch == M.nu_tau_notmassless[next_nu_nq(M, EV.nq[nu_i])][nu_i]) {
          EVout.nq[nu_i] = next_nu_nq(M, EV.nq[nu_i]);
          relayout();
          break;
        }
      }
    } else if (next_switch == tau_switch_nu_nonrel) {
      for (int nu_i = 1; nu_i <= M.Nu_mass_eigenstates; nu_i++) {
        if (!EV.nu_nonrelativistic[nu_i] && next_switch == M.nu_tau_nonrelativistic[nu_i]) {
          EVout.nu_nonrelativistic[nu_i] = true;
          relayout();
          break;
        }
      }
    } else if (next_switch == tau_switch_nu_massive) {
      for (int nu_i = 1; nu_i <= M.Nu_mass_eigenstates; nu_i++) {
        if (!EV.MassiveNuApprox[nu_i] && next_switch == EV.MassiveNuApproxTime[nu_i]) {
          SwitchToMassiveNuApprox(EV, y, nu_i, ybuf);
          break;
        }
      }
    } else if (next_switch == tau_switch_no_nu_multpoles) {
      ind = 1;
      EVout.no_nu_multpoles = true;
      for (int i = 1; i <= M.Nu_mass_eigenstates; i++) EVout.nq[i] = nqmax;
      relayout();
    } else if (next_switch == tau_switch_no_phot_multpoles) {
      ind = 1;
      EVout.no_phot_multpoles = true;
      relayout();
    }
  }
  GaugeInterface_ScalEv(EV, y, tau, tauend, tol1, ind, c, w, Mw);
}

// camb/cmbmain.f90:634-660
static double GetTauStart(const Model& M, double q) {
  double taustart = 0.001 / q;
  taustart = std::fmin(taustart, 0.1);
  if (M.Num_Nu_massive > 0) {
    double mx = 0;
    for (int i = 1; i <= M.Nu_mass_eigenstates; i++) mx = std::fmax(mx, M.nu_masses[i]);
    taustart = std::fmin(taustart, 1.e-3 / mx / M.adotrad);
  }
  return taustart;
}
double orc_GetTauStart(const Model& M, double q) { return GetTauStart(M, q); }

// camb/equations_galileon.f90:2078-2095: one derivs call with the transfer tap armed; rows 2..Transfer_max / k^2.
// Arr = 1-based row of Transfer_max entries.
static void outtransf(EvolutionVars& EV, const double* y, double tau, double* Arr, std::vector<double>& yprime_buf) {
  double* yprime = yprime_buf.data();
  for (int i = 0; i <= EV.nvar; i++) yprime[i] = 0;
  EV.k2_buf = EV.q2;  // flat (:2066-2067 of the vector twin; derivs sets the same buffers)
  EV.OutputTransfer = Arr;
  derivs(EV, EV.ScalEqsToPropagate, tau, y, yprime);
  EV.OutputTransfer = nullptr;
  for (int e = Transfer_kh + 1; e <= Transfer_max; e++) Arr[e] = Arr[e] / EV.k2_buf;
}

// DoSourcek + CalcScalarSources, camb/cmbmain.f90:662-689, :917-1071.
// Writes Src(q_ix, 1:S, 1:nt) into M.Src (and, with WantTransfer, TransferData(:, q_ix, :) at the transfer redshifts,
// :1017-1030, :1049-1068).  Thread-safe across distinct q_ix; errors and counters are returned per call and merged
// by the caller.
void CalcScalarSourcesForK(Model& M, int q_ix, Counters& cnt, int& err, std::string& errmsg) {
  EvolutionVars EV;
  EV.M = &M;
  EV.cnt = &cnt;
  EV.q = M.Evolve_q.points[q_ix];
  EV.q2 = EV.q * EV.q;
  EV.q_ix = q_ix;
  const double taustart = GetTauStart(M, EV.q);
  GetNumEqns(EV);
  const int nvar = EV.nvar, S = M.SourceNum, nk = M.Evolve_q.npoints, nt = M.TimeSteps.npoints;
  std::vector<double> w((size_t)nvar * 9, 0.0), y(nvar + 2, 0.0), ybuf(nvar + 2, 0.0), yprime(nvar + 2, 0.0);
  double c[25];
  for (int i = 0; i < 25; i++) c[i] = 0;
  double sources[4];
  initial(EV, y.data(), taustart);
  double tau = taustart;
  int ind = 1;
  double tol1 = base_tol / std::exp(M.P.AccuracyBoost - 1);
  // local error sink (GlobalError writes into a Model); use a tiny Model-like proxy
  struct ErrSink : Model {};
  static thread_local ErrSink* sink = nullptr;
  if (!sink) sink = new ErrSink();
  sink->error_flag = 0;
  const bool WantTransfer = M.P.WantTransfer != 0;
  const int ntf = WantTransfer ? M.P.transfer_num_redshifts : 0;
  const double* tautf = M.tautf.data();  // 1-based
  auto trow = [&](int itf) { return M.TransferData.data() + (size_t)Transfer_max * ((q_ix - 1) + (size_t)M.num_q_trans * (itf - 1)) - 1; };
  int itf = 1;
  if (WantTransfer) {
    if (M.P.transfer_high_precision) tol1 = tol1 / 100;
    while (itf <= ntf && M.TimeSteps.points[2] > tautf[itf]) {  // :1022-1029
      GaugeInterface_EvolveScal(EV, tau, y.data(), tautf[itf], tol1, ind, c, w.data(), *sink, ybuf);
      if (sink->error_flag) {
        err = sink->error_flag;
        errmsg = sink->error_msg;
        return;
      }
      outtransf(EV, y.data(), tau, trow(itf), yprime);
      itf = itf + 1;
    }
  }
  for (int j = 2; j <= nt; j++) {
    const double tauend = M.TimeSteps.points[j];
    if ((EV.q * tauend > M.max_etak_scalar && tauend > M.taurend) && !M.WantLateTime &&
        (!WantTransfer || tau > tautf[ntf])) {
      for (int s = 1; s <= S; s++) M.Src[(q_ix - 1) + (size_t)nk * ((s - 1) + (size_t)S * (j - 1))] = 0;
    } else {
      GaugeInterface_EvolveScal(EV, tau, y.data(), tauend, tol1, ind, c, w.data(), *sink, ybuf);
      if (sink->error_flag) {
        err = sink->error_flag;
        errmsg = sink->error_msg;
        return;
      }
      output(EV, y.data(), j, tau, sources, S, yprime);
      for (int s = 1; s <= S; s++) M.Src[(q_ix - 1) + (size_t)nk * ((s - 1) + (size_t)S * (j - 1))] = sources[s];
      // transfer functions, :1049-1068 (the `goto 101` loop)
      while (WantTransfer && itf <= ntf) {
        if (j < nt) {
          if (tauend < tautf[itf] && M.TimeSteps.points[j + 1] > tautf[itf]) {
            GaugeInterface_EvolveScal(EV, tau, y.data(), tautf[itf], tol1, ind, c, w.data(), *sink, ybuf);
            if (sink->error_flag) {
              err = sink->error_flag;
              errmsg = sink->error_msg;
              return;
            }
          }
        }
        bool again = false;
        if (std::fabs(tau - tautf[itf]) < 1.e-5) {
          outtransf(EV, y.data(), tau, trow(itf), yprime);
          itf = itf + 1;
          if (j < nt)
            if (itf <= ntf && M.TimeSteps.points[j + 1] > tautf[itf]) again = true;
        }
        if (!again) break;
      }
    }
  }
  err = 0;
}

// One iteration of TransferOut's loop + GetTransfer, camb/cmbmain.f90:1164-1177, :1181-1201: a wavenumber beyond the
// C_l source grid, evolved through the transfer redshifts only.
void GetTransferForK(Model& M, int q_ix, Counters& cnt, int& err, std::string& errmsg) {
  EvolutionVars EV;
  EV.M = &M;
  EV.cnt = &cnt;
  EV.TransferOnly = true;
  EV.q = M.q_trans[q_ix];
  EV.q2 = EV.q * EV.q;
  EV.q_ix = q_ix;
  double tau = GetTauStart(M, EV.q);
  GetNumEqns(EV);
  const int nvar = EV.nvar;
  std::vector<double> w((size_t)nvar * 9, 0.0), y(nvar + 2, 0.0), ybuf(nvar + 2, 0.0), yprime(nvar + 2, 0.0);
  double c[25];
  for (int i = 0; i < 25; i++) c[i] = 0;
  double atol = base_tol / std::exp(M.P.AccuracyBoost - 1);
  if (M.P.transfer_high_precision) atol = atol / 10000;
  int ind = 1;
  initial(EV, y.data(), tau);
  struct ErrSink : Model {};
  static thread_local ErrSink* sink = nullptr;
  if (!sink) sink = new ErrSink();
  sink->error_flag = 0;
  for (int i = 1; i <= M.P.transfer_num_redshifts; i++) {
    GaugeInterface_EvolveScal(EV, tau, y.data(), M.tautf[i], atol, ind, c, w.data(), *sink, ybuf);
    if (sink->error_flag) {
      err = sink->error_flag;
      errmsg = sink->error_msg;
      return;
    }
    outtransf(EV, y.data(), tau, M.TransferData.data() + (size_t)Transfer_max * ((q_ix - 1) + (size_t)M.num_q_trans * (i - 1)) - 1,
              yprime);
  }
  err = 0;
}

}  // namespace orc
