// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// CPU oracle: a plain C++ restatement of the scalar/flat CMB transfer path of
// ClementLeloup/cosmomc_galileon (the Galileon fork of CAMB).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, link or call anything in oracle/.  The CUDA product in
// cosmomc_galileon_b200/ never includes or links these files.
//
// PARITY STATUS: the Galileon terms are PINNED against the reference's own object code: oracle/Makefile target
// `ref` compiles the unmodified camb/galileon.cc (where it lies under /root/reference) into oracle/_ref/
// libgalref.so against stand-in GSL headers (oracle/ref/gsl_shim) and tests/test_oracle_ref.py compares
// GetdHdX_, grhogal_, gpresgal_, Chigal_, qgal_, Pigal_, dphisecond_, pigalprime_ (1e-13; incl. the hard zero below
// a = 9.99999e-7, 0/1/3 mass eigenstates) and the whole arrays_ background (bit-identical h(a), x(a), c5 closure)
// with this restatement.  The FORTRAN part of the path (equations_galileon.f90, cmbmain.f90, modules.f90 ...)
// cannot be built in this image (no Fortran compiler) and the tree holds no golden vector for
// the Galileon C_l: that part stays pinned by (i) known-answer tests of each numerical
// primitive (scipy spherical_jn, natural cubic splines, dverk on closed-form ODEs), (ii) the in-code invariants of
// the reference (Friedmann closure, h(1)=x(1)=1, stability checks), (iii) the one C_l file the reference ships for
// this path, data/base_plikHM_TT_lowTEB.minimum.theory_cl (LENSED LCDM): with the lensing post-step
// (orc_lensing.cpp) the oracle's lensed TT agrees with it to < 1.5e-3 at every 10th l in 30..1500, EE to < 4e-3
// (tests/test_oracle_golden.py).  See DESIGN.md.
//
// Third-party arithmetic not vendored in the reference: GSL (version unpinned;
// camb/Makefile_main:64).  gsl_odeiv_step_rkf45 + gsl_odeiv_control_y_new and
// gsl_interp_cspline are restated from their published algorithms in
// orc_background.cpp.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace orc {

// ---------------------------------------------------------------- constants
// camb/constants.f90:11-63
constexpr double const_pi = 3.1415926535897932384626433832795;
constexpr double const_twopi = 2.0 * const_pi;
constexpr double const_fourpi = 4.0 * const_pi;
constexpr double c_light = 2.99792458e8;
constexpr double h_P = 6.62606896e-34;
constexpr double G_newton = 6.6738e-11;
constexpr double sigma_thomson = 6.6524616e-29;
constexpr double sigma_boltz = 5.6704e-8;
constexpr double k_B = 1.3806504e-23;
constexpr double eV = 1.60217646e-19;
constexpr double m_p = 1.672621637e-27;
constexpr double m_H = 1.673575e-27;
constexpr double m_e = 9.10938215e-31;
constexpr double mass_ratio_He_H = 3.9715;
constexpr double Gyr = 3.1556926e16;
constexpr double Mpc = 3.085678e22;
constexpr double MPC_in_sec = Mpc / c_light;
constexpr double barssc0 = k_B / m_p / (c_light * c_light);
constexpr double kappa = 8.0 * const_pi * G_newton;
// a_rad = 8 pi^5 k_B^4 / 15 / c^3 / h_P^3 (camb/constants.f90:43)
inline double a_rad_f() {
  const double pi5 = const_pi * const_pi * const_pi * const_pi * const_pi;
  const double kb4 = k_B * k_B * k_B * k_B;
  return 8.0 * pi5 * kb4 / 15 / (c_light * c_light * c_light) / (h_P * h_P * h_P);
}
inline double Compton_CT_f() {
  return MPC_in_sec * (8.0 / 3.0) * (sigma_thomson / (m_e * c_light)) * a_rad_f();
}
constexpr double COBE_CMBTemp = 2.7255;
// camb/utils.F90:957 (AMLutils pi, used by cmbmain/recfast)
constexpr double pi_aml = 3.14159265358979323846264338328;

// camb/modules.f90:226-231
constexpr double base_tol = 1.0e-4;
constexpr double spl_large = 1.e40;
constexpr int max_nu = 5;
constexpr int lmin = 2;
constexpr int lmax_extrap_highl = 8000;
constexpr int nthermo = 20000;  // camb/modules.f90:2705
constexpr int nrhopn = 2000;    // camb/modules.f90:1554
constexpr int nqmax0 = 80;
constexpr int max_l_evolve = 256;  // camb/equations_galileon.f90:192
constexpr int GAL_NB = 29999;      // camb/galileon.cc:72
constexpr int max_transfer_redshifts = 150;  // camb/modules.f90:64
// camb/modules.f90:1876-1884 (rows of TransferData)
enum {
  Transfer_kh = 1, Transfer_cdm, Transfer_b, Transfer_g, Transfer_r, Transfer_nu, Transfer_tot, Transfer_nonu,
  Transfer_tot_de, Transfer_Weyl, Transfer_Newt_vel_cdm, Transfer_Newt_vel_baryon, Transfer_vel_baryon_cdm,
  Transfer_max = Transfer_vel_baryon_cdm
};

// error ids, camb/constants.f90:70-79
enum {
  error_reionization = 1,
  error_recombination = 2,
  error_inital_power = 3,
  error_evolution = 4,
  error_unsupported_params = 5,
  error_background_evolution = 6,
  error_conservation = 7,
  error_norm_lensing = 8
};

// ---------------------------------------------------------------- numerics
// camb/subroutines.f90
void splini(double* g, int n);
void splder(const double* y, double* dy, int n, const double* g);
void spline(const double* x, const double* y, int n, double d11, double d1n, double* d2);
void splint(const double* y, double* z, int n);
typedef double (*scalar_fn)(void* ctx, double x);
double rombint(scalar_fn f, void* ctx, double a, double b, double tol);
double rombint2(scalar_fn f, void* ctx, double a, double b, double tol, int maxit, int minsteps);
typedef void (*deriv_fn)(void* ctx, int n, double x, const double* y, double* yp);
// c is 1-based (size >= 25); w is nw x 9 column-major (w[(j-1)*nw + (k-1)] = w(k,j)).
void dverk(void* ctx, int n, deriv_fn fcn, double& x, double* y, double xend, double tol, int& ind,
           double* c, int nw, double* w);

// ---------------------------------------------------------------- Ranges (camb/utils.F90:8-485)
struct Region {
  int start_index = 0;
  int steps = 0;
  bool IsLog = false;
  double Low = 0, High = 0, delta = 0, delta_max = 0, delta_min = 0;
};
struct Regions {
  int count = 0;
  int npoints = 0;
  double Lowest = 0, Highest = 0;
  Region R[101];  // 1-based
  bool has_dpoints = false;
  std::vector<double> points, dpoints;  // 1-based (index 0 unused)
};
void Ranges_Init(Regions& R);
void Ranges_Add_delta(Regions& R, double t_start, double t_end, double t_approx_delta, bool IsLog = false);
void Ranges_Add(Regions& R, double t_start, double t_end, int nstep, bool IsLog = false);
void Ranges_GetArray(Regions& R, bool want_dpoints = true);
void Ranges_Getdpoints(Regions& R, bool half_ends = true);
int Ranges_IndexOf(const Regions& R, double tau);

// ---------------------------------------------------------------- parameters
// Subset of CAMBparams (camb/modules.f90:98-167) + the module-level accuracy knobs
// (camb/modules.f90:202-214) needed by the scalar flat path.
struct Params {
  // cosmology
  double H0 = 70, omegab = 0, omegac = 0, omegav = 0, omegan = 0;
  double TCMB = COBE_CMBTemp, YHe = 0.24;
  double Num_Nu_massless = 2.046;
  int Num_Nu_massive = 1;
  int Nu_mass_eigenstates = 1;
  int share_delta_neff = 1;
  double Nu_mass_degeneracies[max_nu] = {0, 0, 0, 0, 0};
  double Nu_mass_fractions[max_nu] = {1, 0, 0, 0, 0};
  int Nu_mass_numbers[max_nu] = {1, 0, 0, 0, 0};
  // Galileon (camb/modules.f90:164-165)
  int use_galileon = 0;
  double c2 = 0, c3 = 0, c4 = 0, cG = 0;
  // primordial (camb/power_tilt.f90:60-92)
  double ScalarPowerAmp = 2.1e-9, an = 0.96, n_run = 0, n_runrun = 0, k_0_scalar = 0.05;
  // reionization (camb/reionization.f90:64-77)
  int Reionization = 1, use_optical_depth = 1;
  double optical_depth = 0.09, redshift = 11, delta_redshift = 1.5, fraction = -1;
  double helium_redshift = 3.5, helium_delta_redshift = 0.5, helium_redshiftstart = 5.0;
  // recombination (camb/recfast.f90:257-282); fudge already includes the Hswitch shift
  double RECFAST_fudge = 1.14 - (1.14 - (1.105 + 0.02));
  double RECFAST_fudge_He = 0.86;
  int RECFAST_Heswitch = 6;
  int RECFAST_Hswitch = 1;
  // run control
  int Max_l = 2500;
  double Max_eta_k = 5000;  // before the HighAccuracyDefault rule of camb/camb.f90:130
  int DoLensing = 0;
  int AccuratePolarization = 1, AccurateReionization = 1;
  int MassiveNuMethod = 1;  // Nu_trunc
  int DoLateRadTruncation = 1;
  int HighAccuracyDefault = 1;
  double AccuracyBoost = 1, lSampleBoost = 1;
  float lAccuracyBoost = 1.f;
  int use_spline_template = 1;
  int nthreads = 0;  // 0: OpenMP default
  // matter transfer functions (TransferParams, camb/modules.f90:75-90); kmax in 1/Mpc, redshifts descending
  int WantTransfer = 0;
  int transfer_high_precision = 0;  // only F is restated
  double transfer_kmax = 0.9;
  int transfer_k_per_logint = 0;
  int transfer_num_redshifts = 1;
  double transfer_redshifts[max_transfer_redshifts] = {0};
  // lensing amplitudes (camb/cmbmain.f90:122, camb/lensing.f90:52; set by source/Calculator_CAMB.f90:130-131)
  double ALens = 1, ALens_Fiducial = 0;
};

// ---------------------------------------------------------------- model state
struct GalBackground {
  // camb/galileon.cc:80-92 globals
  double om = 0, orad = 0, h0 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, cG = 0, c0 = 0, q = 0;
  double noghost_previous = 0, cs2_previous = 0;
  int n = 0;                       // nb+2 spline knots
  std::vector<double> a, h, x;     // ascending a
  std::vector<double> ch, cx;      // gsl cspline "c" arrays (= half second derivative)
  bool ready = false;
};

struct MassiveNuTables {  // camb/modules.f90:1538-1583
  double dlnam = 0;
  std::vector<double> r1, p1, dr1, dp1, ddr1, ddp1;
  int nqmax = 0;
  double nu_q[nqmax0 + 1];           // 1-based
  double nu_int_kernel[nqmax0 + 1];  // 1-based
  bool ready = false;
};

struct Thermo {  // camb/modules.f90:2700-2726
  std::vector<double> tb, cs2, xe, dcs2, dotmu, ddotmu, sdotmu, emmu, demmu, dddotmu, ddddotmu,
      scaleFactor;  // all 1-based, size nthermo+1
  double tauminn = 0, dlntau = 0, Maxtau = 0;
  std::vector<double> vis, dvis, ddvis, expmmu, dopac, opac;  // 1-based over TimeSteps
  double tight_tau = 0, actual_opt_depth = 0, matter_verydom_tau = 0;
};

struct Recfast {  // camb/recfast.f90
  static constexpr int Nz = 10000;
  std::vector<double> zrec, xrec, dxrec;  // 1-based
  // RECDATA scalars
  double Lalpha, Lalpha_He, DeltaB, CDB, DeltaB_He, CDB_He, CB1, CB1_He1, CB1_He2, CR, CK, CK_He, CL,
      CL_He, CT, Bfact, H_frac, fu, mu_H, mu_T, Tnow, HO, OmegaK, OmegaT, z_eq, Nnow, fHe;
  double recombination_saha_z = 0;
};

struct lSamples {
  int l0 = 0;
  std::vector<int> l;  // 1-based
};

struct Counters {  // algorithmic work, counted for the exact input
  double n_trial = 0, n_derivs = 0, n_out = 0, sum_n_trial = 0, flop_ode = 0;
  double n_iter = 0;
};

struct Model {
  Params P;
  // derived (camb/modules.f90:171-193)
  bool flat = true;
  double omegak = 0, curv = 0, r = 1, tau0 = 0, chi0 = 0, scale = 1;
  double grhom = 0, grhog = 0, grhor = 0, grhob = 0, grhoc = 0, grhov = 0, grhornomass = 0, grhok = 0;
  double taurst = 0, dtaurec = 0, taurend = 0, tau_maxvis = 0, adotrad = 0;
  double grhormass[max_nu + 1] = {0};  // 1-based
  double nu_masses[max_nu + 1] = {0};  // 1-based
  double akthom = 0, fHe = 0, Nnow = 0;
  int Nu_mass_eigenstates = 0, Num_Nu_massive = 0;
  double Nu_mass_degeneracies[max_nu + 1] = {0};
  double Max_eta_k = 0;
  // reion history (camb/reionization.f90:79-86)
  double reion_tau_start = 0, reion_tau_complete = 0, reion_fHe = 0, WindowVarMid = 0, WindowVarDelta = 0;
  double reion_fraction = 0, reion_redshift = 0;
  bool reionization = true;
  GalBackground gal;
  MassiveNuTables nu;
  Thermo th;
  Recfast rec;
  Regions TimeSteps, Evolve_q, q_int, BessRanges;
  lSamples lSamp;
  // GaugeInterface module state (camb/equations_galileon.f90:267-274)
  double polfac[max_l_evolve + 1], denl[max_l_evolve + 1];
  double epsw = 0;
  double nu_tau_notmassless[nqmax0 + 2][max_nu + 1];
  double nu_tau_nonrelativistic[max_nu + 1], nu_tau_massive[max_nu + 1];
  // cmbmain module state (camb/cmbmain.f90:85-122)
  bool WantLateTime = false;
  int SourceNum = 2;
  int C_last = 3;
  double qmin = 0, qmax = 0, dtaurec_q = 0, max_etak_scalar = 0;
  int maximum_l = 0;
  double maximum_qeta = 3000;
  int max_bessels_l_index = 1000000;
  double max_bessels_etak = 2000000;
  int kmaxfile = 0, num_xx = 0;
  std::vector<double> Src, ddSrc;  // Fortran Src(nk,S,nt): idx = (k-1) + nk*((s-1) + S*(t-1))
  std::vector<double> ajl, ajlpr;  // (num_xx, nl) column-major
  std::vector<double> Delta_p_l_k; // (S, nl, nq) column-major
  std::vector<double> iCl_scalar;  // (nl, C_last)
  std::vector<double> Cl_scalar;   // (lmin..Max_l, C_last): idx = (l-lmin) + (Max_l-lmin+1)*(type-1)
  std::vector<double> Cl_lensed;   // (lmin..lmax_lensed, 4) TT,EE,BB,TE : idx = (l-lmin) + nLl*(type-1)
  int lmax_lensed = 0;
  std::vector<double> highL_template;  // (lmin..8000, 4) TT,EE,TE,PP ; idx=(l-lmin)+7999*(t-1)
  int error_flag = 0;
  std::string error_msg;
  Counters cnt;
  std::vector<double> per_k_trials;  // per source-k trial-step counts (for load-balance studies)
  // MatterTransferData (camb/modules.f90:1896-1915); TransferData kept in FP64 (the reference stores default real)
  std::vector<double> tautf;         // 1-based, camb/cmbmain.f90:783-793
  int num_q_trans = 0;
  std::vector<double> q_trans;       // 1-based: Evolve_q points, then the transfer-only wavenumbers
  std::vector<double> TransferData;  // (Transfer_max, num_q_trans, nz): idx = (e-1) + Transfer_max*((ik-1) + num_q_trans*(iz-1))
  std::vector<double> sigma_8;       // per transfer redshift
};

inline void GlobalError(Model& M, const char* msg, int id) {
  M.error_flag = id;
  M.error_msg = msg;
}

// ---------------------------------------------------------------- background / tables
int CAMBParams_Set(Model& M);  // camb/modules.f90:260 (+ init_massive_nu, init_background, Reionization_Init)
double dtauda(const Model& M, double a);  // camb/equations_galileon.f90:107
double DeltaTime(const Model& M, double a1, double a2, double in_tol = -1);
void Nu_background(const Model& M, double am, double& rhonu, double& pnu);
void Nu_rho(const Model& M, double am, double& rhonu);
double Nu_drho(const Model& M, double am, double adotoa, double rhonu);
double Nu_dp(const Model& M, double am, double adotoa, double pnu);
// camb/galileon.cc C-ABI terms (arguments by value here)
int gal_arrays(Model& M, double omegar, double omegam, double H0in, double c2in, double c3in, double c4in,
               double cGin);
double gal_GetX(const Model& M, double a);
double gal_GetH(const Model& M, double a);
void gal_GetdHdX(const Model& M, double a, double hcamb, double xcamb, double& dh, double& dx);
double gal_grhogal(const Model& M, double a, double hcamb, double xcamb);
double gal_gpresgal(const Model& M, double a, double hcamb, double xcamb, double dh, double dx);
double gal_Chigal(const Model& M, double a, double hcamb, double xcamb, double dgrho, double eta, double dphi,
                  double dphiprime, double k);
double gal_qgal(const Model& M, double a, double hcamb, double xcamb, double dgq, double eta, double dphi,
                double dphiprime, double k);
double gal_Pigal(const Model& M, double a, double hcamb, double xcamb, double dh, double dx, double dgrho,
                 double dgq, double dgpi, double eta, double dphi, double k);
double gal_dphisecond(const Model& M, double a, double hcamb, double xcamb, double dh, double dx, double dgrho,
                      double dgq, double eta, double dphi, double dphiprime, double k, double deltafprime);
double gal_pigalprime(const Model& M, double a, double hcamb, double xcamb, double dh, double dx, double dgrho,
                      double dgq, double dgpi, double pidot, double eta, double dphi, double dphiprime, double k,
                      double grho, double gpres);
void gal_tracker(const Model& M, double a, double& h, double& x);  // analytic KAT, galileon.cc:457-486

// ---------------------------------------------------------------- thermo
void Recombination_init(Model& M);
double Recombination_xe(const Model& M, double a);
double Reionization_xe(const Model& M, double a, double tau, double xe_recomb, bool have_start);
void Reionization_Init(Model& M);
void inithermo(Model& M, double taumin, double taumax);
void thermo(const Model& M, double tau, double& cs2b, double& opacity, double* dopacity);
double Thermo_OpacityToTime(const Model& M, double opacity);

// ---------------------------------------------------------------- equations
struct EvolutionVars;  // orc_equations.cpp
void GaugeInterface_Init(Model& M);

// ---------------------------------------------------------------- cmbmain
void initlval(Model& M, int max_l);
void cmb_InitVars(Model& M);
void SetkValuesForSources(Model& M);
void CalcAllSources(Model& M);            // hot loop 1
void InitTransfer(Model& M);              // camb/cmbmain.f90:508-632
void TransferOut(Model& M);               // camb/cmbmain.f90:1151-1201
void Transfer_Get_sigmas(Model& M);       // camb/modules.f90:2304-2370, :2450-2472 (R = 8/h Mpc, total matter)
void InitSourceInterpolation(Model& M);
void GenerateBessels(Model& M);
void bjl(int L, double X, double& JL);
void SetkValuesForInt(Model& M);
void IntegrateAll(Model& M);              // hot loop 2
void CalcScalCls(Model& M);
void InterpolateCls(Model& M);
double ScalarPower(const Model& M, double k);
int run_all(Model& M);                    // CAMB_GetResults for the scalar flat path
int run_model(Model& M, double* stage_times, bool reuse_bessels);
int load_highL_template(Model& M, const char* path);
int lens_Cls(Model& M);                   // camb/lensing.f90:78-528 (curved-sky correlation function method)

}  // namespace orc
