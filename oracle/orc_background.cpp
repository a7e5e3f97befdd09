// TEST INFRASTRUCTURE (oracle) -- see orc.hpp header.
// Background stage: CAMBParams_Set, massive-neutrino tables, dtauda/DeltaTime, the Galileon
// background integrator + spline tables + perturbation source terms of camb/galileon.cc, and
// the tanh reionization history of camb/reionization.f90.
//
// GSL substitutes (GSL is an external, unpinned dependency of the reference; Makefile_main:64):
//   * gsl_odeiv_step_rkf45 + gsl_odeiv_control_y_new(1e-18,1e-16) + gsl_odeiv_evolve_apply
//     -> rkf45_* below, restated from the published Fehlberg 4(5) tableau and GSL's documented
//        step-size controller (S=0.9, shrink if r>1.1 by S/r^(1/5) floored at 0.2, grow if
//        r<0.5 by S/r^(1/6) capped to [1,5]).
//   * gsl_interp_cspline (natural) + gsl_spline_eval with a NULL accelerator
//     -> cspline_init / cspline_eval below (tridiagonal solve + plain binary search).
#include "orc.hpp"

namespace orc {

// Fortran default-real literals (no _dl suffix) are single precision; keep their rounding.
#define SP(x) ((double)(x##f))

// ------------------------------------------------------------------ massive neutrinos
// camb/modules.f90:1544-1560
static const double nu_const = 7.0 / 120 * (const_pi * const_pi * const_pi * const_pi);
static const double nu_const2 = 5.0 / 7 / (const_pi * const_pi);
static const double zeta3 = 1.2020569031595942853997;
static const double zeta5 = 1.0369277551433699263313;
static const double zeta7 = 1.0083492773819228268397;
static const double am_min = 0.01;
static const double am_max = 600.0;
static const double am_minp = am_min * SP(1.1);
static const double am_maxp = am_max * SP(0.9);

// camb/modules.f90:1683-1716
static void nuRhoPres(double am, double& rhonu, double& pnu) {
  const double qmax = 30.0;
  const int nq = 100;
  double dum1[nq + 2], dum2[nq + 2];
  double adq = qmax / nq;
  dum1[1] = 0;
  dum2[1] = 0;
  for (int i = 1; i <= nq; i++) {
    double q = i * adq;
    double aq = am / q;
    double v = 1.0 / std::sqrt(1.0 + aq * aq);
    double aqdn = adq * q * q * q / (std::exp(q) + 1.0);
    dum1[i + 1] = aqdn / v;
    dum2[i + 1] = aqdn * v;
  }
  splint(dum1, &rhonu, nq + 1);
  splint(dum2, &pnu, nq + 1);
  rhonu = (rhonu + dum1[nq + 1] / adq) / nu_const;
  pnu = (pnu + dum2[nq + 1] / adq) / nu_const / 3.0;
}

// camb/modules.f90:1595-1680
static void Nu_init(Model& M) {
  const Params& P = M.P;
  for (int i = 1; i <= M.Nu_mass_eigenstates; i++)
    M.nu_masses[i] = nu_const / (1.5 * zeta3) * M.grhom / M.grhor * P.omegan * P.Nu_mass_fractions[i - 1] /
                     M.Nu_mass_degeneracies[i];
  MassiveNuTables& T = M.nu;
  if (T.ready) return;
  T.r1.assign(nrhopn + 1, 0);
  T.p1 = T.dr1 = T.dp1 = T.ddr1 = T.ddp1 = T.r1;
  T.nqmax = 3;
  if (P.AccuracyBoost > 1) T.nqmax = 4;
  if (P.AccuracyBoost > 2) T.nqmax = 5;
  if (P.AccuracyBoost > 3) T.nqmax = (int)std::lround(P.AccuracyBoost * 10);
  for (int i = 0; i <= nqmax0; i++) T.nu_q[i] = T.nu_int_kernel[i] = 0;
  if (T.nqmax == 3) {
    const double q[3] = {SP(0.913201), SP(3.37517), SP(7.79184)};
    const double w[3] = {SP(0.0687359), SP(3.31435), SP(2.29911)};
    for (int i = 0; i < 3; i++) T.nu_q[i + 1] = q[i], T.nu_int_kernel[i + 1] = w[i];
  } else if (T.nqmax == 4) {
    const double q[4] = {SP(0.7), SP(2.62814), SP(5.90428), SP(12.0)};
    const double w[4] = {SP(0.0200251), SP(1.84539), SP(3.52736), SP(0.289427)};
    for (int i = 0; i < 4; i++) T.nu_q[i + 1] = q[i], T.nu_int_kernel[i + 1] = w[i];
  } else if (T.nqmax == 5) {
    const double q[5] = {SP(0.583165), SP(2.0), SP(4.0), SP(7.26582), SP(13.0)};
    const double w[5] = {SP(0.0081201), SP(0.689407), SP(2.8063), SP(2.05156), SP(0.126817)};
    for (int i = 0; i < 5; i++) T.nu_q[i + 1] = q[i], T.nu_int_kernel[i + 1] = w[i];
  } else {
    // dq = (12 + nqmax/5)/real(nqmax): integer division, then default-real conversion
    double dq = (12 + T.nqmax / 5) / (double)(float)T.nqmax;
    for (int i = 1; i <= T.nqmax; i++) {
      double q = (i - 0.5) * dq;
      T.nu_q[i] = q;
      double dlfdlq = -q / (1.0 + std::exp(-q));
      T.nu_int_kernel[i] = dq * q * q * q / (std::exp(q) + 1.0) * (-0.25 * dlfdlq);
    }
  }
  for (int i = 1; i <= nqmax0; i++) T.nu_int_kernel[i] = T.nu_int_kernel[i] / nu_const;
  T.dlnam = -(std::log(am_min / am_max)) / (nrhopn - 1);
  for (int i = 1; i <= nrhopn; i++) {
    double am = am_min * std::exp((i - 1) * T.dlnam);
    double rhonu, pnu;
    nuRhoPres(am, rhonu, pnu);
    T.r1[i] = std::log(rhonu);
    T.p1[i] = std::log(pnu);
  }
  std::vector<double> sd(nrhopn + 1);
  splini(sd.data(), nrhopn);
  splder(T.r1.data(), T.dr1.data(), nrhopn, sd.data());
  splder(T.p1.data(), T.dp1.data(), nrhopn, sd.data());
  splder(T.dr1.data(), T.ddr1.data(), nrhopn, sd.data());
  splder(T.dp1.data(), T.ddp1.data(), nrhopn, sd.data());
  T.ready = true;
}

static inline double hermite4(const std::vector<double>& f, const std::vector<double>& df, int i, double d) {
  // the cubic used at camb/modules.f90:1746-1749 (and :1781, :1808, :1841)
  return f[i] + d * (df[i] + d * (3.0 * (f[i + 1] - f[i]) - 2.0 * df[i] - df[i + 1] +
                                  d * (df[i] + df[i + 1] + 2.0 * (f[i] - f[i + 1]))));
}

// camb/modules.f90:1721-1754
void Nu_background(const Model& M, double am, double& rhonu, double& pnu) {
  const MassiveNuTables& T = M.nu;
  if (am <= am_minp) {
    rhonu = 1.0 + nu_const2 * am * am;
    pnu = (2 - rhonu) / 3.0;
    return;
  } else if (am >= am_maxp) {
    rhonu = 3 / (2 * nu_const) * (zeta3 * am + (15 * zeta5) / 2 / am);
    pnu = 900.0 / 120.0 / nu_const * (zeta5 - 63.0 / 4 * zeta7 / (am * am)) / am;
    return;
  }
  double d = std::log(am / am_min) / T.dlnam + 1.0;
  int i = (int)d;
  d = d - i;
  rhonu = std::exp(hermite4(T.r1, T.dr1, i, d));
  pnu = std::exp(hermite4(T.p1, T.dp1, i, d));
}

// camb/modules.f90:1757-1786
void Nu_rho(const Model& M, double am, double& rhonu) {
  const MassiveNuTables& T = M.nu;
  if (am <= am_minp) {
    rhonu = 1.0 + nu_const2 * am * am;
    return;
  } else if (am >= am_maxp) {
    rhonu = 3 / (2 * nu_const) * (zeta3 * am + (15 * zeta5) / 2 / am);
    return;
  }
  double d = std::log(am / am_min) / T.dlnam + 1.0;
  int i = (int)d;
  d = d - i;
  rhonu = std::exp(hermite4(T.r1, T.dr1, i, d));
}

// camb/modules.f90:1790-1817
double Nu_drho(const Model& M, double am, double adotoa, double rhonu) {
  const MassiveNuTables& T = M.nu;
  if (am < am_minp) return 2 * nu_const2 * am * am * adotoa;
  if (am > am_maxp) return 3 / (2 * nu_const) * (zeta3 * am - (15 * zeta5) / 2 / am) * adotoa;
  double d = std::log(am / am_min) / T.dlnam + 1.0;
  int i = (int)d;
  d = d - i;
  double r = hermite4(T.dr1, T.ddr1, i, d);
  return rhonu * adotoa * r / T.dlnam;
}

// camb/modules.f90:1821-1850
double Nu_dp(const Model& M, double am, double adotoa, double pnu) {
  const MassiveNuTables& T = M.nu;
  if (am < am_minp) return 2 * adotoa * (1.0 - nu_const2 * am * am) / 3.0 - 4.0 / 3.0 * adotoa;
  if (am > am_maxp) return -900.0 / 120.0 / nu_const * (zeta5 - 189.0 / 4 * zeta7 / (am * am)) * adotoa / am;
  double d = std::log(am / am_min) / T.dlnam + 1.0;
  int i = (int)d;
  d = d - i;
  double r = hermite4(T.dp1, T.ddp1, i, d);
  return pnu * adotoa * r / T.dlnam;
}

// ------------------------------------------------------------------ dtauda / times
// camb/equations_galileon.f90:107-156
double dtauda(const Model& M, double a) {
  double a2 = a * a, grhoa2;
  if (M.P.use_galileon && a >= 1e-10) {
    double a4 = a2 * a2;
    double hbar = gal_GetH(M, a);
    grhoa2 = a4 * M.grhom * (hbar * hbar);
  } else {
    grhoa2 = M.grhok * a2 + (M.grhoc + M.grhob) * a + M.grhog + M.grhornomass;
    grhoa2 = grhoa2 + M.grhov * (a2 * a2);  // w_lam = -1
    if (M.Num_Nu_massive != 0) {
      for (int nu_i = 1; nu_i <= M.Nu_mass_eigenstates; nu_i++) {
        double rhonu;
        Nu_rho(M, a * M.nu_masses[nu_i], rhonu);
        grhoa2 = grhoa2 + rhonu * M.grhormass[nu_i];
      }
    }
  }
  return std::sqrt(3 / grhoa2);
}

static double dtauda_cb(void* ctx, double a) { return dtauda(*(const Model*)ctx, a); }

// camb/modules.f90:557-580
double DeltaTime(const Model& M, double a1, double a2, double in_tol) {
  double atol = in_tol > 0 ? in_tol : base_tol / 1000 / std::exp(M.P.AccuracyBoost - 1);
  return rombint(dtauda_cb, (void*)&M, a1, a2, atol);
}

// ------------------------------------------------------------------ reionization
// camb/reionization.f90:64-98 ; xe_recomb absent => xstart = 0
static const double Rionization_zexp = 1.5;
static const double Reionization_maxz = 50.0;
static const double Reionization_tol = 1e-5;

double Reionization_xe(const Model& M, double a, double /*tau*/, double xe_recomb, bool have_start) {
  double xstart = have_start ? xe_recomb : 0.0;
  double xod = (M.WindowVarMid - 1.0 / std::pow(a, Rionization_zexp)) / M.WindowVarDelta;
  double tgh = xod > 100 ? 1.0 : std::tanh(xod);
  double xe = (M.reion_fraction - xstart) * (tgh + 1.0) / 2.0 + xstart;
  if (a > (1 / (1 + M.P.helium_redshiftstart))) {  // include_helium_fullreion = .true.
    xod = (1 + M.P.helium_redshift - 1.0 / a) / M.P.helium_delta_redshift;
    tgh = xod > 100 ? 1.0 : std::tanh(xod);
    xe = xe + M.reion_fHe * (tgh + 1.0) / 2.0;
  }
  return xe;
}

// camb/reionization.f90:141-149
static void Reionization_SetParamsForZre(Model& M) {
  M.WindowVarMid = std::pow(1.0 + M.reion_redshift, Rionization_zexp);
  M.WindowVarDelta = Rionization_zexp * std::pow(1.0 + M.reion_redshift, Rionization_zexp - 1.0) * M.P.delta_redshift;
}

// camb/reionization.f90:261-271
static double doptdepth_dz(void* ctx, double z) {
  const Model& M = *(const Model*)ctx;
  double a = 1.0 / (1.0 + z);
  return Reionization_xe(M, a, 0, 0, false) * M.akthom * dtauda(M, a);
}

// camb/reionization.f90:273-284
static double Reionization_GetOptDepth(Model& M) {
  return rombint2(doptdepth_dz, &M, 0.0, Reionization_maxz, Reionization_tol, 20,
                  (int)std::lround(Reionization_maxz / M.P.delta_redshift * 5));
}

// camb/reionization.f90:286-322
static void Reionization_zreFromOptDepth(Model& M) {
  double try_b = 0, try_t = Reionization_maxz, tau = 0;
  int i = 0;
  for (;;) {
    i++;
    M.reion_redshift = (try_t + try_b) / 2;
    Reionization_SetParamsForZre(M);
    tau = Reionization_GetOptDepth(M);
    if (tau > M.P.optical_depth)
      try_t = M.reion_redshift;
    else
      try_b = M.reion_redshift;
    if (std::fabs(try_b - try_t) < SP(2e-3) / 1.0) break;
    if (i > 100) {
      GlobalError(M, "Reionization_zreFromOptDepth: failed to converge", error_reionization);
      return;
    }
  }
  if (std::fabs(tau - M.P.optical_depth) > SP(0.002))
    GlobalError(M, "Reionization_zreFromOptDepth: Did not converge to optical depth", error_reionization);
}

// camb/reionization.f90:153-208
void Reionization_Init(Model& M) {
  const Params& P = M.P;
  M.reion_fHe = P.YHe / (mass_ratio_He_H * (1.0 - P.YHe));
  M.reion_tau_start = M.tau0;
  M.reion_tau_complete = M.tau0;
  M.reionization = P.Reionization != 0;
  M.reion_fraction = P.fraction;
  M.reion_redshift = P.redshift;
  if (M.reionization) {
    if ((P.use_optical_depth && P.optical_depth < SP(0.001)) || (!P.use_optical_depth && P.redshift < SP(0.001)))
      M.reionization = false;
  }
  if (M.reionization) {
    if (M.reion_fraction == -1.0) M.reion_fraction = 1.0 + M.reion_fHe;
    if (P.use_optical_depth) {
      // Reionization_SetFromOptDepth, :328-367
      M.reion_redshift = 0;
      if (P.optical_depth != 0)
        Reionization_zreFromOptDepth(M);
      else
        M.reionization = false;
      if (M.error_flag) return;
    }
    Reionization_SetParamsForZre(M);
    double astart = 1.0 / (1.0 + M.reion_redshift + P.delta_redshift * 8);
    M.reion_tau_start = std::fmax(0.05, rombint(dtauda_cb, &M, 0.0, astart, 1e-3));
    M.reion_tau_complete =
        std::fmin(M.tau0, M.reion_tau_start +
                              rombint(dtauda_cb, &M, astart,
                                      1.0 / (1.0 + std::fmax(0.0, M.reion_redshift - P.delta_redshift * 8)), 1e-3));
  }
}

// ------------------------------------------------------------------ CAMBParams_Set
// camb/modules.f90:260-484 (scalar-only, flat or not flagged; WantTransfer = F)
int CAMBParams_Set(Model& M) {
  Params& P = M.P;
  M.error_flag = 0;
  M.error_msg.clear();
  // camb/camb.f90:130-131
  M.Max_eta_k = P.Max_eta_k;
  if (P.HighAccuracyDefault) M.Max_eta_k = std::fmax(std::min(P.Max_l, 3000) * 2.5, P.Max_eta_k);
  M.Num_Nu_massive = P.Num_Nu_massive;
  M.Nu_mass_eigenstates = P.Nu_mass_eigenstates;
  double Num_Nu_massless = P.Num_Nu_massless;
  int mass_numbers[max_nu];
  for (int i = 0; i < max_nu; i++) {
    M.Nu_mass_degeneracies[i + 1] = P.Nu_mass_degeneracies[i];
    mass_numbers[i] = P.Nu_mass_numbers[i];
  }
  if (P.omegan == 0 && M.Num_Nu_massive != 0) {
    if (P.share_delta_neff)
      Num_Nu_massless = Num_Nu_massless + M.Num_Nu_massive;
    else {
      double s = 0;
      for (int i = 1; i <= M.Nu_mass_eigenstates; i++) s += M.Nu_mass_degeneracies[i];
      Num_Nu_massless = Num_Nu_massless + s;
    }
    M.Num_Nu_massive = 0;
    for (int i = 0; i < max_nu; i++) mass_numbers[i] = 0;
  }
  double nu_massless_degeneracy = Num_Nu_massless;
  if (M.Num_Nu_massive > 0) {
    if (M.Nu_mass_eigenstates == 0) {
      GlobalError(M, "Have Num_nu_massive>0 but no nu_mass_eigenstates", error_unsupported_params);
      return M.error_flag;
    }
    if (M.Nu_mass_eigenstates == 1 && mass_numbers[0] == 0) mass_numbers[0] = M.Num_Nu_massive;
    bool allzero = true;
    for (int i = 0; i < M.Nu_mass_eigenstates; i++) allzero = allzero && mass_numbers[i] == 0;
    if (allzero)
      for (int i = 0; i < max_nu; i++) mass_numbers[i] = 1;
    if (P.share_delta_neff) {
      double fractional_number = Num_Nu_massless + M.Num_Nu_massive;
      int actual_massless = (int)(Num_Nu_massless + 1e-6);
      double neff_i = fractional_number / (actual_massless + M.Num_Nu_massive);
      nu_massless_degeneracy = neff_i * actual_massless;
      for (int i = 1; i <= M.Nu_mass_eigenstates; i++) M.Nu_mass_degeneracies[i] = mass_numbers[i - 1] * neff_i;
    }
  } else {
    M.Nu_mass_eigenstates = 0;
  }
  M.omegak = 1 - (P.omegab + P.omegac + P.omegav + P.omegan);  // GetOmegak, equations_galileon.f90:51-59
  M.flat = std::fabs(M.omegak) <= 5e-7;
  if (!M.flat) {
    GlobalError(M, "oracle restates the flat path only", error_unsupported_params);
    return M.error_flag;
  }
  M.curv = 0;
  M.r = 1.0;
  const double c = c_light;
  M.grhom = 3 * (P.H0 * P.H0) / (c * c) * (1000.0 * 1000.0);
  M.grhog = kappa / (c * c) * 4 * sigma_boltz / (c * c * c) * (P.TCMB * P.TCMB * P.TCMB * P.TCMB) * (Mpc * Mpc);
  M.grhor = 7.0 / 8 * std::pow(4.0 / 11, 4.0 / 3) * M.grhog;
  M.grhornomass = M.grhor * nu_massless_degeneracy;
  for (int i = 0; i <= max_nu; i++) M.grhormass[i] = 0;
  double sum_rm = 0;
  for (int nu_i = 1; nu_i <= M.Nu_mass_eigenstates; nu_i++) {
    M.grhormass[nu_i] = M.grhor * M.Nu_mass_degeneracies[nu_i];
    sum_rm += M.grhormass[nu_i];
  }
  M.grhoc = M.grhom * P.omegac;
  M.grhob = M.grhom * P.omegab;
  M.grhov = M.grhom * P.omegav;
  M.grhok = M.grhom * M.omegak;
  M.adotrad = std::sqrt((M.grhog + M.grhornomass + sum_rm) / 3);
  M.Nnow = P.omegab * (1 - P.YHe) * M.grhom * (c * c) / kappa / m_H / (Mpc * Mpc);
  M.akthom = sigma_thomson * M.Nnow * Mpc;
  M.fHe = P.YHe / (mass_ratio_He_H * (1.0 - P.YHe));
  // init_massive_nu (modules.f90:1856-1866)
  for (int i = 0; i <= max_nu; i++) M.nu_masses[i] = 0;
  if (P.omegan != 0) Nu_init(M);
  // init_background (equations_galileon.f90:63-104)
  if (P.use_galileon) {
    double grhorad = M.grhog + M.grhornomass;
    double omegar = grhorad / M.grhom;
    double omegam = P.omegab + P.omegac;
    double hub = P.H0 / c * 1000;
    int status = gal_arrays(M, omegar, omegam, hub, P.c2, P.c3, P.c4, P.cG);
    if (status != 0) {
      GlobalError(M, "Didn't compute background successfully", error_background_evolution);
      return M.error_flag;
    }
  }
  M.tau0 = DeltaTime(M, 0.0, 1.0);  // TimeOfz(0)
  Reionization_Init(M);
  if (M.error_flag) return M.error_flag;
  M.chi0 = M.tau0 / M.r;  // rofChi flat
  M.scale = M.chi0 * M.r / M.tau0;
  return 0;
}

// ------------------------------------------------------------------ Galileon background
#define GAL_CONSTS(M)                                                                                       \
  const GalBackground& G_ = (M).gal;                                                                        \
  const double c2 = G_.c2, c3 = G_.c3, c4 = G_.c4, c5 = G_.c5, cG = G_.cG, h0 = G_.h0, orad = G_.orad;      \
  const double h02 = h0 * h0;                                                                               \
  const int nu_n = (M).Nu_mass_eigenstates;                                                                 \
  (void)c2; (void)c3; (void)c4; (void)c5; (void)cG; (void)h0; (void)orad; (void)h02; (void)nu_n

// camb/galileon.cc:254-302 : d(hbar, dpi/da, pi)/da
static void gal_rhs(const Model& M, double a, const double y[3], double f[3]) {
  GAL_CONSTS(M);
  double prod = a * y[0] * y[1];
  double prod2 = prod * prod, prod3 = prod * prod2, prod4 = prod * prod3, prod5 = prod * prod4;
  double h = y[0], h2 = h * h, h3 = h * h2, h4 = h2 * h2, a2 = a * a;
  double alpha = c2 / 6.0 * prod - 3 * c3 * h * prod2 + 15 * c4 * h2 * prod3 - 17.5 * c5 * h3 * prod4 -
                 3.0 * cG * h2 * prod;
  double gamma = c2 / 3.0 * h * prod - c3 * h2 * prod2 + 2.5 * c5 * h4 * prod4 - 2.0 * cG * h3 * prod;
  double beta = c2 / 6.0 * h2 - 2 * c3 * h3 * prod + 9 * c4 * h4 * prod2 - 10 * c5 * h4 * h * prod3 - cG * h4;
  double sigma = 2.0 * h + 2.0 * c3 * prod3 - 15.0 * c4 * h * prod4 + 21.0 * c5 * h2 * prod5 + 6.0 * cG * h * prod2;
  double lambda = 3.0 * h2 + orad / (a2 * a2) + c2 / 2.0 * prod2 - 2.0 * c3 * h * prod3 + 7.5 * c4 * h2 * prod4 -
                  9.0 * c5 * h3 * prod5 - cG * h2 * prod2;
  for (int i = 1; i <= nu_n; i++) {
    double rhonu = 0, pnu = 0, am = a * M.nu_masses[i];
    Nu_background(M, am, rhonu, pnu);
    lambda += M.grhormass[i] * pnu / (a2 * a2 * h0 * h0);
  }
  double omega = 2 * c3 * h2 * prod2 - 12 * c4 * h3 * prod3 + 15 * c5 * h4 * prod4 + 4.0 * cG * h3 * prod;
  double denom = sigma * beta - alpha * omega;
  f[0] = (omega * gamma - lambda * beta) / (a * denom);
  f[1] = (alpha * lambda - sigma * gamma) / (a2 * denom) - 2.0 * y[1] / a;
  f[2] = y[1];
}

// camb/galileon.cc:130-245 : ghost / Laplace stability checks at one output point (c0 = 0)
static int gal_stability(Model& M, double a, const double y[3]) {
  GAL_CONSTS(M);
  GalBackground& G = M.gal;
  const double c0 = 0;
  double xgal = a * y[1];
  double prod = xgal * y[0];
  double prod2 = prod * prod, prod3 = prod * prod2, prod4 = prod * prod3, prod5 = prod * prod4;
  double h = y[0], h2 = h * h, h3 = h * h2, h4 = h2 * h2;
  double alpha = c0 * h + c2 / 6.0 * prod - 3 * c3 * h * prod2 + 15 * c4 * h2 * prod3 - 17.5 * c5 * h3 * prod4 -
                 3.0 * cG * h2 * prod;
  double gamma = 2 * c0 * h2 + c2 / 3.0 * h * prod - c3 * h2 * prod2 + 2.5 * c5 * h4 * prod4 - 2.0 * cG * h3 * prod;
  double beta = -2 * c3 * h3 * prod + c2 / 6.0 * h2 + 9 * c4 * h4 * prod2 - 10 * c5 * h4 * h * prod3 - cG * h4;
  double sigma = 2.0 * (1.0 - 2 * c0 * y[2]) * h - 2.0 * c0 * prod + 2.0 * c3 * prod3 - 15.0 * c4 * h * prod4 +
                 21.0 * c5 * h2 * prod5 + 6.0 * cG * h * prod2;
  double lambda = 3.0 * (1 - 2 * c0 * y[2]) * h2 + orad / (a * a * a * a) - 2.0 * c0 * h * prod -
                  2.0 * c3 * h * prod3 + c2 / 2.0 * prod2 + 7.5 * c4 * h2 * prod4 - 9.0 * c5 * h3 * prod5 -
                  cG * h2 * prod2;
  double omega = -2 * c0 * h2 + 2 * c3 * h2 * prod2 - 12 * c4 * h3 * prod3 + 15 * c5 * h4 * prod4 + 4.0 * cG * h3 * prod;
  double denom = sigma * beta - alpha * omega;
  for (int i = 1; i <= nu_n; i++) {
    double rhonu = 0, pnu = 0, am = a * M.nu_masses[i];
    Nu_background(M, am, rhonu, pnu);
    lambda += M.grhormass[i] * pnu / (a * a * a * a * h0 * h0);
  }
  double dhdlna = (omega * gamma - lambda * beta) / denom;
  double dxdlna = (alpha * lambda - sigma * gamma) / denom - xgal;
  double k1 = -6 * c4 * h * prod2 * (dhdlna * xgal + h * dxdlna + prod / 3.0) - 2 * c0 +
              c5 * h2 * prod3 * (12 * h * dxdlna + 15 * dhdlna * xgal + 3 * prod) +
              2.0 * cG * (dhdlna * prod + h2 * dxdlna + h * prod);
  double k2 = -0.5 * c2 + 6 * c3 * h * prod - 27 * c4 * h2 * prod2 + 30 * c5 * h3 * prod3 + 3.0 * cG * h2;
  double k3 = -(1.0 - 2.0 * c0 * y[2]) - 0.5 * c4 * prod4 - 3 * c5 * h * prod4 * (h * dxdlna + dhdlna * xgal) +
              cG * prod2;
  double k4 = -2.0 * (1.0 - 2.0 * c0 * y[2]) + 3.0 * c4 * prod4 - 6 * c5 * h * prod5 - 2.0 * cG * prod2;
  double k5 = 2 * c3 * prod2 - 12 * c4 * h * prod3 - 2.0 * c0 + 15.0 * c5 * h2 * prod4 + 4.0 * cG * h * prod;
  double k6 = 0.5 * c2 - 2.0 * c3 * (h2 * dxdlna + dhdlna * prod + 2 * h * prod) +
              c4 * (12 * h3 * prod * dxdlna + 18 * h * prod2 * dhdlna + 13 * h2 * prod2) -
              c5 * (18 * h2 * h2 * prod2 * dxdlna + 30 * h2 * prod3 * dhdlna + 12 * h3 * prod3) -
              cG * (2.0 * h * dhdlna + 3.0 * h2);
  double noghost = k2 + 1.5 * k5 * k5 / k4;
  double cs2 = (4 * k1 * k4 * k5 - 2 * k3 * k5 * k5 - 2 * k4 * k4 * k6) / (k4 * (2 * k4 * k2 + 3 * k5 * k5));
  if (a == 1) {
    G.noghost_previous = noghost;
    G.cs2_previous = cs2;
  }
  double noghostt = 0.5 - 0.75 * c4 * prod4 + 1.5 * c5 * prod5 * h + 0.5 * cG * prod2 - c0 * y[2];
  double ct2 = (0.5 + 0.25 * c4 * prod4 + 1.5 * c5 * prod4 * h * (h * dxdlna + dhdlna * xgal) - 0.5 * cG * prod2 -
                c0 * y[2]) /
               (0.5 - 0.75 * c4 * prod4 + 1.5 * c5 * prod5 * h + 0.5 * cG * prod2 - c0 * y[2]);
  if (noghost > -1.0) {
    if (noghost > -1e-8) return 1;
  }
  if (cs2 < 0) return 2;
  if (noghostt < 0) return 3;
  if (ct2 < 0) return 4;
  if (noghost > -1.0 && noghost < -1e-8 && std::fabs((noghost - G.noghost_previous) / noghost) > 0.25) return 1;
  G.noghost_previous = noghost;
  if (std::fabs((cs2 - G.cs2_previous) / G.cs2_previous) > 1) return 2;
  G.cs2_previous = cs2;
  return 0;
}

// Fehlberg 4(5) step; yerr = h * sum ec_i k_i ; y advanced with the 5th-order weights.
static void rkf45_step(const Model& M, double t, double h, double y[3], double yerr[3], const double k1[3],
                       double dydt_out[3]) {
  static const double ah[] = {1.0 / 4.0, 3.0 / 8.0, 12.0 / 13.0, 1.0, 1.0 / 2.0};
  static const double b3[] = {3.0 / 32.0, 9.0 / 32.0};
  static const double b4[] = {1932.0 / 2197.0, -7200.0 / 2197.0, 7296.0 / 2197.0};
  static const double b5[] = {8341.0 / 4104.0, -32832.0 / 4104.0, 29440.0 / 4104.0, -845.0 / 4104.0};
  static const double b6[] = {-6080.0 / 20520.0, 41040.0 / 20520.0, -28352.0 / 20520.0, 9295.0 / 20520.0,
                              -5643.0 / 20520.0};
  static const double cc1 = 902880.0 / 7618050.0, cc3 = 3953664.0 / 7618050.0, cc4 = 3855735.0 / 7618050.0,
                      cc5 = -1371249.0 / 7618050.0, cc6 = 277020.0 / 7618050.0;
  static const double ec[] = {0.0, 1.0 / 360.0, 0.0, -128.0 / 4275.0, -2197.0 / 75240.0, 1.0 / 50.0, 2.0 / 55.0};
  double k2[3], k3[3], k4[3], k5[3], k6[3], yt[3];
  for (int i = 0; i < 3; i++) yt[i] = y[i] + ah[0] * h * k1[i];
  gal_rhs(M, t + ah[0] * h, yt, k2);
  for (int i = 0; i < 3; i++) yt[i] = y[i] + h * (b3[0] * k1[i] + b3[1] * k2[i]);
  gal_rhs(M, t + ah[1] * h, yt, k3);
  for (int i = 0; i < 3; i++) yt[i] = y[i] + h * (b4[0] * k1[i] + b4[1] * k2[i] + b4[2] * k3[i]);
  gal_rhs(M, t + ah[2] * h, yt, k4);
  for (int i = 0; i < 3; i++) yt[i] = y[i] + h * (b5[0] * k1[i] + b5[1] * k2[i] + b5[2] * k3[i] + b5[3] * k4[i]);
  gal_rhs(M, t + ah[3] * h, yt, k5);
  for (int i = 0; i < 3; i++)
    yt[i] = y[i] + h * (b6[0] * k1[i] + b6[1] * k2[i] + b6[2] * k3[i] + b6[3] * k4[i] + b6[4] * k5[i]);
  gal_rhs(M, t + ah[4] * h, yt, k6);
  for (int i = 0; i < 3; i++) {
    const double d_i = cc1 * k1[i] + cc3 * k3[i] + cc4 * k4[i] + cc5 * k5[i] + cc6 * k6[i];
    y[i] += h * d_i;
  }
  gal_rhs(M, t + h, y, dydt_out);
  for (int i = 0; i < 3; i++)
    yerr[i] = h * (ec[1] * k1[i] + ec[3] * k3[i] + ec[4] * k4[i] + ec[5] * k5[i] + ec[6] * k6[i]);
}

// One gsl_odeiv_evolve_apply call: advance (t,y) by one accepted step towards t1.
static void rkf45_evolve_apply(const Model& M, double& t, double t1, double& h, double y[3]) {
  const double eps_abs = 1e-18, eps_rel = 1e-16, S = 0.9, ord = 5.0;
  const double t0 = t;
  double h0 = h;
  const double dt = t1 - t0;
  double y0[3] = {y[0], y[1], y[2]}, yerr[3], dydt_in[3], dydt_out[3];
  gal_rhs(M, t0, y, dydt_in);
  for (;;) {
    bool final_step;
    if ((dt >= 0.0 && h0 > dt) || (dt < 0.0 && h0 < dt)) {
      h0 = dt;
      final_step = true;
    } else
      final_step = false;
    rkf45_step(M, t0, h0, y, yerr, dydt_in, dydt_out);
    t = final_step ? t1 : t0 + h0;
    const double h_old = h0;
    double rmax = 2.2250738585072014e-308;
    for (int i = 0; i < 3; i++) {
      const double D0 = eps_rel * std::fabs(y[i]) + eps_abs;
      const double r = std::fabs(yerr[i]) / std::fabs(D0);
      rmax = std::fmax(r, rmax);
    }
    if (rmax > 1.1) {
      double r = S / std::pow(rmax, 1.0 / ord);
      if (r < 0.2) r = 0.2;
      h0 = r * h_old;
      const double t_curr = t, t_next = t + h0;
      if (std::fabs(h0) < std::fabs(h_old) && t_next != t_curr) {
        for (int i = 0; i < 3; i++) y[i] = y0[i];
        continue;  // retry from (t0,y0)
      }
      h0 = h_old;
    } else if (rmax < 0.5) {
      double r = S / std::pow(rmax, 1.0 / (ord + 1.0));
      if (r > 5.0) r = 5.0;
      if (r < 1.0) r = 1.0;
      h0 = r * h_old;
    }
    break;
  }
  h = h0;
}

// gsl_interp_cspline (natural) : c[i] = y''(x_i)/2
static void cspline_init(const std::vector<double>& xa, const std::vector<double>& ya, std::vector<double>& c) {
  const int n = (int)xa.size();
  const int max_index = n - 1, sys = max_index - 1;
  c.assign(n, 0.0);
  std::vector<double> g(sys), diag(sys), off(sys), al(sys), ga(sys), z(sys);
  for (int i = 0; i < sys; i++) {
    const double h_i = xa[i + 1] - xa[i], h_ip1 = xa[i + 2] - xa[i + 1];
    const double yd_i = ya[i + 1] - ya[i], yd_ip1 = ya[i + 2] - ya[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    off[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g[i] = 3.0 * (yd_ip1 * g_ip1 - yd_i * g_i);
  }
  al[0] = diag[0];
  ga[0] = off[0] / al[0];
  for (int i = 1; i < sys - 1; i++) {
    al[i] = diag[i] - off[i - 1] * ga[i - 1];
    ga[i] = off[i] / al[i];
  }
  if (sys > 1) al[sys - 1] = diag[sys - 1] - off[sys - 2] * ga[sys - 2];
  z[0] = g[0];
  for (int i = 1; i < sys; i++) z[i] = g[i] - ga[i - 1] * z[i - 1];
  for (int i = 0; i < sys; i++) z[i] = z[i] / al[i];
  c[sys] = z[sys - 1];
  for (int i = sys - 2; i >= 0; i--) c[i + 1] = z[i] - ga[i] * c[i + 2];
}

static inline double cspline_eval(const std::vector<double>& xa, const std::vector<double>& ya,
                                  const std::vector<double>& c, double x) {
  int lo = 0, hi = (int)xa.size() - 1;  // gsl_interp_bsearch(xa, x, 0, size-1)
  while (hi > lo + 1) {
    int i = (hi + lo) / 2;
    if (xa[i] > x)
      hi = i;
    else
      lo = i;
  }
  const double dx = xa[lo + 1] - xa[lo], dy = ya[lo + 1] - ya[lo];
  const double delx = x - xa[lo];
  const double b_i = (dy / dx) - dx * (c[lo + 1] + 2.0 * c[lo]) / 3.0;
  const double d_i = (c[lo + 1] - c[lo]) / (3.0 * dx);
  return ya[lo] + delx * (b_i + delx * (c[lo] + delx * d_i));
}

// camb/galileon.cc:503-688 (arrays_) with calcHubbleGalileon :332-438 inlined; tracker = 0.
int gal_arrays(Model& M, double omegar, double omegam, double H0in, double c2in, double c3in, double c4in,
               double cGin) {
  GalBackground& G = M.gal;
  G.ready = false;
  const int nb = GAL_NB;
  G.om = omegam;
  G.orad = omegar;
  G.h0 = H0in;
  G.c2 = c2in;
  G.cG = cGin;
  const int neig = M.Nu_mass_eigenstates;
  const double grhom = 3 * (G.h0 * G.h0);
  if (c3in == 0 && c4in == 0) {
    G.c3 = 0.5 * (-1. + G.om + G.orad + G.c2 / 6. - 3 * G.cG);
    for (int i = 1; i <= neig; i++) {
      double rhonu = 0;
      Nu_rho(M, M.nu_masses[i], rhonu);
      G.c3 += 0.5 * rhonu * M.grhormass[i] / grhom;
    }
    G.c4 = 0;
    G.c5 = 0;
  } else if (c4in == 0) {
    G.c3 = c3in;
    // the reference reads the *previous* global c4 on the right-hand side here (galileon.cc:546)
    G.c4 = (1. - G.om - G.orad - G.c2 / 6. + 2 * G.c3 - 7.5 * G.c4 + 3 * G.cG) / 7.5;
    for (int i = 1; i <= neig; i++) {
      double rhonu = 0;
      Nu_rho(M, M.nu_masses[i], rhonu);
      G.c4 += -rhonu * M.grhormass[i] / (7.5 * grhom);
    }
    G.c5 = 0;
  } else {
    G.c3 = c3in;
    G.c4 = c4in;
    G.c5 = (-1. + G.om + G.orad + G.c2 / 6. - 2 * G.c3 + 7.5 * G.c4 - 3 * G.cG) / 7.;
    for (int i = 1; i <= neig; i++) {
      double rhonu = 0, pnu = 0;
      Nu_background(M, M.nu_masses[i], rhonu, pnu);
      G.c5 += rhonu * M.grhormass[i] / (7. * grhom);
    }
  }
  const double amin = 1e-10, amax = 1.;
  G.q = std::pow(amin / amax, 1. / nb);
  std::vector<double> intvar(nb + 1), hubble(nb + 1), xgal(nb + 1);
  for (int i = 0; i <= nb; i++) intvar[i] = amax * std::pow(G.q, i);

  // calcHubbleGalileon
  int status = 0;
  {
    double y[3] = {1.0, 1.0, 0.0};
    double a = 1;
    double h = -1e-6;
    for (int i = 0; i <= nb; ++i) {
      const double acurrtarg = intvar[i];
      while (a > acurrtarg) rkf45_evolve_apply(M, a, acurrtarg, h, y);
      if (std::isnan(y[0]) || std::isnan(y[1]) || std::isnan(y[2])) {
        status = 8;
        break;
      }
      if (int stab = gal_stability(M, a, y)) {
        if (getenv("ORC_DEBUG")) fprintf(stderr, "gal stability fail code %d at i=%d a=%.6g h=%.6g x=%.6g\n", stab, i, a, y[0], a * y[1]);
        status = 7;
        break;
      }
      double h2 = y[0] * y[0];
      double x1 = a * y[1], x2 = x1 * x1, x3 = x2 * x1;
      double a3 = a * a * a;
      double OmegaP = (0.5 * G.c2 * x2 - 6.0 * G.c3 * h2 * x3 + 22.5 * G.c4 * h2 * h2 * x2 * x2 -
                       21.0 * G.c5 * h2 * h2 * h2 * x3 * x2 - 9.0 * G.cG * h2 * x2) /
                      3.0;
      double OmegaM = 1 - OmegaP - G.orad / (a3 * a * h2);
      for (int j = 1; j <= neig; j++) {
        double rhonu = 0, pnu = 0, am = intvar[i] * M.nu_masses[j];
        Nu_background(M, am, rhonu, pnu);
        OmegaM -= M.grhormass[j] * rhonu / (a3 * a * 3. * h2 * (G.h0 * G.h0));
      }
      double OmTest = G.om / (a3 * h2);
      if (std::fabs((OmegaM - OmTest) / OmTest) > 1e-4) {
        if (getenv("ORC_DEBUG")) fprintf(stderr, "gal closure fail at i=%d a=%.6g OmegaM=%.8g OmTest=%.8g\n", i, a, OmegaM, OmTest);
        status = 4;
        break;
      }
      hubble[i] = y[0];
      xgal[i] = y[1];
    }
  }
  if (status != 0) return status;
  G.n = nb + 2;
  G.a.assign(G.n, 0);
  G.h.assign(G.n, 0);
  G.x.assign(G.n, 0);
  for (int i = 0; i <= nb; i++) {
    G.a[i] = intvar[nb - i];
    G.h[i] = hubble[nb - i];
    G.x[i] = xgal[nb - i];
  }
  G.a[nb + 1] = 1. / G.q;
  G.h[nb + 1] = (G.h[nb] - G.h[nb - 1]) / (G.a[nb] - G.a[nb - 1]) * (G.a[nb + 1] - G.a[nb - 1]) + G.h[nb - 1];
  G.x[nb + 1] = (G.x[nb] - G.x[nb - 1]) / (G.a[nb] - G.a[nb - 1]) * (G.a[nb + 1] - G.a[nb - 1]) + G.x[nb - 1];
  cspline_init(G.a, G.h, G.ch);
  cspline_init(G.a, G.x, G.cx);
  G.ready = true;
  return status;
}

// camb/galileon.cc:691-717 / :719-744
double gal_GetX(const Model& M, double a) {
  if (a < 0.0 || a > 1.1) {
    fprintf(stderr, "Forbidden value of a : %.12f\n", a);
    std::abort();
  }
  return a * cspline_eval(M.gal.a, M.gal.x, M.gal.cx, a);
}
double gal_GetH(const Model& M, double a) {
  if (a < 0.0 || a > 1.1) {
    fprintf(stderr, "Forbidden value of a : %.12f\n", a);
    std::abort();
  }
  return cspline_eval(M.gal.a, M.gal.h, M.gal.ch, a);
}

// camb/galileon.cc:457-486 (analytic tracker; known-answer test only)
void gal_tracker(const Model& M, double a, double& h, double& x) {
  GAL_CONSTS(M);
  double op = c2 / 6.0 - 2.0 * c3 + 7.5 * c4 - 7.0 * c5;
  double a2 = a * a;
  double omega = G_.om / (a2 * a) + orad / (a2 * a2);
  for (int j = 1; j <= nu_n; j++) {
    double rhonu = 0;
    Nu_rho(M, a * M.nu_masses[j], rhonu);
    omega += M.grhormass[j] * rhonu / (a2 * a2 * 3. * (h0 * h0));
  }
  h = std::sqrt(0.5 * (omega - 3 * cG + std::sqrt(4 * op + 12 * cG + (3 * cG - omega) * (3 * cG - omega))));
  x = 1.0 / (a * h * h);
}

// ------------------------------------------------------------------ perturbation source terms
// Closed forms of camb/galileon.cc; arguments by value; hc = conformal Hubble a*hbar in units
// of H0, xc = a*dpi/da.
// camb/galileon.cc:746-787
void gal_GetdHdX(const Model& M, double a, double hc, double xc, double& dh, double& dx) {
  GAL_CONSTS(M);
  double a2 = a*a;
  double xgal2 = xc*xc;
  double xgal3 = xgal2*xc;
  double xgal4 = xgal2*xgal2;
  double xgal5 = xgal3*xgal2;
  double h = hc/a; // here H=dlna/dt
  double h2 = h*h;
  double h3 = h2*h;
  double h4 = h2*h2;
  double h5 = h3*h2;
  double h6 = h3*h3;
  double h7 = h4*h3;
  double h8 = h4*h4;
  double alpha = c2/6*h*xc-3*c3*h3*xgal2 + 15*c4*h5*xgal3 - 17.5*c5*h7*xgal4 - 3*cG*h3*xc;
  double gamma = c2/3*h2*xc-c3*h4*xgal2 + 2.5*c5*h8*xgal4 - 2*cG*h4*xc;
  double beta = c2/6*h2 -2*c3*h4*xc + 9*c4*h6*xgal2 - 10*c5*h8*xgal3 - cG*h4;
  double sigma = 2*h + 2*c3*h3*xgal3 - 15*c4*h5*xgal4 + 21*c5*h7*xgal5 + 6*cG*h3*xgal2;
  double lambda = 3*h2 + orad/(a2*a2) + c2/2*h2*xgal2 - 2*c3*h4*xgal3 + 7.5*c4*h6*xgal4 - 9*c5*h8*xgal5 - cG*h4*xgal2;
  if(nu_n>0){
    for(int i=0; i<nu_n; i++){
      double rhonu = 0;
      double pnu = 0;
      double am = a*M.nu_masses[i + 1];
      Nu_background(M, am, rhonu, pnu);
      lambda += M.grhormass[i + 1]*pnu/(a2*a2*h0*h0);
    }
  }
  double omega = 2*c3*h4*xgal2 - 12*c4*h6*xgal3 + 15*c5*h8*xgal4 + 4*cG*h4*xc;
  dh = (omega*gamma-lambda*beta)/(sigma*beta-alpha*omega); // dh/dlna
  dx = -xc+(alpha*lambda-sigma*gamma)/(sigma*beta-alpha*omega); // dx/dlna
}

// camb/galileon.cc:809-829
double gal_grhogal(const Model& M, double a, double hc, double xc) {
  GAL_CONSTS(M);
  double xgal2 = xc*xc;
  double xgal3 = xgal2*xc;
  double xgal4 = xgal2*xgal2;
  double xgal5 = xgal2*xgal3;
  double h = hc/a; // here H=dlna/dt
  double h2 = h*h;
  double h4 = h2*h2;
  double grhogal = 0;
  if(a >= 9.99999e-7) grhogal = h0*h0*a*a*(c2/2*h2*xgal2 - 6*c3*h4*xgal3 + 22.5*c4*h4*h2*xgal4 - 21*c5*h4*h4*xgal5 - 9*cG*h4*xgal2);
  return grhogal;
}

// camb/galileon.cc:831-850
double gal_gpresgal(const Model& M, double a, double hc, double xc, double dh, double dx) {
  GAL_CONSTS(M);
  double xgal2 = xc*xc;
  double xgal4 = xgal2*xgal2;
  double h = hc/a; // here H=dlna/dt
  double h2 = h*h;
  double h3 = h2*h;
  double h4 = h2*h2;
  double h6 = h4*h2;
  double gpresgal = 0;
  if(a >= 9.99999e-7) gpresgal = h0*h0*a*a*(c2/2*h2*xgal2 + 2*c3*h3*xgal2*(dh*xc+dx*h) - c4*(4.5*h6*xgal4 + 12*h6*xgal2*xc*dx + 15*h4*h*xgal4*dh) + 3*c5*h6*h*xgal4*(5*h*dx+7*dh*xc+2*h*xc) + cG*(6*h3*xgal2*dh + 4*h4*xc*dx + 3*h4*xgal2));
  return gpresgal;
}

// camb/galileon.cc:852-895
double gal_Chigal(const Model& M, double a, double hc, double xc, double dgrho, double eta, double dphi,
                  double dphip, double k) {
  GAL_CONSTS(M);
  const double k2 = k * k;
  double a2 = a*a;
  double a4 = a2*a2;
  double a6 = a4*a2;
  double xgal2 = xc*xc;
  double xgal3 = xgal2*xc;
  double xgal4 = xgal2*xgal2;
  double xgal5 = xgal3*xgal2;
  double h2 = hc*hc;
  double h3 = h2*hc;
  double h4 = h2*h2;
  double h6 = h4*h2;
  double alpha_Z = -2*c3/a2*xgal3*h2 
    + 15*c4/a4*xgal4*h4 
    - 21*c5/a6*xgal5*h6 
    - 6*cG/a2*xgal2*h2;
  double alpha_eta = 1.5*c4/a4*xgal4*h4
    - 3*c5/a6*xgal5*h6
    - cG/a2*xgal2*h2;
  double beta = 1./(1-0.5*alpha_Z);
  double ChitildeG = c2*h0*xc*hc*dphip 
    - c3/a2*(18*h0*xgal2*h3*dphip + 2*k2*xgal2*h2*dphi) 
    + c4/a4*(90*h0*xgal3*h4*hc*dphip + 12*k2*xgal3*h4*dphi) 
    - c5/a6*(105*h0*xgal4*h6*hc*dphip + 15*k2*xgal4*h6*dphi) 
    - cG/a2*(18*h0*xc*h3*dphip + 4*k2*xc*h2*dphi);
  double ChiG = 0;
  if (a >= 9.99999e-7) ChiG = beta*(ChitildeG + 0.5*alpha_Z*dgrho + (alpha_Z - 2*alpha_eta)*k*eta);
  return ChiG;
}

// camb/galileon.cc:897-936
double gal_qgal(const Model& M, double a, double hc, double xc, double dgq, double eta, double dphi,
                double dphip, double k) {
  GAL_CONSTS(M);
  double a2 = a*a;
  double a4 = a2*a2;
  double a6 = a4*a2;
  double xgal2 = xc*xc;
  double xgal3 = xgal2*xc;
  double xgal4 = xgal2*xgal2;
  double h2 = hc*hc;
  double h3 = h2*hc;
  double h4 = h2*h2;
  double h6 = h4*h2;
  double alpha = c4/a4*xgal4*h4 
    - 2*c5/a6*xgal4*xc*h6 
    - cG/(1.5*a2)*xgal2*h2;
  double beta = 1./(1-1.5*alpha);
  double qtildeG = c2*k*h0*xc*hc*dphi 
    - c3/a2*k*(6*h0*xgal2*h3*dphi - 2*xgal2*h2*dphip) 
    + c4/a4*k*(-12*xgal3*h4*dphip + 18*h0*xgal3*h4*hc*dphi) 
    - c5/a6*k*(-15*xgal4*h6*dphip + 15*h0*xgal4*h6*hc*dphi) 
    - cG/a2*k*(-4*xc*h2*dphip + 6*h0*xc*h3*dphi);
  double qG = 0;
  if(a >= 9.99999e-7) qG = beta*(qtildeG + 1.5*alpha*dgq);
  return qG;
}

// camb/galileon.cc:938-985
double gal_Pigal(const Model& M, double a, double hc, double xc, double dh, double dx, double dgrho,
                 double dgq, double dgpi, double eta, double dphi, double k) {
  GAL_CONSTS(M);
  const double k2 = k * k;
  double a2 = a*a;
  double a4 = a2*a2;
  double a6 = a4*a2;
  double xgal2 = xc*xc;
  double xgal3 = xgal2*xc;
  double xgal4 = xgal2*xgal2;
  double xgal5 = xgal3*xgal2;
  double h2 = hc*hc;
  double h3 = h2*hc;
  double h4 = h2*h2;
  double h5 = h3*h2;
  double h6 = h3*h3;
  double hprime = a*dh + hc;
  double xhprime = dx*hc + xc*hprime; // Derivative of the product (xgal*H)'
  double alpha_sigprime = c4/a4*xgal4*h4
    - 3*c5/a6*xgal4*h5*xhprime;
  double alpha_sig = c4/a4*(3*xgal4*h4 - 6*xgal3*h3*xhprime)
    - c5/a6*(-3*xgal5*h5*hprime + 12*xgal5*h6 - 15*xgal4*h5*xhprime)
    + 2*cG/a2*xc*hc*xhprime;
  double alpha_phi = -c4/a4*xgal4*h4
    - c5/a6*(6*xgal4*h5*xhprime - 6*xgal5*h6)
    + 2*cG/a2*xgal2*h2;
  double beta_pi = 1./(1+0.5*alpha_phi - alpha_sigprime);
  double PitildeG = k2*(c4/a4*(4*xgal3*h4*dphi - 6*xgal2*h3*xhprime*dphi)
				- c5/a6*(-12*xgal3*h5*xhprime*dphi + 12*xgal4*h6*dphi - 3*xgal4*h5*hprime*dphi)
				+ 2*cG/a2*hc*xhprime*dphi);
  double PiG = 0;
  if(a >= 9.99999e-7) PiG = beta_pi*(PitildeG + (alpha_sigprime - 0.5*alpha_phi)*dgpi + 0.5*(2*alpha_sigprime + alpha_sig - alpha_phi)*(dgrho + 3*h0*hc/k*dgq) + (alpha_sig + alpha_sigprime)*k*eta);
  return PiG;
}

// camb/galileon.cc:987-1082
double gal_dphisecond(const Model& M, double a, double hc, double xc, double dh, double dx, double dgrho,
                      double dgq, double eta, double dphi, double dphip, double k, double deltafprime) {
  GAL_CONSTS(M);
  const double k2 = k * k;
  double a2 = a*a;
  double a4 = a2*a2;
  double a6 = a4*a2;
  double xgal2 = xc*xc;
  double xgal3 = xgal2*xc;
  double xgal4 = xgal2*xgal2;
  double xgal5 = xgal3*xgal2;
  double h2 = hc*hc;
  double h3 = h2*hc;
  double h4 = h2*h2;
  double h5 = h3*h2;
  double h6 = h3*h3;
  double h7 = h4*h3;
  double h8 = h4*h4;
  double hprime = a*dh + hc;
  double xhprime = dx*hc + xc*hprime; // Derivative of the product (xgal*H)'
  double alpha_gammasecond = c2
    - 12*c3/a2*xc*h2
    + 54*c4/a4*xgal2*h4
    - 60*c5/a6*xgal3*h6
    - 6*cG/a2*h2;
  double alpha_gammaprime = 2*c2
    - c3/a2*(12*hc*xhprime + 12*xc*hc*hprime)
    + c4/a4*(-108*xgal2*h4 + 108*xc*h3*xhprime + 108*xgal2*h3*hprime)
    - c5/a6*(-240*xgal3*h6 + 180*xgal2*h5*xhprime + 180*xgal3*h5*hprime)
    - 12*cG/a2*hc*hprime;
  double alpha_gamma = c2
    - c3/a2*(4*xc*h2 + 4*hc*xhprime)
    + c4/a4*(-10*xgal2*h4 + 24*xc*h3*xhprime + 12*xgal2*h3*hprime)
    - c5/a6*(-36*xgal3*h6 + 36*xgal2*h5*xhprime + 24*xgal3*h5*hprime)
    - cG/a2*(2*h2 + 4*hc*hprime);
  double alpha_Z = c2*xc
    - c3/a2*(6*xgal2*h2 + 4*xc*hc*xhprime)
    + c4/a4*(-6*xgal3*h4 + 36*xgal2*h3*xhprime + 12*xgal3*h3*hprime)
    - c5/a6*(-45*xgal4*h6 + 60*xgal3*h5*xhprime + 30*xgal4*h5*hprime)
    - cG/a2*(6*xc*h2 + 4*hc*xhprime + 4*xc*hc*hprime);
  double alpha_Zprime = -2*c3/a2*xgal2*h2
    + 12*c4/a4*xgal3*h4
    - 15*c5/a6*xgal4*h6
    - 4*cG/a2*xc*h2;
  double alpha_eta = c4/a4*(-4*xgal3*h4 + 6*xgal2*h3*xhprime)
    - c5/a6*(-12*xgal4*h6 + 3*xgal4*h5*hprime + 12*xgal3*h5*xhprime)
    - 2*cG/a2*hc*xhprime;
  double dotdeltaf = deltafprime;
  double chiprimehat = c2*(-2*h02*xc*h2*dphip + h02*xc*hc*hprime*dphip + h02*h2*dx*dphip)
    - c3/a2*(-72*h02*xgal2*h4*dphip + 54*h02*xgal2*h3*hprime*dphip + 36*h02*xc*h4*dx*dphip
	     + k2*(2*xgal2*h2*dphip - 8*h0*xgal2*h3*dphi + 4*h0*xgal2*h2*hprime*dphi + 4*h0*xc*h3*dx*dphi))
    + c4/a4*(-540*h02*xgal3*h6*dphip + 450*h02*xgal3*h5*hprime*dphip + 270*h02*xgal2*h6*dx*dphip
	     + k2*(12*xgal3*h4*dphip - 72*h0*xgal3*h5*dphi + 48*h0*xgal3*h4*hprime*dphi + 36*h0*xgal2*h5*dx*dphi))
    - c5/a6*(-840*h02*xgal4*h8*dphip + 735*h02*xgal4*h7*hprime*dphip + 420*h02*xgal3*h8*dx*dphip
	     + k2*(15*xgal4*h6*dphip - 120*h0*xgal4*h7*dphi + 90*h0*xgal4*h6*hprime*dphi + 60*h0*xgal3*h7*dx*dphi))
    - cG/a2*(-72*h02*xc*h4*dphip + 54*h02*xc*h3*hprime*dphip + 18*h02*h4*dx*dphip
	     + k2*(4*xc*h2*dphip - 16*h0*xc*h3*dphi + 8*h0*xc*h2*hprime*dphi + 4*h0*h3*dx*dphi))
    ;
  double beta_gammasecond = c2*h0*xc*hc - 18*c3*h0/a2*xgal2*h3 + 90*c4*h0/a4*xgal3*h5 - 105*c5*h0/a6*xgal4*h7 - 18*cG*h0/a2*xc*h3;
  double beta_Z = -2*c3/a2*xgal3*h2
    + 15*c4/a4*xgal4*h4
    - 21*c5/a6*xgal5*h6
    - 6*cG/a2*xgal2*h2;
  double beta_eta = 1.5*c4/a4*xgal4*h4
    - 3*c5/a6*xgal5*h6
    - cG/a2*xgal2*h2;
  double beta_Z_prime = -c3*h0/a2*(-4*xgal3*h3 + 4*xgal3*h2*hprime + 6*xgal2*h3*dx)
    + c4*h0/a4*(-60*xgal4*h5 + 60*xgal4*h4*hprime + 60*xgal3*h5*dx)
    - c5*h0/a6*(-126*xgal5*h7 + 126*xgal5*h6*hprime + 105*xgal4*h7*dx)
    - cG*h0/a2*(-12*xgal2*h3 + 12*xgal2*h2*hprime + 12*xc*h3*dx);
  double beta_eta_prime = c4*h0/a4*(-6*xgal4*h5 + 6*xgal4*h4*hprime + 6*xgal3*h5*dx)
    - c5*h0/a6*(-18*xgal5*h7 + 18*xgal5*h6*hprime + 15*xgal4*h7*dx)
    - cG*h0/a2*(-2*xgal2*h3 + 2*xgal2*h2*hprime + 2*xc*h3*dx);
  double ksi = alpha_Zprime/(h0*hc*(2-beta_Z));
  double dphisecond = 0;
  if(a >= 9.99999e-7) dphisecond = -(alpha_gammaprime*h0*hc*dphip + alpha_gamma*k2*dphi + (0.5*alpha_Z + 2*ksi*(h0*hc - 0.5*(h0*hprime + beta_Z*h0*hc) + 0.25*(beta_Z_prime + beta_Z*h0*hprime)))*dgrho + ksi*(1-beta_eta)*k*dgq + ksi*dotdeltaf + ksi*chiprimehat + (alpha_Z - 2*alpha_eta + 2*ksi*(2*beta_eta*h0*hc - h0*hprime - beta_Z*h0*hc - beta_eta_prime + 0.5*(beta_Z_prime + beta_Z*h0*hprime)))*k*eta)/(alpha_gammasecond + ksi*beta_gammasecond);
  return dphisecond;
}

// camb/galileon.cc:1084-1174
double gal_pigalprime(const Model& M, double a, double hc, double xc, double dh, double dx, double dgrho,
                      double dgq, double dgpi, double pidot, double eta, double dphi, double dphip, double k,
                      double grho, double gpres) {
  GAL_CONSTS(M);
  const double k2 = k * k;
  // pow(hprime,2), pow(dx,2), pow(den,2) of the reference written as products
  double hoft = hc/a;
  double a2 = a*a;
  double a4 = a2*a2;
  double a6 = a4*a2;
  double hoft2 = hoft*hoft;
  double hoft3 = hoft2*hoft;
  double hoft4 = hoft3*hoft;
  double hoft5 = hoft4*hoft;
  double hoft6 = hoft5*hoft;
  double hoft7 = hoft6*hoft;
  double hoft8 = hoft7*hoft;
  double xgal2 = xc*xc;
  double xgal3 = xgal2*xc;
  double xgal4 = xgal2*xgal2;
  double xgal5 = xgal3*xgal2;
  double h2 = hc*hc;
  double h3 = h2*hc;
  double h4 = h2*h2;
  double h5 = h3*h2;
  double h6 = h3*h3;
  double h7 = h4*h3;
  double alpha = c2/6*hoft*xc-3*c3*hoft3*xgal2 + 15*c4*hoft5*xgal3 - 17.5*c5*hoft7*xgal4 - 3*cG*hoft3*xc;
  double gamma = c2/3*hoft2*xc-c3*hoft4*xgal2 + 2.5*c5*hoft8*xgal4 - 2*cG*hoft4*xc;
  double beta = c2/6*hoft2 -2*c3*hoft4*xc + 9*c4*hoft6*xgal2 - 10*c5*hoft8*xgal3 - cG*hoft4;
  double delta = 2*hoft + 2*c3*hoft3*xgal3 - 15*c4*hoft5*xgal4 + 21*c5*hoft7*xgal5 + 6*cG*hoft3*xgal2;
  double lambda = 3*hoft2 + orad/(a2*a2) + c2/2*hoft2*xgal2 - 2*c3*hoft4*xgal3 + 7.5*c4*hoft6*xgal4 - 9*c5*hoft8*xgal5 - cG*hoft4*xgal2;
  double omega = 2*c3*hoft4*xgal2 - 12*c4*hoft6*xgal3 + 15*c5*hoft8*xgal4 + 4*cG*hoft4*xc;
  double hprime = a*dh + hc;
  const double hprime2 = hprime * hprime, dx2 = dx * dx;
  double xhprime = dx*hc + xc*hprime; // Derivative of the product (xgal*H)'
  double alpha_prime = (c2/6*xc - 9*c3*hoft2*xgal2 + 75*c4*hoft4*xgal3 - 122.5*c5*hoft6*xgal4 - 9*cG*hoft2*xc)*(hprime-hc)/a + (c2/6*hoft - 6*c3*hoft3*xc + 45*c4*hoft5*xgal2 - 70*c5*hoft7*xgal3 - 3*cG*hoft3)*dx;
  double gamma_prime = (2.*c2/3*hoft*xc - 4*c3*hoft3*xgal2 + 20*c5*hoft7*xgal4 - 8*cG*hoft3*xc)*(hprime-hc)/a + (c2/3*hoft2 - 2*c3*hoft4*xc + 10*c5*hoft8*xgal3 - 2*cG*hoft4)*dx;
  double beta_prime = (c2/3*hoft - 8*c3*hoft3*xc + 54*c4*hoft5*xgal2 - 80*c5*hoft7*xgal3 - 4*cG*hoft3)*(hprime-hc)/a + (-2*c3*hoft4 + 18*c4*hoft6*xc - 30*c5*hoft8*xgal2)*dx;
  double delta_prime = (2 + 6*c3*hoft2*xgal3 - 75*c4*hoft4*xgal4 + 147*c5*hoft6*xgal5 + 18*cG*hoft2*xgal2)*(hprime-hc)/a + (6*c3*hoft3*xgal2 - 60*c4*hoft5*xgal3 + 105*c5*hoft7*xgal4 + 12*cG*hoft3*xc)*dx;
  double lambda_prime = -4*orad/a4 + (6*hoft + c2*hoft*xgal2 - 8*c3*hoft3*xgal3 + 45*c4*hoft5*xgal4 - 72*c5*hoft7*xgal5 - 4*cG*hoft3*xgal2)*(hprime-hc)/a + (c2*hoft2*xc - 6*c3*hoft4*xgal2 + 30*c4*hoft6*xgal3 - 45*c5*hoft8*xgal4 - 2*cG*hoft4*xc)*dx;
  if(nu_n>0){
    for(int i=0; i<nu_n; i++){
      double pnu = 0;
      double rhonu = 0;
      double am = a*M.nu_masses[i + 1];
      Nu_background(M, am, rhonu, pnu);
      lambda += M.grhormass[i + 1]*pnu/(a2*a2*h0*h0);
      double dpnu = Nu_dp(M, am, hc, pnu);
      lambda_prime += (dpnu-4*pnu*hc)*M.grhormass[i + 1]/(a2*a2*h0*h0);
    }
  }
  double omega_prime = (8*c3*hoft3*xgal2 - 72*c4*hoft5*xgal3 + 120*c5*hoft7*xgal4 + 16*cG*hoft3*xc)*(hprime-hc)/a + (4*c3*hoft4*xc - 36*c4*hoft6*xgal2 + 60*c5*hoft8*xgal3 + 4*cG*hoft4)*dx;
  const double den = delta*beta-alpha*omega, den2 = den * den;
  double xprimedot = h0*hc*(-dx + ((alpha_prime*lambda + alpha*lambda_prime - delta_prime*gamma - delta*gamma_prime)*(delta*beta-alpha*omega) - (alpha*lambda-delta*gamma)*(delta_prime*beta + delta*beta_prime - alpha_prime*omega - alpha*omega_prime))/den2);
  double hprimedot = a*h0*hc*((omega*gamma-lambda*beta)/(delta*beta-alpha*omega) + ((omega_prime*gamma + omega*gamma_prime - lambda_prime*beta - lambda*beta_prime)*(delta*beta-alpha*omega) - (omega*gamma-lambda*beta)*(delta_prime*beta + delta*beta_prime - alpha_prime*omega - alpha*omega_prime))/den2) + h0*hc*hprime;
  double piprimetilde = k2*c4/a4*(4*xgal3*h4*dphip - 6*xgal3*h3*hprime*dphip - 6*xgal2*h4*dx*dphip - 24*h0*xgal3*h5*dphi + 52*h0*xgal3*h4*hprime*dphi - 18*h0*xgal3*h3*hprime2*dphi + 48*h0*xgal2*h5*dx*dphi - 42*h0*xgal2*h4*hprime*dx*dphi - 12*h0*xc*h5*dx2*dphi - 6*xgal3*h3*hprimedot*dphi - 6*xgal2*h4*xprimedot*dphi)
    - k2*c5/a6*(12*xgal4*h6*dphip - 15*xgal4*h5*hprime*dphip - 12*xgal3*h6*dx*dphip - 96*h0*xgal4*h7*dphi + 192*h0*xgal4*h6*hprime*dphi - 75*h0*xgal4*h5*hprime2*dphi + 144*h0*xgal3*h7*dx*dphi - 132*h0*xgal3*h6*hprime*dx*dphi - 36*h0*xgal2*h7*dx2*dphi - 15*xgal4*h5*hprimedot*dphi - 12*xgal3*h6*xprimedot*dphi)
    - k2*cG/a2*(-2*xc*hc*hprime*dphip - 2*h2*dx*dphip + 8*h0*xc*h2*hprime*dphi - 2*h0*xc*hc*hprime2*dphi + 8*h0*h3*dx*dphi - 6*h0*h2*hprime*dx*dphi - 2*xc*hc*hprimedot*dphi - 2*h2*xprimedot*dphi);
  double alpha_sig = c4/a4*(3*xgal4*h4 - 6*xgal3*h3*xhprime)
    - c5/a6*(-3*xgal5*h5*hprime + 12*xgal5*h6 - 15*xgal4*h5*xhprime)
    + 2*cG/a2*xc*hc*xhprime;
  double alpha_sigprime = c4/a4*xgal4*h4
    - 3*c5/a6*xgal4*h5*xhprime;
  double alpha_phi = -c4/a4*xgal4*h4
    - c5/a6*(6*xgal4*h5*xhprime - 6*xgal5*h6)
    + 2*cG/a2*xgal2*h2;
  double alpha_sig_prime = c4/a4*(-12*h0*xgal4*h5 + 36*h0*xgal4*h4*hprime - 18*h0*xgal4*h3*hprime2 + 36*h0*xgal3*h5*dx - 48*h0*xgal3*h4*hprime*dx - 18*h0*xgal2*h5*dx2 - 6*xgal4*h3*hprimedot - 6*xgal3*h4*xprimedot)
    - c5/a6*(-72*h0*xgal5*h7 + 180*h0*xgal5*h6*hprime - 90*h0*xgal5*h5*hprime2 + 150*h0*xgal4*h7*dx - 180*h0*xgal4*h6*hprime*dx - 60*h0*xgal3*h7*dx2 - 18*xgal5*h5*hprimedot - 15*xgal4*h6*xprimedot)
    - cG/a2*(4*h0*xgal2*h2*hprime - 2*h0*xgal2*hc*hprime2 + 4*h0*xc*h3*dx - 8*h0*xc*h2*hprime*dx - 2*h0*h3*dx2 - 2*xgal2*hc*hprimedot - 2*xc*h2*xprimedot);
  double alpha_sigprime_prime = c4/a4*(-4*h0*xgal4*h5 + 4*h0*xgal4*h4*hprime + 4*h0*xgal3*h5*dx)
    - c5/a6*(-18*h0*xgal5*h6*hprime + 15*h0*xgal5*h5*hprime2 - 18*h0*xgal4*h7*dx + 33*h0*xgal4*h6*hprime*dx + 12*h0*xgal3*h7*dx2 + 3*xgal5*h5*hprimedot + 3*xgal4*h6*xprimedot);
  double alpha_phi_prime = c4/a4*(4*h0*xgal4*h5 - 4*h0*xgal4*h4*hprime - 4*h0*xgal3*h5*dx)
    - c5/a6*(36*h0*xgal5*h7 - 72*h0*xgal5*h6*hprime + 30*h0*xgal5*h5*hprime2 - 66*h0*xgal4*h7*dx + 66*h0*xgal4*h6*hprime*dx + 24*h0*xgal3*h7*dx2 + 6*xgal5*h5*hprimedot + 6*xgal4*h6*xprimedot)
    - cG/a2*(4*h0*xgal2*h3 - 4*h0*xgal2*h2*hprime - 4*h0*xc*h3*dx);
  double piGdot = 0;
  if(a >= 9.99999e-7) piGdot = (piprimetilde + (alpha_sigprime - 0.5*alpha_phi)*pidot + (-3.5*h0*hc*alpha_sigprime + 0.5*h0*hprime*alpha_sigprime + 0.25*(alpha_phi - alpha_sigprime)*(grho+gpres)/(h0*hc) + 0.5*h0*hprime*alpha_sig + 0.5*alpha_sig_prime + alpha_sigprime_prime - 2*h0*hc*alpha_sig + 1.5*h0*hc*alpha_phi - 0.5*alpha_phi_prime)*dgrho + (alpha_sig_prime + alpha_sigprime_prime + alpha_sigprime*h0*hprime + alpha_sig*h0*hprime - 3*h0*hc*alpha_sigprime - 3*h0*hc*alpha_sig + 0.5*(alpha_phi - alpha_sigprime)*(grho+gpres)/(h0*hc))*k*eta + 0.5/k*(3*h02*hc*hprime*alpha_sigprime + 3*h02*hc*hprime*alpha_sig + 1.5*(alpha_phi - alpha_sigprime)*(grho+gpres) + 3*h0*hc*alpha_sig_prime - 12*h02*h2*alpha_sig - 21*h02*h2*alpha_sigprime - 3*h0*hc*alpha_phi_prime + 9*h02*h2*alpha_phi + 6*h0*hc*alpha_sigprime_prime + k2*(alpha_phi - alpha_sigprime))*dgq + (-2*h0*hc*alpha_sigprime + h0*hc*alpha_phi - 0.5*alpha_phi_prime + alpha_sigprime_prime - h0*hc*alpha_sig)*dgpi)/(1. - alpha_sigprime + 0.5*alpha_phi);
  return piGdot;
}

}  // namespace orc
