// TEST INFRASTRUCTURE (oracle) -- see orc.hpp header.
// Lensed C_l by the curved-sky correlation-function method: lens_Cls / CorrFuncFullSky / CorrFuncFullSkyImpl of
// camb/lensing.f90:78-528 (lensing_method = 1, no tensors, AccurateBB = F, ALens_Fiducial = 0).
#include "orc.hpp"

namespace orc {

#define SP(x) ((double)(x##f))
static inline long nint(double x) { return std::lround(x); }

// camb/lensing.f90:107-528.  Needs M.Cl_scalar with C_Phi (DoLensing run) and the high-L template (4 columns).
int lens_Cls(Model& M) {
  const Params& P = M.P;
  const double pi = pi_aml;
  if (M.C_last < 4) {
    GlobalError(M, "lens_Cls needs the lensing potential (DoLensing)", error_unsupported_params);
    return M.error_flag;
  }
  const int Max_l = P.Max_l;
  // CorrFuncFullSky :95-104
  int lmax_extrap = Max_l - 100 + 450;
  if (P.HighAccuracyDefault) lmax_extrap = lmax_extrap + 300;
  lmax_extrap = std::min(lmax_extrap_highl, lmax_extrap);
  const int lmax = std::max(lmax_extrap, Max_l);
  if (lmax > Max_l && M.highL_template.empty()) {
    GlobalError(M, "lens_Cls needs the high-L template", error_unsupported_params);
    return M.error_flag;
  }
  const lSamples& lS = M.lSamp;
  int max_lensed_ix = lS.l0 - 1;
  while (lS.l[max_lensed_ix] > Max_l - 100) max_lensed_ix = max_lensed_ix - 1;
  const int lmax_lensed = lS.l[max_lensed_ix];
  M.lmax_lensed = lmax_lensed;
  M.Cl_lensed.assign((size_t)4 * (lmax_lensed - lmin + 1), 0.0);
  int npoints = (int)(Max_l * 2 * P.AccuracyBoost);
  double dtheta = pi / npoints;
  if (Max_l > 3500) dtheta = dtheta / SP(1.3);
  const int apodize_point_width = (int)nint(SP(0.003) / dtheta);
  npoints = (int)(pi / dtheta);
  const double range_fac = std::fmax(1.0, 32 / P.AccuracyBoost);
  npoints = (int)(npoints / range_fac);
  const int interp_fac = std::max(1, std::min((int)nint(10 / P.AccuracyBoost), (int)(range_fac * 2) - 1));
  std::vector<double> lfacs(lmax + 1), lfacs2(lmax + 1), lrootfacs(lmax + 1), theta_cut(lmax + 1), ls(lmax + 1);
  int jmax = 0;
  for (int l = lmin; l <= lmax; l++) {
    if (l <= 15 || ((l - 15) % interp_fac) == interp_fac / 2) {
      jmax = jmax + 1;
      ls[jmax] = l;
    }
    lfacs[l] = (double)(l * (l + 1));
    lfacs2[l] = (double)((l + 2) * (l - 1));
    lrootfacs[l] = std::sqrt(lfacs[l] * lfacs2[l]);
  }
  for (int l = 2; l <= lmax; l++) theta_cut[l] = 0.244949 / std::sqrt(3.0 * lfacs[l] - 8.0);
  std::vector<double> roots_(lmax + 6);
  double* roots = roots_.data() + 1;  // roots(-1:lmax+4)
  roots[-1] = 0;
  for (int l = 0; l <= lmax + 4; l++) roots[l] = std::sqrt((double)l);
  const int nL = Max_l - lmin + 1, nT = lmax_extrap_highl - lmin + 1;
  auto Cls = [&](int l, int type) { return M.Cl_scalar[(size_t)(type - 1) * nL + (l - lmin)]; };
  auto Tmpl = [&](int l, int col) { return M.highL_template[(size_t)col * nT + (l - lmin)]; };  // 0 TT,1 EE,2 TE,3 PP
  std::vector<double> Cphil3(lmax + 1), CTT(lmax + 1), CTE(lmax + 1), CEE(lmax + 1);
  for (int l = lmin; l <= Max_l; l++) {
    Cphil3[l] = Cls(l, 4) * (2 * l + 1) * (l + 1) / ((double)l * (double)l * (double)l) / (4 * pi);
    const double fac = (2 * l + 1) / (4 * pi) * 2 * pi / (l * (l + 1));
    CTT[l] = Cls(l, 1) * fac;
    CEE[l] = Cls(l, 2) * fac;
    CTE[l] = Cls(l, 3) * fac;
  }
  if (Cphil3[10] > SP(1e-7)) {
    GlobalError(M, "You need to normalize realistically to use lensing", error_norm_lensing);
    return M.error_flag;
  }
  if (lmax > Max_l) {
    int l = Max_l;
    double sc = (2 * l + 1) / (4 * pi) * 2 * pi / (l * (l + 1));
    const double fac2 = CTT[Max_l] / (sc * Tmpl(Max_l, 0));
    const double fac = Cphil3[Max_l] / (sc * Tmpl(Max_l, 3));
    for (l = Max_l + 1; l <= lmax; l++) {
      sc = (2 * l + 1) / (4 * pi) * 2 * pi / (l * (l + 1));
      Cphil3[l] = Tmpl(l, 3) * fac * sc;
      CTT[l] = Tmpl(l, 0) * fac2 * sc;
      CEE[l] = Tmpl(l, 1) * fac2 * sc;
      CTE[l] = Tmpl(l, 2) * fac2 * sc;
    }
    if (Cphil3[Max_l + 1] > SP(1e-7)) {  // :243-249
      GlobalError(M, "You need to normalize the high-L template so it is dimensionless", error_norm_lensing);
      return M.error_flag;
    }
  }
  if (M.P.ALens_Fiducial > 0) {  // :252-257
    for (int l = 2; l <= lmax; l++) {
      const double sc = (2 * l + 1) / (4 * pi) * 2 * pi / (l * (l + 1));
      Cphil3[l] = sc * Tmpl(l, 3) * M.P.ALens_Fiducial;
    }
  }
  std::vector<double> lens_contrib((size_t)4 * (lmax_lensed + 1), 0.0);  // [type 1..4][l]
  auto LC = [&](int t, int l) -> double& { return lens_contrib[(size_t)(t - 1) * (lmax_lensed + 1) + l]; };
  std::vector<double> Pl(lmax + 1), dP(lmax + 1), d_11(lmax + 1), d_m11(lmax + 1), d_22(lmax + 1), d_2m2(lmax + 1),
      d_20(lmax + 1);
  std::vector<double> cc((size_t)4 * (jmax + 1));
  for (int i = 1; i <= npoints - 1; i++) {
    const double theta = i * dtheta;
    const double x = std::cos(theta), sinth = std::sin(theta), halfsinth = sinth / 2;
    double pmm = 1, pmmp1 = x, Cg2 = 0, sigmasq = 0;
    for (int l = 2; l <= lmax; l++) {
      Pl[l] = ((2 * l - 1) * x * pmmp1 - (l - 1) * pmm) / l;
      dP[l] = l * (pmmp1 - x * Pl[l]) / (sinth * sinth);
      pmm = pmmp1;
      pmmp1 = Pl[l];
      const double llp1 = lfacs[l];
      const double fac1 = (1 - x), fac2 = (1 + x), fac = fac1 / fac2;
      d_11[l] = fac1 * dP[l] / llp1 + Pl[l];
      d_m11[l] = fac2 * dP[l] / llp1 - Pl[l];
      sigmasq = sigmasq + (1 - d_11[l]) * Cphil3[l];
      Cg2 = Cg2 + d_m11[l] * Cphil3[l];
      d_22[l] = (((4 * x - 8) / fac2 + llp1) * Pl[l] + 4 * fac * (fac2 + (x - 2) / llp1) * dP[l]) / lfacs2[l];
      if (theta > theta_cut[l]) {
        d_2m2[l] = ((llp1 - (4 * x + 8) / fac1) * Pl[l] + 4 / fac * (-fac1 + (x + 2) / llp1) * dP[l]) / lfacs2[l];
      } else {
        const double th2 = theta * theta;
        d_2m2[l] = lfacs[l] * lfacs2[l] * (th2 * th2) * (1.0 / 384.0 - (3.0 * lfacs[l] - 8.0) / 23040.0 * th2);
      }
      d_20[l] = (2 * x * dP[l] - llp1 * Pl[l]) / lrootfacs[l];
    }
    for (int j = 1; j <= jmax; j++) {
      const int l = (int)ls[j];
      double fac1 = (1 - x), fac2 = (1 + x);
      const double llp1 = lfacs[l];
      const double rootllp1 = roots[l] * roots[l + 1];
      const double rootfac1 = roots[l + 2] * roots[l - 1];
      const double rootfac2 = roots[l + 3] * roots[l - 2];
      const double dm11 = d_m11[l], d11 = d_11[l];
      const double d2m2 = d_2m2[l], d22 = d_22[l], d20 = d_20[l];
      const double d1m2 = sinth / rootfac1 * (dP[l] - 2 / fac1 * dm11);
      const double d12 = sinth / rootfac1 * (dP[l] - 2 / fac2 * d11);
      double d1m3, d2m3, d3m3, d13, d23, d33;
      if (l < 3) {
        d1m3 = d2m3 = d3m3 = d13 = d23 = d33 = 0;
      } else {
        const double sinfac = 4 / sinth;
        d1m3 = (-(x + 0.5) * d1m2 * sinfac - lfacs2[l] * dm11 / rootfac1) / rootfac2;
        d2m3 = (-fac2 * d2m2 * sinfac - rootfac1 * d1m2) / rootfac2;
        d3m3 = (-(x + 1.5) * d2m3 * sinfac - rootfac1 * d1m3) / rootfac2;
        d13 = ((x - 0.5) * d12 * sinfac - lfacs2[l] * d11 / rootfac1) / rootfac2;
        d23 = (-fac1 * d22 * sinfac + rootfac1 * d12) / rootfac2;
        d33 = (-(x - 1.5) * d23 * sinfac - rootfac1 * d13) / rootfac2;
      }
      (void)d33;
      double d04, d2m4, d4m4, rootfac3;
      if (l < 4) {
        d04 = d2m4 = d4m4 = rootfac3 = 0;
      } else {
        rootfac3 = roots[l - 3] * roots[l + 4];
        d04 = ((-llp1 + (18 * x * x + 6) / (sinth * sinth)) * d20 - 6 * x * lfacs2[l] * dP[l] / lrootfacs[l]) /
              (rootfac2 * rootfac3);
        d2m4 = (-(6 * x + 4) * d2m3 / sinth - rootfac2 * d2m2) / rootfac3;
        d4m4 = (-7 / 5.0 * (llp1 - 6) * d2m2 + 12 / 5.0 * (-llp1 + (9 * x + 26) / fac1) * d3m3) / (llp1 - 12);
      }
      const double X000 = std::exp(-llp1 * sigmasq / 4);
      const double X022 = X000 * (1 + sigmasq);
      const double X220 = lrootfacs[l] / 4 * X000;
      const double X121 = -0.5 * rootfac1 * X000;
      const double X132 = -0.5 * rootfac2 * X000;
      const double X242 = 0.25 * rootfac2 * rootfac3 * X022;
      const double dX000 = -llp1 / 4 * X000;
      const double dX022 = (1 - llp1 / 4) * X022;
      fac1 = dX000 * dX000;
      const double fac3 = X220 * X220;
      const double Cg2sq = Cg2 * Cg2;
      double fac = ((X000 * X000 - 1) + Cg2sq * fac1) * Pl[l] + Cg2sq * fac3 * d2m2 + 8 / llp1 * fac1 * Cg2 * dm11;
      cc[(size_t)0 * (jmax + 1) + j] = CTT[l] * fac;
      fac2 = (Cg2 * dX022) * (Cg2 * dX022) + (X022 * X022 - 1);
      fac = 2 * Cg2 * X121 * X132 * d13 + fac2 * d22 + Cg2sq * X242 * X220 * d04;
      cc[(size_t)1 * (jmax + 1) + j] = CEE[l] * fac;
      fac = (fac3 * Pl[l] + X242 * X242 * d4m4) * Cg2sq / 2 + Cg2 * (X121 * X121 * dm11 + X132 * X132 * d3m3) + fac2 * d2m2;
      cc[(size_t)2 * (jmax + 1) + j] = CEE[l] * fac;
      fac = (X000 * X022 - 1) * d20 + 2 * dX000 * Cg2 * (X121 * d11 + X132 * d1m3) / rootllp1 +
            Cg2sq * (X220 / 2 * d2m4 * X242 + (fac3 / 2 + dX022 * dX000) * d20);
      cc[(size_t)3 * (jmax + 1) + j] = CTE[l] * fac;
    }
    double corr[4];
    for (int t = 0; t < 4; t++) {
      double s1 = 0, s2 = 0;
      for (int j = 1; j <= std::min(14, jmax); j++) s1 += cc[(size_t)t * (jmax + 1) + j];
      for (int j = 15; j <= jmax; j++) s2 += cc[(size_t)t * (jmax + 1) + j];
      corr[t] = s1 + interp_fac * s2;
    }
    if (i > npoints - apodize_point_width * 3) {
      const int d = i - npoints + apodize_point_width * 3;  // default-real arithmetic in the reference (:468)
      const double w = (double)std::exp(-(float)(d * d) / (float)(2 * apodize_point_width * apodize_point_width));
      for (int t = 0; t < 4; t++) corr[t] = corr[t] * w;
    }
    for (int l = lmin; l <= lmax_lensed; l++) {
      LC(1, l) = LC(1, l) + corr[0] * Pl[l] * sinth;
      const double T2 = corr[1] * d_22[l];
      const double T4 = corr[2] * d_2m2[l];
      LC(2, l) = LC(2, l) + (T2 + T4) * halfsinth;
      LC(3, l) = LC(3, l) + (T2 - T4) * halfsinth;
      LC(4, l) = LC(4, l) + corr[3] * d_20[l] * sinth;
    }
  }
  const int nLl = lmax_lensed - lmin + 1;
  for (int l = lmin; l <= lmax_lensed; l++) {
    const double fac = l * (l + 1) / (2 * pi) * dtheta * 2 * pi;
    M.Cl_lensed[(size_t)0 * nLl + (l - lmin)] = LC(1, l) * fac + Cls(l, 1);
    M.Cl_lensed[(size_t)1 * nLl + (l - lmin)] = LC(2, l) * fac + Cls(l, 2);
    M.Cl_lensed[(size_t)2 * nLl + (l - lmin)] = LC(3, l) * fac;
    M.Cl_lensed[(size_t)3 * nLl + (l - lmin)] = LC(4, l) * fac + Cls(l, 3);
  }
  return 0;
}

}  // namespace orc
