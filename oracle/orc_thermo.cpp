// TEST INFRASTRUCTURE (oracle) -- see orc.hpp header.
// Thermal history: RECFAST 1.5.2 (camb/recfast.f90) and the ThermoData module
// (camb/modules.f90:2698-3202): inithermo, thermo, SetTimeSteps, DoThermoSpline.
#include "orc.hpp"

namespace orc {

#define SP(x) ((double)(x##f))

namespace {
// camb/recfast.f90:200-236 (RECDATA)
const double Lambda = 8.2245809, Lambda_He = 51.3;
const double L_H_ion = 1.096787737e7, L_H_alpha = 8.225916453e6, L_He1_ion = 1.98310772e7, L_He2_ion = 4.389088863e7,
             L_He_2s = 1.66277434e7, L_He_2p = 1.71134891e7;
const double A2P_s = 1.798287e9, A2P_t = 177.58, L_He_2Pt = 1.690871466e7, L_He_2St = 1.5985597526e7,
             L_He2St_ion = 3.8454693845e6, sigma_He_2Ps = 1.436289e-22, sigma_He_2Pt = 1.484872e-22;
const double bigH = 100.0e3 / Mpc;
const double not4 = mass_ratio_He_H;
const double zinitial = 1e4, zfinal = 0.0;
const int Nz = Recfast::Nz;
const double delta_z = (zinitial - zfinal) / Nz;
const double AGauss1 = -0.14, AGauss2 = 0.079, zGauss1 = 7.28, zGauss2 = 6.73, wGauss1 = 0.18, wGauss2 = 0.33;
const double Pi = pi_aml, G = G_newton, C = c_light;

struct IonCtx {
  Model* M;
};

// camb/recfast.f90:778-1012
void ION(void* vctx, int /*Ndim*/, double z, const double* y, double* f) {
  Model& M = *((IonCtx*)vctx)->M;
  Recfast& R = M.rec;
  const double fHe = R.fHe, Nnow = R.Nnow, Tnow = R.Tnow;
  double a_PPB = 4.309, b_PPB = -0.6166, c_PPB = 0.6703, d_PPB = 0.5300;
  double a_VF = std::pow(10.0, -16.744), b_VF = 0.711;
  double T_0 = std::pow(10.0, 0.477121), T_1 = std::pow(10.0, 5.114);
  double a_trip = std::pow(10.0, -16.306), b_trip = 0.761;
  double x_H = y[0], x_He = y[1];
  double x = x_H + fHe * x_He;
  double Tmat = y[2];
  double n = Nnow * ((1.0 + z) * (1.0 + z) * (1.0 + z));
  double n_He = fHe * Nnow * ((1.0 + z) * (1.0 + z) * (1.0 + z));
  double Trad = Tnow * (1.0 + z);
  double Hz = 1 / dtauda(M, 1 / (1.0 + z)) * ((1.0 + z) * (1.0 + z)) / MPC_in_sec;
  double Rdown = 1.e-19 * a_PPB * std::pow(Tmat / 1.e4, b_PPB) / (1.0 + c_PPB * std::pow(Tmat / 1.e4, d_PPB));
  double Rup = Rdown * std::pow(R.CR * Tmat, 1.5) * std::exp(-R.CDB / Tmat);
  double sq_0 = std::sqrt(Tmat / T_0), sq_1 = std::sqrt(Tmat / T_1);
  double Rdown_He = a_VF / (sq_0 * std::pow(1.0 + sq_0, 1.0 - b_VF));
  Rdown_He = Rdown_He / std::pow(1.0 + sq_1, 1.0 + b_VF);
  double Rup_He = Rdown_He * std::pow(R.CR * Tmat, 1.5) * std::exp(-R.CDB_He / Tmat);
  Rup_He = 4.0 * Rup_He;
  double He_Boltz = (R.Bfact / Tmat) > 680.0 ? std::exp(680.0) : std::exp(R.Bfact / Tmat);
  double K;
  if (!M.P.RECFAST_Hswitch)
    K = R.CK / Hz;
  else {
    double g1 = (std::log(1.0 + z) - zGauss1) / wGauss1, g2 = (std::log(1.0 + z) - zGauss2) / wGauss2;
    K = R.CK / Hz * (1.0 + AGauss1 * std::exp(-(g1 * g1)) + AGauss2 * std::exp(-(g2 * g2)));
  }
  double Rdown_trip = a_trip / (sq_0 * std::pow(1.0 + sq_0, 1.0 - b_trip));
  Rdown_trip = Rdown_trip / std::pow(1.0 + sq_1, 1.0 + b_trip);
  double Rup_trip = Rdown_trip * std::exp(-h_P * C * L_He2St_ion / (k_B * Tmat));
  Rup_trip = Rup_trip * std::pow(R.CR * Tmat, 1.5) * (4.0 / 3.0);
  int Heflag = ((x_He < 5.e-9) || (x_He > 0.98)) ? 0 : M.P.RECFAST_Heswitch;
  double K_He, CfHe_t = 0;
  if (Heflag == 0) {
    K_He = R.CK_He / Hz;
  } else {
    double tauHe_s = A2P_s * R.CK_He * 3.0 * n_He * (1.0 - x_He) / Hz;
    double pHe_s = (1.0 - std::exp(-tauHe_s)) / tauHe_s;
    K_He = 1.0 / (A2P_s * pHe_s * 3.0 * n_He * (1.0 - x_He));
    if (((Heflag == 2) || (Heflag >= 5)) && x_H < 0.9999999) {
      double Doppler = 2.0 * k_B * Tmat / (m_H * not4 * C * C);
      Doppler = C * L_He_2p * std::sqrt(Doppler);
      double gamma_2Ps = 3.0 * A2P_s * fHe * (1.0 - x_He) * C * C /
                         (std::sqrt(Pi) * sigma_He_2Ps * 8.0 * Pi * Doppler * (1.0 - x_H)) /
                         ((C * L_He_2p) * (C * L_He_2p));
      double pb = 0.36, qb = M.P.RECFAST_fudge_He;
      double AHcon = A2P_s / (1.0 + pb * std::pow(gamma_2Ps, qb));
      K_He = 1.0 / ((A2P_s * pHe_s + AHcon) * 3.0 * n_He * (1.0 - x_He));
    }
    if (Heflag >= 3) {
      double tauHe_t = A2P_t * n_He * (1.0 - x_He) * 3.0;
      tauHe_t = tauHe_t / (8.0 * Pi * Hz * (L_He_2Pt * L_He_2Pt * L_He_2Pt));
      double pHe_t = (1.0 - std::exp(-tauHe_t)) / tauHe_t;
      double CL_PSt = h_P * C * (L_He_2Pt - L_He_2St) / k_B;
      if ((Heflag == 3) || (Heflag == 5) || (x_H > 0.99999)) {
        CfHe_t = A2P_t * pHe_t * std::exp(-CL_PSt / Tmat);
        CfHe_t = CfHe_t / (Rup_trip + CfHe_t);
      } else {
        double Doppler = 2.0 * k_B * Tmat / (m_H * not4 * C * C);
        Doppler = C * L_He_2Pt * std::sqrt(Doppler);
        double gamma_2Pt = 3.0 * A2P_t * fHe * (1.0 - x_He) * C * C /
                           (std::sqrt(Pi) * sigma_He_2Pt * 8.0 * Pi * Doppler * (1.0 - x_H)) /
                           ((C * L_He_2Pt) * (C * L_He_2Pt));
        double pb = 0.66, qb = 0.9;
        double AHcon = A2P_t / (1.0 + pb * std::pow(gamma_2Pt, qb)) / 3.0;
        CfHe_t = (A2P_t * pHe_t + AHcon) * std::exp(-CL_PSt / Tmat);
        CfHe_t = CfHe_t / (Rup_trip + CfHe_t);
      }
    }
  }
  double Trad4 = Trad * Trad * Trad * Trad;
  double timeTh = (1.0 / (R.CT * Trad4)) * (1.0 + x + fHe) / x;
  double timeH = 2. / (3. * R.HO * std::pow(1.0 + z, 1.5));
  if (x_H > SP(0.99)) {
    f[0] = 0.0;
  } else if (x_H > 0.985) {
    f[0] = (x * x_H * n * Rdown - Rup * (1.0 - x_H) * std::exp(-R.CL / Tmat)) / (Hz * (1.0 + z));
    R.recombination_saha_z = z;
  } else {
    f[0] = ((x * x_H * n * Rdown - Rup * (1.0 - x_H) * std::exp(-R.CL / Tmat)) *
            (1.0 + K * Lambda * n * (1.0 - x_H))) /
           (Hz * (1.0 + z) * (1.0 / R.fu + K * Lambda * n * (1.0 - x_H) / R.fu + K * Rup * n * (1.0 - x_H)));
  }
  if (x_He < SP(1.e-15)) {
    f[1] = 0.0;
  } else {
    f[1] = ((x * x_He * n * Rdown_He - Rup_He * (1 - x_He) * std::exp(-R.CL_He / Tmat)) *
            (1 + K_He * Lambda_He * n_He * (1.0 - x_He) * He_Boltz)) /
           (Hz * (1 + z) * (1 + K_He * (Lambda_He + Rup_He) * n_He * (1.0 - x_He) * He_Boltz));
    if (Heflag >= 3) {
      f[1] = f[1] + (x * x_He * n * Rdown_trip -
                     (1.0 - x_He) * 3.0 * Rup_trip * std::exp(-h_P * C * L_He_2St / (k_B * Tmat))) *
                        CfHe_t / (Hz * (1.0 + z));
    }
  }
  if (timeTh < R.H_frac * timeH) {
    double dHdz = (R.HO * R.HO / 2.0 / Hz) * (4.0 * ((1.0 + z) * (1.0 + z) * (1.0 + z)) / (1.0 + R.z_eq) * R.OmegaT +
                                              3.0 * R.OmegaT * ((1.0 + z) * (1.0 + z)) + 2.0 * R.OmegaK * (1.0 + z));
    double epsilon = Hz * (1.0 + x + fHe) / (R.CT * (Trad * Trad * Trad) * x);
    f[2] = Tnow + epsilon * ((1.0 + fHe) / (1.0 + fHe + x)) * ((f[0] + fHe * f[1]) / x) - epsilon * dHdz / Hz +
           3.0 * epsilon / (1.0 + z);
  } else {
    f[2] = R.CT * Trad4 * x / (1.0 + x + fHe) * (Tmat - Trad) / (Hz * (1.0 + z)) + 2.0 * Tmat / (1.0 + z);
  }
}

// camb/recfast.f90:725-776
void get_init(const Recfast& R, double z, double& x_H0, double& x_He0, double& x0) {
  const double fHe = R.fHe;
  if (z > 8000.0) {
    x_H0 = 1.0;
    x_He0 = 1.0;
    x0 = 1.0 + 2.0 * fHe;
  } else if (z > 3500.0) {
    x_H0 = 1.0;
    x_He0 = 1.0;
    double rhs = std::exp(1.5 * std::log(R.CR * R.Tnow / (1.0 + z)) - R.CB1_He2 / (R.Tnow * (1.0 + z))) / R.Nnow;
    rhs = rhs * 1.0;
    x0 = 0.5 * (std::sqrt((rhs - 1.0 - fHe) * (rhs - 1.0 - fHe) + 4.0 * (1.0 + 2.0 * fHe) * rhs) - (rhs - 1.0 - fHe));
  } else if (z > 2000.0) {
    x_H0 = 1.0;
    double rhs = std::exp(1.5 * std::log(R.CR * R.Tnow / (1.0 + z)) - R.CB1_He1 / (R.Tnow * (1.0 + z))) / R.Nnow;
    rhs = rhs * 4.0;
    x_He0 = 0.5 * (std::sqrt((rhs - 1.0) * (rhs - 1.0) + 4.0 * (1.0 + fHe) * rhs) - (rhs - 1.0));
    x0 = x_He0;
    x_He0 = (x0 - 1.0) / fHe;
  } else {
    double rhs = std::exp(1.5 * std::log(R.CR * R.Tnow / (1.0 + z)) - R.CB1 / (R.Tnow * (1.0 + z))) / R.Nnow;
    x_H0 = 0.5 * (std::sqrt(rhs * rhs + 4.0 * rhs) - rhs);
    x_He0 = 0.0;
    x0 = x_H0;
  }
}
}  // namespace

// camb/recfast.f90:460-722
void Recombination_init(Model& M) {
  const Params& P = M.P;
  Recfast& R = M.rec;
  R.zrec.assign(Nz + 1, 0);
  R.xrec.assign(Nz + 1, 0);
  R.dxrec.assign(Nz + 1, 0);
  R.recombination_saha_z = 0.0;
  R.Tnow = P.TCMB;
  double z = zinitial;
  R.OmegaT = P.omegac + P.omegab;
  R.OmegaK = 1.0 - R.OmegaT - P.omegav;
  double H = P.H0 / 100.0;
  R.HO = H * bigH;
  const double Yp = P.YHe;
  R.mu_H = 1.0 / (1.0 - Yp);
  R.mu_T = not4 / (not4 - (not4 - 1.0) * Yp);
  R.fHe = Yp / (not4 * (1.0 - Yp));
  R.Nnow = 3.0 * R.HO * R.HO * P.omegab / (8.0 * Pi * G * R.mu_H * m_H);
  double fnu = (21.0 / 8.0) * std::pow(4.0 / 11.0, 4.0 / 3.0);
  double Tnow4 = R.Tnow * R.Tnow * R.Tnow * R.Tnow;
  R.z_eq = (3.0 * ((R.HO * C) * (R.HO * C)) / (8.0 * Pi * G * a_rad_f() * (1.0 + fnu) * Tnow4)) * (P.omegab + P.omegac);
  R.z_eq = R.z_eq - 1.0;
  R.Lalpha = 1.0 / L_H_alpha;
  R.Lalpha_He = 1.0 / L_He_2p;
  R.DeltaB = h_P * C * (L_H_ion - L_H_alpha);
  R.CDB = R.DeltaB / k_B;
  R.DeltaB_He = h_P * C * (L_He1_ion - L_He_2s);
  R.CDB_He = R.DeltaB_He / k_B;
  R.CB1 = h_P * C * L_H_ion / k_B;
  R.CB1_He1 = h_P * C * L_He1_ion / k_B;
  R.CB1_He2 = h_P * C * L_He2_ion / k_B;
  R.CR = 2.0 * Pi * (m_e / h_P) * (k_B / h_P);
  R.CK = (R.Lalpha * R.Lalpha * R.Lalpha) / (8.0 * Pi);
  R.CK_He = (R.Lalpha_He * R.Lalpha_He * R.Lalpha_He) / (8.0 * Pi);
  R.CL = C * h_P / (k_B * R.Lalpha);
  R.CL_He = C * h_P / (k_B / L_He_2s);
  R.CT = Compton_CT_f() / MPC_in_sec;
  R.Bfact = h_P * C * (L_He_2p - L_He_2s) / k_B;
  R.H_frac = 1e-3;
  R.fu = P.RECFAST_fudge;
  double y[4];
  y[2] = R.Tnow * (1.0 + z);
  double x_H0, x_He0, x0;
  get_init(R, z, x_H0, x_He0, x0);
  y[0] = x_H0;
  y[1] = x_He0;
  y[3] = y[2];
  int ind = 1;
  const int nw = 3;
  double cw[25];
  for (int i = 0; i < 25; i++) cw[i] = 0;
  double w[3 * 9];
  for (int i = 0; i < 27; i++) w[i] = 0;
  IonCtx ctx{&M};
  const double tol = 1e-5;
  for (int i = 1; i <= Nz; i++) {
    double zstart = zinitial - (double)(i - 1) * delta_z;
    double zend = zinitial - (double)i * delta_z;
    z = zend;
    if (zend > 8000.0) {
      x_H0 = 1.0;
      x_He0 = 1.0;
      x0 = 1.0 + 2.0 * R.fHe;
      y[0] = x_H0;
      y[1] = x_He0;
      y[2] = R.Tnow * (1.0 + z);
    } else if (z > 5000.0) {
      x_H0 = 1.0;
      x_He0 = 1.0;
      double rhs = std::exp(1.5 * std::log(R.CR * R.Tnow / (1.0 + z)) - R.CB1_He2 / (R.Tnow * (1.0 + z))) / R.Nnow;
      rhs = rhs * 1.0;
      x0 = 0.5 * (std::sqrt((rhs - 1.0 - R.fHe) * (rhs - 1.0 - R.fHe) + 4.0 * (1.0 + 2.0 * R.fHe) * rhs) -
                  (rhs - 1.0 - R.fHe));
      y[0] = x_H0;
      y[1] = x_He0;
      y[2] = R.Tnow * (1.0 + z);
    } else if (z > 3500.0) {
      x_H0 = 1.0;
      x_He0 = 1.0;
      x0 = x_H0 + R.fHe * x_He0;
      y[0] = x_H0;
      y[1] = x_He0;
      y[2] = R.Tnow * (1.0 + z);
    } else if (y[1] > SP(0.99)) {
      x_H0 = 1.0;
      double rhs = std::exp(1.5 * std::log(R.CR * R.Tnow / (1.0 + z)) - R.CB1_He1 / (R.Tnow * (1.0 + z))) / R.Nnow;
      rhs = rhs * 4.0;
      x_He0 = 0.5 * (std::sqrt((rhs - 1.0) * (rhs - 1.0) + 4.0 * (1.0 + R.fHe) * rhs) - (rhs - 1.0));
      x0 = x_He0;
      x_He0 = (x0 - 1.0) / R.fHe;
      y[0] = x_H0;
      y[1] = x_He0;
      y[2] = R.Tnow * (1.0 + z);
    } else if (y[0] > 0.99) {
      double rhs = std::exp(1.5 * std::log(R.CR * R.Tnow / (1.0 + z)) - R.CB1 / (R.Tnow * (1.0 + z))) / R.Nnow;
      x_H0 = 0.5 * (std::sqrt(rhs * rhs + 4.0 * rhs) - rhs);
      dverk(&ctx, 3, ION, zstart, y, zend, tol, ind, cw, nw, w);
      y[0] = x_H0;
      x0 = y[0] + R.fHe * y[1];
    } else {
      dverk(&ctx, nw, ION, zstart, y, zend, tol, ind, cw, nw, w);
      x0 = y[0] + R.fHe * y[1];
    }
    R.zrec[i] = zend;
    R.xrec[i] = x0;
  }
  spline(R.zrec.data(), R.xrec.data(), Nz, 1.0e40, 1.0e40, R.dxrec.data());
}

// camb/recfast.f90:434-458
double Recombination_xe(const Model& M, double a) {
  const Recfast& R = M.rec;
  double z = 1 / a - 1;
  if (z >= R.zrec[1]) return R.xrec[1];
  if (z <= R.zrec[Nz]) return R.xrec[Nz];
  double zst = (zinitial - z) / delta_z;
  int ihi = (int)zst;
  int ilo = ihi + 1;
  double az = zst - (int)zst;
  double bz = 1 - az;
  return az * R.xrec[ilo] + bz * R.xrec[ihi] +
         ((az * az * az - az) * R.dxrec[ilo] + (bz * bz * bz - bz) * R.dxrec[ihi]) / 6.0;
}

// ------------------------------------------------------------------ ThermoData
static inline double herm(const std::vector<double>& f, const std::vector<double>& df, int i, double d) {
  return f[i] + d * (df[i] + d * (3 * (f[i + 1] - f[i]) - 2 * df[i] - df[i + 1] +
                                  d * (df[i] + df[i + 1] + 2 * (f[i] - f[i + 1]))));
}

// camb/modules.f90:2728-2769
void thermo(const Model& M, double tau, double& cs2b, double& opacity, double* dopacity) {
  const Thermo& T = M.th;
  double d = std::log(tau / T.tauminn) / T.dlntau + 1.0;
  int i = (int)d;
  d = d - i;
  if (i < 1) {
    fprintf(stderr, "thermo out of bounds\n");
    std::abort();
  } else if (i >= nthermo) {
    cs2b = T.cs2[nthermo] + (d + i - nthermo) * T.dcs2[nthermo];
    opacity = T.dotmu[nthermo] + (d - nthermo) * T.ddotmu[nthermo];
    if (dopacity) {
      fprintf(stderr, "thermo: shouldn't happen\n");
      std::abort();
    }
  } else {
    cs2b = herm(T.cs2, T.dcs2, i, d);
    opacity = herm(T.dotmu, T.ddotmu, i, d);
    if (dopacity) *dopacity = herm(T.ddotmu, T.dddotmu, i, d) / (tau * T.dlntau);
  }
}

// camb/modules.f90:2771-2783
double Thermo_OpacityToTime(const Model& M, double opacity) {
  int j = 1;
  while (M.th.dotmu[j] > opacity) j++;
  return std::exp((j - 1) * M.th.dlntau) * M.th.tauminn;
}

// camb/modules.f90:3106-3146
static void SetTimeSteps(Model& M) {
  Thermo& T = M.th;
  Ranges_Init(M.TimeSteps);
  Ranges_Add_delta(M.TimeSteps, M.taurst, M.taurend, M.dtaurec);
  double dtau0 = T.Maxtau / 500.0 / M.P.AccuracyBoost;
  Ranges_Add_delta(M.TimeSteps, M.taurend, M.tau0, dtau0);
  if (M.reionization) {
    int nri0 = (int)(50 * M.P.AccuracyBoost);
    Ranges_Add(M.TimeSteps, M.reion_tau_start, M.reion_tau_complete, nri0);
  }
  Ranges_GetArray(M.TimeSteps);
  int nstep = M.TimeSteps.npoints;
  T.vis.assign(nstep + 1, 0);
  T.dvis = T.ddvis = T.expmmu = T.dopac = T.opac = T.vis;
}

// camb/modules.f90:3159-3202
static void DoThermoSpline(Model& M, int j2, double tau) {
  Thermo& T = M.th;
  double d = std::log(tau / T.tauminn) / T.dlntau + 1.0;
  int i = (int)d;
  d = d - i;
  double ddopac;
  if (i < nthermo) {
    T.opac[j2] = herm(T.dotmu, T.ddotmu, i, d);
    T.dopac[j2] = herm(T.ddotmu, T.dddotmu, i, d) / (tau * T.dlntau);
    ddopac = (herm(T.dddotmu, T.ddddotmu, i, d) - (T.dlntau * T.dlntau) * tau * T.dopac[j2]) /
             ((tau * T.dlntau) * (tau * T.dlntau));
    T.expmmu[j2] = herm(T.emmu, T.demmu, i, d);
  } else {
    T.opac[j2] = T.dotmu[nthermo];
    T.dopac[j2] = T.ddotmu[nthermo];
    ddopac = T.dddotmu[nthermo];
    T.expmmu[j2] = T.emmu[nthermo];
  }
  const double op = T.opac[j2];
  T.vis[j2] = op * T.expmmu[j2];
  T.dvis[j2] = T.expmmu[j2] * (op * op + T.dopac[j2]);
  T.ddvis[j2] = T.expmmu[j2] * (op * op * op + 3 * op * T.dopac[j2] + ddopac);
}

// camb/modules.f90:2787-3103 (derived-parameter block :3045-3101 omitted: it does not feed C_l)
void inithermo(Model& M, double taumin, double taumax) {
  const Params& P = M.P;
  Thermo& T = M.th;
  Recombination_init(M);
  if (M.error_flag) return;
  const int nt = nthermo;
  for (std::vector<double>* v : {&T.tb, &T.cs2, &T.xe, &T.dcs2, &T.dotmu, &T.ddotmu, &T.sdotmu, &T.emmu, &T.demmu,
                                 &T.dddotmu, &T.ddddotmu, &T.scaleFactor})
    v->assign(nt + 1, 0.0);
  T.Maxtau = taumax;
  T.tight_tau = 0;
  T.actual_opt_depth = 0;
  int ncount = 0;
  double thomc0 = Compton_CT_f() * (P.TCMB * P.TCMB * P.TCMB * P.TCMB);
  T.tauminn = 0.05 * taumin;
  T.dlntau = std::log(M.tau0 / T.tauminn) / (nt - 1);
  double last_dotmu = 0;
  T.matter_verydom_tau = 0;
  double a_verydom = P.AccuracyBoost * 5 * (M.grhog + M.grhornomass) / (M.grhoc + M.grhob);
  double tau01 = T.tauminn;
  double adot0 = M.adotrad;
  double a0 = M.adotrad * T.tauminn;
  double a02 = a0 * a0;
  T.tb[1] = P.TCMB / a0;
  double xe0 = 1.0, x1 = 0.0, x2 = 1.0;
  T.xe[1] = xe0 + 0.25 * P.YHe / (1.0 - P.YHe) * (x1 + 2 * x2);
  double barssc = barssc0 * (1.0 - 0.75 * P.YHe + (1.0 - P.YHe) * T.xe[1]);
  T.cs2[1] = 4.0 / 3.0 * barssc * T.tb[1];
  T.dotmu[1] = T.xe[1] * M.akthom / a02;
  T.sdotmu[1] = 0;
  const bool DerivedParameters = true;
  for (int i = 2; i <= nt; i++) {
    double tau = T.tauminn * std::exp((i - 1) * T.dlntau);
    double dtau = tau - tau01;
    double a = a0 + adot0 * dtau;
    T.scaleFactor[i] = a;
    double a2 = a * a;
    double adot = 1 / dtauda(M, a);
    if (T.matter_verydom_tau == 0 && a > a_verydom) T.matter_verydom_tau = tau;
    a = a0 + 2.0 * dtau / (1.0 / adot0 + 1.0 / adot);
    double tg0 = P.TCMB / a0;
    double ahalf = 0.5 * (a0 + a);
    double adothalf = 0.5 * (adot0 + adot);
    double fe = (1.0 - P.YHe) * T.xe[i - 1] / (1.0 - 0.75 * P.YHe + (1.0 - P.YHe) * T.xe[i - 1]);
    double thomc = thomc0 * fe / adothalf / (ahalf * ahalf * ahalf);
    double etc = std::exp(-thomc * (a - a0));
    double a2t = a0 * a0 * (T.tb[i - 1] - tg0) * etc - P.TCMB / thomc * (1.0 - etc);
    T.tb[i] = P.TCMB / a + a2t / (a * a);
    if (M.reionization && tau > M.reion_tau_start) {
      if (ncount == 0) ncount = i - 1;
      T.xe[i] = Reionization_xe(M, a, tau, T.xe[ncount], true);
      if (P.AccurateReionization && DerivedParameters) {
        T.dotmu[i] = (Recombination_xe(M, a) - T.xe[i]) * M.akthom / a2;
        if (last_dotmu != 0)
          T.actual_opt_depth = T.actual_opt_depth - 2.0 * dtau / (1.0 / T.dotmu[i] + 1.0 / last_dotmu);
        last_dotmu = T.dotmu[i];
      }
    } else {
      T.xe[i] = Recombination_xe(M, a);
    }
    double dtbdla = -2.0 * T.tb[i] - thomc * adothalf / adot * (a * T.tb[i] - P.TCMB);
    barssc = barssc0 * (1.0 - 0.75 * P.YHe + (1.0 - P.YHe) * T.xe[i]);
    T.cs2[i] = barssc * T.tb[i] * (1 - dtbdla / T.tb[i] / 3.0);
    T.dotmu[i] = T.xe[i] * M.akthom / a2;
    if (T.tight_tau == 0 && 1 / (tau * T.dotmu[i]) > SP(0.005)) T.tight_tau = tau;
    if (tau < SP(0.001))
      T.sdotmu[i] = 0;
    else
      T.sdotmu[i] = T.sdotmu[i - 1] + 2.0 * dtau / (1.0 / T.dotmu[i] + 1.0 / T.dotmu[i - 1]);
    a0 = a;
    tau01 = tau;
    adot0 = adot;
  }
  for (int j1 = 1; j1 <= nt; j1++) {
    if (T.sdotmu[j1] - T.sdotmu[nt] < -69)
      T.emmu[j1] = 1.e-30;
    else {
      T.emmu[j1] = std::exp(T.sdotmu[j1] - T.sdotmu[nt]);
      if (!P.AccurateReionization && T.actual_opt_depth == 0 && T.xe[j1] < SP(1e-3))
        T.actual_opt_depth = -T.sdotmu[j1] + T.sdotmu[nt];
    }
  }
  int iv = 0;
  double vfi = 0.0, cf1;
  int ns;
  if (ncount == 0) {
    cf1 = 1.0;
    ns = nt;
  } else {
    cf1 = std::exp(T.sdotmu[nt] - T.sdotmu[ncount]);
    ns = ncount;
  }
  double maxvis = 0;
  for (int j1 = 1; j1 <= ns; j1++) {
    double vis = T.emmu[j1] * T.dotmu[j1];
    double tau = T.tauminn * std::exp((j1 - 1) * T.dlntau);
    vfi = vfi + vis * cf1 * T.dlntau * tau;
    if ((iv == 0) && (vfi > 1.0e-7 / P.AccuracyBoost)) {
      M.taurst = 9.0 / 10.0 * tau;
      iv = 1;
    } else if (iv == 1) {
      if (vis > maxvis) {
        maxvis = vis;
        M.tau_maxvis = tau;
      }
      if (vfi > SP(0.995)) {
        M.taurend = tau;
        iv = 2;
        break;
      }
    }
  }
  if (iv != 2) {
    GlobalError(M, "inithermo: failed to find end of recombination", error_reionization);
    return;
  }
  M.dtaurec = std::fmin(M.dtaurec, M.taurst / 40) / P.AccuracyBoost;
  if (M.reionization) M.taurend = std::fmin(M.taurend, M.reion_tau_start);
  std::vector<double> sd(nt + 1);
  splini(sd.data(), nt);
  splder(T.cs2.data(), T.dcs2.data(), nt, sd.data());
  splder(T.dotmu.data(), T.ddotmu.data(), nt, sd.data());
  splder(T.ddotmu.data(), T.dddotmu.data(), nt, sd.data());
  splder(T.dddotmu.data(), T.ddddotmu.data(), nt, sd.data());
  splder(T.emmu.data(), T.demmu.data(), nt, sd.data());
  SetTimeSteps(M);
  for (int j2 = 1; j2 <= M.TimeSteps.npoints; j2++) DoThermoSpline(M, j2, M.TimeSteps.points[j2]);
}

}  // namespace orc
