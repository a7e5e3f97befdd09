// Galileon perturbation source terms, restructured for the device.
//
// Same closed forms as camb/galileon.cc:746-1081 (GetdHdX, grhogal, gpresgal, Chigal, qgal, Pigal, dphisecond;
// arXiv:1208.0600 eqs. 40-48 in CAMB's synchronous-gauge variables), but organised around what changes between
// calls: everything that depends only on the background point (a, H = a*hbar, x) -- inverse powers of a, the
// monomials x^n H^m, the alpha/beta coefficient polynomials -- is evaluated ONCE per right-hand-side call into a
// GalFast record; the terms themselves are then short linear forms in the perturbation variables.  The reference
// divides by a^2, a^4, a^6 inside every monomial (~50 FP64 divisions per RHS call) and re-derives the same
// powers in each of its seven functions; here there are 6 divisions per call.  Results agree with the
// reference-order evaluation (oracle/orc_background.cpp) to rounding (tests/test_galileon_fast.py).
#pragma once

#ifndef GC_HD
#ifdef __CUDACC__
#define GC_HD __host__ __device__ __forceinline__
#else
#define GC_HD inline
#endif
#endif

namespace gc {

struct GalConst {  // per-model constants (camb/galileon.cc:80-92)
  double c2, c3, c4, c5, cG, h0, orad;
};

// the six polynomials of the background system at one point (galileon.cc:763-783), kept for pigalprime
struct GalDyn {
  double g, alpha, gamma, beta, delta, lambda, omega;
};

// linear forms in the perturbation variables (coefficients depend on the background point and k only)
struct GalPiLin {  // Pigal = d*dphi + pi*dgpi + g*dgrho + q*dgq + e*eta
  double d, pi, g, q, e;
};
struct GalPhi2Lin {  // dphisecond = p*dphip + d*dphi + g*dgrho + q*dgq + f*dotdeltaf + e*eta
  double p, d, g, q, f, e;
};
struct GalPiDotLin {  // pigalprime = p*dphip + d*dphi + pid*pidot + g*dgrho + q*dgq + pi*dgpi + e*eta
  double p, d, pid, g, q, pi, e;
};

struct GalFast {
  bool on;  // a >= 9.99999e-7: the reference hard-zeroes every perturbation term below (galileon.cc:822 ...)
  double a, H, X, k, k2;
  double dh, dx, hprime, xhprime;  // d hbar/dlna, d x/dlna, (a hbar)', (x a hbar)' in the reference's notation
  double grhogal, gpresgal;
  // shared monomial combinations
  double C3, C4, C5, CG;  // c_n / a^(2,4,6,2)
  double x2, x3, x4, x5, h2, h3, h4, h5, h6, h7, h8;
  double aZ, aEta;  // Chigal's alpha_Z, alpha_eta (= dphisecond's beta_Z, beta_eta)
  double A;         // h0 * (c2 x H - 18 C3 x^2 H^3 + 90 C4 x^3 H^5 - 105 C5 x^4 H^7 - 18 CG x H^3)
  double Bk;        // -2 C3 x^2 H^2 + 12 C4 x^3 H^4 - 15 C5 x^4 H^6 - 4 CG x H^2
  // Chigal = chi_p*dphip + chi_d*dphi + chi_g*dgrho + chi_e*eta ; qgal = q_d*dphi + q_p*dphip + q_g*dgq
  double chi_p, chi_d, chi_g, chi_e, q_d, q_p, q_g;
};

// GetdHdX (galileon.cc:746-787) in g = hbar = H/a: d hbar/dlna and d x/dlna from the 2x2 background system.
// lam_nu = sum_i grhormass_i * pnu_i (massive-neutrino pressure term of lambda).
GC_HD void gal_fast_dhdx(const GalConst& G, double a, double g, double X, double lam_nu, double& dh, double& dx,
                         GalDyn* dyn = nullptr) {
  const double c2 = G.c2, c3 = G.c3, c4 = G.c4, c5 = G.c5, cG = G.cG, h0 = G.h0;
  const double a2 = a * a, ia4 = 1.0 / (a2 * a2);
  const double x2 = X * X, x3 = x2 * X, x4 = x2 * x2, x5 = x3 * x2;
  const double g2 = g * g, g3 = g2 * g, g4 = g2 * g2, g5 = g3 * g2, g6 = g3 * g3, g7 = g4 * g3, g8 = g4 * g4;
  const double alpha = c2 / 6 * g * X - 3 * c3 * g3 * x2 + 15 * c4 * g5 * x3 - 17.5 * c5 * g7 * x4 - 3 * cG * g3 * X;
  const double gamma = c2 / 3 * g2 * X - c3 * g4 * x2 + 2.5 * c5 * g8 * x4 - 2 * cG * g4 * X;
  const double beta = c2 / 6 * g2 - 2 * c3 * g4 * X + 9 * c4 * g6 * x2 - 10 * c5 * g8 * x3 - cG * g4;
  const double sigma = 2 * g + 2 * c3 * g3 * x3 - 15 * c4 * g5 * x4 + 21 * c5 * g7 * x5 + 6 * cG * g3 * x2;
  const double lambda = 3 * g2 + G.orad * ia4 + c2 / 2 * g2 * x2 - 2 * c3 * g4 * x3 + 7.5 * c4 * g6 * x4 - 9 * c5 * g8 * x5 -
                        cG * g4 * x2 + lam_nu * ia4 / (h0 * h0);
  const double omega = 2 * c3 * g4 * x2 - 12 * c4 * g6 * x3 + 15 * c5 * g8 * x4 + 4 * cG * g4 * X;
  const double iden = 1.0 / (sigma * beta - alpha * omega);
  dh = (omega * gamma - lambda * beta) * iden;
  dx = -X + (alpha * lambda - sigma * gamma) * iden;
  if (dyn) *dyn = GalDyn{g, alpha, gamma, beta, sigma, lambda, omega};
}

// everything else that depends on the background point only, for given (dh, dx): grhogal (:809-829), gpresgal
// (:831-850), the shared monomials and the Chigal / qgal coefficient sets.
GC_HD void gal_fast_fill(const GalConst& G, double a, double g, double X, double dh, double dx, double k, GalFast& F) {
  const double c2 = G.c2, c3 = G.c3, c4 = G.c4, c5 = G.c5, cG = G.cG, h0 = G.h0;
  F.a = a;
  F.X = X;
  F.k = k;
  F.k2 = k * k;
  const double H = a * g;
  F.H = H;
  F.dh = dh;
  F.dx = dx;
  const double a2 = a * a;
  const double ia2 = 1.0 / a2, ia4 = ia2 * ia2, ia6 = ia4 * ia2;
  const double x2 = X * X, x3 = x2 * X, x4 = x2 * x2, x5 = x3 * x2;
  F.on = a >= 9.99999e-7;
  if (F.on) {
    const double g2 = g * g, g3 = g2 * g, g4 = g2 * g2, g5 = g3 * g2, g6 = g3 * g3, g7 = g4 * g3, g8 = g4 * g4;
    const double pre = h0 * h0 * a2;
    F.grhogal = pre * (c2 / 2 * g2 * x2 - 6 * c3 * g4 * x3 + 22.5 * c4 * g6 * x4 - 21 * c5 * g8 * x5 - 9 * cG * g4 * x2);
    F.gpresgal = pre * (c2 / 2 * g2 * x2 + 2 * c3 * g3 * x2 * (dh * X + dx * g) -
                        c4 * (4.5 * g6 * x4 + 12 * g6 * x3 * dx + 15 * g5 * x4 * dh) +
                        3 * c5 * g7 * x4 * (5 * g * dx + 7 * dh * X + 2 * g * X) +
                        cG * (6 * g3 * x2 * dh + 4 * g4 * X * dx + 3 * g4 * x2));
  } else {
    F.grhogal = 0;
    F.gpresgal = 0;
  }
  F.hprime = a * dh + H;
  F.xhprime = dx * H + X * F.hprime;
  if (!F.on) {
    F.chi_p = F.chi_d = F.chi_g = F.chi_e = F.q_d = F.q_p = F.q_g = 0;
    return;
  }
  const double h2 = H * H, h3 = h2 * H, h4 = h2 * h2, h5 = h3 * h2, h6 = h3 * h3, h7 = h4 * h3, h8 = h4 * h4;
  const double C3 = c3 * ia2, C4 = c4 * ia4, C5 = c5 * ia6, CG = cG * ia2;
  F.C3 = C3; F.C4 = C4; F.C5 = C5; F.CG = CG;
  F.x2 = x2; F.x3 = x3; F.x4 = x4; F.x5 = x5;
  F.h2 = h2; F.h3 = h3; F.h4 = h4; F.h5 = h5; F.h6 = h6; F.h7 = h7; F.h8 = h8;
  // Chigal, galileon.cc:852-895
  const double aZ = -2 * C3 * x3 * h2 + 15 * C4 * x4 * h4 - 21 * C5 * x5 * h6 - 6 * CG * x2 * h2;
  const double aEta = 1.5 * C4 * x4 * h4 - 3 * C5 * x5 * h6 - CG * x2 * h2;
  const double A = h0 * (c2 * X * H - 18 * C3 * x2 * h3 + 90 * C4 * x3 * h5 - 105 * C5 * x4 * h7 - 18 * CG * X * h3);
  const double Bk = -2 * C3 * x2 * h2 + 12 * C4 * x3 * h4 - 15 * C5 * x4 * h6 - 4 * CG * X * h2;
  F.aZ = aZ; F.aEta = aEta; F.A = A; F.Bk = Bk;
  const double chi_beta = 1.0 / (1 - 0.5 * aZ);
  F.chi_p = chi_beta * A;
  F.chi_d = chi_beta * F.k2 * Bk;
  F.chi_g = chi_beta * 0.5 * aZ;
  F.chi_e = chi_beta * (aZ - 2 * aEta) * k;
  // qgal, galileon.cc:897-936
  const double qa = C4 * x4 * h4 - 2 * C5 * x5 * h6 - CG / 1.5 * x2 * h2;
  const double q_beta = 1.0 / (1 - 1.5 * qa);
  const double Qd = h0 * (c2 * X * H - 6 * C3 * x2 * h3 + 18 * C4 * x3 * h5 - 15 * C5 * x4 * h7 - 6 * CG * X * h3);
  F.q_d = q_beta * k * Qd;
  F.q_p = -q_beta * k * Bk;
  F.q_g = q_beta * 1.5 * qa;
}

// background part in one call: hbar = H/a and x from the spline, dh and dx from the background system
GC_HD void gal_fast_background(const GalConst& G, double a, double hbar, double X, double lam_nu, double k, GalFast& F,
                               GalDyn* dyn = nullptr) {
  double dh, dx;
  gal_fast_dhdx(G, a, hbar, X, lam_nu, dh, dx, dyn);
  gal_fast_fill(G, a, hbar, X, dh, dx, k, F);
}

GC_HD double gal_fast_Chigal(const GalFast& F, double dgrho, double eta, double dphi, double dphip) {
  return F.chi_p * dphip + F.chi_d * dphi + F.chi_g * dgrho + F.chi_e * eta;
}

GC_HD double gal_fast_qgal(const GalFast& F, double dgq, double dphi, double dphip) {
  return F.q_d * dphi + F.q_p * dphip + F.q_g * dgq;
}

// the three anisotropic-stress response coefficients shared by Pigal and pigalprime (galileon.cc:951-963):
// a_sp multiplies sigma', a_s multiplies sigma, a_ph multiplies phi
struct GalStress {
  double a_sp, a_s, a_ph;
};
GC_HD GalStress gal_fast_stress(const GalFast& F) {
  const double C4 = F.C4, C5 = F.C5, CG = F.CG, X = F.X, H = F.H, hp = F.hprime, xhp = F.xhprime;
  const double x2 = F.x2, x3 = F.x3, x4 = F.x4, x5 = F.x5, h2 = F.h2, h3 = F.h3, h4 = F.h4, h5 = F.h5, h6 = F.h6;
  GalStress S;
  S.a_sp = C4 * x4 * h4 - 3 * C5 * x4 * h5 * xhp;
  S.a_s = C4 * (3 * x4 * h4 - 6 * x3 * h3 * xhp) - C5 * (-3 * x5 * h5 * hp + 12 * x5 * h6 - 15 * x4 * h5 * xhp) + 2 * CG * X * H * xhp;
  S.a_ph = -C4 * x4 * h4 - C5 * (6 * x4 * h5 * xhp - 6 * x5 * h6) + 2 * CG * x2 * h2;
  return S;
}

// Pigal (galileon.cc:938-985) as a linear form
GC_HD void gal_fast_Pigal_lin(const GalConst& G, const GalFast& F, GalPiLin& P) {
  if (!F.on) {
    P = GalPiLin{0, 0, 0, 0, 0};
    return;
  }
  const double C4 = F.C4, C5 = F.C5, CG = F.CG, H = F.H, hp = F.hprime, xhp = F.xhprime;
  const double x2 = F.x2, x3 = F.x3, x4 = F.x4, h3 = F.h3, h4 = F.h4, h5 = F.h5, h6 = F.h6;
  const GalStress S = gal_fast_stress(F);
  const double beta_pi = 1.0 / (1 + 0.5 * S.a_ph - S.a_sp);
  const double half = 0.5 * (2 * S.a_sp + S.a_s - S.a_ph);
  P.d = beta_pi * F.k2 *
        (C4 * (4 * x3 * h4 - 6 * x2 * h3 * xhp) - C5 * (-12 * x3 * h5 * xhp + 12 * x4 * h6 - 3 * x4 * h5 * hp) + 2 * CG * H * xhp);
  P.pi = beta_pi * (S.a_sp - 0.5 * S.a_ph);
  P.g = beta_pi * half;
  P.q = beta_pi * half * (3 * G.h0 * H / F.k);
  P.e = beta_pi * (S.a_s + S.a_sp) * F.k;
}
GC_HD double gal_fast_Pigal(const GalConst& G, const GalFast& F, double dgrho, double dgq, double dgpi, double eta, double dphi) {
  GalPiLin P;
  gal_fast_Pigal_lin(G, F, P);
  return P.d * dphi + P.pi * dgpi + P.g * dgrho + P.q * dgq + P.e * eta;
}

// dphisecond (galileon.cc:987-1081) as a linear form.  chiprimehat of the reference is itself linear in
// (dphiprime, dphi); its two coefficients are folded into p and d.
GC_HD void gal_fast_dphisecond_lin(const GalConst& G, const GalFast& F, GalPhi2Lin& D) {
  if (!F.on) {
    D = GalPhi2Lin{0, 0, 0, 0, 0, 0};
    return;
  }
  const double c2 = G.c2, h0 = G.h0, h02 = h0 * h0;
  const double C3 = F.C3, C4 = F.C4, C5 = F.C5, CG = F.CG, X = F.X, H = F.H, hp = F.hprime, xhp = F.xhprime, dx = F.dx;
  const double x2 = F.x2, x3 = F.x3, x4 = F.x4, x5 = F.x5;
  const double h2 = F.h2, h3 = F.h3, h4 = F.h4, h5 = F.h5, h6 = F.h6, h7 = F.h7, h8 = F.h8;
  const double k2 = F.k2, k = F.k;
  const double a_g2 = c2 - 12 * C3 * X * h2 + 54 * C4 * x2 * h4 - 60 * C5 * x3 * h6 - 6 * CG * h2;
  const double a_g1 = 2 * c2 - C3 * (12 * H * xhp + 12 * X * H * hp) + C4 * (-108 * x2 * h4 + 108 * X * h3 * xhp + 108 * x2 * h3 * hp) -
                      C5 * (-240 * x3 * h6 + 180 * x2 * h5 * xhp + 180 * x3 * h5 * hp) - 12 * CG * H * hp;
  const double a_g0 = c2 - C3 * (4 * X * h2 + 4 * H * xhp) + C4 * (-10 * x2 * h4 + 24 * X * h3 * xhp + 12 * x2 * h3 * hp) -
                      C5 * (-36 * x3 * h6 + 36 * x2 * h5 * xhp + 24 * x3 * h5 * hp) - CG * (2 * h2 + 4 * H * hp);
  const double a_Z = c2 * X - C3 * (6 * x2 * h2 + 4 * X * H * xhp) + C4 * (-6 * x3 * h4 + 36 * x2 * h3 * xhp + 12 * x3 * h3 * hp) -
                     C5 * (-45 * x4 * h6 + 60 * x3 * h5 * xhp + 30 * x4 * h5 * hp) - CG * (6 * X * h2 + 4 * H * xhp + 4 * X * H * hp);
  const double a_eta = C4 * (-4 * x3 * h4 + 6 * x2 * h3 * xhp) - C5 * (-12 * x4 * h6 + 3 * x4 * h5 * hp + 12 * x3 * h5 * xhp) - 2 * CG * H * xhp;
  // chiprimehat = cp_p * dphiprime + cp_d * dphi
  const double cp_p = c2 * h02 * (-2 * X * h2 + X * H * hp + h2 * dx) -
                      C3 * (h02 * (-72 * x2 * h4 + 54 * x2 * h3 * hp + 36 * X * h4 * dx) + k2 * (2 * x2 * h2)) +
                      C4 * (h02 * (-540 * x3 * h6 + 450 * x3 * h5 * hp + 270 * x2 * h6 * dx) + k2 * (12 * x3 * h4)) -
                      C5 * (h02 * (-840 * x4 * h8 + 735 * x4 * h7 * hp + 420 * x3 * h8 * dx) + k2 * (15 * x4 * h6)) -
                      CG * (h02 * (-72 * X * h4 + 54 * X * h3 * hp + 18 * h4 * dx) + k2 * (4 * X * h2));
  const double cp_d = h0 * k2 *
                      (-C3 * (-8 * x2 * h3 + 4 * x2 * h2 * hp + 4 * X * h3 * dx) + C4 * (-72 * x3 * h5 + 48 * x3 * h4 * hp + 36 * x2 * h5 * dx) -
                       C5 * (-120 * x4 * h7 + 90 * x4 * h6 * hp + 60 * x3 * h7 * dx) - CG * (-16 * X * h3 + 8 * X * h2 * hp + 4 * h3 * dx));
  const double b_Z = F.aZ, b_eta = F.aEta;
  const double b_Zp = h0 * (-C3 * (-4 * x3 * h3 + 4 * x3 * h2 * hp + 6 * x2 * h3 * dx) + C4 * (-60 * x4 * h5 + 60 * x4 * h4 * hp + 60 * x3 * h5 * dx) -
                            C5 * (-126 * x5 * h7 + 126 * x5 * h6 * hp + 105 * x4 * h7 * dx) - CG * (-12 * x2 * h3 + 12 * x2 * h2 * hp + 12 * X * h3 * dx));
  const double b_ep = h0 * (C4 * (-6 * x4 * h5 + 6 * x4 * h4 * hp + 6 * x3 * h5 * dx) - C5 * (-18 * x5 * h7 + 18 * x5 * h6 * hp + 15 * x4 * h7 * dx) -
                            CG * (-2 * x2 * h3 + 2 * x2 * h2 * hp + 2 * X * h3 * dx));
  const double hH = h0 * H, hhp = h0 * hp;
  const double ksi = F.Bk / (hH * (2 - b_Z));
  const double mden = -1.0 / (a_g2 + ksi * F.A);
  D.p = mden * (a_g1 * hH + ksi * cp_p);
  D.d = mden * (a_g0 * k2 + ksi * cp_d);
  D.g = mden * (0.5 * a_Z + 2 * ksi * (hH - 0.5 * (hhp + b_Z * hH) + 0.25 * (b_Zp + b_Z * hhp)));
  D.q = mden * ksi * (1 - b_eta) * k;
  D.f = mden * ksi;
  D.e = mden * (a_Z - 2 * a_eta + 2 * ksi * (2 * b_eta * hH - hhp - b_Z * hH - b_ep + 0.5 * (b_Zp + b_Z * hhp))) * k;
}
GC_HD double gal_fast_dphisecond(const GalConst& G, const GalFast& F, double dgrho, double dgq, double eta, double dphi,
                                 double dphip, double dotdeltaf) {
  GalPhi2Lin D;
  gal_fast_dphisecond_lin(G, F, D);
  return D.p * dphip + D.d * dphi + D.g * dgrho + D.q * dgq + D.f * dotdeltaf + D.e * eta;
}

// pigalprime (galileon.cc:1084-1174): conformal-time derivative of the Galileon anisotropic stress, only needed by
// the ISW term of output().  Organisation here:
//   1. the background system (alpha ... omega of GalDyn) is differentiated once more: every polynomial P(g, x) has
//      dP/dlna = P_g * dh + P_x * dx (the reference's (hprime - hc)/a IS dh), which gives xprimedot = d(dx)/dtau and
//      hprimedot = d(hprime)/dtau in units of H0;
//   2. time derivatives of the three stress coefficients (GalStress);
//   3. the result as a linear form in (dphiprime, dphi, pidot, dgrho, dgq, dgpi, eta).
// lam_nu_prime = sum_i grhormass_i * (dpnu_i - 4 pnu_i H) with dpnu_i = Nu_dp(a m_i, H, pnu_i) (the reference hands
// Nu_dp the H0-unit Hubble rate H as its "adotoa", galileon.cc:1133-1137); grho_plus_gpres = grho + gpres.
GC_HD void gal_fast_pigalprime_lin(const GalConst& G, const GalFast& F, const GalDyn& B, double lam_nu_prime,
                                   double grho_plus_gpres, GalPiDotLin& L) {
  if (!F.on) {
    L = GalPiDotLin{0, 0, 0, 0, 0, 0, 0};
    return;
  }
  const double c2 = G.c2, c3 = G.c3, c4 = G.c4, c5 = G.c5, cG = G.cG, h0 = G.h0, h02 = h0 * h0;
  const double a = F.a, X = F.X, H = F.H, dh = F.dh, dx = F.dx, hp = F.hprime, xhp = F.xhprime;
  const double x2 = F.x2, x3 = F.x3, x4 = F.x4, x5 = F.x5;
  const double g = B.g, g2 = g * g, g3 = g2 * g, g4 = g2 * g2, g5 = g3 * g2, g6 = g3 * g3, g7 = g4 * g3, g8 = g4 * g4;
  const double ia4 = 1.0 / (a * a * a * a);
  // ---- 1. d/dlna of the six polynomials: partial derivatives with respect to g and x
  const double al_g = c2 / 6 * X - 9 * c3 * g2 * x2 + 75 * c4 * g4 * x3 - 122.5 * c5 * g6 * x4 - 9 * cG * g2 * X;
  const double al_x = c2 / 6 * g - 6 * c3 * g3 * X + 45 * c4 * g5 * x2 - 70 * c5 * g7 * x3 - 3 * cG * g3;
  const double ga_g = 2. * c2 / 3 * g * X - 4 * c3 * g3 * x2 + 20 * c5 * g7 * x4 - 8 * cG * g3 * X;
  const double ga_x = c2 / 3 * g2 - 2 * c3 * g4 * X + 10 * c5 * g8 * x3 - 2 * cG * g4;
  const double be_g = c2 / 3 * g - 8 * c3 * g3 * X + 54 * c4 * g5 * x2 - 80 * c5 * g7 * x3 - 4 * cG * g3;
  const double be_x = -2 * c3 * g4 + 18 * c4 * g6 * X - 30 * c5 * g8 * x2;
  const double de_g = 2 + 6 * c3 * g2 * x3 - 75 * c4 * g4 * x4 + 147 * c5 * g6 * x5 + 18 * cG * g2 * x2;
  const double de_x = 6 * c3 * g3 * x2 - 60 * c4 * g5 * x3 + 105 * c5 * g7 * x4 + 12 * cG * g3 * X;
  const double la_g = 6 * g + c2 * g * x2 - 8 * c3 * g3 * x3 + 45 * c4 * g5 * x4 - 72 * c5 * g7 * x5 - 4 * cG * g3 * x2;
  const double la_x = c2 * g2 * X - 6 * c3 * g4 * x2 + 30 * c4 * g6 * x3 - 45 * c5 * g8 * x4 - 2 * cG * g4 * X;
  const double om_g = 8 * c3 * g3 * x2 - 72 * c4 * g5 * x3 + 120 * c5 * g7 * x4 + 16 * cG * g3 * X;
  const double om_x = 4 * c3 * g4 * X - 36 * c4 * g6 * x2 + 60 * c5 * g8 * x3 + 4 * cG * g4;
  const double al_p = al_g * dh + al_x * dx, ga_p = ga_g * dh + ga_x * dx, be_p = be_g * dh + be_x * dx;
  const double de_p = de_g * dh + de_x * dx, om_p = om_g * dh + om_x * dx;
  const double la_p = -4 * G.orad * ia4 + la_g * dh + la_x * dx + lam_nu_prime * ia4 / h02;
  const double den = B.delta * B.beta - B.alpha * B.omega, iden2 = 1.0 / (den * den);
  const double den_p = de_p * B.beta + B.delta * be_p - al_p * B.omega - B.alpha * om_p;
  const double nx = B.alpha * B.lambda - B.delta * B.gamma;
  const double nx_p = al_p * B.lambda + B.alpha * la_p - de_p * B.gamma - B.delta * ga_p;
  const double nh = B.omega * B.gamma - B.lambda * B.beta;
  const double nh_p = om_p * B.gamma + B.omega * ga_p - la_p * B.beta - B.lambda * be_p;
  const double hH = h0 * H, hhp = h0 * hp;
  const double xpd = hH * (-dx + (nx_p * den - nx * den_p) * iden2);                    // d(dx)/dtau
  const double hpd = a * hH * (nh / den + (nh_p * den - nh * den_p) * iden2) + hH * hp;  // d(hprime)/dtau
  // ---- 2. stress coefficients and their conformal-time derivatives
  const double C4 = F.C4, C5 = F.C5, CG = F.CG;
  const double h2 = F.h2, h3 = F.h3, h4 = F.h4, h5 = F.h5, h6 = F.h6, h7 = F.h7;
  const double hp2 = hp * hp, dx2 = dx * dx, hpdx = hp * dx;
  const GalStress S = gal_fast_stress(F);
  const double as_t = C4 * (h0 * (-12 * x4 * h5 + 36 * x4 * h4 * hp - 18 * x4 * h3 * hp2 + 36 * x3 * h5 * dx - 48 * x3 * h4 * hpdx - 18 * x2 * h5 * dx2) -
                            6 * x4 * h3 * hpd - 6 * x3 * h4 * xpd) -
                      C5 * (h0 * (-72 * x5 * h7 + 180 * x5 * h6 * hp - 90 * x5 * h5 * hp2 + 150 * x4 * h7 * dx - 180 * x4 * h6 * hpdx - 60 * x3 * h7 * dx2) -
                            18 * x5 * h5 * hpd - 15 * x4 * h6 * xpd) -
                      CG * (h0 * (4 * x2 * h2 * hp - 2 * x2 * H * hp2 + 4 * X * h3 * dx - 8 * X * h2 * hpdx - 2 * h3 * dx2) - 2 * x2 * H * hpd -
                            2 * X * h2 * xpd);
  const double asp_t = C4 * h0 * (-4 * x4 * h5 + 4 * x4 * h4 * hp + 4 * x3 * h5 * dx) -
                       C5 * (h0 * (-18 * x5 * h6 * hp + 15 * x5 * h5 * hp2 - 18 * x4 * h7 * dx + 33 * x4 * h6 * hpdx + 12 * x3 * h7 * dx2) +
                             3 * x5 * h5 * hpd + 3 * x4 * h6 * xpd);
  const double aph_t = C4 * h0 * (4 * x4 * h5 - 4 * x4 * h4 * hp - 4 * x3 * h5 * dx) -
                       C5 * (h0 * (36 * x5 * h7 - 72 * x5 * h6 * hp + 30 * x5 * h5 * hp2 - 66 * x4 * h7 * dx + 66 * x4 * h6 * hpdx + 24 * x3 * h7 * dx2) +
                             6 * x5 * h5 * hpd + 6 * x4 * h6 * xpd) -
                       CG * h0 * (4 * x2 * h3 - 4 * x2 * h2 * hp - 4 * X * h3 * dx);
  // ---- 3. the linear form
  const double k = F.k, k2 = F.k2;
  const double inv = 1.0 / (1. - S.a_sp + 0.5 * S.a_ph);
  const double dps = S.a_ph - S.a_sp;  // alpha_phi - alpha_sigprime
  const double w = grho_plus_gpres / hH;
  L.p = inv * k2 * (C4 * (4 * x3 * h4 - 6 * x3 * h3 * hp - 6 * x2 * h4 * dx) - C5 * (12 * x4 * h6 - 15 * x4 * h5 * hp - 12 * x3 * h6 * dx) -
                    CG * (-2 * X * H * hp - 2 * h2 * dx));
  L.d = inv * k2 *
        (C4 * (h0 * (-24 * x3 * h5 + 52 * x3 * h4 * hp - 18 * x3 * h3 * hp2 + 48 * x2 * h5 * dx - 42 * x2 * h4 * hpdx - 12 * X * h5 * dx2) -
               6 * x3 * h3 * hpd - 6 * x2 * h4 * xpd) -
         C5 * (h0 * (-96 * x4 * h7 + 192 * x4 * h6 * hp - 75 * x4 * h5 * hp2 + 144 * x3 * h7 * dx - 132 * x3 * h6 * hpdx - 36 * x2 * h7 * dx2) -
               15 * x4 * h5 * hpd - 12 * x3 * h6 * xpd) -
         CG * (h0 * (8 * X * h2 * hp - 2 * X * H * hp2 + 8 * h3 * dx - 6 * h2 * hpdx) - 2 * X * H * hpd - 2 * h2 * xpd));
  L.pid = inv * (S.a_sp - 0.5 * S.a_ph);
  L.g = inv * (-3.5 * hH * S.a_sp + 0.5 * hhp * S.a_sp + 0.25 * dps * w + 0.5 * hhp * S.a_s + 0.5 * as_t + asp_t - 2 * hH * S.a_s +
               1.5 * hH * S.a_ph - 0.5 * aph_t);
  L.e = inv * (as_t + asp_t + S.a_sp * hhp + S.a_s * hhp - 3 * hH * S.a_sp - 3 * hH * S.a_s + 0.5 * dps * w) * k;
  L.q = inv * (0.5 / k) *
        (3 * h02 * H * hp * S.a_sp + 3 * h02 * H * hp * S.a_s + 1.5 * dps * grho_plus_gpres + 3 * hH * as_t - 12 * h02 * h2 * S.a_s -
         21 * h02 * h2 * S.a_sp - 3 * hH * aph_t + 9 * h02 * h2 * S.a_ph + 6 * hH * asp_t + k2 * dps);
  L.pi = inv * (-2 * hH * S.a_sp + hH * S.a_ph - 0.5 * aph_t + asp_t - hH * S.a_s);
}

}  // namespace gc
