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

// background part: GetdHdX (:746-787), grhogal (:809-829), gpresgal (:831-850) + coefficient set-up.
// hbar = H/a; lam_nu = sum_i grhormass_i * pnu_i (massive-neutrino pressure term of lambda).
GC_HD void gal_fast_background(const GalConst& G, double a, double hbar, double X, double lam_nu, double k, GalFast& F) {
  const double c2 = G.c2, c3 = G.c3, c4 = G.c4, c5 = G.c5, cG = G.cG, h0 = G.h0;
  F.a = a;
  F.X = X;
  F.k = k;
  F.k2 = k * k;
  const double H = a * hbar;
  F.H = H;
  const double a2 = a * a;
  const double ia2 = 1.0 / a2, ia4 = ia2 * ia2, ia6 = ia4 * ia2;
  const double x2 = X * X, x3 = x2 * X, x4 = x2 * x2, x5 = x3 * x2;
  {  // GetdHdX in g = hbar
    const double g = hbar, g2 = g * g, g3 = g2 * g, g4 = g2 * g2, g5 = g3 * g2, g6 = g3 * g3, g7 = g4 * g3, g8 = g4 * g4;
    const double alpha = c2 / 6 * g * X - 3 * c3 * g3 * x2 + 15 * c4 * g5 * x3 - 17.5 * c5 * g7 * x4 - 3 * cG * g3 * X;
    const double gamma = c2 / 3 * g2 * X - c3 * g4 * x2 + 2.5 * c5 * g8 * x4 - 2 * cG * g4 * X;
    const double beta = c2 / 6 * g2 - 2 * c3 * g4 * X + 9 * c4 * g6 * x2 - 10 * c5 * g8 * x3 - cG * g4;
    const double sigma = 2 * g + 2 * c3 * g3 * x3 - 15 * c4 * g5 * x4 + 21 * c5 * g7 * x5 + 6 * cG * g3 * x2;
    const double lambda = 3 * g2 + G.orad * ia4 + c2 / 2 * g2 * x2 - 2 * c3 * g4 * x3 + 7.5 * c4 * g6 * x4 - 9 * c5 * g8 * x5 -
                          cG * g4 * x2 + lam_nu * ia4 / (h0 * h0);
    const double omega = 2 * c3 * g4 * x2 - 12 * c4 * g6 * x3 + 15 * c5 * g8 * x4 + 4 * cG * g4 * X;
    const double iden = 1.0 / (sigma * beta - alpha * omega);
    F.dh = (omega * gamma - lambda * beta) * iden;
    F.dx = -X + (alpha * lambda - sigma * gamma) * iden;
    F.on = a >= 9.99999e-7;
    if (F.on) {
      const double dh = F.dh, dx = F.dx;
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
  }
  F.hprime = a * F.dh + H;
  F.xhprime = F.dx * H + X * F.hprime;
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

GC_HD double gal_fast_Chigal(const GalFast& F, double dgrho, double eta, double dphi, double dphip) {
  return F.chi_p * dphip + F.chi_d * dphi + F.chi_g * dgrho + F.chi_e * eta;
}

GC_HD double gal_fast_qgal(const GalFast& F, double dgq, double dphi, double dphip) {
  return F.q_d * dphi + F.q_p * dphip + F.q_g * dgq;
}

// galileon.cc:938-985
GC_HD double gal_fast_Pigal(const GalConst& G, const GalFast& F, double dgrho, double dgq, double dgpi, double eta, double dphi) {
  if (!F.on) return 0;
  const double C4 = F.C4, C5 = F.C5, CG = F.CG, X = F.X, H = F.H, hprime = F.hprime, xhprime = F.xhprime;
  const double x2 = F.x2, x3 = F.x3, x4 = F.x4, x5 = F.x5, h2 = F.h2, h3 = F.h3, h4 = F.h4, h5 = F.h5, h6 = F.h6;
  const double a_sp = C4 * x4 * h4 - 3 * C5 * x4 * h5 * xhprime;
  const double a_s = C4 * (3 * x4 * h4 - 6 * x3 * h3 * xhprime) - C5 * (-3 * x5 * h5 * hprime + 12 * x5 * h6 - 15 * x4 * h5 * xhprime) +
                     2 * CG * X * H * xhprime;
  const double a_ph = -C4 * x4 * h4 - C5 * (6 * x4 * h5 * xhprime - 6 * x5 * h6) + 2 * CG * x2 * h2;
  const double beta_pi = 1.0 / (1 + 0.5 * a_ph - a_sp);
  const double Pit = F.k2 * dphi * (C4 * (4 * x3 * h4 - 6 * x2 * h3 * xhprime) -
                                    C5 * (-12 * x3 * h5 * xhprime + 12 * x4 * h6 - 3 * x4 * h5 * hprime) + 2 * CG * H * xhprime);
  return beta_pi * (Pit + (a_sp - 0.5 * a_ph) * dgpi + 0.5 * (2 * a_sp + a_s - a_ph) * (dgrho + 3 * G.h0 * H / F.k * dgq) +
                    (a_s + a_sp) * F.k * eta);
}

// galileon.cc:987-1081
GC_HD double gal_fast_dphisecond(const GalConst& G, const GalFast& F, double dgrho, double dgq, double eta, double dphi,
                                 double dphip, double dotdeltaf) {
  if (!F.on) return 0;
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
  const double a_Zp = F.Bk;
  const double a_eta = C4 * (-4 * x3 * h4 + 6 * x2 * h3 * xhp) - C5 * (-12 * x4 * h6 + 3 * x4 * h5 * hp + 12 * x3 * h5 * xhp) - 2 * CG * H * xhp;
  const double hd = h0 * dphi;
  const double chiprimehat =
      c2 * h02 * dphip * (-2 * X * h2 + X * H * hp + h2 * dx) -
      C3 * (h02 * dphip * (-72 * x2 * h4 + 54 * x2 * h3 * hp + 36 * X * h4 * dx) +
            k2 * (2 * x2 * h2 * dphip + hd * (-8 * x2 * h3 + 4 * x2 * h2 * hp + 4 * X * h3 * dx))) +
      C4 * (h02 * dphip * (-540 * x3 * h6 + 450 * x3 * h5 * hp + 270 * x2 * h6 * dx) +
            k2 * (12 * x3 * h4 * dphip + hd * (-72 * x3 * h5 + 48 * x3 * h4 * hp + 36 * x2 * h5 * dx))) -
      C5 * (h02 * dphip * (-840 * x4 * h8 + 735 * x4 * h7 * hp + 420 * x3 * h8 * dx) +
            k2 * (15 * x4 * h6 * dphip + hd * (-120 * x4 * h7 + 90 * x4 * h6 * hp + 60 * x3 * h7 * dx))) -
      CG * (h02 * dphip * (-72 * X * h4 + 54 * X * h3 * hp + 18 * h4 * dx) +
            k2 * (4 * X * h2 * dphip + hd * (-16 * X * h3 + 8 * X * h2 * hp + 4 * h3 * dx)));
  const double b_g2 = F.A;
  const double b_Z = F.aZ, b_eta = F.aEta;
  const double b_Zp = h0 * (-C3 * (-4 * x3 * h3 + 4 * x3 * h2 * hp + 6 * x2 * h3 * dx) + C4 * (-60 * x4 * h5 + 60 * x4 * h4 * hp + 60 * x3 * h5 * dx) -
                            C5 * (-126 * x5 * h7 + 126 * x5 * h6 * hp + 105 * x4 * h7 * dx) - CG * (-12 * x2 * h3 + 12 * x2 * h2 * hp + 12 * X * h3 * dx));
  const double b_ep = h0 * (C4 * (-6 * x4 * h5 + 6 * x4 * h4 * hp + 6 * x3 * h5 * dx) - C5 * (-18 * x5 * h7 + 18 * x5 * h6 * hp + 15 * x4 * h7 * dx) -
                            CG * (-2 * x2 * h3 + 2 * x2 * h2 * hp + 2 * X * h3 * dx));
  const double hH = h0 * H, hhp = h0 * hp;
  const double ksi = a_Zp / (hH * (2 - b_Z));
  const double num = a_g1 * hH * dphip + a_g0 * k2 * dphi +
                     (0.5 * a_Z + 2 * ksi * (hH - 0.5 * (hhp + b_Z * hH) + 0.25 * (b_Zp + b_Z * hhp))) * dgrho +
                     ksi * (1 - b_eta) * k * dgq + ksi * dotdeltaf + ksi * chiprimehat +
                     (a_Z - 2 * a_eta + 2 * ksi * (2 * b_eta * hH - hhp - b_Z * hH - b_ep + 0.5 * (b_Zp + b_Z * hhp))) * k * eta;
  return -num / (a_g2 + ksi * b_g2);
}

}  // namespace gc
