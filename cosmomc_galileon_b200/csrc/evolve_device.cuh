// Per-(k, cosmology) scalar perturbation evolution: device functions.
//
// Restates, for sm_100a, the reference routines
//   GetNumEqns / SetupScalarArrayIndices / CopyScalarVariableArray   camb/equations_galileon.f90:555-946
//   initial / derivs / output                                         camb/equations_galileon.f90:1714-1928,
//                                                                     :2098-2589, :1256-1576
//   the Galileon closed forms                                          camb/galileon.cc:691-1174
//   thermo / massive-neutrino table lookups                            camb/modules.f90:2728-2769, :1721-1850
// One lane owns one (k, cosmology) pair.  A lane's state vector y and the nine dverk work columns live in a
// warp-interleaved global workspace: element i of column c is at ws[(c*NV + i)*32 + lane], so the 32 lanes of
// a warp touch one 256-byte line per element (see DESIGN.md "data layout").
#pragma once
#include <cmath>
#include <cstdint>

#include "galileon_fast.cuh"
#include "model.h"

// functions shared with the host-side galileon.cc replacement (csrc/galileon_host.cu)
#define GC_HD __host__ __device__
// read-only table rows (thermo, background spline, neutrino tables): global, never written by the kernel
#ifdef __CUDA_ARCH__
#define GC_LDG(p) __ldg(p)
#else
#define GC_LDG(p) (*(p))
#endif
#define GC_LANES 32
#define GC_NCOL 12  // column 0 = y, columns 1..9 = dverk w(:,1..9), 10 = y of a pending output, 11 = its y'

__constant__ double c_denl[260];    // 1/(2l+1)            camb/equations_galileon.f90:525-527
__constant__ double c_polfac[260];  // (l+3)(l-1)/(l+1)    camb/equations_galileon.f90:516-520

namespace gc {

// float-rounded literals of the reference's default-real constants (no _dl suffix)
#define GC_SP(x) ((double)(x##f))

struct EV {
  double q, k, k2, ik, ik2;  // ik = 1/k, ik2 = 1/k^2 (per-lane constants)
  int w_ix, r_ix, g_ix, polind, nu_pert_ix, nvar, neq;
  int lmaxg, lmaxnr, lmaxnu, lmaxgpol, lmaxnu_pert;
  int nu_ix[GC_MAX_NU], nq[GC_MAX_NU], lmaxnu_tau[GC_MAX_NU];
  bool TightCoupling, high_ktau, no_nu, no_phot, has_nu_rel;
  bool MassiveNuApprox[GC_MAX_NU], nu_nonrel[GC_MAX_NU];
  double G11[GC_MAX_NU], G30[GC_MAX_NU], MassiveNuApproxTime[GC_MAX_NU];
  double TightSwitchoffTime, pig, pigdot, poltruncfac;
};

// column element accessors (1-based i like the reference)
#define AT(col, i) (col)[((i)-1) * GC_LANES]

__device__ __forceinline__ double denlk(const EV& e, int l) { return c_denl[l] * e.k * l; }
__device__ __forceinline__ double denlk2(const EV& e, int l) { return c_denl[l] * e.k * 1.0 * (l + 1); }
__device__ __forceinline__ double polfack(const EV& e, int l) { return c_polfac[l] * e.k * 1.0 * c_denl[l]; }

__device__ __forceinline__ long nintd(double x) { return lround(x); }

// ---------------------------------------------------------------- table lookups
GC_HD __forceinline__ double herm(double f0, double d0, double f1, double d1, double d) {
  return f0 + d * (d0 + d * (3 * (f1 - f0) - 2 * d0 - d1 + d * (d0 + d1 + 2 * (f0 - f1))));
}

// camb/modules.f90:2728-2769
__device__ inline void thermo(const DevModel& M, double tau, double& cs2b, double& opacity, double* dopacity) {
  double d = log(tau * M.inv_tauminn) * M.inv_dlntau + 1.0;
  int i = (int)d;
  d = d - i;
  const int nth = M.nthermo;
  if (i >= nth) {
    const double* r = M.th + (size_t)(nth - 1) * 5;
    cs2b = GC_LDG(r) + (d + i - nth) * GC_LDG(r + 1);
    opacity = GC_LDG(r + 2) + (d - nth) * GC_LDG(r + 3);
    if (dopacity) *dopacity = 0;
    return;
  }
  if (i < 1) i = 1;  // reference aborts ("thermo out of bounds"); cannot happen for tau >= tauminn
  const double* r0 = M.th + (size_t)(i - 1) * 5;
  const double* r1 = r0 + 5;
  double a0[5], a1[5];
#pragma unroll
  for (int c = 0; c < 5; c++) {
    a0[c] = GC_LDG(r0 + c);
    a1[c] = GC_LDG(r1 + c);
  }
  cs2b = herm(a0[0], a0[1], a1[0], a1[1], d);
  opacity = herm(a0[2], a0[3], a1[2], a1[3], d);
  if (dopacity) *dopacity = herm(a0[3], a0[4], a1[3], a1[4], d) * M.inv_dlntau / tau;
}

// camb/modules.f90:2771-2783 via the running-minimum table (same index as the linear scan)
__device__ inline double Thermo_OpacityToTime(const DevModel& M, double opacity) {
  int lo = 0, hi = M.nthermo - 1;  // find first j (0-based) with pmin[j] <= opacity
  if (M.dotmu_pmin[hi] > opacity) return exp((double)(M.nthermo) * M.dlntau) * M.tauminn;  // off the table
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (M.dotmu_pmin[mid] <= opacity)
      hi = mid;
    else
      lo = mid + 1;
  }
  return exp((double)lo * M.dlntau) * M.tauminn;
}

// massive neutrino tables, camb/modules.f90:1544-1560
#define GC_NU_CONST (7.0 / 120 * (3.1415926535897932384626433832795 * 3.1415926535897932384626433832795 * 3.1415926535897932384626433832795 * 3.1415926535897932384626433832795))
#define GC_NU_CONST2 (5.0 / 7 / (3.1415926535897932384626433832795 * 3.1415926535897932384626433832795))
#define GC_ZETA3 1.2020569031595942853997
#define GC_ZETA5 1.0369277551433699263313
#define GC_ZETA7 1.0083492773819228268397
#define GC_AM_MIN 0.01
#define GC_AM_MAX 600.0
#define GC_AM_MINP (0.01 * GC_SP(1.1))
#define GC_AM_MAXP (600.0 * GC_SP(0.9))

// nu_tab rows: {r1, dr1, p1, dp1, ddr1, ddp1}
// camb/modules.f90:1721-1754
GC_HD inline void Nu_background(const DevModel& M, double am, double& rhonu, double& pnu) {
  if (am <= GC_AM_MINP) {
    rhonu = 1.0 + GC_NU_CONST2 * am * am;
    pnu = (2 - rhonu) / 3.0;
    return;
  } else if (am >= GC_AM_MAXP) {
    rhonu = 3 / (2 * GC_NU_CONST) * (GC_ZETA3 * am + (15 * GC_ZETA5) / 2 / am);
    pnu = 900.0 / 120.0 / GC_NU_CONST * (GC_ZETA5 - 63.0 / 4 * GC_ZETA7 / (am * am)) / am;
    return;
  }
  double d = log(am * 100.0) * M.inv_dlnam + 1.0;
  int i = (int)d;
  d = d - i;
  const double* r0 = M.nu_tab + (size_t)(i - 1) * 6;
  const double* r1 = r0 + 6;
  rhonu = exp(herm(GC_LDG(r0 + 0), GC_LDG(r0 + 1), GC_LDG(r1 + 0), GC_LDG(r1 + 1), d));
  pnu = exp(herm(GC_LDG(r0 + 2), GC_LDG(r0 + 3), GC_LDG(r1 + 2), GC_LDG(r1 + 3), d));
}

// camb/modules.f90:1790-1817
GC_HD inline double Nu_drho(const DevModel& M, double am, double adotoa, double rhonu) {
  if (am < GC_AM_MINP) return 2 * GC_NU_CONST2 * am * am * adotoa;
  if (am > GC_AM_MAXP) return 3 / (2 * GC_NU_CONST) * (GC_ZETA3 * am - (15 * GC_ZETA5) / 2 / am) * adotoa;
  double d = log(am * 100.0) * M.inv_dlnam + 1.0;
  int i = (int)d;
  d = d - i;
  const double* r0 = M.nu_tab + (size_t)(i - 1) * 6;
  const double* r1 = r0 + 6;
  return rhonu * adotoa * herm(GC_LDG(r0 + 1), GC_LDG(r0 + 4), GC_LDG(r1 + 1), GC_LDG(r1 + 4), d) / M.dlnam;
}

// camb/modules.f90:1821-1850
GC_HD inline double Nu_dp(const DevModel& M, double am, double adotoa, double pnu) {
  if (am < GC_AM_MINP) return 2 * adotoa * (1.0 - GC_NU_CONST2 * am * am) / 3.0 - 4.0 / 3.0 * adotoa;
  if (am > GC_AM_MAXP)
    return -900.0 / 120.0 / GC_NU_CONST * (GC_ZETA5 - 189.0 / 4 * GC_ZETA7 / (am * am)) * adotoa / am;
  double d = log(am * 100.0) * M.inv_dlnam + 1.0;
  int i = (int)d;
  d = d - i;
  const double* r0 = M.nu_tab + (size_t)(i - 1) * 6;
  const double* r1 = r0 + 6;
  return pnu * adotoa * herm(GC_LDG(r0 + 3), GC_LDG(r0 + 5), GC_LDG(r1 + 3), GC_LDG(r1 + 5), d) / M.dlnam;
}

// Galileon background spline: hbar(a) and x(a)=a*dpi/da from one row search
// (camb/galileon.cc:691-744; natural cubic spline on the geometric a grid, bisection-equivalent index).
__device__ inline void gal_HX(const DevModel& M, double a, double& hbar, double& xgal) {
  const int n = M.ngal;
  int i = (int)(log(a / M.gal_a0) * M.gal_lnq_inv);
  if (i < 0) i = 0;
  if (i > n - 2) i = n - 2;
  // The grid is geometric, so the guess is the bisection result except within rounding of a node.  Load both rows
  // of the guessed interval at once (ten independent loads, one memory latency) and only fall back to the
  // dependent-load search when the bracket test fails.
  const double* r0 = M.gal + (size_t)i * 5;
  const double* r1 = r0 + 5;
  double v0[5], v1[5];
#pragma unroll
  for (int c = 0; c < 5; c++) {
    v0[c] = GC_LDG(r0 + c);
    v1[c] = GC_LDG(r1 + c);
  }
  if (!(v0[0] <= a && (a < v1[0] || i == n - 2))) {
    while (i > 0 && M.gal[(size_t)i * 5] > a) i--;
    while (i < n - 2 && M.gal[(size_t)(i + 1) * 5] <= a) i++;
    r0 = M.gal + (size_t)i * 5;
    r1 = r0 + 5;
#pragma unroll
    for (int c = 0; c < 5; c++) {
      v0[c] = r0[c];
      v1[c] = r1[c];
    }
  }
  const double dx = v1[0] - v0[0], delx = a - v0[0];
  const double idx = 1.0 / dx, third = 1.0 / 3.0;
  {
    const double dy = v1[1] - v0[1];
    const double b = dy * idx - dx * (v1[2] + 2.0 * v0[2]) * third;
    const double dd = (v1[2] - v0[2]) * idx * third;
    hbar = v0[1] + delx * (b + delx * (v0[2] + delx * dd));
  }
  {
    const double dy = v1[3] - v0[3];
    const double b = dy * idx - dx * (v1[4] + 2.0 * v0[4]) * third;
    const double dd = (v1[4] - v0[4]) * idx * third;
    xgal = a * (v0[3] + delx * (b + delx * (v0[4] + delx * dd)));
  }
}

// camb/equations_galileon.f90:107-156 (general a; used by the Romberg integral below)
__device__ inline double dtauda_full(const DevModel& M, double a) {
  const double a2 = a * a;
  double grhoa2;
  if (M.use_galileon && a >= 1e-10) {
    double hbar, xg;
    gal_HX(M, a, hbar, xg);
    const double a4 = a2 * a2;
    grhoa2 = a4 * M.grhom * (hbar * hbar);
  } else {
    grhoa2 = M.grhok * a2 + (M.grhoc + M.grhob) * a + M.grhog + M.grhornomass;
    grhoa2 = grhoa2 + M.grhov * (a2 * a2);
    if (M.Num_Nu_massive != 0) {
      for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
        double rhonu, pnu;
        Nu_background(M, a * M.nu_masses[nu_i], rhonu, pnu);
        grhoa2 = grhoa2 + rhonu * M.grhormass[nu_i];
      }
    }
  }
  return sqrt(3 / grhoa2);
}

// camb/subroutines.f90:117-176 specialised to f = dtauda
__device__ inline double rombint_dtauda(const DevModel& M, double a, double b, double tol) {
  const int MAXITER = 20, MAXJ = 5;
  double g[MAXJ + 2];
  double h = 0.5 * (b - a);
  double gmax = h * (dtauda_full(M, a) + dtauda_full(M, b));
  g[1] = gmax;
  int nint = 1;
  double error = 1.0e20;
  int i = 0;
  double g0 = 0, g1, fourj;
  for (;;) {
    i = i + 1;
    if (i > MAXITER || (i > 5 && fabs(error) < tol)) break;
    g0 = 0.0;
    for (int kk = 1; kk <= nint; kk++) g0 = g0 + dtauda_full(M, a + (kk + kk - 1) * h);
    g0 = 0.5 * g[1] + h * g0;
    h = 0.5 * h;
    nint = nint + nint;
    int jmax = i < MAXJ ? i : MAXJ;
    fourj = 1.0;
    for (int j = 1; j <= jmax; j++) {
      fourj = 4.0 * fourj;
      g1 = g0 + (g0 - g[j]) / (fourj - 1.0);
      g[j] = g0;
      g0 = g1;
    }
    if (fabs(g0) > tol)
      error = 1.0 - gmax / g0;
    else
      error = gmax;
    gmax = g0;
    g[jmax + 1] = g0;
  }
  return g0;
}

// ---------------------------------------------------------------- Galileon perturbation terms
// Closed forms of camb/galileon.cc (arXiv:1208.0600 eqs. 40-50 in CAMB variables).  hc = a*hbar,
// xc = a*dpi/da, pnu[] = massive-neutrino pressures already looked up by the caller.
struct GalC {
  double c2, c3, c4, c5, cG, h0, orad;
};

// camb/galileon.cc:746-787
GC_HD inline void gal_GetdHdX(const DevModel& M, double a, double hc, double xc, const double* pnu, double& dh,
                                   double& dx) {
  const double c2 = M.c2, c3 = M.c3, c4 = M.c4, c5 = M.c5, cG = M.cG, h0 = M.h0, orad = M.orad;
  const double a2 = a * a;
  const double xgal2 = xc * xc, xgal3 = xgal2 * xc, xgal4 = xgal2 * xgal2, xgal5 = xgal3 * xgal2;
  const double h = hc / a;
  const double h2 = h * h, h3 = h2 * h, h4 = h2 * h2, h5 = h3 * h2, h6 = h3 * h3, h7 = h4 * h3, h8 = h4 * h4;
  const double alpha = c2 / 6 * h * xc - 3 * c3 * h3 * xgal2 + 15 * c4 * h5 * xgal3 - 17.5 * c5 * h7 * xgal4 - 3 * cG * h3 * xc;
  const double gamma = c2 / 3 * h2 * xc - c3 * h4 * xgal2 + 2.5 * c5 * h8 * xgal4 - 2 * cG * h4 * xc;
  const double beta = c2 / 6 * h2 - 2 * c3 * h4 * xc + 9 * c4 * h6 * xgal2 - 10 * c5 * h8 * xgal3 - cG * h4;
  const double sigma = 2 * h + 2 * c3 * h3 * xgal3 - 15 * c4 * h5 * xgal4 + 21 * c5 * h7 * xgal5 + 6 * cG * h3 * xgal2;
  double lambda = 3 * h2 + orad / (a2 * a2) + c2 / 2 * h2 * xgal2 - 2 * c3 * h4 * xgal3 + 7.5 * c4 * h6 * xgal4 -
                  9 * c5 * h8 * xgal5 - cG * h4 * xgal2;
  for (int i = 0; i < M.n_eig; i++) lambda += M.grhormass[i] * pnu[i] / (a2 * a2 * h0 * h0);
  const double omega = 2 * c3 * h4 * xgal2 - 12 * c4 * h6 * xgal3 + 15 * c5 * h8 * xgal4 + 4 * cG * h4 * xc;
  dh = (omega * gamma - lambda * beta) / (sigma * beta - alpha * omega);
  dx = -xc + (alpha * lambda - sigma * gamma) / (sigma * beta - alpha * omega);
}

// camb/galileon.cc:809-829
GC_HD inline double gal_grhogal(const DevModel& M, double a, double hc, double xc) {
  if (!(a >= 9.99999e-7)) return 0;
  const double c2 = M.c2, c3 = M.c3, c4 = M.c4, c5 = M.c5, cG = M.cG, h0 = M.h0;
  const double xgal2 = xc * xc, xgal3 = xgal2 * xc, xgal4 = xgal2 * xgal2, xgal5 = xgal2 * xgal3;
  const double h = hc / a, h2 = h * h, h4 = h2 * h2;
  return h0 * h0 * a * a *
         (c2 / 2 * h2 * xgal2 - 6 * c3 * h4 * xgal3 + 22.5 * c4 * h4 * h2 * xgal4 - 21 * c5 * h4 * h4 * xgal5 -
          9 * cG * h4 * xgal2);
}

// camb/galileon.cc:831-850
GC_HD inline double gal_gpresgal(const DevModel& M, double a, double hc, double xc, double dh, double dx) {
  if (!(a >= 9.99999e-7)) return 0;
  const double c2 = M.c2, c3 = M.c3, c4 = M.c4, c5 = M.c5, cG = M.cG, h0 = M.h0;
  const double xgal2 = xc * xc, xgal4 = xgal2 * xgal2;
  const double h = hc / a, h2 = h * h, h3 = h2 * h, h4 = h2 * h2, h6 = h4 * h2;
  return h0 * h0 * a * a *
         (c2 / 2 * h2 * xgal2 + 2 * c3 * h3 * xgal2 * (dh * xc + dx * h) -
          c4 * (4.5 * h6 * xgal4 + 12 * h6 * xgal2 * xc * dx + 15 * h4 * h * xgal4 * dh) +
          3 * c5 * h6 * h * xgal4 * (5 * h * dx + 7 * dh * xc + 2 * h * xc) +
          cG * (6 * h3 * xgal2 * dh + 4 * h4 * xc * dx + 3 * h4 * xgal2));
}

// camb/galileon.cc:852-895
GC_HD inline double gal_Chigal(const DevModel& M, double a, double hc, double xc, double dgrho, double eta,
                                    double dphi, double dphip, double k) {
  if (!(a >= 9.99999e-7)) return 0;
  const double c2 = M.c2, c3 = M.c3, c4 = M.c4, c5 = M.c5, cG = M.cG, h0 = M.h0;
  const double k2 = k * k;
  const double a2 = a * a, a4 = a2 * a2, a6 = a4 * a2;
  const double xgal2 = xc * xc, xgal3 = xgal2 * xc, xgal4 = xgal2 * xgal2, xgal5 = xgal3 * xgal2;
  const double h2 = hc * hc, h3 = h2 * hc, h4 = h2 * h2, h6 = h4 * h2;
  const double alpha_Z = -2 * c3 / a2 * xgal3 * h2 + 15 * c4 / a4 * xgal4 * h4 - 21 * c5 / a6 * xgal5 * h6 - 6 * cG / a2 * xgal2 * h2;
  const double alpha_eta = 1.5 * c4 / a4 * xgal4 * h4 - 3 * c5 / a6 * xgal5 * h6 - cG / a2 * xgal2 * h2;
  const double beta = 1. / (1 - 0.5 * alpha_Z);
  const double ChitildeG = c2 * h0 * xc * hc * dphip - c3 / a2 * (18 * h0 * xgal2 * h3 * dphip + 2 * k2 * xgal2 * h2 * dphi) +
                           c4 / a4 * (90 * h0 * xgal3 * h4 * hc * dphip + 12 * k2 * xgal3 * h4 * dphi) -
                           c5 / a6 * (105 * h0 * xgal4 * h6 * hc * dphip + 15 * k2 * xgal4 * h6 * dphi) -
                           cG / a2 * (18 * h0 * xc * h3 * dphip + 4 * k2 * xc * h2 * dphi);
  return beta * (ChitildeG + 0.5 * alpha_Z * dgrho + (alpha_Z - 2 * alpha_eta) * k * eta);
}

// camb/galileon.cc:897-936
GC_HD inline double gal_qgal(const DevModel& M, double a, double hc, double xc, double dgq, double eta, double dphi,
                                  double dphip, double k) {
  if (!(a >= 9.99999e-7)) return 0;
  const double c2 = M.c2, c3 = M.c3, c4 = M.c4, c5 = M.c5, cG = M.cG, h0 = M.h0;
  const double a2 = a * a, a4 = a2 * a2, a6 = a4 * a2;
  const double xgal2 = xc * xc, xgal3 = xgal2 * xc, xgal4 = xgal2 * xgal2;
  const double h2 = hc * hc, h3 = h2 * hc, h4 = h2 * h2, h6 = h4 * h2;
  const double alpha = c4 / a4 * xgal4 * h4 - 2 * c5 / a6 * xgal4 * xc * h6 - cG / (1.5 * a2) * xgal2 * h2;
  const double beta = 1. / (1 - 1.5 * alpha);
  const double qtildeG = c2 * k * h0 * xc * hc * dphi - c3 / a2 * k * (6 * h0 * xgal2 * h3 * dphi - 2 * xgal2 * h2 * dphip) +
                         c4 / a4 * k * (-12 * xgal3 * h4 * dphip + 18 * h0 * xgal3 * h4 * hc * dphi) -
                         c5 / a6 * k * (-15 * xgal4 * h6 * dphip + 15 * h0 * xgal4 * h6 * hc * dphi) -
                         cG / a2 * k * (-4 * xc * h2 * dphip + 6 * h0 * xc * h3 * dphi);
  return beta * (qtildeG + 1.5 * alpha * dgq);
}

// camb/galileon.cc:938-985
GC_HD inline double gal_Pigal(const DevModel& M, double a, double hc, double xc, double dh, double dx, double dgrho,
                                   double dgq, double dgpi, double eta, double dphi, double k) {
  if (!(a >= 9.99999e-7)) return 0;
  const double c4 = M.c4, c5 = M.c5, cG = M.cG, h0 = M.h0;
  const double k2 = k * k;
  const double a2 = a * a, a4 = a2 * a2, a6 = a4 * a2;
  const double xgal2 = xc * xc, xgal3 = xgal2 * xc, xgal4 = xgal2 * xgal2, xgal5 = xgal3 * xgal2;
  const double h2 = hc * hc, h3 = h2 * hc, h4 = h2 * h2, h5 = h3 * h2, h6 = h3 * h3;
  const double hprime = a * dh + hc;
  const double xhprime = dx * hc + xc * hprime;
  const double alpha_sigprime = c4 / a4 * xgal4 * h4 - 3 * c5 / a6 * xgal4 * h5 * xhprime;
  const double alpha_sig = c4 / a4 * (3 * xgal4 * h4 - 6 * xgal3 * h3 * xhprime) -
                           c5 / a6 * (-3 * xgal5 * h5 * hprime + 12 * xgal5 * h6 - 15 * xgal4 * h5 * xhprime) +
                           2 * cG / a2 * xc * hc * xhprime;
  const double alpha_phi = -c4 / a4 * xgal4 * h4 - c5 / a6 * (6 * xgal4 * h5 * xhprime - 6 * xgal5 * h6) + 2 * cG / a2 * xgal2 * h2;
  const double beta_pi = 1. / (1 + 0.5 * alpha_phi - alpha_sigprime);
  const double PitildeG =
      k2 * (c4 / a4 * (4 * xgal3 * h4 * dphi - 6 * xgal2 * h3 * xhprime * dphi) -
            c5 / a6 * (-12 * xgal3 * h5 * xhprime * dphi + 12 * xgal4 * h6 * dphi - 3 * xgal4 * h5 * hprime * dphi) +
            2 * cG / a2 * hc * xhprime * dphi);
  return beta_pi * (PitildeG + (alpha_sigprime - 0.5 * alpha_phi) * dgpi +
                    0.5 * (2 * alpha_sigprime + alpha_sig - alpha_phi) * (dgrho + 3 * h0 * hc / k * dgq) +
                    (alpha_sig + alpha_sigprime) * k * eta);
}
// camb/galileon.cc:987-1082
GC_HD inline double gal_dphisecond(const DevModel& M, double a, double hc, double xc, double dh, double dx,
                                        double dgrho, double dgq, double eta, double dphi, double dphip, double k,
                                        double deltafprime) {
  if (!(a >= 9.99999e-7)) return 0;
  const double c2 = M.c2, c3 = M.c3, c4 = M.c4, c5 = M.c5, cG = M.cG, h0 = M.h0, orad = M.orad;
  const double h02 = h0 * h0;
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
  dphisecond = -(alpha_gammaprime*h0*hc*dphip + alpha_gamma*k2*dphi + (0.5*alpha_Z + 2*ksi*(h0*hc - 0.5*(h0*hprime + beta_Z*h0*hc) + 0.25*(beta_Z_prime + beta_Z*h0*hprime)))*dgrho + ksi*(1-beta_eta)*k*dgq + ksi*dotdeltaf + ksi*chiprimehat + (alpha_Z - 2*alpha_eta + 2*ksi*(2*beta_eta*h0*hc - h0*hprime - beta_Z*h0*hc - beta_eta_prime + 0.5*(beta_Z_prime + beta_Z*h0*hprime)))*k*eta)/(alpha_gammasecond + ksi*beta_gammasecond);
  return dphisecond;
}

// camb/galileon.cc:1084-1174
GC_HD inline double gal_pigalprime(const DevModel& M, double a, double hc, double xc, double dh, double dx,
                                        double dgrho, double dgq, double dgpi, double pidot, double eta, double dphi,
                                        double dphip, double k, double grho, double gpres, const double* pnu) {
  if (!(a >= 9.99999e-7)) return 0;
  const double c2 = M.c2, c3 = M.c3, c4 = M.c4, c5 = M.c5, cG = M.cG, h0 = M.h0, orad = M.orad;
  const double h02 = h0 * h0;
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
  for (int i = 0; i < M.n_eig; i++) {
    const double am = a * M.nu_masses[i];
    lambda += M.grhormass[i] * pnu[i] / (a2 * a2 * h0 * h0);
    const double dpnu = Nu_dp(M, am, hc, pnu[i]);
    lambda_prime += (dpnu - 4 * pnu[i] * hc) * M.grhormass[i] / (a2 * a2 * h0 * h0);
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
  piGdot = (piprimetilde + (alpha_sigprime - 0.5*alpha_phi)*pidot + (-3.5*h0*hc*alpha_sigprime + 0.5*h0*hprime*alpha_sigprime + 0.25*(alpha_phi - alpha_sigprime)*(grho+gpres)/(h0*hc) + 0.5*h0*hprime*alpha_sig + 0.5*alpha_sig_prime + alpha_sigprime_prime - 2*h0*hc*alpha_sig + 1.5*h0*hc*alpha_phi - 0.5*alpha_phi_prime)*dgrho + (alpha_sig_prime + alpha_sigprime_prime + alpha_sigprime*h0*hprime + alpha_sig*h0*hprime - 3*h0*hc*alpha_sigprime - 3*h0*hc*alpha_sig + 0.5*(alpha_phi - alpha_sigprime)*(grho+gpres)/(h0*hc))*k*eta + 0.5/k*(3*h02*hc*hprime*alpha_sigprime + 3*h02*hc*hprime*alpha_sig + 1.5*(alpha_phi - alpha_sigprime)*(grho+gpres) + 3*h0*hc*alpha_sig_prime - 12*h02*h2*alpha_sig - 21*h02*h2*alpha_sigprime - 3*h0*hc*alpha_phi_prime + 9*h02*h2*alpha_phi + 6*h0*hc*alpha_sigprime_prime + k2*(alpha_phi - alpha_sigprime))*dgq + (-2*h0*hc*alpha_sigprime + h0*hc*alpha_phi - 0.5*alpha_phi_prime + alpha_sigprime_prime - h0*hc*alpha_sig)*dgpi)/(1. - alpha_sigprime + 0.5*alpha_phi);
  return piGdot;
}


// ---------------------------------------------------------------- state-vector layout
// camb/equations_galileon.f90:298-312
__device__ inline int next_nu_nq(const DevModel& M, int nq) {
  if (nq == 0) return 1;
  const int qq = (int)M.nu_q[nq - 1];
  if (qq >= 10) return M.nqmax;
  return nq + 1;
}

// camb/equations_galileon.f90:555-644
__device__ inline void SetupScalarArrayIndices(const DevModel& M, EV& e, int* max_num_eqns) {
  int neq = 5, maxeq = 5;
  if (!e.no_phot) {
    e.g_ix = 6;
    if (e.TightCoupling)
      neq = neq + 2;
    else {
      neq = neq + (e.lmaxg + 1);
      e.polind = neq - 1;
      neq = neq + e.lmaxgpol - 1;
    }
  }
  if (!e.no_nu) {
    e.r_ix = neq + 1;
    if (e.high_ktau)
      neq = neq + 3;
    else
      neq = neq + (e.lmaxnr + 1);
  }
  maxeq = maxeq + (e.lmaxg + 1) + (e.lmaxnr + 1) + e.lmaxgpol - 1;
  if (M.use_galileon) {
    e.w_ix = neq + 1;
    neq = neq + 2;
    maxeq = maxeq + 2;
  } else {
    e.w_ix = 0;
  }
  if (M.Num_Nu_massive != 0) {
    const int nqmax = M.nqmax;
    e.has_nu_rel = false;
    for (int i = 0; i < M.n_eig; i++)
      if (e.nq[i] != nqmax) e.has_nu_rel = true;
    if (e.has_nu_rel) {
      e.lmaxnu_pert = e.lmaxnu;
      e.nu_pert_ix = neq + 1;
      neq = neq + e.lmaxnu_pert + 1;
      maxeq = maxeq + e.lmaxnu_pert + 1;
    } else {
      e.lmaxnu_pert = 0;
    }
    const double lAB = M.lAccuracyBoost;
    for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
      if (e.high_ktau) {
        e.lmaxnu_tau[nu_i] = (int)(4 * (float)lAB);  // lmaxnu_high_ktau * lAccuracyBoost (default real)
      } else {
        long v = nintd(0.8 * e.q * M.nu_tau_nonrelativistic[nu_i] * lAB);
        if (v > e.lmaxnu) v = e.lmaxnu;
        if (v < 3) v = 3;
        e.lmaxnu_tau[nu_i] = (int)v;
        if (e.nu_nonrel[nu_i]) e.lmaxnu_tau[nu_i] = min(e.lmaxnu_tau[nu_i], (int)nintd(4 * lAB));
      }
      e.lmaxnu_tau[nu_i] = min(e.lmaxnu, e.lmaxnu_tau[nu_i]);
      e.nu_ix[nu_i] = neq + 1;
      if (e.MassiveNuApprox[nu_i])
        neq = neq + 4;
      else
        neq = neq + e.nq[nu_i] * (e.lmaxnu_tau[nu_i] + 1);
      maxeq = maxeq + nqmax * (e.lmaxnu + 1);
    }
  } else {
    e.has_nu_rel = false;
  }
  e.neq = neq;
  if (max_num_eqns) *max_num_eqns = maxeq;
}

// camb/equations_galileon.f90:646-724 ; y, yout are workspace columns
__device__ inline void CopyScalarVariableArray(const DevModel& M, const double* y, double* yout, const EV& e,
                                               const EV& o) {
  for (int i = 1; i <= o.nvar; i++) AT(yout, i) = 0;
  for (int i = 1; i <= 5; i++) AT(yout, i) = AT(y, i);
  if (M.use_galileon) {
    AT(yout, o.w_ix) = AT(y, e.w_ix);
    AT(yout, o.w_ix + 1) = AT(y, e.w_ix + 1);
  }
  int lmax;
  if (!e.no_phot && !o.no_phot) {
    if (e.TightCoupling || o.TightCoupling)
      lmax = 1;
    else
      lmax = min(e.lmaxg, o.lmaxg);
    for (int l = 0; l <= lmax; l++) AT(yout, o.g_ix + l) = AT(y, e.g_ix + l);
    if (!e.TightCoupling && !o.TightCoupling) {
      lmax = min(e.lmaxgpol, o.lmaxgpol);
      for (int l = 2; l <= lmax; l++) AT(yout, o.polind + l) = AT(y, e.polind + l);
    }
  }
  if (!e.no_nu && !o.no_nu) {
    if (e.high_ktau || o.high_ktau)
      lmax = 2;
    else
      lmax = min(e.lmaxnr, o.lmaxnr);
    for (int l = 0; l <= lmax; l++) AT(yout, o.r_ix + l) = AT(y, e.r_ix + l);
  }
  if (M.Num_Nu_massive != 0) {
    for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
      const int ix_off = e.nu_ix[nu_i], ix_off2 = o.nu_ix[nu_i];
      if (e.MassiveNuApprox[nu_i] && o.MassiveNuApprox[nu_i]) {
        for (int l = 0; l < 4; l++) AT(yout, ix_off2 + l) = AT(y, ix_off + l);
      } else if (!e.MassiveNuApprox[nu_i] && !o.MassiveNuApprox[nu_i]) {
        lmax = min(e.lmaxnu_tau[nu_i], o.lmaxnu_tau[nu_i]);
        const int nq = min(e.nq[nu_i], o.nq[nu_i]);
        for (int i = 1; i <= nq; i++) {
          const int ind = ix_off + (i - 1) * (e.lmaxnu_tau[nu_i] + 1);
          const int ind2 = ix_off2 + (i - 1) * (o.lmaxnu_tau[nu_i] + 1);
          for (int l = 0; l <= lmax; l++) AT(yout, ind2 + l) = AT(y, ind + l);
        }
        for (int i = nq + 1; i <= o.nq[nu_i]; i++) {
          lmax = min(o.lmaxnu_tau[nu_i], e.lmaxnr);
          const int ind2 = ix_off2 + (i - 1) * (o.lmaxnu_tau[nu_i] + 1);
          for (int l = 0; l <= lmax; l++) AT(yout, ind2 + l) = AT(y, e.r_ix + l);
          const double r = M.nu_masses[nu_i] / M.nu_q[i - 1];
          const double pert_scale = (r * r) / 2;
          lmax = min(lmax, e.lmaxnu_pert);
          for (int l = 0; l <= lmax; l++) AT(yout, ind2 + l) = AT(yout, ind2 + l) + AT(y, e.nu_pert_ix + l) * pert_scale;
        }
      }
    }
    if (o.has_nu_rel && e.has_nu_rel) {
      lmax = min(o.lmaxnu_pert, e.lmaxnu_pert);
      for (int l = 0; l <= lmax; l++) AT(yout, o.nu_pert_ix + l) = AT(y, e.nu_pert_ix + l);
    }
  }
}

// camb/equations_galileon.f90:793-946 (scalar, flat)
__device__ inline void GetNumEqns(const DevModel& M, EV& e) {
  const double lAB = M.lAccuracyBoost;
  double max_nu_mass = 0;
  if (M.Num_Nu_massive == 0) {
    e.lmaxnu = 0;
  } else {
    const int nqmax = M.nqmax;
    for (int i = 0; i < M.n_eig; i++) max_nu_mass = fmax(max_nu_mass, M.nu_masses[i]);
    for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
      int q_rel = 0;
      for (int j = 0; j < nqmax; j++) {
        if (M.nu_q[j] > M.nu_masses[nu_i] * M.adotrad / e.q) break;
        q_rel = q_rel + 1;
      }
      e.nq[nu_i] = (q_rel >= nqmax - 2) ? nqmax : q_rel;
      e.nu_nonrel[nu_i] = false;
    }
    e.lmaxnu = max(3, (int)nintd(10 * lAB));
    if (max_nu_mass > 700) e.lmaxnu = max(3, (int)nintd(15 * lAB));
  }
  e.high_ktau = false;
  e.TightCoupling = true;
  e.no_phot = false;
  e.no_nu = false;
  for (int i = 0; i < GC_MAX_NU; i++) e.MassiveNuApprox[i] = false;
  if (M.HighAccuracyDefault && M.AccuratePolarization)
    e.lmaxg = max((int)nintd(11 * lAB), 3);
  else
    e.lmaxg = max((int)nintd(8 * lAB), 3);
  e.lmaxnr = max((int)nintd(14 * lAB), 3);
  if (max_nu_mass > 700 && M.HighAccuracyDefault) e.lmaxnr = max((int)nintd(32 * lAB), 3);
  e.lmaxgpol = e.lmaxg;
  if (!M.AccuratePolarization) e.lmaxgpol = max((int)nintd(4 * lAB), 3);
  if (e.q < GC_SP(0.05)) {
    double scal = 1;
    if (M.AccuratePolarization) scal = 4;
    e.lmaxgpol = max(3, (int)nintd(min(8L, nintd(scal * 150 * e.q)) * lAB));
    e.lmaxnr = max(3, (int)nintd(min(7L, nintd(sqrt(scal) * 150 * e.q)) * lAB));
    e.lmaxg = max(3, (int)nintd(min(8L, nintd(sqrt(scal) * 300 * e.q)) * lAB));
    if (M.AccurateReionization) {
      e.lmaxg = e.lmaxg * 4;
      e.lmaxgpol = e.lmaxgpol * 2;
    }
  }
  e.poltruncfac = (double)e.lmaxgpol / max(1, (e.lmaxgpol - 2));
  SetupScalarArrayIndices(M, e, &e.nvar);
}

// ---------------------------------------------------------------- massive neutrino moments
// camb/equations_galileon.f90:1058-1107
__device__ inline void Nu_Integrate_L012(const DevModel& M, const EV& e, const double* y, double a, int nu_i,
                                         double& drhonu, double& fnu, double* dpnu, double* pinu) {
  drhonu = 0;
  fnu = 0;
  if (dpnu) {
    *dpnu = 0;
    *pinu = 0;
  }
  const double am = a * M.nu_masses[nu_i];
  int ind = e.nu_ix[nu_i];
  for (int iq = 1; iq <= e.nq[nu_i]; iq++) {
    const double aq = am * M.inv_nu_q[iq - 1];
    const double e2 = 1.0 + aq * aq;
    const double v = rsqrt(e2), iv = e2 * v;  // v = q/E, 1/v = E/q
    const double ker = M.nu_int_kernel[iq - 1];
    drhonu = drhonu + ker * AT(y, ind) * iv;
    fnu = fnu + ker * AT(y, ind + 1);
    if (dpnu) {
      *dpnu = *dpnu + ker * AT(y, ind) * v;
      *pinu = *pinu + ker * AT(y, ind + 2) * v;
    }
    ind = ind + e.lmaxnu_tau[nu_i] + 1;
  }
  ind = e.nu_pert_ix;
  for (int iq = e.nq[nu_i] + 1; iq <= M.nqmax; iq++) {
    const double aq = am * M.inv_nu_q[iq - 1];
    const double e2 = 1.0 + aq * aq;
    const double v = rsqrt(e2), iv = e2 * v;
    const double r = M.nu_masses[nu_i] * M.inv_nu_q[iq - 1];
    const double pert_scale = (r * r) / 2;
    const double ker = M.nu_int_kernel[iq - 1];
    const double tmp = ker * (AT(y, e.r_ix) + pert_scale * AT(y, ind));
    drhonu = drhonu + tmp * iv;
    fnu = fnu + ker * (AT(y, e.r_ix + 1) + pert_scale * AT(y, ind + 1));
    if (dpnu) {
      *dpnu = *dpnu + tmp * v;
      *pinu = *pinu + ker * (AT(y, e.r_ix + 2) + pert_scale * AT(y, ind + 2)) * v;
    }
  }
  if (dpnu) *dpnu = *dpnu / 3;
}

// camb/equations_galileon.f90:1109-1151
__device__ inline double Nu_pinudot(const DevModel& M, const EV& e, const double* y, const double* ydot, double a,
                                    double adotoa, int nu_i) {
  double pinudot = 0.0;
  int ind = e.nu_ix[nu_i] + 2;
  const double am = a * M.nu_masses[nu_i];
  for (int iq = 1; iq <= e.nq[nu_i]; iq++) {
    const double aq = am / M.nu_q[iq - 1];
    const double aqdot = aq * adotoa;
    const double v = 1.0 / sqrt(1.0 + aq * aq);
    const double vdot = -aq * aqdot / pow(1.0 + aq * aq, 1.5);
    pinudot = pinudot + M.nu_int_kernel[iq - 1] * (AT(ydot, ind) * v + AT(y, ind) * vdot);
    ind = ind + e.lmaxnu_tau[nu_i] + 1;
  }
  ind = e.nu_pert_ix + 2;
  for (int iq = e.nq[nu_i] + 1; iq <= M.nqmax; iq++) {
    const double qq = M.nu_q[iq - 1];
    const double aq = am / qq;
    const double aqdot = aq * adotoa;
    const double r = M.nu_masses[nu_i] / qq;
    const double pert_scale = (r * r) / 2;
    const double v = 1.0 / sqrt(1.0 + aq * aq);
    const double vdot = -aq * aqdot / pow(1.0 + aq * aq, 1.5);
    const double psi2dot = AT(ydot, e.r_ix + 2) + pert_scale * AT(ydot, ind);
    const double psi2 = AT(y, e.r_ix + 2) + pert_scale * AT(y, ind);
    pinudot = pinudot + M.nu_int_kernel[iq - 1] * (psi2dot * v + psi2 * vdot);
  }
  return pinudot;
}

// camb/equations_galileon.f90:1179-1207
__device__ inline void Nu_Intvsq(const DevModel& M, const EV& e, const double* y, double a, int nu_i, double& G11,
                                 double& G30) {
  const double am = a * M.nu_masses[nu_i];
  int ind = e.nu_ix[nu_i];
  G11 = 0.0;
  G30 = 0.0;
  for (int iq = 1; iq <= e.nq[nu_i]; iq++) {
    const double aq = am / M.nu_q[iq - 1];
    const double v = 1.0 / sqrt(1.0 + aq * aq);
    G11 = G11 + M.nu_int_kernel[iq - 1] * AT(y, ind + 1) * (v * v);
    if (e.lmaxnu_tau[nu_i] > 2) G30 = G30 + M.nu_int_kernel[iq - 1] * AT(y, ind + 3) * (v * v);
    ind = ind + e.lmaxnu_tau[nu_i] + 1;
  }
}

// ---------------------------------------------------------------- derivs
// camb/equations_galileon.f90:2098-2589.  ay / ayp are workspace columns.  Writes e.pig / e.pigdot during
// tight coupling exactly as the reference does (its value at a switch is the one left by the LAST call).
__device__ __forceinline__ void derivs(const DevModel* __restrict__ Mp, EV* __restrict__ ep, double tau,
                                    const double* __restrict__ ay, double* __restrict__ ayp) {
  // __restrict__ on every pointer: the state-vector stores below must not force reloads of the model constants
  // (global) or of the lane's layout descriptor (local) after each store.
  const DevModel& M = *Mp;
  EV& e = *ep;
  const bool gal = M.use_galileon != 0;
  const double k = e.k, ik = e.ik, ik2 = e.ik2;
  const double a = AT(ay, 1), a2 = a * a;
  const double ia = 1.0 / a, ia2 = ia * ia;
  const double etak = AT(ay, 2), clxc = AT(ay, 3), clxb = AT(ay, 4), vb = AT(ay, 5);
  double dphi = 0, dphiprime = 0;
  if (gal) {
    dphi = AT(ay, e.w_ix);
    dphiprime = AT(ay, e.w_ix + 1);
  }
  // issue the loads of the low multipoles now: they are independent and their (L2) latencies then overlap with the
  // table lookups below instead of being paid one by one where each value is first used
  double y_r0 = 0, y_r1 = 0, y_r2 = 0, y_g0 = 0, y_g1 = 0, y_g2 = 0, y_g3 = 0, y_e2 = 0, y_e3 = 0, y_r3 = 0;
  if (!e.no_nu) {
    y_r0 = AT(ay, e.r_ix);
    y_r1 = AT(ay, e.r_ix + 1);
    y_r2 = AT(ay, e.r_ix + 2);
    if (!e.high_ktau && e.lmaxnr > 2) y_r3 = AT(ay, e.r_ix + 3);
  }
  if (!e.no_phot) {
    y_g0 = AT(ay, e.g_ix);
    y_g1 = AT(ay, e.g_ix + 1);
    if (!e.TightCoupling) {
      y_g2 = AT(ay, e.g_ix + 2);
      if (e.lmaxg > 2) y_g3 = AT(ay, e.g_ix + 3);
      y_e2 = AT(ay, e.polind + 2);
      if (e.lmaxgpol > 2) y_e3 = AT(ay, e.polind + 3);
    }
  }
  const double grhob_t = M.grhob * ia, grhoc_t = M.grhoc * ia, grhor_t = M.grhornomass * ia2, grhog_t = M.grhog * ia2;
  // massive-neutrino background at this a (shared by MassiveNuVars, GetdHdX and dtauda)
  double rhonu[GC_MAX_NU], pnu[GC_MAX_NU];
  for (int i = 0; i < M.n_eig; i++) Nu_background(M, a * M.nu_masses[i], rhonu[i], pnu[i]);
  double grhov_t = 0, xgal = 0, grhogal_t = 0, hbar = 0;
  GalFast F;
  const GalConst GC{M.c2, M.c3, M.c4, M.c5, M.cG, M.h0, M.orad};
  if (gal) {
    gal_HX(M, a, hbar, xgal);
    double lam_nu = 0;
    for (int i = 0; i < M.n_eig; i++) lam_nu += M.grhormass[i] * pnu[i];
    gal_fast_background(GC, a, hbar, xgal, lam_nu, k, F);
    grhogal_t = F.grhogal;
  } else {
    grhov_t = M.grhov * a2;
  }
  double cs2, opacity, dopacity = 0;
  thermo(M, tau, cs2, opacity, e.TightCoupling ? &dopacity : nullptr);
  double gpres = 0;
  gpres = gpres + (grhog_t + grhor_t) / 3;
  if (gal)
    gpres = gpres + F.gpresgal;
  else
    gpres = gpres + grhov_t * (-1.0);
  double grho_matter = grhob_t + grhoc_t;
  double dgrho_matter = grhob_t * clxb + grhoc_t * clxc;
  double dgq = grhob_t * vb;
  double wnu_arr[GC_MAX_NU];
  if (M.Num_Nu_massive > 0) {  // MassiveNuVars, :1209-1252
    for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
      const double grhormass_t = M.grhormass[nu_i] * ia2;
      const double irho = 1.0 / rhonu[nu_i];
      double clxnu, qnu;
      if (e.MassiveNuApprox[nu_i]) {
        clxnu = AT(ay, e.nu_ix[nu_i]);
        qnu = AT(ay, e.nu_ix[nu_i] + 2);
      } else {
        Nu_Integrate_L012(M, e, ay, a, nu_i, clxnu, qnu, nullptr, nullptr);
        qnu = qnu * irho;
        clxnu = clxnu * irho;
      }
      const double grhonu_t = grhormass_t * rhonu[nu_i];
      const double gpnu_t = grhormass_t * pnu[nu_i];
      grho_matter = grho_matter + grhonu_t;
      gpres = gpres + gpnu_t;
      dgrho_matter = dgrho_matter + grhonu_t * clxnu;
      dgq = dgq + grhonu_t * qnu;
      wnu_arr[nu_i] = pnu[nu_i] * irho;
    }
  }
  // dtauda(a), :107-156
  double grhoa2;
  if (gal && a >= 1e-10) {
    const double a4 = a2 * a2;
    grhoa2 = a4 * M.grhom * (hbar * hbar);
  } else {
    grhoa2 = M.grhok * a2 + (M.grhoc + M.grhob) * a + M.grhog + M.grhornomass;
    grhoa2 = grhoa2 + M.grhov * (a2 * a2);
    if (M.Num_Nu_massive != 0)
      for (int nu_i = 0; nu_i < M.n_eig; nu_i++) grhoa2 = grhoa2 + rhonu[nu_i] * M.grhormass[nu_i];
  }
  const double adotoa = sqrt(grhoa2 * (1.0 / 3.0)) * ia;  // 1/(a dtauda(a))
  const double iadotoa = 1.0 / adotoa;
  const double cothxor = 1.0 / tau;
  double dgrho = dgrho_matter;
  double z = 0, dz, clxr, qr, pir, clxg, qg, pig = 0, dgrhogal_t;
  if (e.no_nu) {
    if (gal) {
      dgrhogal_t = gal_fast_Chigal(F, dgrho, etak, dphi, dphiprime);
      z = (0.5 * (dgrho + dgrhogal_t) * ik + etak) * iadotoa;
      dz = -adotoa * z - 0.5 * (dgrho + dgrhogal_t) * ik;
    } else {
      z = (0.5 * dgrho * ik + etak) * iadotoa;
      dz = -adotoa * z - 0.5 * dgrho * ik;
    }
    clxr = -4 * dz * ik;
    qr = -4.0 / 3 * z;
    pir = 0;
  } else {
    clxr = y_r0;
    qr = y_r1;
    pir = y_r2;
  }
  if (e.no_phot) {
    if (!e.no_nu) {
      if (gal) {
        dgrhogal_t = gal_fast_Chigal(F, dgrho, etak, dphi, dphiprime);
        z = (0.5 * (dgrho + dgrhogal_t) * ik + etak) * iadotoa;
        dz = -adotoa * z - 0.5 * (dgrho + dgrhogal_t) * ik;
      } else {
        z = (0.5 * dgrho * ik + etak) * iadotoa;
        dz = -adotoa * z - 0.5 * dgrho * ik;
      }
      clxg = -4 * dz * ik - 4 * ik * opacity * (vb + z);
      qg = -4.0 / 3 * z;
    } else {
      clxg = clxr - 4 * ik * opacity * (vb + z);
      qg = qr;
    }
    pig = 0;
  } else {
    clxg = y_g0;
    qg = y_g1;
    if (!e.TightCoupling) pig = y_g2;
  }
  dgrho = dgrho + grhog_t * clxg + grhor_t * clxr;
  dgq = dgq + grhog_t * qg + grhor_t * qr;
  const double photbar = grhog_t / grhob_t;
  const double pb43 = 4.0 / 3 * photbar;
  AT(ayp, 1) = adotoa * a;
  if (gal) {
    dgrhogal_t = gal_fast_Chigal(F, dgrho, etak, dphi, dphiprime);
    dgrho = dgrho + dgrhogal_t;
    dgq = dgq + gal_fast_qgal(F, dgq, dphi, dphiprime);
  }
  z = (0.5 * dgrho * ik + etak) * iadotoa;
  const double sigma = (z + 1.5 * dgq * ik2);
  AT(ayp, 2) = 0.5 * dgq;
  const double clxcdot = -k * z;
  AT(ayp, 3) = clxcdot;
  const double clxbdot = -k * (z + vb);
  AT(ayp, 4) = clxbdot;
  const double clxgdot = -k * (4.0 / 3.0 * z + qg);
  double vbdot, qgdot, pigdot;
  if (e.TightCoupling) {
    const double adotdota = (adotoa * adotoa - gpres) / 2;
    const double iop = 1.0 / opacity, iop2 = iop * iop, ipb = 1.0 / (1 + pb43);
    const double c3245 = 32.0 / 45.0;
    pig = c3245 * iop * k * (sigma + vb);
    const double slip = -(2 * adotoa * ipb + dopacity * iop) * (vb - 3.0 / 4 * qg) +
                        (-adotdota * vb - k / 2 * adotoa * clxg + k * (cs2 * clxbdot - clxgdot / 4)) * (iop * ipb);
    double dgs = grhog_t * pig + grhor_t * pir;
    if (gal) dgs = dgs + gal_fast_Pigal(GC, F, dgrho, dgq, dgs, etak, dphi);
    const double sigmadot = -2 * adotoa * sigma - dgs * ik + etak;
    qgdot = k * (clxg / 4.0 - pig / 2.0) + opacity * slip;
    const double dop = dopacity * iop2;
    pig = c3245 * iop * k * (sigma + 3.0 * qg / 4.0) * (1 + (dop * (11.0 / 6.0))) +
          (c3245 * iop2) * k * (sigmadot + 3.0 * qgdot / 4.0) * (-11.0 / 6.0);
    pigdot = -c3245 * dop * k * (sigma + 3.0 * qg / 4.0) * (1 + dop * (11.0 / 6.0)) +
             (c3245 * iop) * k * (sigmadot + 3.0 * qgdot / 4.0) * (1 + (11.0 / 6.0) * dop);
    e.pigdot = pigdot;
    vbdot = (-adotoa * vb + cs2 * k * clxb + k / 4 * pb43 * (clxg - 2 * 1.0 * pig)) * ipb;
    vbdot = vbdot + pb43 * ipb * slip;
    e.pig = pig;
  } else {
    vbdot = -adotoa * vb + cs2 * k * clxb - photbar * opacity * (4.0 / 3 * vb - qg);
  }
  AT(ayp, 5) = vbdot;
  if (!e.no_phot) {
    AT(ayp, e.g_ix) = clxgdot;
    qgdot = 4.0 / 3 * (-vbdot - adotoa * vb + cs2 * k * clxb) / pb43 + denlk(e, 1) * clxg - denlk2(e, 1) * pig;
    AT(ayp, e.g_ix + 1) = qgdot;
    if (!e.TightCoupling) {
      const double E2 = y_e2;
      const double polter = pig / 10 + 9.0 / 15 * E2;
      int ix = e.g_ix + 2;
      if (e.lmaxg > 2) {
        pigdot = denlk(e, 2) * qg - denlk2(e, 2) * y_g3 - opacity * (pig - polter) + 8.0 / 15.0 * k * sigma;
        AT(ayp, ix) = pigdot;
#pragma unroll 4
        for (int l = 3; l <= e.lmaxg - 1; l++) {
          ix = ix + 1;
          AT(ayp, ix) = (denlk(e, l) * AT(ay, ix - 1) - denlk2(e, l) * AT(ay, ix + 1)) - opacity * AT(ay, ix);
        }
        ix = ix + 1;
        AT(ayp, ix) = k * AT(ay, ix - 1) - (e.lmaxg + 1) * cothxor * AT(ay, ix) - opacity * AT(ay, ix);
      } else {
        pigdot = denlk(e, 2) * qg - opacity * (pig - polter) + 8.0 / 15.0 * k * sigma;
        AT(ayp, ix) = pigdot;
      }
      ix = e.polind + 2;
      if (e.lmaxgpol > 2) {
        AT(ayp, ix) = -opacity * (y_e2 - polter) - k / 3.0 * y_e3;
#pragma unroll 4
        for (int l = 3; l <= e.lmaxgpol - 1; l++) {
          ix = ix + 1;
          AT(ayp, ix) = -opacity * AT(ay, ix) + (denlk(e, l) * AT(ay, ix - 1) - polfack(e, l) * AT(ay, ix + 1));
        }
        ix = ix + 1;
        AT(ayp, ix) = -opacity * AT(ay, ix) + k * e.poltruncfac * AT(ay, ix - 1) - (e.lmaxgpol + 3) * cothxor * AT(ay, ix);
      } else {
        AT(ayp, ix) = -opacity * (AT(ay, ix) - polter);
      }
    }
  }
  // clxrdot is read uninitialised by the reference when no_nu_multpoles (:2517 vs :2482); pinned to 0 like
  // the oracle (the value of its defining expression under qr = -4z/3).
  double clxrdot = 0;
  if (!e.no_nu) {
    clxrdot = -k * (4.0 / 3.0 * z + qr);
    AT(ayp, e.r_ix) = clxrdot;
    AT(ayp, e.r_ix + 1) = denlk(e, 1) * clxr - denlk2(e, 1) * pir;
    if (e.high_ktau) {
      AT(ayp, e.r_ix + 2) = -3 * pir * cothxor - clxrdot;
    } else {
      int ix = e.r_ix + 2;
      if (e.lmaxnr > 2) {
        AT(ayp, ix) = denlk(e, 2) * qr - denlk2(e, 2) * y_r3 + 8.0 / 15.0 * k * sigma;
#pragma unroll 4
        for (int l = 3; l <= e.lmaxnr - 1; l++) {
          ix = ix + 1;
          AT(ayp, ix) = (denlk(e, l) * AT(ay, ix - 1) - denlk2(e, l) * AT(ay, ix + 1));
        }
        ix = ix + 1;
        AT(ayp, ix) = k * AT(ay, ix - 1) - (e.lmaxnr + 1) * cothxor * AT(ay, ix);
      } else {
        AT(ayp, ix) = denlk(e, 2) * qr + 8.0 / 15.0 * k * sigma;
      }
    }
  }
  if (gal) {
    const double dotdeltaf = grhob_t * (clxbdot - 3 * adotoa * clxb) + grhoc_t * (clxcdot - 3 * adotoa * clxc) +
                             grhor_t * (clxrdot - 4 * adotoa * clxr) + grhog_t * (clxgdot - 4 * adotoa * clxg);
    AT(ayp, e.w_ix) = dphiprime;
    AT(ayp, e.w_ix + 1) = gal_fast_dphisecond(GC, F, dgrho, dgq, etak, dphi, dphiprime, dotdeltaf);
  }
  if (M.Num_Nu_massive == 0) return;
  for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
    if (e.MassiveNuApprox[nu_i]) {
      const double ia3 = ia * ia2;
      const double G11_t = e.G11[nu_i] * ia3;
      const double G30_t = e.G30[nu_i] * ia3;
      const int o = e.nu_ix[nu_i];
      const double w = wnu_arr[nu_i];
      AT(ayp, o) = -k * z * (w + 1) + 3 * adotoa * (w * AT(ay, o) - AT(ay, o + 1)) - k * AT(ay, o + 2);
      AT(ayp, o + 1) = (3 * w - 2) * adotoa * AT(ay, o + 1) - 5.0 / 3 * k * z * w - k / 3 * G11_t;
      AT(ayp, o + 2) = (3 * w - 1) * adotoa * AT(ay, o + 2) - k * (2.0 / 3 * 1.0 * AT(ay, o + 3) - AT(ay, o + 1));
      AT(ayp, o + 3) = (3 * w - 2) * adotoa * AT(ay, o + 3) + 2 * w * k * sigma - k / 5 * (3 * 1.0 * G30_t - 2 * G11_t);
    } else {
      int ind = e.nu_ix[nu_i];
      for (int i = 1; i <= e.nq[nu_i]; i++) {
        const double aq = a * M.nu_masses[nu_i] * M.inv_nu_q[i - 1];
        const double v = rsqrt(1.0 + aq * aq);
        AT(ayp, ind) = -k * (4.0 / 3.0 * z + v * AT(ay, ind + 1));
        ind = ind + 1;
        AT(ayp, ind) = v * (denlk(e, 1) * AT(ay, ind - 1) - denlk2(e, 1) * AT(ay, ind + 1));
        ind = ind + 1;
        if (e.lmaxnu_tau[nu_i] == 2) {
          AT(ayp, ind) = -AT(ayp, ind - 2) - 3 * cothxor * AT(ay, ind);
        } else {
          AT(ayp, ind) = v * (denlk(e, 2) * AT(ay, ind - 1) - denlk2(e, 2) * AT(ay, ind + 1)) + k * 8.0 / 15.0 * sigma;
#pragma unroll 4
          for (int l = 3; l <= e.lmaxnu_tau[nu_i] - 1; l++) {
            ind = ind + 1;
            AT(ayp, ind) = v * (denlk(e, l) * AT(ay, ind - 1) - denlk2(e, l) * AT(ay, ind + 1));
          }
          ind = ind + 1;
          AT(ayp, ind) = k * v * AT(ay, ind - 1) - (e.lmaxnu_tau[nu_i] + 1) * cothxor * AT(ay, ind);
        }
        ind = ind + 1;
      }
    }
  }
  if (e.has_nu_rel) {
    int ind = e.nu_pert_ix;
    AT(ayp, ind) = +k * a2 * qr - k * AT(ay, ind + 1);
    int ind2 = e.r_ix;
#pragma unroll 4
    for (int l = 1; l <= e.lmaxnu_pert - 1; l++) {
      ind = ind + 1;
      ind2 = ind2 + 1;
      AT(ayp, ind) = -a2 * (denlk(e, l) * AT(ay, ind2 - 1) - denlk2(e, l) * AT(ay, ind2 + 1)) +
                     (denlk(e, l) * AT(ay, ind - 1) - denlk2(e, l) * AT(ay, ind + 1));
    }
    ind = ind + 1;
    ind2 = ind2 + 1;
    AT(ayp, ind) = k * (AT(ay, ind - 1) - a2 * AT(ay, ind2 - 1)) - (e.lmaxnu_pert + 1) * cothxor * AT(ay, ind);
  }
}

// ---------------------------------------------------------------- output
// camb/equations_galileon.f90:1256-1576, everything after the derivs call at :1297.
// ts = this output time's row {tau, dtau, vis, dvis, ddvis, expmmu, opac, dopac}.
static __device__ __noinline__ void output_sources(const DevModel* __restrict__ Mp, const EV* __restrict__ ep,
                                            const double* __restrict__ y, const double* __restrict__ yprime,
                                            const double* __restrict__ ts, double tau, double* __restrict__ sources) {
  const DevModel& M = *Mp;
  const EV& e = *ep;
  const bool gal = M.use_galileon != 0;
  const double vis = ts[2], dvis = ts[3], ddvis = ts[4], expmmu = ts[5], opac = ts[6], dopac = ts[7];
  double pol2 = 0, pol3 = 0, polprime2 = 0, polprime3 = 0;
  const bool localpol = e.TightCoupling || e.no_phot;
  if (!localpol) {
    pol2 = AT(y, e.polind + 2);
    pol3 = AT(y, e.polind + 3);
    polprime2 = AT(yprime, e.polind + 2);
    polprime3 = AT(yprime, e.polind + 3);
  }
  const double k = e.k, k2 = e.k2;
  const double a = AT(y, 1), a2 = a * a, etak = AT(y, 2), clxc = AT(y, 3), clxb = AT(y, 4), vb = AT(y, 5),
               vbdot = AT(yprime, 5);
  double dphi = 0, dphiprime = 0;
  if (gal) {
    dphi = AT(y, e.w_ix);
    dphiprime = AT(y, e.w_ix + 1);
  }
  const double grhob_t = M.grhob / a, grhoc_t = M.grhoc / a, grhor_t = M.grhornomass / a2, grhog_t = M.grhog / a2;
  const double grhov_t = M.grhov * pow(a, 2.0);
  double grho = grhob_t + grhoc_t + grhor_t + grhog_t;
  double gpres = (grhog_t + grhor_t) / 3;
  double rhonu[GC_MAX_NU], pnu[GC_MAX_NU];
  for (int i = 0; i < M.n_eig; i++) Nu_background(M, a * M.nu_masses[i], rhonu[i], pnu[i]);
  double xgal = 0, hub = 0, dh = 0, dx = 0, hbar = 0;
  if (gal) {
    gal_HX(M, a, hbar, xgal);
    hub = a * hbar;
    gal_GetdHdX(M, a, hub, xgal, pnu, dh, dx);
    grho = grho + gal_grhogal(M, a, hub, xgal);
    gpres = gpres + gal_gpresgal(M, a, hub, xgal, dh, dx);
  } else {
    grho = grho + grhov_t;
    gpres = gpres + grhov_t * (-1.0);
  }
  // dtauda(a)
  double grhoa2;
  if (gal && a >= 1e-10) {
    grhoa2 = (a2 * a2) * M.grhom * (hbar * hbar);
  } else {
    grhoa2 = M.grhok * a2 + (M.grhoc + M.grhob) * a + M.grhog + M.grhornomass;
    grhoa2 = grhoa2 + M.grhov * (a2 * a2);
    if (M.Num_Nu_massive != 0)
      for (int nu_i = 0; nu_i < M.n_eig; nu_i++) grhoa2 = grhoa2 + rhonu[nu_i] * M.grhormass[nu_i];
  }
  const double dtauda_a = sqrt(3 / grhoa2);
  double dgrho = grhob_t * clxb + grhoc_t * clxc;
  double dgq = grhob_t * vb;
  double dgpi = 0, dgpi_diff = 0, pidot_sum = 0;
  if (M.Num_Nu_massive != 0) {  // MassiveNuVarsOut, :992-1056
    double grhonu = 0, dgrhonu = 0;
    for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
      const double grhormass_t = M.grhormass[nu_i] / (a * a);
      double clxnu, qnu, pinu, dpnu, pinudot;
      if (e.MassiveNuApprox[nu_i]) {
        clxnu = AT(y, e.nu_ix[nu_i]);
        qnu = AT(y, e.nu_ix[nu_i] + 2);
        pinu = AT(y, e.nu_ix[nu_i] + 3);
        pinudot = AT(yprime, e.nu_ix[nu_i] + 3);
      } else {
        Nu_Integrate_L012(M, e, y, a, nu_i, clxnu, qnu, &dpnu, &pinu);
        qnu = qnu / rhonu[nu_i];
        clxnu = clxnu / rhonu[nu_i];
        pinu = pinu / rhonu[nu_i];
        const double adotoa_n = 1 / (a * dtauda_a);
        const double rhonudot = Nu_drho(M, a * M.nu_masses[nu_i], adotoa_n, rhonu[nu_i]);
        pinudot = Nu_pinudot(M, e, y, yprime, a, adotoa_n, nu_i);
        pinudot = pinudot / rhonu[nu_i] - rhonudot / rhonu[nu_i] * pinu;
      }
      const double grhonu_t = grhormass_t * rhonu[nu_i];
      const double gpnu_t = grhormass_t * pnu[nu_i];
      grhonu = grhonu + grhonu_t;
      gpres = gpres + gpnu_t;
      dgrhonu = dgrhonu + grhonu_t * clxnu;
      dgq = dgq + grhonu_t * qnu;
      dgpi = dgpi + grhonu_t * pinu;
      dgpi_diff = dgpi_diff + pinu * (3 * gpnu_t - grhonu_t);
      pidot_sum = pidot_sum + grhonu_t * pinudot;
    }
    grho = grho + grhonu;
    dgrho = dgrho + dgrhonu;
  }
  const double adotoa = 1. / (a * dtauda_a);
  double z, dz, clxr, qr, pir, pirdot, clxg, qg, pig, pigdot = 0, octg, octgprime, qgdot, dgrhogal_t;
  if (e.no_nu) {
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
    clxr = AT(y, e.r_ix);
    qr = AT(y, e.r_ix + 1);
    pir = AT(y, e.r_ix + 2);
    pirdot = AT(yprime, e.r_ix + 2);
  }
  if (e.no_phot) {
    if (gal) {
      dgrhogal_t = gal_Chigal(M, a, hub, xgal, dgrho, etak, dphi, dphiprime, k);
      z = (0.5 * (dgrho + dgrhogal_t) / k + etak) / adotoa;
      dz = -adotoa * z - 0.5 * (dgrho + dgrhogal_t) / k;
    } else {
      z = (0.5 * dgrho / k + etak) / adotoa;
      dz = -adotoa * z - 0.5 * dgrho / k;
    }
    clxg = -4 * dz / k - 4 / k * opac * (vb + z);
    qg = -4.0 / 3 * z;
    pig = 0;
    pigdot = 0;
    octg = 0;
    octgprime = 0;
    qgdot = -4 * dz / 3;
  } else {
    if (e.TightCoupling) {
      pig = e.pig;
      octg = (3.0 / 7.0) * pig * (e.k / opac);
      pol2 = e.pig / 4 + e.pigdot * (1.0 / opac) * (-5.0 / 8.0);
      pol3 = (3.0 / 7.0) * (e.k / opac) * pol2;
      octgprime = 0;
    } else {
      pig = AT(y, e.g_ix + 2);
      pigdot = AT(yprime, e.g_ix + 2);
      octg = AT(y, e.g_ix + 3);
      octgprime = AT(yprime, e.g_ix + 3);
    }
    clxg = AT(y, e.g_ix);
    qg = AT(y, e.g_ix + 1);
    qgdot = AT(yprime, e.g_ix + 1);
  }
  dgrho = dgrho + grhog_t * clxg + grhor_t * clxr;
  dgq = dgq + grhog_t * qg + grhor_t * qr;
  dgpi = dgpi + grhor_t * pir + grhog_t * pig;
  double dgpigal_t = 0;
  if (gal) {
    dgrhogal_t = gal_Chigal(M, a, hub, xgal, dgrho, etak, dphi, dphiprime, k);
    dgrho = dgrho + dgrhogal_t;
    dgq = dgq + gal_qgal(M, a, hub, xgal, dgq, etak, dphi, dphiprime, k);
    dgpigal_t = gal_Pigal(M, a, hub, xgal, dh, dx, dgrho, dgq, dgpi, etak, dphi, k);
    dgpi = dgpi + dgpigal_t;
  }
  z = (0.5 * dgrho / k + etak) / adotoa;
  const double sigma = (z + 1.5 * dgq / k2) / 1.0;
  const double polter = 0.1 * pig + 9.0 / 15.0 * pol2;
  const double x = k * (M.tau0 - tau);
  const double divfac = x * x;
  if (e.TightCoupling) {
    pigdot = e.pigdot;
    polprime2 = (pigdot / 4.0) * (1 + (5.0 / 2.0) * (dopac / (opac * opac)));
  }
  pidot_sum = pidot_sum + grhog_t * pigdot + grhor_t * pirdot;
  double diff_rhopi;
  if (gal) {
    diff_rhopi = pidot_sum - (4 * (dgpi - dgpigal_t) + dgpi_diff) * adotoa;
    diff_rhopi = diff_rhopi + gal_pigalprime(M, a, hub, xgal, dh, dx, dgrho, dgq, dgpi, diff_rhopi, etak, dphi, dphiprime,
                                             k, grho, gpres, pnu);
  } else {
    diff_rhopi = pidot_sum - (4 * dgpi + dgpi_diff) * adotoa;
  }
  const double Kf1 = 1.0, Kf2 = 1.0;
  const double kk = k * k;
  const double ISW = (4.0 / 3.0 * k * Kf1 * sigma + (-2.0 / 3.0 * sigma - 2.0 / 3.0 * etak / adotoa) * k - diff_rhopi / kk -
                      1.0 / adotoa * dgrho / 3.0 + (3.0 * gpres + 5.0 * grho) * sigma / k / 3.0 - 2.0 / k * adotoa / Kf1 * etak) *
                     expmmu;
  sources[0] =
      ISW +
      ((-9.0 / 160.0 * pig - 27.0 / 80.0 * pol2) / kk * opac +
       (11.0 / 10.0 * sigma - 3.0 / 8.0 * Kf2 * pol3 + vb - 9.0 / 80.0 * Kf2 * octg + 3.0 / 40.0 * qg) / k -
       (-180.0 * polprime2 - 30.0 * pigdot) / kk / 160.0) *
          dvis +
      (-(9.0 * pigdot + 54.0 * polprime2) / kk * opac / 160.0 + pig / 16.0 + clxg / 4.0 + 3.0 / 8.0 * pol2 +
       (-21.0 / 5.0 * adotoa * sigma - 3.0 / 8.0 * Kf2 * polprime3 + vbdot + 3.0 / 40.0 * qgdot - 9.0 / 80.0 * Kf2 * octgprime) / k +
       (-9.0 / 160.0 * dopac * pig - 21.0 / 10.0 * dgpi - 27.0 / 80.0 * dopac * pol2) / kk) *
          vis +
      (3.0 / 16.0 * ddvis * pig + 9.0 / 8.0 * ddvis * pol2) / kk + 21.0 / 10.0 / k / Kf1 * vis * etak;
  if (x > 0.0)
    sources[1] = vis * polter * (15.0 / 8.0) / divfac;
  else
    sources[1] = 0;
  if (M.SourceNum > 2) {
    if (tau > M.tau_maxvis && M.tau0 - tau > 0.1) {
      const double phi = -(dgrho + 3 * dgq * adotoa / k) / (k2 * Kf1 * 2) - dgpi / k2 / 2;
      sources[2] = -2 * phi * (tau - M.tau_maxvis) / ((M.tau0 - M.tau_maxvis) * (M.tau0 - tau));
    } else {
      sources[2] = 0;
    }
  }
}

// ---------------------------------------------------------------- initial conditions
// camb/equations_galileon.f90:1714-1928 (adiabatic mode), y = workspace column 0
__device__ inline void initial(const DevModel& M, EV& e, double* y, double tau) {
  e.k = e.q;
  e.k2 = e.q * e.q;
  e.ik = 1.0 / e.k;
  e.ik2 = e.ik * e.ik;
  const double k = e.k;
  double ep;
  const double ep0 = 1.0e-2;
  if (e.k > M.epsw) {
    if (e.k > M.epsw * 5) {
      ep = ep0 * 5 / M.AccuracyBoost;
      if (M.HighAccuracyDefault) ep = ep * GC_SP(0.65);
    } else {
      ep = ep0;
    }
  } else {
    ep = ep0;
  }
  ep = ep * 2;
  e.TightSwitchoffTime = fmin(M.tight_tau, Thermo_OpacityToTime(M, e.k / ep));
  for (int i = 1; i <= e.nvar; i++) AT(y, i) = 0;
  const double x = k * tau, x2 = x * x, x3 = x2 * x;
  double rhomass = 0;
  for (int i = 0; i < M.n_eig; i++) rhomass += M.grhormass[i];
  const double grhonu = rhomass + M.grhornomass;
  const double om = (M.grhob + M.grhoc) / sqrt(3 * (M.grhog + grhonu));
  const double omtau = om * tau;
  const double Rv = grhonu / (grhonu + M.grhog);
  const double Rp15 = 4 * Rv + 15;
  const double a = tau * M.adotrad * (1 + omtau / 4);
  const double chi = 1, Kf1 = 1.0;
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
  AT(y, 1) = a;
  AT(y, 2) = -(-i_eta) * k / 2;
  AT(y, 3) = -i_clxc;
  AT(y, 4) = -i_clxb;
  AT(y, 5) = -i_vb;
  AT(y, e.g_ix) = -i_clxg;
  AT(y, e.g_ix + 1) = -i_qg;
  AT(y, e.r_ix) = -i_clxr;
  AT(y, e.r_ix + 1) = -i_qr;
  AT(y, e.r_ix + 2) = -i_pir;
  if (e.lmaxnr > 2) AT(y, e.r_ix + 3) = -i_aj3r;
  if (M.Num_Nu_massive == 0) return;
  for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
    e.MassiveNuApproxTime[nu_i] = M.nu_tau_massive[nu_i];
    const double a_massive = 20000 * k / M.nu_masses[nu_i] * M.AccuracyBoost * M.lAccuracyBoost;
    if (a_massive >= GC_SP(0.99)) {
      e.MassiveNuApproxTime[nu_i] = M.tau0 + 1;
    } else if (a_massive > 17.0 / M.nu_masses[nu_i] * M.AccuracyBoost) {
      e.MassiveNuApproxTime[nu_i] = fmax(e.MassiveNuApproxTime[nu_i], rombint_dtauda(M, 0.0, a_massive, 0.01));
    }
    int ind = e.nu_ix[nu_i];
    for (int i = 1; i <= e.nq[nu_i]; i++) {
      AT(y, ind) = AT(y, e.r_ix);
      AT(y, ind + 1) = AT(y, e.r_ix + 1);
      AT(y, ind + 2) = AT(y, e.r_ix + 2);
      if (e.lmaxnu_tau[nu_i] > 2) AT(y, ind + 3) = -i_aj3r;
      ind = ind + e.lmaxnu_tau[nu_i] + 1;
    }
  }
}

}  // namespace gc
