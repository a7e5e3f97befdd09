// Per-(k, cosmology) scalar perturbation evolution: device functions.
//
// Restates, for sm_100a, the reference routines
//   GetNumEqns / SetupScalarArrayIndices / CopyScalarVariableArray   camb/equations_galileon.f90:555-946
//   initial / derivs / output                                         camb/equations_galileon.f90:1714-1928,
//                                                                     :2098-2589, :1256-1576
//   the Galileon closed forms                                          camb/galileon.cc:691-1174
//   thermo / massive-neutrino table lookups                            camb/modules.f90:2728-2769, :1721-1850
// One work item = one (k, cosmology) pair, owned by an OCTET of eight lanes (csrc/evolve.cu).  The state vector y
// and the Runge-Kutta columns of an item live in shared memory, unit stride: element i of a column is col[i-1].
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdint>

#include "galileon_fast.cuh"
#include "model.h"

// functions shared with the host-side galileon.cc replacement (csrc/galileon_host.cu)
#define GC_HOSTDEV __host__ __device__
// read-only table rows (thermo, background spline, neutrino tables): global, never written by the kernel
#ifdef __CUDA_ARCH__
#define GC_LDG(p) __ldg(p)
#else
#define GC_LDG(p) (*(p))
#endif

__constant__ double c_denl[260];    // 1/(2l+1)            camb/equations_galileon.f90:525-527
__constant__ double c_polfac[260];  // (l+3)(l-1)/(l+1)    camb/equations_galileon.f90:516-520

namespace gc {

// float-rounded literals of the reference's default-real constants (no _dl suffix)
#define GC_SP(x) ((double)(x##f))

// eigenstates the per-item records are sized for: GC_MAX_NU (= the reference's max_nu) in the general kernels, 1 in the
// -DGC_EV_NU=1 compilation the launcher picks when no model of the batch has more than one mass eigenstate (smaller
// records: the lane-per-item kernel then leaves 60 KB instead of 28 KB of each SM to the L1 cache)
#ifndef GC_EV_NU
#define GC_EV_NU GC_MAX_NU
#endif

struct EV {
  double q, k, k2, ik, ik2;  // ik = 1/k, ik2 = 1/k^2 (per-lane constants)
  // 16/8-bit layout fields: the record sits in shared memory once per work item (csrc/evolve.cu) or per lane
  // (csrc/evolve_batch.cu: 256 records per SM)
  short w_ix, r_ix, g_ix, polind, nu_pert_ix, nvar, neq;
  short lmaxg, lmaxnr, lmaxnu, lmaxgpol, lmaxnu_pert;
  short nu_ix[GC_EV_NU];
  unsigned char nq[GC_EV_NU], lmaxnu_tau[GC_EV_NU];
  bool TightCoupling, high_ktau, no_nu, no_phot, has_nu_rel;
#ifdef GC_WITH_TRANSFER
  bool TransferOnly;  // a wavenumber beyond the source grid, camb/cmbmain.f90:1165
#endif
  bool MassiveNuApprox[GC_EV_NU], nu_nonrel[GC_EV_NU];
  double G11[GC_EV_NU], G30[GC_EV_NU], MassiveNuApproxTime[GC_EV_NU];
  double TightSwitchoffTime, pig, pigdot, poltruncfac;
};

// column element accessors (1-based i like the reference).  GC_STRIDE = 1: unit-stride columns (shared memory, the
// eight-lane kernel csrc/evolve.cu); GC_STRIDE = 32: lane-interleaved columns of the lane-per-item batch kernel
// (csrc/evolve_batch.cu defines it before including this file)
#ifndef GC_STRIDE
#define GC_STRIDE 1
#endif
#define AT(col, i) (col)[((i)-1) * GC_STRIDE]

__device__ __forceinline__ double denlk(const EV& e, int l) { return c_denl[l] * e.k * l; }
__device__ __forceinline__ double denlk2(const EV& e, int l) { return c_denl[l] * e.k * 1.0 * (l + 1); }
__device__ __forceinline__ double polfack(const EV& e, int l) { return c_polfac[l] * e.k * 1.0 * c_denl[l]; }

__device__ __forceinline__ long nintd(double x) { return lround(x); }

// ---------------------------------------------------------------- table lookups
GC_HOSTDEV __forceinline__ double herm(double f0, double d0, double f1, double d1, double d) {
  return f0 + d * (d0 + d * (3 * (f1 - f0) - 2 * d0 - d1 + d * (d0 + d1 + 2 * (f0 - f1))));
}

// camb/modules.f90:2728-2769
__device__ inline void thermo(const DevModel& M, double tau, double& cs2b, double& opacity, double* dopacity) {
  const int nth = M.nthermo;
  double d = log(tau * M.inv_tauminn) * M.inv_dlntau + 1.0;
  // The last node is a branch point (below it the Hermite cubic, from it on the reference's end-of-table formula, which
  // is NOT its continuation), and tau = tau0 sits exactly on it: there the index has to come out of the reference's own
  // arithmetic, two divisions instead of the reciprocal multiplies.
  if (d > nth - 0.5) d = log(tau / M.tauminn) / M.dlntau + 1.0;
  int i = (int)d;
  d = d - i;
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
template <class ModelT>
GC_HOSTDEV inline void Nu_background(const ModelT& M, double am, double& rhonu, double& pnu) {
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
GC_HOSTDEV inline double Nu_drho(const DevModel& M, double am, double adotoa, double rhonu) {
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
template <class ModelT>
GC_HOSTDEV inline double Nu_dp(const ModelT& M, double am, double adotoa, double pnu) {
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

// ---------------------------------------------------------------- state-vector layout
// camb/equations_galileon.f90:298-312
__device__ inline int next_nu_nq(const DevModel& M, int nq) {
  if (nq == 0) return 1;
  const int qq = (int)M.nu_q[nq - 1];
  if (qq >= 10) return M.nqmax;
  return nq + 1;
}

// camb/equations_galileon.f90:555-644
__device__ inline void SetupScalarArrayIndices(const DevModel& M, EV& e, short* max_num_eqns) {
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
  if (max_num_eqns) *max_num_eqns = (short)maxeq;
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
  for (int i = 0; i < GC_EV_NU; i++) e.MassiveNuApprox[i] = false;
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
#ifdef GC_WITH_TRANSFER
  if (e.TransferOnly) {  // :872-875
    e.lmaxgpol = min((int)e.lmaxgpol, (int)nintd(5 * lAB));
    e.lmaxg = min((int)e.lmaxg, (int)nintd(6 * lAB));
  }
#endif
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
  double rhonu[GC_EV_NU], pnu[GC_EV_NU];
  for (int i = 0; i < M.n_eig; i++) Nu_background(M, a * M.nu_masses[i], rhonu[i], pnu[i]);
  double xgal = 0, hub = 0, hbar = 0;
  GalFast F;
  GalDyn B;
  const GalConst GC{M.c2, M.c3, M.c4, M.c5, M.cG, M.h0, M.orad};
  if (gal) {
    gal_HX(M, a, hbar, xgal);
    hub = a * hbar;
    double lam_nu = 0;
    for (int i = 0; i < M.n_eig; i++) lam_nu += M.grhormass[i] * pnu[i];
    gal_fast_background(GC, a, hbar, xgal, lam_nu, k, F, &B);
    grho = grho + F.grhogal;
    gpres = gpres + F.gpresgal;
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
      dgrhogal_t = gal_fast_Chigal(F, dgrho, etak, dphi, dphiprime);
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
      dgrhogal_t = gal_fast_Chigal(F, dgrho, etak, dphi, dphiprime);
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
    dgrhogal_t = gal_fast_Chigal(F, dgrho, etak, dphi, dphiprime);
    dgrho = dgrho + dgrhogal_t;
    dgq = dgq + gal_fast_qgal(F, dgq, dphi, dphiprime);
    dgpigal_t = gal_fast_Pigal(GC, F, dgrho, dgq, dgpi, etak, dphi);
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
    // pigalprime_ (camb/galileon.cc:1084-1174) as a linear form; its massive-neutrino term hands Nu_dp the H0-unit
    // Hubble rate `hub` as adotoa, like the reference does (:1133-1137)
    double lam_nu_prime = 0;
    for (int i = 0; i < M.n_eig; i++)
      lam_nu_prime += M.grhormass[i] * (Nu_dp(M, a * M.nu_masses[i], hub, pnu[i]) - 4 * pnu[i] * hub);
    GalPiDotLin PL;
    gal_fast_pigalprime_lin(GC, F, B, lam_nu_prime, grho + gpres, PL);
    diff_rhopi = diff_rhopi + (PL.p * dphiprime + PL.d * dphi + PL.pid * diff_rhopi + PL.g * dgrho + PL.q * dgq + PL.pi * dgpi +
                               PL.e * etak);
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

#ifdef GC_WITH_TRANSFER
// ---------------------------------------------------------------- matter transfer functions
// outtransf (camb/equations_galileon.f90:2078-2095): the transfer tap of derivs (:2325-2353) evaluated on the state y
// at time tau; out[0..12] = TransferData(Transfer_kh .. Transfer_vel_baryon_cdm), rows 2..13 already divided by k^2.
// Only the part of derivs above the tap is evaluated (the derivatives themselves are not needed).  During tight
// coupling the reference reads pig before assigning it (:2331 vs :2375); 0 here and in the oracle.
static __device__ __noinline__ void transfer_outputs(const DevModel* __restrict__ Mp, const EV* __restrict__ ep,
                                                     const double* __restrict__ y, double tau, double* __restrict__ out) {
  const DevModel& M = *Mp;
  const EV& e = *ep;
  const bool gal = M.use_galileon != 0;
  const double k = e.k, k2 = e.k2;
  const double a = AT(y, 1), a2 = a * a, etak = AT(y, 2), clxc = AT(y, 3), clxb = AT(y, 4), vb = AT(y, 5);
  double dphi = 0, dphiprime = 0;
  if (gal) {
    dphi = AT(y, e.w_ix);
    dphiprime = AT(y, e.w_ix + 1);
  }
  const double grhob_t = M.grhob / a, grhoc_t = M.grhoc / a, grhor_t = M.grhornomass / a2, grhog_t = M.grhog / a2;
  double rhonu[GC_EV_NU], pnu[GC_EV_NU];
  for (int i = 0; i < M.n_eig; i++) Nu_background(M, a * M.nu_masses[i], rhonu[i], pnu[i]);
  double xgal = 0, hbar = 0;
  GalFast F;
  const GalConst GC{M.c2, M.c3, M.c4, M.c5, M.cG, M.h0, M.orad};
  if (gal) {
    gal_HX(M, a, hbar, xgal);
    double lam_nu = 0;
    for (int i = 0; i < M.n_eig; i++) lam_nu += M.grhormass[i] * pnu[i];
    gal_fast_background(GC, a, hbar, xgal, lam_nu, k, F);
  }
  double cs2, opacity;
  thermo(M, tau, cs2, opacity, nullptr);
  double grho_matter = grhob_t + grhoc_t;
  double dgrho_matter = grhob_t * clxb + grhoc_t * clxc;
  double dgq = grhob_t * vb;
  double dgpi_nu = 0, grhonu = 0, dgrhonu = 0;
  if (M.Num_Nu_massive > 0) {  // MassiveNuVars (:1209-1252) and the two outputs of MassiveNuVarsOut the tap asks for (:2335)
    for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
      const double grhormass_t = M.grhormass[nu_i] / a2;
      double clxnu, qnu, dpnu, pinu;
      if (e.MassiveNuApprox[nu_i]) {
        clxnu = AT(y, e.nu_ix[nu_i]);
        qnu = AT(y, e.nu_ix[nu_i] + 2);
        pinu = AT(y, e.nu_ix[nu_i] + 3);
      } else {
        Nu_Integrate_L012(M, e, y, a, nu_i, clxnu, qnu, &dpnu, &pinu);
        qnu = qnu / rhonu[nu_i];
        clxnu = clxnu / rhonu[nu_i];
        pinu = pinu / rhonu[nu_i];
      }
      const double grhonu_t = grhormass_t * rhonu[nu_i];
      grho_matter = grho_matter + grhonu_t;
      dgrho_matter = dgrho_matter + grhonu_t * clxnu;
      dgq = dgq + grhonu_t * qnu;
      grhonu = grhonu + grhonu_t;
      dgrhonu = dgrhonu + grhonu_t * clxnu;
      dgpi_nu = dgpi_nu + grhonu_t * pinu;
    }
  }
  // dtauda(a), :107-156
  double grhoa2;
  if (gal && a >= 1e-10) {
    grhoa2 = (a2 * a2) * M.grhom * (hbar * hbar);
  } else {
    grhoa2 = M.grhok * a2 + (M.grhoc + M.grhob) * a + M.grhog + M.grhornomass;
    grhoa2 = grhoa2 + M.grhov * (a2 * a2);
    if (M.Num_Nu_massive != 0)
      for (int nu_i = 0; nu_i < M.n_eig; nu_i++) grhoa2 = grhoa2 + rhonu[nu_i] * M.grhormass[nu_i];
  }
  const double adotoa = 1. / (a * sqrt(3 / grhoa2));
  double dgrho = dgrho_matter;
  double z = 0, dz, clxr, qr, pir, clxg, qg, pig = 0, dgrhogal_t;
  if (e.no_nu) {
    if (gal) {
      dgrhogal_t = gal_fast_Chigal(F, dgrho, etak, dphi, dphiprime);
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
    clxr = AT(y, e.r_ix);
    qr = AT(y, e.r_ix + 1);
    pir = AT(y, e.r_ix + 2);
  }
  if (e.no_phot) {
    if (!e.no_nu) {
      if (gal) {
        dgrhogal_t = gal_fast_Chigal(F, dgrho, etak, dphi, dphiprime);
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
    clxg = AT(y, e.g_ix);
    qg = AT(y, e.g_ix + 1);
    if (!e.TightCoupling) pig = AT(y, e.g_ix + 2);
  }
  dgrho = dgrho + grhog_t * clxg + grhor_t * clxr;
  dgq = dgq + grhog_t * qg + grhor_t * qr;
  if (gal) {
    dgrhogal_t = gal_fast_Chigal(F, dgrho, etak, dphi, dphiprime);
    dgrho = dgrho + dgrhogal_t;
    dgq = dgq + gal_fast_qgal(F, dgq, dphi, dphiprime);
  }
  z = (0.5 * dgrho / k + etak) / adotoa;
  const double sigma = (z + 1.5 * dgq / k2);
  double dgpi = grhor_t * pir + grhog_t * pig;
  double clxnu_all = 0;
  if (M.Num_Nu_massive != 0) {
    dgpi = dgpi + dgpi_nu;
    clxnu_all = dgrhonu / grhonu;
  }
  if (gal) dgpi = dgpi + gal_fast_Pigal(GC, F, dgrho, dgq, dgpi, etak, dphi);
  const double ik2 = 1.0 / k2;
  out[0] = k / M.h0_100;
  out[1] = clxc * ik2;
  out[2] = clxb * ik2;
  out[3] = clxg * ik2;
  out[4] = clxr * ik2;
  out[5] = clxnu_all * ik2;
  out[6] = (dgrho_matter / grho_matter) * ik2;
  out[7] = ((grhob_t * clxb + grhoc_t * clxc) / (grhob_t + grhoc_t)) * ik2;
  out[8] = (dgrho / grho_matter) * ik2;
  out[9] = (-(dgrho + 3 * dgq * adotoa / k) / 2 - dgpi / 2) * ik2;
  out[10] = (-k * sigma / adotoa) * ik2;
  out[11] = (-k * (vb + sigma) / adotoa) * ik2;
  out[12] = vb * ik2;
}
#endif  // GC_WITH_TRANSFER

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
