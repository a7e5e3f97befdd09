// K1 "evolve_sources": per-(k, cosmology) integration of the Einstein-Boltzmann + Galileon system with the
// Verner 6(5) pair of dverk, sampling the CMB sources on the output time grid.
//
// Replaces the OpenMP loop camb/cmbmain.f90:201-206 (DoSourcek :662 -> CalcScalarSources :917 ->
// GaugeInterface_EvolveScal camb/equations_galileon.f90:314 -> dverk camb/subroutines.f90:370 -> derivs/output).
//
// Execution model (DESIGN.md "K1"): one lane per (k, cosmology) work item, FP64 throughout.  Each lane runs the
// reference's control flow (output-time loop, approximation switches, adaptive step accept/reject) as an explicit
// scalar state machine (advance) that stops whenever it needs one of three vector operations, which the warp then
// executes as ONE converged call for all of its lanes: the right-hand side (derivs), a stage combination of the
// Runge-Kutta tableau (combine) or the source evaluation of an output time (flush_output).  Lanes of different
// cosmologies need different numbers of steps; they are kept in the same Runge-Kutta stage by letting a lane start a
// trial step only every 8th round (PH_DV_LOOP), so that only the short scalar bookkeeping diverges.
#include <algorithm>
#include <cstddef>
#include <cstdio>

// -DGC_WITH_TRANSFER compiles this file a second time with the matter-transfer taps of CP%WantTransfer
// (camb/cmbmain.f90:1017-1068, :1151-1201) in the control flow, in its own namespace and under entry-point names ending
// in _tr: the default kernels stay exactly as they are without the taps (the extra state-machine phases cost registers
// in the hot loop: +14 % on a 1024-point batch when compiled in unconditionally).
// -DGC_EV_NU=1 compiles it a third time with the per-item records sized for one massive-neutrino eigenstate (entry points
// ending in _nu1): the launcher's choice when no model of the batch has more (the usual CosmoMC set-up).
#if defined(GC_WITH_TRANSFER)
#define gc gc_tr
#define GC_ENTRY(name) name##_tr
#elif defined(GC_EV_NU)
#define gc gc_nu1
#define GC_ENTRY(name) name##_nu1
#else
#define GC_ENTRY(name) name
#endif

// lane-interleaved columns: element i of a column is col[(i-1)*32] (the warp touches one 256-byte line per element)
#define GC_STRIDE 32
#define GC_LANES 32
#define GC_NCOL 12  // column 0 = y, columns 1..9 = dverk w(:,1..9), 10 = y of a pending output, 11 = its y'
#include "evolve_device.cuh"

// elements of a multipole hierarchy per pass: their loads are issued together, so one exposed L2 latency per pass
#ifndef GC_HIER_UNROLL
#define GC_HIER_UNROLL 4
#endif
constexpr int kHierUnroll = GC_HIER_UNROLL;
#ifndef GC_EARLY_FIELD_EQ
#define GC_EARLY_FIELD_EQ 1
#endif

#ifndef GC_STATE_CG
#define GC_STATE_CG 0
#endif
// the derivative of the last stage (k8) is stored in k2's column (dead after row 6 of the tableau): 9 live columns
// per lane instead of 10 in the L2-resident workspace
#ifndef GC_K8_IN_K2
#define GC_K8_IN_K2 1
#endif
// high multipoles of every hierarchy evaluated ahead of the Einstein / Galileon part of derivs() (see there)
#ifndef GC_EARLY_HIER
#define GC_EARLY_HIER 1
#endif
// stage combinations with two batches of loads in flight (see combine_uniform)
#ifndef GC_COMBINE_PIPE
#define GC_COMBINE_PIPE 0
#endif
#ifndef GC_COMBINE_NC4
#define GC_COMBINE_NC4 6  // rows reading at least this many columns use batches of 4 elements instead of 8
#endif
#ifndef GC_PIPE_MIN_NC
#define GC_PIPE_MIN_NC 1
#endif
#ifndef GC_PIPE_MAX_NC
#define GC_PIPE_MAX_NC 7
#endif

namespace gc {

// Read-only view of a workspace column for derivs(): with GC_STATE_CG the element loads bypass L1 (ld.global.cg), which
// then holds only table rows, per-model constants and spilled registers instead of the once-read state stream.
struct InCol {
  const double* __restrict__ p;
  __device__ __forceinline__ double operator[](int i) const {
#if GC_STATE_CG
    return __ldcg(p + i);
#else
    return p[i];
#endif
  }
  __device__ __forceinline__ operator const double*() const { return p; }
};

// One pass over a run of `cnt` consecutive multipoles of a hierarchy, first column index ix0, first multipole l0:
// y'(ix) = f(l, y(ix-1), y(ix), y(ix+1)).  All loads of a batch of U equations are issued together (U + 2 independent
// loads, one exposed memory latency) before any arithmetic; the expressions themselves are the caller's, unchanged.
// The last equation of a run is the truncation formula, which does not read y(ix+1): that element is not loaded.
// skip_l: a multipole whose equation the caller evaluates elsewhere (nothing is stored for it).
#ifndef GC_HIER_BATCH
#define GC_HIER_BATCH 6
#endif
// the two halves of a batch: the U + 2 loads of batch c (c = 0, U, 2U ...) of a run, and its U equations
template <int U>
__device__ __forceinline__ void hier_load(const InCol ay, int ix0, int cnt, int c, double (&v)[U + 2]) {
#pragma unroll
  for (int u = 0; u < U + 2; u++) v[u] = (c < cnt && c + u <= cnt) ? AT(ay, ix0 + c + u - 1) : 0.0;
}
template <int U, class F>
__device__ __forceinline__ void hier_eval(double* __restrict__ ayp, int ix0, int l0, int cnt, int c, int skip_l,
                                          const double (&v)[U + 2], F f) {
#pragma unroll
  for (int u = 0; u < U; u++) {
    const int l = l0 + c + u;
    if (c + u < cnt && l != skip_l) AT(ayp, ix0 + c + u) = f(l, v[u], v[u + 1], v[u + 2]);
  }
}
// batches c0, c0 + U, ... of a run, one after the other
template <class F>
__device__ __forceinline__ void hier_run(const InCol ay, double* __restrict__ ayp, int ix0, int l0, int cnt, int skip_l, int c0, F f) {
  constexpr int U = GC_HIER_BATCH;
  for (int c = c0; c < cnt; c += U) {
    double v[U + 2];
    hier_load<U>(ay, ix0, cnt, c, v);
    hier_eval<U>(ayp, ix0, l0, cnt, c, skip_l, v, f);
  }
}

// The two thermo-table rows of the lane's last look-up, kept in its shared-memory record: consecutive right-hand sides of
// an item fall into the same interval of the 20 000-row table most of the time, and with 32 models per warp every fresh
// look-up is 32 scattered cache lines (usually from HBM: the tables of a batch are 1.7 GB).  thermo() of evolve_device.cuh
// with the rows taken from here when the interval is the cached one; same arithmetic.
struct ThermoRows {
  int i;  // 1-based lower row of the interval (0 = nothing cached)
  double v[10];
};
// row 5 of the tableau formed by the pass of row 4 (same columns): seven stage-combination passes per step instead of eight
#ifndef GC_FUSE_ROW5
#define GC_FUSE_ROW5 (GC_K8_IN_K2)
#endif
// the background-spline rows ahead of the hierarchy runs: 1 = loaded into registers there (measured: 4436 vs 4189 ms, the
// twenty registers held across the runs spill), 2 = prefetched into L2 there, 0 = read where gal_HX() needs them
#ifndef GC_EARLY_GAL_ROWS
#define GC_EARLY_GAL_ROWS 0
#endif
#ifndef GC_LOCKSTEP_EVERY
#define GC_LOCKSTEP_EVERY 1
#endif
#ifndef GC_THERMO_CACHE
#define GC_THERMO_CACHE 1
#endif
__device__ __forceinline__ void thermo_cached(const DevModel& M, ThermoRows* __restrict__ tc, double tau, double& cs2b, double& opacity,
                                              double* dopacity) {
  const int nth = M.nthermo;
  double d = log(tau * M.inv_tauminn) * M.inv_dlntau + 1.0;
  if (d > nth - 0.5) d = log(tau / M.tauminn) / M.dlntau + 1.0;  // see thermo()
  int i = (int)d;
  d = d - i;
  if (i >= nth) {
    const double* r = M.th + (size_t)(nth - 1) * 5;
    cs2b = GC_LDG(r) + (d + i - nth) * GC_LDG(r + 1);
    opacity = GC_LDG(r + 2) + (d - nth) * GC_LDG(r + 3);
    if (dopacity) *dopacity = 0;
    return;
  }
  if (i < 1) i = 1;
  double a[10];
  if (tc->i == i) {
#pragma unroll
    for (int c = 0; c < 10; c++) a[c] = tc->v[c];
  } else {
    const double* r0 = M.th + (size_t)(i - 1) * 5;
#pragma unroll
    for (int c = 0; c < 10; c++) a[c] = GC_LDG(r0 + c);  // rows i and i + 1 are adjacent
#pragma unroll
    for (int c = 0; c < 10; c++) tc->v[c] = a[c];
    tc->i = i;
  }
  cs2b = herm(a[0], a[1], a[5], a[6], d);
  opacity = herm(a[2], a[3], a[7], a[8], d);
  if (dopacity) *dopacity = herm(a[3], a[4], a[8], a[9], d) * M.inv_dlntau / tau;
}

// gal_HX() of evolve_device.cuh in two halves, so that derivs() can issue the ten loads of the two background-spline rows
// as soon as it knows a and spend their latency (usually an HBM access: 32 models per warp, 1.2 MB of rows per model) on
// the hierarchy runs; same index rule, same arithmetic.
__device__ __forceinline__ int gal_rows_load(const DevModel& M, double a, double (&v0)[5], double (&v1)[5]) {
  const int n = M.ngal;
  int i = (int)(log(a / M.gal_a0) * M.gal_lnq_inv);
  if (i < 0) i = 0;
  if (i > n - 2) i = n - 2;
  const double* r0 = M.gal + (size_t)i * 5;
#pragma unroll
  for (int c = 0; c < 5; c++) {
    v0[c] = GC_LDG(r0 + c);
    v1[c] = GC_LDG(r0 + 5 + c);
  }
  return i;
}
__device__ __forceinline__ void gal_rows_eval(const DevModel& M, double a, int i, double (&v0)[5], double (&v1)[5], double& hbar,
                                              double& xgal) {
  const int n = M.ngal;
  if (!(v0[0] <= a && (a < v1[0] || i == n - 2))) {  // the guess missed the interval by rounding: dependent-load search
    while (i > 0 && M.gal[(size_t)i * 5] > a) i--;
    while (i < n - 2 && M.gal[(size_t)(i + 1) * 5] <= a) i++;
    const double* r0 = M.gal + (size_t)i * 5;
#pragma unroll
    for (int c = 0; c < 5; c++) {
      v0[c] = r0[c];
      v1[c] = r0[5 + c];
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

// Nu_Integrate_L012 (camb/equations_galileon.f90:1058-1107; evolve_device.cuh) for derivs(): density and flux moments
// only, the loads of all momentum bins issued before the sums (same terms, same order).
__device__ __forceinline__ void nu_moments_01(const DevModel& M, const EV& e, const InCol y, double a, int nu_i, double& drhonu,
                                              double& fnu) {
  const double am = a * M.nu_masses[nu_i];
  const int nq = e.nq[nu_i], nqmax = M.nqmax;
  const int ind = e.nu_ix[nu_i], step = e.lmaxnu_tau[nu_i] + 1, indp = e.nu_pert_ix;
  double y0[GC_MAX_NQ], y1[GC_MAX_NQ];
#pragma unroll
  for (int iq = 0; iq < GC_MAX_NQ; iq++) {
    const bool own = iq < nq, pert = !own && iq < nqmax;
    y0[iq] = own ? AT(y, ind + iq * step) : 0.0;
    y1[iq] = own ? AT(y, ind + iq * step + 1) : 0.0;
    (void)pert;
  }
  double r0 = 0, r1 = 0, p0 = 0, p1 = 0;
  if (nq < nqmax) {
    r0 = AT(y, e.r_ix);
    r1 = AT(y, e.r_ix + 1);
    p0 = AT(y, indp);
    p1 = AT(y, indp + 1);
  }
  drhonu = 0;
  fnu = 0;
#pragma unroll
  for (int iq = 0; iq < GC_MAX_NQ; iq++) {
    if (iq < nq) {
      const double aq = am * M.inv_nu_q[iq];
      const double e2 = 1.0 + aq * aq;
      const double v = rsqrt(e2), iv = e2 * v;
      const double ker = M.nu_int_kernel[iq];
      drhonu = drhonu + ker * y0[iq] * iv;
      fnu = fnu + ker * y1[iq];
    }
  }
#pragma unroll
  for (int iq = 0; iq < GC_MAX_NQ; iq++) {
    if (iq >= nq && iq < nqmax) {
      const double aq = am * M.inv_nu_q[iq];
      const double e2 = 1.0 + aq * aq;
      const double v = rsqrt(e2), iv = e2 * v;
      const double r = M.nu_masses[nu_i] * M.inv_nu_q[iq];
      const double pert_scale = (r * r) / 2;
      const double ker = M.nu_int_kernel[iq];
      const double tmp = ker * (r0 + pert_scale * p0);
      drhonu = drhonu + tmp * iv;
      fnu = fnu + ker * (r1 + pert_scale * p1);
    }
  }
}

// ---------------------------------------------------------------- derivs
// camb/equations_galileon.f90:2098-2589.  ay / ayp are workspace columns.  Writes e.pig / e.pigdot during
// tight coupling exactly as the reference does (its value at a switch is the one left by the LAST call).
__device__ __forceinline__ void derivs(const DevModel* __restrict__ Mp, EV* __restrict__ ep, ThermoRows* __restrict__ tc, double tau,
                                    const double* __restrict__ ay_, double* __restrict__ ayp) {
  const InCol ay{ay_};
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
  double cs2, opacity, dopacity = 0;
#if GC_THERMO_CACHE
  thermo_cached(M, tc, tau, cs2, opacity, e.TightCoupling ? &dopacity : nullptr);
#else
  thermo(M, tau, cs2, opacity, e.TightCoupling ? &dopacity : nullptr);
#endif
  const double cothxor = 1.0 / tau;
#if GC_EARLY_GAL_ROWS == 1
  double gal_v0[5], gal_v1[5];
  int gal_i = 0;
  if (gal) gal_i = gal_rows_load(M, a, gal_v0, gal_v1);  // consumed after the hierarchy runs below
#elif GC_EARLY_GAL_ROWS == 2
  if (gal) {  // the two rows gal_HX() will read, asked into L2 now
    int i = (int)(log(a / M.gal_a0) * M.gal_lnq_inv);
    i = i < 0 ? 0 : i > M.ngal - 2 ? M.ngal - 2 : i;
    const double* r0 = M.gal + (size_t)i * 5;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(r0));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(r0 + 9));
  }
#endif
#if GC_EARLY_HIER
  // The equations of the high multipoles (l >= 3 of photons, polarisation, massless neutrinos; l = 1 and l >= 3 of every
  // momentum bin of the massive neutrinos; the perturbative massive-neutrino hierarchy) read nothing but y, k, the opacity
  // and tau: they are evaluated HERE, before the Einstein / Galileon part, while few registers are live -- whole runs of
  // multipoles per exposed memory latency instead of a latency per statement under the register pressure further down.
  // Every expression is the one of camb/equations_galileon.f90:2437-2589; the stores go to distinct elements of y'.
  {
    const bool hp = !e.no_phot && !e.TightCoupling;
    const int lmg = e.lmaxg, lme = e.lmaxgpol, lmr = e.lmaxnr;
    const int ng = (hp && lmg > 2) ? lmg - 2 : 0, ne = (hp && lme > 2) ? lme - 2 : 0;
    const int nr = (!e.no_nu && !e.high_ktau && lmr > 2) ? lmr - 2 : 0;
    auto fg = [&](int l, double ym, double y0, double yp) {
      return l < lmg ? (denlk(e, l) * ym - denlk2(e, l) * yp) - opacity * y0 : k * ym - (lmg + 1) * cothxor * y0 - opacity * y0;
    };
    auto fe = [&](int l, double ym, double y0, double yp) {
      return l < lme ? -opacity * y0 + (denlk(e, l) * ym - polfack(e, l) * yp)
                     : -opacity * y0 + k * e.poltruncfac * ym - (lme + 3) * cothxor * y0;
    };
    auto fr = [&](int l, double ym, double y0, double yp) {
      return l < lmr ? (denlk(e, l) * ym - denlk2(e, l) * yp) : k * ym - (lmr + 1) * cothxor * y0;
    };
    // one hierarchy after the other (issuing the first batches of all of them together measured 12 % SLOWER)
    hier_run(ay, ayp, e.g_ix + 3, 3, ng, -1, 0, fg);
    hier_run(ay, ayp, e.polind + 3, 3, ne, -1, 0, fe);
    hier_run(ay, ayp, e.r_ix + 3, 3, nr, -1, 0, fr);
  }
  if (M.Num_Nu_massive != 0) {
    for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
      if (e.MassiveNuApprox[nu_i]) continue;
      const int lm = e.lmaxnu_tau[nu_i];
      int ind = e.nu_ix[nu_i];
      for (int i = 1; i <= e.nq[nu_i]; i++) {
        const double aq = a * M.nu_masses[nu_i] * M.inv_nu_q[i - 1];
        const double v = rsqrt(1.0 + aq * aq);
        // l = 1 .. lm of this momentum bin; l = 0 and l = 2 need z and sigma and follow below.  With lm = 2 only l = 1.
        hier_run(ay, ayp, ind + 1, 1, lm == 2 ? 1 : lm, 2, 0, [&](int l, double ym, double y0, double yp) {
          return (l < lm || lm == 2) ? v * (denlk(e, l) * ym - denlk2(e, l) * yp) : k * v * ym - (lm + 1) * cothxor * y0;
        });
        ind = ind + lm + 1;
      }
    }
  }
#endif
  const double grhob_t = M.grhob * ia, grhoc_t = M.grhoc * ia, grhor_t = M.grhornomass * ia2, grhog_t = M.grhog * ia2;
  // massive-neutrino background at this a (shared by MassiveNuVars, GetdHdX and dtauda)
  double rhonu[GC_EV_NU], pnu[GC_EV_NU];
  for (int i = 0; i < M.n_eig; i++) Nu_background(M, a * M.nu_masses[i], rhonu[i], pnu[i]);
  double grhov_t = 0, xgal = 0, hbar = 0;
  // Galileon terms as linear forms in the perturbations (galileon_fast.cuh).  All of their coefficients are formed HERE,
  // from the background record F of this (a, hbar, x): F itself (45 doubles) is then dead, and only the 7 + 6 (+ 5 in
  // tight coupling) coefficients stay live through the hierarchy loops below instead of being spilled to local memory.
  double chi_p = 0, chi_d = 0, chi_g = 0, chi_e = 0, q_d = 0, q_p = 0, q_g = 0, gpresgal_t = 0;
  GalPhi2Lin D2{0, 0, 0, 0, 0, 0};
  GalPiLin P2{0, 0, 0, 0, 0};
  if (gal) {
#if GC_EARLY_GAL_ROWS == 1
    gal_rows_eval(M, a, gal_i, gal_v0, gal_v1, hbar, xgal);
#else
    gal_HX(M, a, hbar, xgal);
#endif
    double lam_nu = 0;
    for (int i = 0; i < M.n_eig; i++) lam_nu += M.grhormass[i] * pnu[i];
    const GalConst GC{M.c2, M.c3, M.c4, M.c5, M.cG, M.h0, M.orad};
    GalFast F;
    gal_fast_background(GC, a, hbar, xgal, lam_nu, k, F);
    gpresgal_t = F.gpresgal;
    chi_p = F.chi_p; chi_d = F.chi_d; chi_g = F.chi_g; chi_e = F.chi_e;
    q_d = F.q_d; q_p = F.q_p; q_g = F.q_g;
    gal_fast_dphisecond_lin(GC, F, D2);
    if (e.TightCoupling) gal_fast_Pigal_lin(GC, F, P2);
  } else {
    grhov_t = M.grhov * a2;
  }
  double gpres = 0;
  gpres = gpres + (grhog_t + grhor_t) / 3;
  if (gal)
    gpres = gpres + gpresgal_t;
  else
    gpres = gpres + grhov_t * (-1.0);
  double grho_matter = grhob_t + grhoc_t;
  double dgrho_matter = grhob_t * clxb + grhoc_t * clxc;
  double dgq = grhob_t * vb;
  double wnu_arr[GC_EV_NU];
  if (M.Num_Nu_massive > 0) {  // MassiveNuVars, :1209-1252
    for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
      const double grhormass_t = M.grhormass[nu_i] * ia2;
      const double irho = 1.0 / rhonu[nu_i];
      double clxnu, qnu;
      if (e.MassiveNuApprox[nu_i]) {
        clxnu = AT(ay, e.nu_ix[nu_i]);
        qnu = AT(ay, e.nu_ix[nu_i] + 2);
      } else {
#if GC_EARLY_HIER
        nu_moments_01(M, e, ay, a, nu_i, clxnu, qnu);
#else
        Nu_Integrate_L012(M, e, ay, a, nu_i, clxnu, qnu, nullptr, nullptr);
#endif
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
  double dgrho = dgrho_matter;
  double z = 0, dz, clxr, qr, pir, clxg, qg, pig = 0, dgrhogal_t;
  if (e.no_nu) {
    if (gal) {
      dgrhogal_t = (chi_p * dphiprime + chi_d * dphi + chi_g * dgrho + chi_e * etak);
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
        dgrhogal_t = (chi_p * dphiprime + chi_d * dphi + chi_g * dgrho + chi_e * etak);
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
    dgrhogal_t = (chi_p * dphiprime + chi_d * dphi + chi_g * dgrho + chi_e * etak);
    dgrho = dgrho + dgrhogal_t;
    dgq = dgq + (q_d * dphi + q_p * dphiprime + q_g * dgq);
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
    if (gal) dgs = dgs + (P2.d * dphi + P2.pi * dgs + P2.g * dgrho + P2.q * dgq + P2.e * etak);
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
#if GC_EARLY_FIELD_EQ
  // The Galileon field equation (camb/equations_galileon.f90:2516-2521) here, ahead of the multipole hierarchies: all of
  // its inputs are known, and its 6 + 12 operands need not stay live through the loops below (the stores are to distinct
  // elements of y', so their order is free).
  const double clxrdot = e.no_nu ? 0.0 : -k * (4.0 / 3.0 * z + qr);
  if (gal) {
    const double dotdeltaf = grhob_t * (clxbdot - 3 * adotoa * clxb) + grhoc_t * (clxcdot - 3 * adotoa * clxc) +
                             grhor_t * (clxrdot - 4 * adotoa * clxr) + grhog_t * (clxgdot - 4 * adotoa * clxg);
    AT(ayp, e.w_ix) = dphiprime;
    AT(ayp, e.w_ix + 1) = D2.p * dphiprime + D2.d * dphi + D2.g * dgrho + D2.q * dgq + D2.f * dotdeltaf + D2.e * etak;
  }
#endif
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
#if !GC_EARLY_HIER
#pragma unroll(kHierUnroll)
        for (int l = 3; l <= e.lmaxg - 1; l++) {
          ix = ix + 1;
          AT(ayp, ix) = (denlk(e, l) * AT(ay, ix - 1) - denlk2(e, l) * AT(ay, ix + 1)) - opacity * AT(ay, ix);
        }
        ix = ix + 1;
        AT(ayp, ix) = k * AT(ay, ix - 1) - (e.lmaxg + 1) * cothxor * AT(ay, ix) - opacity * AT(ay, ix);
#endif
      } else {
        pigdot = denlk(e, 2) * qg - opacity * (pig - polter) + 8.0 / 15.0 * k * sigma;
        AT(ayp, ix) = pigdot;
      }
      ix = e.polind + 2;
      if (e.lmaxgpol > 2) {
        AT(ayp, ix) = -opacity * (y_e2 - polter) - k / 3.0 * y_e3;
#if !GC_EARLY_HIER
#pragma unroll(kHierUnroll)
        for (int l = 3; l <= e.lmaxgpol - 1; l++) {
          ix = ix + 1;
          AT(ayp, ix) = -opacity * AT(ay, ix) + (denlk(e, l) * AT(ay, ix - 1) - polfack(e, l) * AT(ay, ix + 1));
        }
        ix = ix + 1;
        AT(ayp, ix) = -opacity * AT(ay, ix) + k * e.poltruncfac * AT(ay, ix - 1) - (e.lmaxgpol + 3) * cothxor * AT(ay, ix);
#endif
      } else {
        AT(ayp, ix) = -opacity * (AT(ay, ix) - polter);
      }
    }
  }
  // clxrdot is read uninitialised by the reference when no_nu_multpoles (:2517 vs :2482); pinned to 0 like
  // the oracle (the value of its defining expression under qr = -4z/3).
#if !GC_EARLY_FIELD_EQ
  double clxrdot = 0;
#endif
  if (!e.no_nu) {
#if !GC_EARLY_FIELD_EQ
    clxrdot = -k * (4.0 / 3.0 * z + qr);
#endif
    AT(ayp, e.r_ix) = clxrdot;
    AT(ayp, e.r_ix + 1) = denlk(e, 1) * clxr - denlk2(e, 1) * pir;
    if (e.high_ktau) {
      AT(ayp, e.r_ix + 2) = -3 * pir * cothxor - clxrdot;
    } else {
      int ix = e.r_ix + 2;
      if (e.lmaxnr > 2) {
        AT(ayp, ix) = denlk(e, 2) * qr - denlk2(e, 2) * y_r3 + 8.0 / 15.0 * k * sigma;
#if !GC_EARLY_HIER
#pragma unroll(kHierUnroll)
        for (int l = 3; l <= e.lmaxnr - 1; l++) {
          ix = ix + 1;
          AT(ayp, ix) = (denlk(e, l) * AT(ay, ix - 1) - denlk2(e, l) * AT(ay, ix + 1));
        }
        ix = ix + 1;
        AT(ayp, ix) = k * AT(ay, ix - 1) - (e.lmaxnr + 1) * cothxor * AT(ay, ix);
#endif
      } else {
        AT(ayp, ix) = denlk(e, 2) * qr + 8.0 / 15.0 * k * sigma;
      }
    }
  }
#if !GC_EARLY_FIELD_EQ
  if (gal) {
    const double dotdeltaf = grhob_t * (clxbdot - 3 * adotoa * clxb) + grhoc_t * (clxcdot - 3 * adotoa * clxc) +
                             grhor_t * (clxrdot - 4 * adotoa * clxr) + grhog_t * (clxgdot - 4 * adotoa * clxg);
    AT(ayp, e.w_ix) = dphiprime;
    AT(ayp, e.w_ix + 1) = D2.p * dphiprime + D2.d * dphi + D2.g * dgrho + D2.q * dgq + D2.f * dotdeltaf + D2.e * etak;
  }
#endif
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
#if GC_EARLY_HIER
      // l = 0 and l = 2 of every momentum bin (the other multipoles were formed ahead of the Einstein part): the loads of
      // all bins first
      const int lm = e.lmaxnu_tau[nu_i];
      double y1[GC_MAX_NQ], y2[GC_MAX_NQ], y3[GC_MAX_NQ];
#pragma unroll
      for (int i = 0; i < GC_MAX_NQ; i++) {
        const bool p = i < e.nq[nu_i];
        const int b = ind + i * (lm + 1);
        y1[i] = p ? AT(ay, b + 1) : 0.0;
        y2[i] = (p && lm == 2) ? AT(ay, b + 2) : 0.0;
        y3[i] = (p && lm != 2) ? AT(ay, b + 3) : 0.0;
      }
#pragma unroll
      for (int i = 0; i < GC_MAX_NQ; i++) {
        if (i < e.nq[nu_i]) {
          const int b = ind + i * (lm + 1);
          const double aq = a * M.nu_masses[nu_i] * M.inv_nu_q[i];
          const double v = rsqrt(1.0 + aq * aq);
          const double d0 = -k * (4.0 / 3.0 * z + v * y1[i]);
          AT(ayp, b) = d0;
          if (lm == 2)
            AT(ayp, b + 2) = -d0 - 3 * cothxor * y2[i];
          else
            AT(ayp, b + 2) = v * (denlk(e, 2) * y1[i] - denlk2(e, 2) * y3[i]) + k * 8.0 / 15.0 * sigma;
        }
      }
#else
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
#pragma unroll(kHierUnroll)
          for (int l = 3; l <= e.lmaxnu_tau[nu_i] - 1; l++) {
            ind = ind + 1;
            AT(ayp, ind) = v * (denlk(e, l) * AT(ay, ind - 1) - denlk2(e, l) * AT(ay, ind + 1));
          }
          ind = ind + 1;
          AT(ayp, ind) = k * v * AT(ay, ind - 1) - (e.lmaxnu_tau[nu_i] + 1) * cothxor * AT(ay, ind);
        }
        ind = ind + 1;
      }
#endif
    }
  }
  if (e.has_nu_rel) {
    int ind = e.nu_pert_ix;
    AT(ayp, ind) = +k * a2 * qr - k * AT(ay, ind + 1);
    int ind2 = e.r_ix;
#if GC_EARLY_HIER
    {  // l = 1 .. lmaxnu_pert: batches of 4 equations, their 2 x 6 loads first
      const int lm = e.lmaxnu_pert;
      for (int c = 0; c < lm; c += 4) {
        double p[6], r[6];
#pragma unroll
        for (int u = 0; u < 6; u++) {
          const bool in = c + u <= lm;  // the truncation equation reads nothing beyond its own element
          p[u] = in ? AT(ay, ind + c + u) : 0.0;
          r[u] = in ? AT(ay, ind2 + c + u) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int l = 1 + c + u;
          if (l < lm)
            AT(ayp, ind + l) = -a2 * (denlk(e, l) * r[u] - denlk2(e, l) * r[u + 2]) + (denlk(e, l) * p[u] - denlk2(e, l) * p[u + 2]);
          else if (l == lm)
            AT(ayp, ind + l) = k * (p[u] - a2 * r[u]) - (lm + 1) * cothxor * p[u + 1];
        }
      }
    }
#else
#pragma unroll(kHierUnroll)
    for (int l = 1; l <= e.lmaxnu_pert - 1; l++) {
      ind = ind + 1;
      ind2 = ind2 + 1;
      AT(ayp, ind) = -a2 * (denlk(e, l) * AT(ay, ind2 - 1) - denlk2(e, l) * AT(ay, ind2 + 1)) +
                     (denlk(e, l) * AT(ay, ind - 1) - denlk2(e, l) * AT(ay, ind + 1));
    }
    ind = ind + 1;
    ind2 = ind2 + 1;
    AT(ayp, ind) = k * (AT(ay, ind - 1) - a2 * AT(ay, ind2 - 1)) - (e.lmaxnu_pert + 1) * cothxor * AT(ay, ind);
#endif
  }
}

}  // namespace gc

// batch mode: ONE CTA of 7 warps per SM (255 registers, 170-194 KB of lane records); its warps run derivs() in lock-step
// (see the main loop), so an SM fetches one instruction stream.  Measured on 1024 distinct points (evolve ms): 2 CTAs x 4
// warps 5026, 2 x 3: 5597, 1 x 4: 6691, 1 x 6: 5220, 1 x 7: 4744, 1 x 8: 4811, 3 x 4 (168 registers): 6607, 4 x 4 (128): 9468.
#ifndef GC_EVOLVE_THREADS
#define GC_EVOLVE_THREADS 224
#endif
#ifndef GC_EVOLVE_MINBLOCKS
#define GC_EVOLVE_MINBLOCKS 1
#endif

namespace gc {

enum Phase : int {
  PH_NEXT_OUTPUT = 0,
  PH_EVOLVE,
  PH_DO_SWITCH,
  PH_DV_ENTER,
  PH_DV_LOOP,
  PH_DV_K1,
  PH_DV_STAGE,
  PH_DV_FINAL,
  PH_OUT_MARK,
  PH_OUT_REQ,
  PH_OUT_DONE,
#ifdef GC_WITH_TRANSFER
  PH_TF_CHECK,    // label 101 of camb/cmbmain.f90:1049: is a transfer redshift due before the next time step?
  PH_TF_ARRIVED,  // at (or within 1e-5 of) a transfer time: outtransf
#endif
  PH_DONE
};

#ifndef GC_HOT_IN_SMEM
#define GC_HOT_IN_SMEM 1
#endif
#define GC_HOT_BYTES (offsetof(DevModel, grhor))
static_assert(GC_HOT_BYTES % 8 == 0, "hot block must be a whole number of doubles");

struct Lane {
  EV e;
  const DevModel* M;
  double* ws;  // lane base of the warp workspace
  int NV;
  int kidx;
  int phase, stage, j, ind, after_switch;
  double tau, tauend, target;
  // dverk communication (c(13), c(14), c(17), c(18), c(19), c(20), c(21), c(23))
  double hmin, hmag, xtrial, htrial, est, c20, c21, nfail;
  // pending RHS request
  double req_tau;
  int req_in, req_out;
  bool want_derivs;
  // pending stage combination (row of the tableau, -1 = none) and its scale h/D; results of a final row
  int crow;
  double errn, ynorm_new;
  // max|y_i| of the accepted state (dverk's c(12) input), kept up to date by the final row
  double ynorm;
  bool ynorm_valid;
  int ysel;  // which physical column (0 or 9) currently holds y; the other one is the stage input w9
  // deferred output (see flush_output): the sources of output time out_j are formed later, together with the other
  // lanes' pending outputs, from y (stashed in column 10 by the next combine()), y' (column 11) and these scalars
  bool out_pending, out_stashed, flush_now, out_need_k1, stash_k1, waiting;
  int out_j, resume_phase;
#ifdef GC_WITH_TRANSFER
  // matter transfer functions: next transfer redshift (0-based), kind of the stop being integrated to (0 none, 1 in the
  // time-step loop: taken if within 1e-5, 2 before the loop / transfer-only wavenumber: always taken)
  short itf, tf_stop;
#endif
  double out_pig, out_pigdot;
  int status;
  double tol1;    // M.tol1 (read by every pass of advance(); the model record itself is a scattered global load)
  ThermoRows th;  // see thermo_cached()
  // work counters
  unsigned n_derivs, sum_n, n_out, n_shared;  // per lane: < 1e5 evaluations of < 200 equations
#if GC_HOT_IN_SMEM
  // private copy of the model's hot block (the prefix of DevModel that derivs() and its helpers read): in batch mode
  // the 32 lanes of a warp belong to 32 different models, so reading the constants from the DevModel array costs 32
  // cache lines per load; from here it is a conflict-free shared-memory access.
  double hot[GC_HOT_BYTES / 8];
#endif
};

// element i of column c of this lane: ws[(c*NV + i-1)*32]  (ws already points at the lane's slot)
// shared-memory stride between the Lane records of consecutive threads: sizeof(Lane) rounded up to an odd multiple of 8
#define GC_LANE_STRIDE ((((sizeof(gc::Lane) + 7) / 8) | 1) * 8)
// the CTA's lane records + 1 KB must fit the 227 KB of dynamic shared memory of an SM
static_assert((size_t)GC_EVOLVE_THREADS * GC_LANE_STRIDE + 1024 <= 227 * 1024, "Lane records too large for the SM");

// the lane's workspace base as a pointer the compiler knows to be global (it is read from a shared-memory record;
// without the hint every access through it is a generic LD/ST)
__device__ __forceinline__ double* ws_of(const Lane& L) {
  double* q = L.ws;
  __builtin_assume(__isGlobal(q));
  return q;
}
// logical column c of the lane (0 = y, 1..8 = k1..k8, 9 = stage input / trial solution).  y and w9 trade places
// when a trial step is accepted (L.ysel), so accepting a step costs no copy.
__device__ __forceinline__ int phys_col(const Lane& L, int c) {
  if (c == 0) return L.ysel ? 9 : 0;
  if (c == 9) return L.ysel ? 0 : 9;
  return c;
}
__device__ __forceinline__ double* col(const Lane& L, int c) {
  const int p = phys_col(L, c);
  double* q = L.ws + (size_t)p * L.NV * GC_LANES;
  __builtin_assume(__isGlobal(q));  // the pointer comes out of a shared-memory record: without the hint every access is a generic LD/ST
  return q;
}

// Verner 6(5) tableau of camb/subroutines.f90:930-1019 as rows over the derivative columns w1..w8:
//   rows 0..6  stage inputs   w9 = y + temp*(sum_j c_row[r][j]*w_j)            (:930-993)
//   row  7     trial solution w9 = y + temp*(sum_j c_row[7][j]*w_j)  and the error vector
//              w2 = (sum_j c_err[j]*w_j)/1398169080000                         (:995-1019)
// every sum runs left to right over ascending j with the reference's signs, so each partial sum equals the Fortran
// expression's.  c_rowmask[r] = the columns row r reads.
__constant__ double c_row[8][8] = {
    {233028180000.0, 0, 0, 0, 0, 0, 0, 0},
    {74569017600.0, 298276070400.0, 0, 0, 0, 0, 0, 0},
    {1165140900000.0, -3728450880000.0, 3495422700000.0, 0, 0, 0, 0, 0},
    {-3604654659375.0, 12816549900000.0, -9284716546875.0, 1237962206250.0, 0, 0, 0, 0},
    {3355605792000.0, -11185352640000.0, 9172628850000.0, -427218330000.0, 482505408000.0, 0, 0, 0},
    {-770204740536.0, 2311639545600.0, -1322092233000.0, -453006781920.0, 326875481856.0, 0, 0, 0},
    {2845924389000.0, -9754668000000.0, 7897110375000.0, -192082660000.0, 400298976000.0, 0, 201586000000.0, 0},
    {104862681000.0, 0, 545186250000.0, 446637345000.0, 188806464000.0, 0, 15076875000.0, 97599465000.0}};
__constant__ double c_err[8] = {8738556750.0,  0, 9735468750.0, -9709507500.0, 8582112000.0, 95329710000.0,
                                -15076875000.0, -97599465000.0};
__constant__ unsigned c_rowmask[8] = {0x01, 0x03, 0x07, 0x0f, 0x1f, 0x1f, 0x5f, 0xfd};

// Stage combination of a warp whose lanes all form the SAME tableau row (the normal state of an aligned warp): the
// columns the row reads are known per instantiation, the loop body is branch-free and small, and the compiler
// unrolls/pipelines it so that several batches of (NC+1) x U independent loads are in flight (U = 8 elements for the
// rows that read up to five columns, 4 for the last two).  Runs over the full batches below nend; the generic loop of combine() finishes the tail.
//   ROW0: w9 = y + (temp*w1)*c with the reference's association (camb/subroutines.f90:930), plus the copies a pending
//         output needs;  FIN: trial solution + error vector of the last row (:995-1019).
// loads of the derivative columns in the stage combinations: streamed once per row, never reused from L1 -> keep them
// out of it (cache-global); in the last row they are dead afterwards (cache-streaming = evict first)
#ifndef GC_LDK_MODE
#define GC_LDK_MODE 1
#endif
#if GC_LDK_MODE == 1
#define GC_LDK(fin, p) ((fin) ? __ldcs(p) : __ldcg(p))
#elif GC_LDK_MODE == 2
#define GC_LDK(fin, p) __ldcg(p)
#else
#define GC_LDK(fin, p) (*(p))
#endif

#if GC_STATE_CG
#define GC_LDY(p) __ldcg(p)
#else
#define GC_LDY(p) (*(p))
#endif

template <int NC, int U, bool ROW0, bool FIN, bool DUAL>
__device__ __forceinline__ int combine_uniform(int i, int nend, const double* __restrict__ y, double* __restrict__ w9,
                                               const double* const* kc, const double* cf, const double* ce, double temp,
                                               bool stash, bool stash_k1, double* __restrict__ ys,
                                               double* __restrict__ ks, double& errn, double& ynew, double* __restrict__ w8,
                                               const double* cf2) {
  auto load = [&](int i0, double (&yv)[U], double (&kv)[NC][U]) {
#pragma unroll
    for (int u = 0; u < U; u++) yv[u] = GC_LDY(&AT(y, i0 + u));
#pragma unroll
    for (int c = 0; c < NC; c++)
#pragma unroll
      for (int u = 0; u < U; u++) kv[c][u] = GC_LDK(FIN, &AT(kc[c], i0 + u));
  };
  auto comp = [&](int i0, const double (&yv)[U], const double (&kv)[NC][U]) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      double v;
      if (ROW0) {
        v = yv[u] + temp * kv[0][u] * 233028180000.0;
      } else {
        double acc = kv[0][u] * cf[0];
#pragma unroll
        for (int c = 1; c < NC; c++) acc = acc + kv[c][u] * cf[c];
        v = yv[u] + temp * acc;
      }
      AT(w9, i0 + u) = v;
      if (DUAL) {  // the row after this one reads the same columns: its stage input from the same loads (see GC_FUSE_ROW5)
        double acc2 = kv[0][u] * cf2[0];
#pragma unroll
        for (int c = 1; c < NC; c++) acc2 = acc2 + kv[c][u] * cf2[c];
        AT(w8, i0 + u) = yv[u] + temp * acc2;
      }
      if (ROW0) {
        if (stash) AT(ys, i0 + u) = yv[u];
        if (stash_k1) AT(ks, i0 + u) = kv[0][u];
      }
      if (FIN) {
        double a2 = kv[0][u] * ce[0];
#pragma unroll
        for (int c = 1; c < NC; c++) a2 = a2 + kv[c][u] * ce[c];
        const double w2 = a2 / 1398169080000.0;
        errn = fmax(errn, fabs(w2) / fmax(1.0, fabs(yv[u])));
        ynew = fmax(ynew, fabs(v));
      }
    }
  };
#if GC_COMBINE_PIPE
  if (NC >= GC_PIPE_MIN_NC && NC <= GC_PIPE_MAX_NC) {
    // two batches in flight: the loads of the next batch are issued before the arithmetic and the stores of the current
    // one, so the column stream never drains between batches (one exposed memory latency per pass, not one per batch)
    if (i + U - 1 > nend) return i;
    double yc[U], kcur[NC][U];
    load(i, yc, kcur);
    for (; i + 2 * U - 1 <= nend; i += U) {
      double yn[U], kn[NC][U];
      load(i + U, yn, kn);
      comp(i, yc, kcur);
#pragma unroll
      for (int u = 0; u < U; u++) {
        yc[u] = yn[u];
#pragma unroll
        for (int c = 0; c < NC; c++) kcur[c][u] = kn[c][u];
      }
    }
    comp(i, yc, kcur);
    return i + U;
  }
#endif
#pragma unroll(U == 4 ? 2 : 1)
  for (; i + U - 1 <= nend; i += U) {
    double yv[U], kv[NC][U];
    load(i, yv, kv);
    comp(i, yv, kv);
  }
  return i;
}

// elements per batch of the uniform loops for a row that reads NC columns
__host__ __device__ constexpr int combine_batch(int NC) {
#if GC_COMBINE_PIPE
  if (NC >= GC_PIPE_MIN_NC && NC <= GC_PIPE_MAX_NC) return NC <= 2 ? 8 : NC <= 5 ? 4 : 2;
  return NC >= 6 ? 4 : 8;
#else
  return NC >= GC_COMBINE_NC4 ? 4 : 8;
#endif
}

// Row R of the tableau through combine_uniform: the compact column list of the row (ascending j, as the sums run) and its
// coefficients are compile-time selections -- registers and constant-bank operands, no run-time-indexed local arrays.
// Columns: rows 0..5 read k1..k(R+1) (row 5: k1..k5), row 6 skips k6, row 7 (trial solution + error) skips k2.
template <int R>
__device__ __forceinline__ int combine_uniform_row(int i, int nend, const double* __restrict__ y, double* __restrict__ w9,
                                                   const double* const* kc, double temp, bool stash, bool stash_k1,
                                                   double* __restrict__ ys, double* __restrict__ ks, double& errn, double& ynew,
                                                   double* __restrict__ w8) {
  constexpr int NC = R == 0 ? 1 : R == 1 ? 2 : R == 2 ? 3 : R == 3 ? 4 : R <= 5 ? 5 : R == 6 ? 6 : 7;
  constexpr int U = combine_batch(NC);
  const double* kk[NC];
  double cfk[NC], cek[NC], cf2k[NC];
#pragma unroll
  for (int j = 0; j < NC; j++) {
    const int c = R == 6 ? (j < 5 ? j : 6) : R == 7 ? (j == 0 ? 0 : j + 1) : j;
    kk[j] = kc[c];
    cfk[j] = c_row[R][c];
    cek[j] = R == 7 ? c_err[c] : 0.0;
    cf2k[j] = R == 4 ? c_row[5][c] : 0.0;
  }
  return combine_uniform<NC, U, R == 0, R == 7, (GC_FUSE_ROW5 && R == 4)>(i, nend, y, w9, kk, cfk, cek, temp, stash, stash_k1, ys, ks,
                                                                          errn, ynew, w8, cf2k);
}

// ONE stage-combination routine for every row of the tableau, executed by the whole warp together: each lane brings
// its own row (L.crow), length and scale, so lanes that sit in different Runge-Kutta stages -- the normal state of
// a warp whose 32 lanes integrate 32 different cosmologies -- share one pass over the columns instead of running
// one specialised loop per stage present in the warp one after the other.  4 elements x 4 columns of independent
// loads are in flight per batch (the columns live in L2/HBM).
__device__ __forceinline__ void combine(Lane& L) {
  const bool on = L.crow >= 0;
  const int r = on ? L.crow : 0;
  const int n = on ? L.e.neq : 0;
  const unsigned cmask = on ? c_rowmask[r] : 0u;
  const int nmax = __reduce_max_sync(0xffffffffu, n);
  const unsigned umask = __reduce_or_sync(0xffffffffu, cmask);
  const bool fin = on && r == 7;
  const bool any_fin = __any_sync(0xffffffffu, fin);
  double cf[8], ce[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    cf[j] = c_row[r][j];
    ce[j] = fin ? c_err[j] : 0.0;
  }
  const double* __restrict__ y = col(L, 0);
  double* __restrict__ w9 = col(L, 9);
  const double* kc[8];
#pragma unroll
  for (int j = 0; j < 8; j++) kc[j] = ws_of(L) + (size_t)(j + 1) * L.NV * GC_LANES;
#if GC_K8_IN_K2
  kc[7] = kc[1];  // k8 lives in k2's column: row 6 is the last to read k2, row 7 the only one to read k8
#endif
  const double temp = L.htrial / 1398169080000.0;
  double errn = 0.0, ynew = 0.0;
  // GC_FUSE_ROW5: row 5 of the tableau (the input of stage 7) reads k1..k5 like row 4 and does not need k6, so the pass of
  // row 4 forms it as well, into column 8 (free: k8 lives in k2's column), and the state machine skips the pass of row 5
  const bool dual = GC_FUSE_ROW5 && on && r == 4;
  const bool any_dual = GC_FUSE_ROW5 && __any_sync(0xffffffffu, dual);
  double* __restrict__ w8 = ws_of(L) + (size_t)8 * L.NV * GC_LANES;
  // a lane whose last output is still pending keeps a copy of that y (this is the first pass over y since then)
  const bool stash = on && L.out_pending && !L.out_stashed;
  double* __restrict__ ys = ws_of(L) + (size_t)10 * L.NV * GC_LANES;
  const bool stash_k1 = stash && L.stash_k1;  // row 0 only: kv[0] is the k1 the output shares
  double* __restrict__ ks = ws_of(L) + (size_t)11 * L.NV * GC_LANES;
  // warp-uniform fast path: every lane forms the same row (the normal state of an aligned warp) -> no per-lane column
  // predicates; full batches below the shortest lane's length need no range predicates either
  const bool same_row = __all_sync(0xffffffffu, !on || cmask == umask);
  const int nmin = __reduce_min_sync(0xffffffffu, on ? n : 0x7fffffff);
  int i = 1;
  if (same_row && nmin >= 4 && nmin != 0x7fffffff) {
    if (on) {  // padding lanes of a short k group sit this out
      switch (r) {  // warp-uniform here
        case 0: i = combine_uniform_row<0>(i, nmin, y, w9, kc, temp, stash, stash_k1, ys, ks, errn, ynew, w8); break;
        case 1: i = combine_uniform_row<1>(i, nmin, y, w9, kc, temp, false, false, ys, ks, errn, ynew, w8); break;
        case 2: i = combine_uniform_row<2>(i, nmin, y, w9, kc, temp, false, false, ys, ks, errn, ynew, w8); break;
        case 3: i = combine_uniform_row<3>(i, nmin, y, w9, kc, temp, false, false, ys, ks, errn, ynew, w8); break;
        case 4: i = combine_uniform_row<4>(i, nmin, y, w9, kc, temp, false, false, ys, ks, errn, ynew, w8); break;
        case 5: i = combine_uniform_row<5>(i, nmin, y, w9, kc, temp, false, false, ys, ks, errn, ynew, w8); break;
        case 6: i = combine_uniform_row<6>(i, nmin, y, w9, kc, temp, false, false, ys, ks, errn, ynew, w8); break;
        default: i = combine_uniform_row<7>(i, nmin, y, w9, kc, temp, false, false, ys, ks, errn, ynew, w8); break;
      }
    }
    {  // first element the uniform loops did not cover (same for every lane)
      const int ub = combine_batch(__popc(umask));
      i = 1 + (nmin / ub) * ub;
    }
  }
  for (; i <= nmax; i += 4) {
    // all loads of the batch first (up to 4 + 8 x 4 independent loads in flight), then the arithmetic, then the stores
    double yv[4], kv[8][4];
    bool p[4];
    {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        p[u] = i + u <= n;
        yv[u] = p[u] ? GC_LDY(&AT(y, i + u)) : 0.0;
      }
#pragma unroll
      for (int c = 0; c < 8; c++) {
        if (umask & (1u << c)) {
          const bool mine = (cmask >> c) & 1u;
#pragma unroll
          for (int u = 0; u < 4; u++) kv[c][u] = (mine && p[u]) ? GC_LDK(false, &AT(kc[c], i + u)) : 0.0;
        } else {
#pragma unroll
          for (int u = 0; u < 4; u++) kv[c][u] = 0.0;
        }
      }
    }
    double a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 8; c++) {
      if (umask & (1u << c)) {
#pragma unroll
        for (int u = 0; u < 4; u++) a1[u] = a1[u] + kv[c][u] * cf[c];
      }
    }
    if (any_fin) {
#pragma unroll
      for (int c = 0; c < 8; c++)
#pragma unroll
        for (int u = 0; u < 4; u++) a2[u] = a2[u] + kv[c][u] * ce[c];
    }
    if (any_dual) {  // lanes in row 4 form row 5 as well (its columns are 0..4)
      double a3[4] = {0, 0, 0, 0};
#pragma unroll
      for (int c = 0; c < 5; c++)
#pragma unroll
        for (int u = 0; u < 4; u++) a3[u] = a3[u] + kv[c][u] * (dual ? c_row[5][c] : 0.0);
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (dual && p[u]) AT(w8, i + u) = yv[u] + temp * a3[u];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (p[u]) {
        // row 0 keeps the reference's association y + (temp*w1)*c  (camb/subroutines.f90:930)
        const double v = (r == 0) ? yv[u] + temp * kv[0][u] * 233028180000.0 : yv[u] + temp * a1[u];
        AT(w9, i + u) = v;
        if (stash) AT(ys, i + u) = yv[u];
        if (stash_k1) AT(ks, i + u) = kv[0][u];
        if (fin) {
          const double w2 = a2[u] / 1398169080000.0;
          errn = fmax(errn, fabs(w2) / fmax(1.0, fabs(yv[u])));
          ynew = fmax(ynew, fabs(v));
        }
      }
    }
  }
  if (fin) {
    L.errn = errn;
    L.ynorm_new = ynew;
  }
  if (stash) {
    L.out_stashed = true;
    L.stash_k1 = false;
  }
  L.crow = -1;
}

// combine() for warps that hold few work items (up to 16 models in flight: one k group = the low lanes of a warp, the
// rest padding): the idle lanes help.  The warp splits into groups of `gsz` lanes {o, o + nown, o + 2 nown, ...} around
// each owner lane o < nown; every lane of a group reads the owner's lane record from shared memory and forms the
// elements i = sub + 1, sub + 1 + gsz, ... of the owner's stage input -- each element by one lane, in the same
// left-to-right order, so the result is bit-identical to combine()'s.  Single model: 110 elements in one batch of 4
// per lane instead of 28 sequential batches on one lane.
__device__ __forceinline__ void combine_group(unsigned char* smem_base, int nown) {
  __syncwarp();  // the owners' records were written by their own lanes (advance)
  const int lane = threadIdx.x & 31;
  const int owner = lane % nown, sub = lane / nown, gsz = 32 / nown;
  Lane& O = *reinterpret_cast<Lane*>(smem_base + (size_t)((threadIdx.x & ~31) + owner) * GC_LANE_STRIDE);
  const bool on = O.crow >= 0;
  const int r = on ? O.crow : 0;
  const int n = on ? O.e.neq : 0;
  const unsigned cmask = on ? c_rowmask[r] : 0u;
  const bool fin = on && r == 7;
  const double* __restrict__ y = col(O, 0);
  double* __restrict__ w9 = col(O, 9);
  const double temp = O.htrial / 1398169080000.0;
  const bool stash = on && O.out_pending && !O.out_stashed;
  const bool stash_k1 = stash && O.stash_k1;
  double* __restrict__ ys = ws_of(O) + (size_t)10 * O.NV * GC_LANES;
  double* __restrict__ ks = ws_of(O) + (size_t)11 * O.NV * GC_LANES;
  double* __restrict__ w8 = ws_of(O) + (size_t)8 * O.NV * GC_LANES;
  const int nmax = __reduce_max_sync(0xffffffffu, n);
  double errn = 0.0, ynew = 0.0;
  for (int i0 = 1 + sub; i0 <= nmax; i0 += 4 * gsz) {
    double yv[4], kv[8][4];
    bool p[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      p[u] = i0 + u * gsz <= n;
      yv[u] = p[u] ? GC_LDY(&AT(y, i0 + u * gsz)) : 0.0;
    }
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const bool mine = (cmask >> c) & 1u;
      const double* kcol = ws_of(O) + (size_t)((GC_K8_IN_K2 && c == 7) ? 2 : c + 1) * O.NV * GC_LANES;
#pragma unroll
      for (int u = 0; u < 4; u++) kv[c][u] = (mine && p[u]) ? GC_LDK(false, &AT(kcol, i0 + u * gsz)) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (p[u]) {
        const int i = i0 + u * gsz;
        double a1 = 0.0;
#pragma unroll
        for (int c = 0; c < 8; c++) a1 = a1 + kv[c][u] * c_row[r][c];
        const double v = (r == 0) ? yv[u] + temp * kv[0][u] * 233028180000.0 : yv[u] + temp * a1;
        AT(w9, i) = v;
        if (GC_FUSE_ROW5 && r == 4) {  // row 5 from the same loads (see combine())
          double a3 = 0.0;
#pragma unroll
          for (int c = 0; c < 5; c++) a3 = a3 + kv[c][u] * c_row[5][c];
          AT(w8, i) = yv[u] + temp * a3;
        }
        if (stash) AT(ys, i) = yv[u];
        if (stash_k1) AT(ks, i) = kv[0][u];
        if (fin) {
          double a2 = 0.0;
#pragma unroll
          for (int c = 0; c < 8; c++) a2 = a2 + kv[c][u] * c_err[c];
          const double w2 = a2 / 1398169080000.0;
          errn = fmax(errn, fabs(w2) / fmax(1.0, fabs(yv[u])));
          ynew = fmax(ynew, fabs(v));
        }
      }
    }
  }
  // max over the lanes of a group (they differ in the bits above log2(nown))
  for (int o = nown; o < 32; o <<= 1) {
    errn = fmax(errn, __shfl_xor_sync(0xffffffffu, errn, o));
    ynew = fmax(ynew, __shfl_xor_sync(0xffffffffu, ynew, o));
  }
  __syncwarp();  // every helper has read the owner's record before the owner changes it
  if (sub == 0 && on) {
    if (fin) {
      O.errn = errn;
      O.ynorm_new = ynew;
    }
    if (stash) {
      O.out_stashed = true;
      O.stash_k1 = false;
    }
    O.crow = -1;
  }
  __syncwarp();
}

// Deferred output_sources(): called from the main loop when ANY lane of the warp must have its pending output out
// of the way (it reached its next output time, a switch or the end); every lane with a pending output joins, so a
// warp whose lanes reach their output times in different iterations still runs the ~2 kFLOP source evaluation about
// once per output interval instead of once per lane.  Nothing the evaluation reads changes in between: y is
// stashed, y' sits in its own column, the layout descriptor only changes at switches (which flush first), and the
// tight-coupling scalars derivs() leaves in EV (camb/equations_galileon.f90:2407,2417) are saved and put back.
__device__ __forceinline__ void flush_output(Lane& L) {
  const DevModel& M = *L.M;
  const int S = M.SourceNum;
  const double pig = L.e.pig, pigdot = L.e.pigdot;
  L.e.pig = L.out_pig;
  L.e.pigdot = L.out_pigdot;
  double src[3] = {0, 0, 0};
  // not yet copied aside (the flush came before the step's first combine): y and the shared k1 are still in place
  const double* y = L.out_stashed ? L.ws + (size_t)10 * L.NV * GC_LANES : col(L, 0);
  const double* yp = (!L.out_stashed && L.stash_k1) ? col(L, 1) : col(L, 11);
  L.stash_k1 = false;
  output_sources(L.M, &L.e, y, yp, M.ts + (size_t)(L.out_j - 1) * 8, M.ts[(size_t)(L.out_j - 1) * 8], src);
  for (int s = 0; s < S; s++) M.Src[((size_t)(L.out_j - 1) * S + s) * M.nk + L.kidx] = src[s];
  L.e.pig = pig;
  L.e.pigdot = pigdot;
  L.out_pending = false;
}

struct Switches {
  double ktau, no_nu, no_phot, nu_massless, nu_nonrel, nu_massive, next;
};

// camb/equations_galileon.f90:328-364
__device__ inline void compute_switches(const Lane& L, Switches& s) {
  const DevModel& M = *L.M;
  const EV& e = L.e;
  const double noSwitch = M.tau0 + 1;
  s.ktau = s.no_nu = s.no_phot = s.nu_massless = s.nu_nonrel = s.nu_massive = noSwitch;
  if (!e.high_ktau && !e.no_nu) s.ktau = max(20, e.lmaxnr - 4) / e.k;
  if (M.Num_Nu_massive != 0) {
    for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
      if (e.nq[nu_i] != M.nqmax)
        s.nu_massless = fmin(s.nu_massless, M.nu_tau_notmassless[next_nu_nq(M, e.nq[nu_i]) - 1][nu_i]);
      else if (!e.nu_nonrel[nu_i])
        s.nu_nonrel = fmin(M.nu_tau_nonrelativistic[nu_i], s.nu_nonrel);
      else if (M.NuMethod == 1 && !e.MassiveNuApprox[nu_i])
        s.nu_massive = fmin(s.nu_massive, e.MassiveNuApproxTime[nu_i]);
    }
  }
  if (M.DoLateRadTruncation) {
    if (!e.no_nu) s.no_nu = fmax(15 / e.k * M.AccuracyBoost, fmin(M.taurend, M.matter_verydom_tau));
    if (!e.no_phot && (e.k > GC_SP(0.03) * M.AccuracyBoost)) s.no_phot = fmax(15 / e.k, M.taurend) * M.AccuracyBoost;
  }
  double n = fmin(s.ktau, s.nu_massless);
  n = fmin(n, e.TightSwitchoffTime);
  n = fmin(n, s.nu_massive);
  n = fmin(n, s.no_nu);
  n = fmin(n, s.no_phot);
  n = fmin(n, s.nu_nonrel);
  n = fmin(n, noSwitch);
  s.next = n;
}

// the switch bodies of camb/equations_galileon.f90:370-463 (y in column 0, scratch column 9)
__device__ inline void do_switch(Lane& L) {
  const DevModel& M = *L.M;
  EV& e = L.e;
  Switches s;
  compute_switches(L, s);
  const double next_switch = s.next;
  const double noSwitch = M.tau0 + 1;
  double* y = col(L, 0);
  double* yout = col(L, 9);
  EV o = e;
  bool relayout = false;
  bool tight = false;
  if (next_switch == e.TightSwitchoffTime) {
    o.TightCoupling = false;
    o.TightSwitchoffTime = noSwitch;
    relayout = true;
    tight = true;
    L.ind = 1;
  } else if (next_switch == s.ktau) {
    o.high_ktau = true;
    relayout = true;
  } else if (next_switch == s.nu_massless) {
    for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
      if (e.nq[nu_i] != M.nqmax && next_switch == M.nu_tau_notmassless[next_nu_nq(M, e.nq[nu_i]) - 1][nu_i]) {
        o.nq[nu_i] = next_nu_nq(M, e.nq[nu_i]);
        relayout = true;
        break;
      }
    }
  } else if (next_switch == s.nu_nonrel) {
    for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
      if (!e.nu_nonrel[nu_i] && next_switch == M.nu_tau_nonrelativistic[nu_i]) {
        o.nu_nonrel[nu_i] = true;
        relayout = true;
        break;
      }
    }
  } else if (next_switch == s.nu_massive) {
    for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
      if (!e.MassiveNuApprox[nu_i] && next_switch == e.MassiveNuApproxTime[nu_i]) {
        // SwitchToMassiveNuApprox, :949-990
        const double a = AT(y, 1), a2 = a * a;
        o.MassiveNuApprox[nu_i] = true;
        SetupScalarArrayIndices(M, o, nullptr);
        CopyScalarVariableArray(M, y, yout, e, o);
        double rhonu, pnu, clxnu, qnu, dpnu, pinu;
        Nu_background(M, a * M.nu_masses[nu_i], rhonu, pnu);
        Nu_Integrate_L012(M, e, y, a, nu_i, clxnu, qnu, &dpnu, &pinu);
        dpnu = dpnu / rhonu;
        qnu = qnu / rhonu;
        clxnu = clxnu / rhonu;
        pinu = pinu / rhonu;
        AT(yout, o.nu_ix[nu_i]) = clxnu;
        AT(yout, o.nu_ix[nu_i] + 1) = dpnu;
        AT(yout, o.nu_ix[nu_i] + 2) = qnu;
        AT(yout, o.nu_ix[nu_i] + 3) = pinu;
        Nu_Intvsq(M, e, y, a, nu_i, o.G11[nu_i], o.G30[nu_i]);
        o.G11[nu_i] = o.G11[nu_i] * a2 * a / rhonu;
        o.G30[nu_i] = o.G30[nu_i] * a2 * a / rhonu;
        e = o;
        for (int i = 1; i <= e.nvar; i++) AT(y, i) = AT(yout, i);
        break;
      }
    }
  } else if (next_switch == s.no_nu) {
    L.ind = 1;
    o.no_nu = true;
    for (int i = 0; i < M.n_eig; i++) o.nq[i] = M.nqmax;
    relayout = true;
  } else if (next_switch == s.no_phot) {
    L.ind = 1;
    o.no_phot = true;
    relayout = true;
  }
  if (relayout) {
    SetupScalarArrayIndices(M, o, nullptr);
    CopyScalarVariableArray(M, y, yout, e, o);
    e = o;
    for (int i = 1; i <= e.nvar; i++) AT(y, i) = AT(yout, i);
  }
  if (tight) {
    AT(y, e.g_ix + 2) = e.pig;
    double cs2, opacity, dopacity;
    thermo(M, L.tau, cs2, opacity, &dopacity);
    const double op2 = opacity * opacity;
    AT(y, e.g_ix + 3) = (3.0 / 7.0) * AT(y, e.g_ix + 2) * (e.k / opacity) * (1.0 + dopacity / op2) +
                        (3.0 / 7.0) * e.pigdot * (e.k / op2) * (-1.0);
    AT(y, e.polind + 2) = e.pig / 4 + e.pigdot * (1.0 / opacity) * (-5.0 / 8.0 - (25.0 / 16.0) * dopacity / op2) +
                          e.pig * ((e.k / opacity) * (e.k / opacity)) * (-5.0 / 56.0);
    AT(y, e.polind + 3) = (3.0 / 7.0) * (e.k / opacity) * AT(y, e.polind + 2) * (1.0 + dopacity / op2) +
                          (3.0 / 7.0) * (e.k / op2) * ((e.pigdot / 4.0) * (1.0 + (5.0 / 2.0) * dopacity / op2)) * (-1.0);
  }
}

// Run the lane's control flow until it needs derivs() (returns with L.req_* set) or finishes.
__device__ __forceinline__ void advance(Lane& L, int wphase, bool align) {
  const DevModel& M = *L.M;
  EV& e = L.e;
  const int S = M.SourceNum;
  const double tol = L.tol1;
  for (;;) {
    switch (L.phase) {
      case PH_NEXT_OUTPUT: {
#ifdef GC_WITH_TRANSFER
        if (e.TransferOnly) {  // GetTransfer, camb/cmbmain.f90:1195-1199: through the transfer times only
          if (L.itf >= M.ntf) {
            L.phase = PH_DONE;
            return;
          }
          L.tauend = M.tautf[L.itf];
          L.tf_stop = 2;
          L.phase = PH_EVOLVE;
          break;
        }
        if (M.WantTransfer && L.j == 1 && L.itf < M.ntf && M.ts[8] > M.tautf[L.itf]) {  // :1022-1029
          L.tauend = M.tautf[L.itf];
          L.tf_stop = 2;
          L.phase = PH_EVOLVE;
          break;
        }
#endif
        const int jn = L.j + 1;
        const bool stops = jn > M.nt;
        double tauend = 0;
        bool zero = false;
        if (!stops) {
          tauend = M.ts[(size_t)(jn - 1) * 8];
          zero = (e.q * tauend > M.max_etak_scalar && tauend > M.taurend) && !M.WantLateTime;  // camb/cmbmain.f90:1034-1036
#ifdef GC_WITH_TRANSFER
          zero = zero && (!M.WantTransfer || L.tau > M.tautf[M.ntf - 1]);
#endif
        }
        if ((stops || zero) && L.out_need_k1) {  // no further step from this state: evaluate y' for the output now
          L.resume_phase = PH_NEXT_OUTPUT;
          L.phase = PH_OUT_REQ;
          break;
        }
        if (stops) {
          if (L.out_pending) {  // the last output must be out before the lane retires
            L.flush_now = true;
            return;
          }
          L.j = jn;
          L.phase = PH_DONE;
          return;
        }
        L.j = jn;
        if (zero) {
          for (int s = 0; s < S; s++) M.Src[((size_t)(L.j - 1) * S + s) * M.nk + L.kidx] = 0;
          break;
        }
        L.tauend = tauend;
        L.phase = PH_EVOLVE;
        break;
      }
      case PH_EVOLVE: {  // camb/equations_galileon.f90:314-469, one pass of the (tail-)recursion
        Switches s;
        compute_switches(L, s);
        const double smallTime = fmin(L.tau, 1 / e.k) / 100;
        if (s.next < L.tauend) {
          if (s.next > L.tau + smallTime) {
            L.target = s.next;
            L.after_switch = 1;
            L.phase = PH_DV_ENTER;
          } else {
            L.phase = PH_DO_SWITCH;
          }
        } else {
          L.target = L.tauend;
          L.after_switch = 0;
          L.phase = PH_DV_ENTER;
        }
        break;
      }
      case PH_DO_SWITCH: {
        if (L.out_need_k1) {  // the next k1 will be taken with the new layout: evaluate y' for the output now
          L.resume_phase = PH_DO_SWITCH;
          L.phase = PH_OUT_REQ;
          break;
        }
        if (L.out_pending) {  // the switch re-lays y out and changes the descriptor the pending output needs
          L.flush_now = true;
          return;
        }
        do_switch(L);
        L.ynorm_valid = false;
        L.phase = PH_EVOLVE;
        break;
      }
      case PH_DV_ENTER: {  // camb/subroutines.f90:700-770 (entry logic of dverk)
        if (L.ind == 3) {
          if (L.c21 != 0.0 && (L.tau != L.c20 || L.target == L.c20)) {
            L.status = 101;
            L.phase = PH_DONE;
            return;
          }
          L.c21 = 0.0;
        } else {  // ind = 1
          L.c20 = L.tau;
          L.c21 = 0.0;
          L.nfail = 0.0;
          L.hmag = 0.0;
          L.est = 0.0;
        }
        L.phase = PH_DV_LOOP;
        break;
      }
      case PH_DV_LOOP: {
        // Warp alignment: every trial step is 8 RHS evaluations, so lanes that all start their steps in rounds with
        // wphase == 0 sit in the same Runge-Kutta stage in every round -- same tableau row in combine(), same
        // branch of this state machine -- no matter how many steps each lane needs.  A lane that fell out of step
        // (explicit output evaluation, rejected step) sits out until the next such round.
        if (align && wphase != 0) {
          L.waiting = true;
          return;
        }
        if (L.ind != 6) {
          L.req_tau = L.tau;
          L.req_in = 0;
          L.req_out = 1;
          L.want_derivs = true;
          L.phase = PH_DV_K1;
          return;
        }
        L.phase = PH_DV_K1;  // rejected step: k1 is still valid (camb/subroutines.f90:1095-1110)
        if (align) {
          L.waiting = true;
          return;
        }
        break;
      }
      case PH_DV_K1: {  // step-size selection, camb/subroutines.f90:800-925, then stage 2 input
        if (L.out_need_k1) {
          // the k1 just evaluated at (tau, y) IS the y' that output() evaluates at :1297 -- same state, same
          // descriptor -- so the output shares it; combine() copies it (and y) aside for flush_output()
          L.out_pig = e.pig;
          L.out_pigdot = e.pigdot;
          L.out_need_k1 = false;
          L.stash_k1 = true;
          L.n_shared++;
          L.sum_n += e.neq;
        }
        const int n = e.neq;
        const double x = L.tau, xend = L.target;
        double temp = L.ynorm;
        if (!L.ynorm_valid) {  // first step after initial() or a switch; otherwise the final row keeps it current
          const double* y = col(L, 0);
          temp = 0.0;
          for (int i = 1; i <= n; i++) temp = fmax(temp, fabs(AT(y, i)));
          L.ynorm = temp;
          L.ynorm_valid = true;
        }
        const double c12 = fmin(temp, 1.0);
        L.hmin = 10.0 * fmax(1.e-35, 1.3877787807814457e-17 * fmax(c12 / tol, fabs(x)));  // rreb = 2^-56
        const double hmax = 2.0;
        if (L.hmin > hmax) {
          L.status = 102;
          L.phase = PH_DONE;
          return;
        }
        if (L.ind <= 2) {
          L.hmag = hmax * pow(tol, 1.0 / 6.0);
        } else if (L.nfail <= 1.0) {
          temp = 2.0 * L.hmag;
          const double r = 2.0 / .9;
          if (tol < (r * r * r * r * r * r) * L.est) temp = .9 * pow(tol / L.est, 1.0 / 6.0) * L.hmag;
          L.hmag = fmax(temp, .5 * L.hmag);
        } else {
          L.hmag = .5 * L.hmag;
        }
        L.hmag = fmin(L.hmag, hmax);
        L.hmag = fmax(L.hmag, L.hmin);
        if (L.hmag >= fabs(xend - x)) {
          L.hmag = fabs(xend - x);
          L.xtrial = xend;
        } else {
          L.hmag = fmin(L.hmag, .5 * fabs(xend - x));
          L.xtrial = x + copysign(L.hmag, xend - x);
        }
        L.htrial = L.xtrial - x;
        L.crow = 0;
        L.req_tau = x + L.htrial / 6.0;
        L.req_in = 9;
        L.req_out = 2;
        L.want_derivs = true;
        L.stage = 2;
        L.phase = PH_DV_STAGE;
        return;
      }
      case PH_DV_STAGE: {
        const int s = L.stage;  // the stage whose derivative has just been written (2..8)
        const double x = L.tau;
        if (s < 8) {
          // input of stage s+1 : row s-1 of the tableau, formed by the warp-wide combine() -- except the input of stage 7
          // (row 5), which the pass of row 4 has already left in column 8 (GC_FUSE_ROW5)
          L.crow = (GC_FUSE_ROW5 && s == 6) ? -1 : s - 1;
          double frac;
          switch (s) {
            case 2: frac = 4.0 / 15.0; break;
            case 3: frac = 2.0 / 3.0; break;
            case 4: frac = 5.0 / 6.0; break;
            case 5: frac = 1.0; break;
            case 6: frac = 1.0 / 15.0; break;
            default: frac = 1.0; break;
          }
          if (s == 5 || s == 7)
            L.req_tau = x + L.htrial;
          else if (s == 6)
            L.req_tau = x + L.htrial / 15.0;
          else
            L.req_tau = x + L.htrial * frac;
          L.req_in = (GC_FUSE_ROW5 && s == 6) ? 8 : 9;
          L.req_out = (GC_K8_IN_K2 && s == 7) ? 2 : s + 1;
          L.want_derivs = true;
          L.stage = s + 1;
          return;
        }
        // s == 8: trial solution and error vector (row 7), then PH_DV_FINAL
        L.crow = 7;
        L.phase = PH_DV_FINAL;
        return;
      }
      case PH_DV_FINAL: {  // accept/reject, camb/subroutines.f90:1021-1110
        const int n = e.neq;
        L.est = L.errn * L.hmag * 1.0;
        L.ind = 5;
        if (L.est > tol) L.ind = 6;
        if (L.ind != 6) {
          L.tau = L.xtrial;
          L.ysel ^= 1;  // y <- w9 : the columns trade places
          L.ynorm = L.ynorm_new;
          L.nfail = 0.0;
          if (L.tau == L.target) {
            L.ind = 3;
            L.c20 = L.target;
            L.c21 = 1.0;
#ifdef GC_WITH_TRANSFER
            L.phase = L.after_switch ? PH_DO_SWITCH : (L.tf_stop ? PH_TF_ARRIVED : PH_OUT_MARK);
#else
            L.phase = L.after_switch ? PH_DO_SWITCH : PH_OUT_MARK;
#endif
            break;
          }
        } else {
          L.nfail = L.nfail + 1.0;
          if (!(L.hmag > L.hmin)) {
            L.status = 4;  // dverk ind=-3 -> error_evolution (camb/equations_galileon.f90:285-292)
            L.phase = PH_DONE;
            return;
          }
        }
        L.phase = PH_DV_LOOP;
        break;
      }
      case PH_OUT_MARK: {  // an output time has been reached (camb/cmbmain.f90:1038 -> output)
        if (L.out_pending) {  // one pending output per lane
          L.flush_now = true;
          return;
        }
        L.out_pending = true;
        L.out_stashed = false;
        L.out_need_k1 = true;
        L.out_j = L.j;
        L.n_out++;
#ifdef GC_WITH_TRANSFER
        L.phase = (M.WantTransfer && L.itf < M.ntf) ? PH_TF_CHECK : PH_NEXT_OUTPUT;
#else
        L.phase = PH_NEXT_OUTPUT;
#endif
        break;
      }
#ifdef GC_WITH_TRANSFER
      case PH_TF_CHECK: {  // camb/cmbmain.f90:1049-1056
        if (!(M.WantTransfer && L.itf < M.ntf)) {
          L.phase = PH_NEXT_OUTPUT;
          break;
        }
        const double ttf = M.tautf[L.itf];
        if (L.j < M.nt && M.ts[(size_t)(L.j - 1) * 8] < ttf && M.ts[(size_t)L.j * 8] > ttf) {
          L.tauend = ttf;
          L.tf_stop = 1;
          L.phase = PH_EVOLVE;
          break;
        }
        L.tf_stop = 1;
        L.phase = PH_TF_ARRIVED;
        break;
      }
      case PH_TF_ARRIVED: {  // camb/cmbmain.f90:1026, :1059-1066, :1198 -> outtransf
        const int kind = L.tf_stop;
        L.tf_stop = 0;
        bool again = false;
        if (kind == 2 || fabs(L.tau - M.tautf[L.itf]) < 1.e-5) {
          transfer_outputs(L.M, &e, col(L, 0), L.tau, M.Transfer + ((size_t)L.itf * (M.nk + M.nk_tr) + L.kidx) * GC_TRANSFER_MAX);
          L.n_derivs++;  // the derivs call of outtransf, camb/equations_galileon.f90:2090
          L.sum_n += e.neq;
          L.itf++;
          if (kind == 1 && L.j < M.nt && L.itf < M.ntf && M.ts[(size_t)L.j * 8] > M.tautf[L.itf]) again = true;
        }
        L.phase = again ? PH_TF_CHECK : PH_NEXT_OUTPUT;
        break;
      }
#endif
      case PH_OUT_REQ: {  // explicit y' evaluation, camb/equations_galileon.f90:1296-1297
        double* yp = col(L, 11);
        for (int i = 1; i <= e.nvar; i++) AT(yp, i) = 0;
        L.req_tau = L.tau;
        L.req_in = 0;
        L.req_out = 11;
        L.want_derivs = true;
        L.phase = PH_OUT_DONE;
        return;
      }
      case PH_OUT_DONE: {  // the sources themselves are formed by flush_output()
        L.out_pig = e.pig;
        L.out_pigdot = e.pigdot;
        L.out_need_k1 = false;
        L.phase = L.resume_phase;
        break;
      }
      default:
        return;
    }
  }
}

// items[i] = (model index, k index); workspace = nwarps * GC_NCOL * NV * 32 doubles.
__global__ void __launch_bounds__(GC_EVOLVE_THREADS, GC_EVOLVE_MINBLOCKS) evolve_batch_kernel(const DevModel* __restrict__ models, const int2* __restrict__ items,
                                                             int nitems, double* __restrict__ workspace, int NV,
                                                             int* __restrict__ status,
                                                             unsigned long long* __restrict__ counters) {
  // The per-lane control state (layout descriptor EV, dverk communication, pending request) lives in SHARED memory:
  // as a local variable it is a 1.2 KB stack frame per thread (470 KB per SM at 12 warps), which thrashes L1 and
  // turned every field access into a long-scoreboard stall.  Stride = odd multiple of 8 B -> conflict-free.
  extern __shared__ unsigned char gc_smem_raw[];
  Lane& L = *reinterpret_cast<Lane*>(gc_smem_raw + (size_t)threadIdx.x * GC_LANE_STRIDE);
  // Persistent CTAs: a CTA takes blocks of blockDim.x items off a counter, in the launcher's expensive-first order, until
  // none is left.  Its warps keep ONE workspace slot each for the whole launch (slot = resident warp, not item), so the
  // integrator state of everything resident is a compact range at the start of the workspace -- which the launcher
  // can name in an L2 access-policy window.
  __shared__ int s_block;
  const int nblocks = (nitems + (int)blockDim.x - 1) / (int)blockDim.x;
  const int lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // workspace slot
  for (;;) {
  __syncthreads();  // every warp is done with the previous block (and has read s_block)
  if (threadIdx.x == 0) s_block = (int)atomicAdd(&counters[16], 1ULL);
  __syncthreads();
  const int block = s_block;
  if (block >= nblocks) break;
  const int tid = block * blockDim.x + threadIdx.x;
  const bool have = tid < nitems && items[tid < nitems ? tid : 0].x >= 0;  // padding items carry model = -1
  L.phase = PH_DONE;
  L.status = 0;
  L.n_derivs = L.sum_n = L.n_out = L.n_shared = 0;
  L.want_derivs = false;
  L.out_pending = L.out_stashed = L.flush_now = L.out_need_k1 = L.stash_k1 = L.waiting = false;
  L.resume_phase = PH_DONE;
  L.crow = -1;
#ifdef GC_WITH_TRANSFER
  L.itf = 0;
  L.tf_stop = 0;
  L.e.TransferOnly = false;
#endif
  L.ysel = 0;
  L.ynorm_valid = false;
  L.ynorm = L.ynorm_new = L.errn = 0;
  L.NV = NV;
  L.ws = workspace;
  L.e.neq = 0;
  if (have) {
    const int2 it = items[tid];
    L.M = models + it.x;
    const DevModel& M = *L.M;
#if GC_HOT_IN_SMEM
    {
      const double* src = reinterpret_cast<const double*>(L.M);
      for (int i = 0; i < (int)(GC_HOT_BYTES / 8); i++) L.hot[i] = src[i];
    }
#endif
    L.kidx = it.y;
    L.tol1 = M.tol1;
    L.th.i = 0;
    L.NV = NV;
    L.ws = workspace + (size_t)gwarp * GC_NCOL * NV * GC_LANES + lane;
    // DoSourcek, camb/cmbmain.f90:662-689
#ifdef GC_WITH_TRANSFER
    // k indices beyond the source grid are the transfer-only wavenumbers (:1164-1177)
    L.e.TransferOnly = it.y >= M.nk;
    L.e.q = L.e.TransferOnly ? M.k_tr[it.y - M.nk] : M.k[it.y];
#else
    L.e.q = M.k[it.y];
#endif
    L.e.pig = 0;
    L.e.pigdot = 0;
    for (int i = 0; i < GC_EV_NU; i++) {
      L.e.nu_ix[i] = 0;
      L.e.nq[i] = 0;
      L.e.lmaxnu_tau[i] = 0;
      L.e.nu_nonrel[i] = false;
      L.e.G11[i] = L.e.G30[i] = L.e.MassiveNuApproxTime[i] = 0;
    }
    L.e.nu_pert_ix = 0;
    L.e.lmaxnu_pert = 0;
    L.e.polind = 0;
    GetNumEqns(M, L.e);
    if (L.e.nvar > NV) {
      L.status = 103;
    } else {
      // GetTauStart, camb/cmbmain.f90:634-660
      double taustart = 0.001 / L.e.q;
      taustart = fmin(taustart, 0.1);
      if (M.Num_Nu_massive > 0) {
        double mx = 0;
        for (int i = 0; i < M.n_eig; i++) mx = fmax(mx, M.nu_masses[i]);
        taustart = fmin(taustart, 1.e-3 / mx / M.adotrad);
      }
      // w = 0 (camb/cmbmain.f90:946)
      for (int c = 1; c < GC_NCOL; c++) {
        double* p = col(L, c);
        for (int i = 1; i <= L.e.nvar; i++) AT(p, i) = 0;
      }
      initial(M, L.e, col(L, 0), taustart);
      L.tau = taustart;
      L.ind = 1;
      L.j = 1;
      L.c20 = 0;
      L.c21 = 0;
      L.nfail = 0;
      L.hmag = 0;
      L.est = 0;
      L.hmin = 0;
      // Src(:,:,1) = 0 (never written by the time loop, camb/cmbmain.f90:1031)
#ifdef GC_WITH_TRANSFER
      if (!L.e.TransferOnly)
#endif
        for (int s = 0; s < M.SourceNum; s++) M.Src[(size_t)s * M.nk + L.kidx] = 0;
      L.phase = PH_NEXT_OUTPUT;
    }
  }
  // All warps of the CTA run the RHS evaluation together: the kernel's hot loop is ~8k straight-line
  // instructions (far beyond the 32 KB L1.5 instruction cache), so warps that reach derivs() at the same time
  // share one instruction stream from L2 instead of fetching it once per warp.
#ifdef GC_PROFILE_PHASES
  long long pw[5] = {0, 0, 0, 0, 0};
#endif
  const unsigned have_mask = __ballot_sync(0xffffffffu, have);
  const bool align = __popc(have_mask) > 1;
  // few items in this warp (they are its low lanes): the padding lanes help in combine_group()
  int nown = 0;
  if (have_mask && __popc(have_mask) <= 16 && have_mask == (1u << __popc(have_mask)) - 1u) {
    nown = 1;
    while (nown < __popc(have_mask)) nown <<= 1;
  }
  int wphase = 0;
  for (;; wphase = (wphase + 1) & 7) {
    // control flow up to the next RHS request.  A lane normally needs one pass: advance() leaves a stage
    // combination (and the request that follows it) pending, combine() forms it for the whole warp.  After the last
    // stage the accept/reject logic needs the combination's result first, hence a second pass; a lane that has to
    // wait for flush_output() takes one more.
    L.waiting = false;
#ifdef GC_PROFILE_PHASES
    long long t0 = clock64(), t_adv = 0, t_flush = 0, t_comb = 0;
#endif
    for (;;) {
      const bool run = !L.want_derivs && L.crow < 0 && L.phase != PH_DONE && !L.waiting;
      if (!__any_sync(0xffffffffu, run)) break;
#ifdef GC_PROFILE_PHASES
      long long ta = clock64();
#endif
      if (run) advance(L, wphase, align);
#ifdef GC_PROFILE_PHASES
      __syncwarp();
      long long tb = clock64();
      t_adv += tb - ta;
#endif
      if (__any_sync(0xffffffffu, L.flush_now)) {
        if (L.out_pending && !L.out_need_k1) flush_output(L);
        L.flush_now = false;
      }
#ifdef GC_PROFILE_PHASES
      __syncwarp();
      long long tc = clock64();
      t_flush += tc - tb;
#endif
      if (__any_sync(0xffffffffu, L.crow >= 0)) {
        if (nown) combine_group(gc_smem_raw, nown);
        else combine(L);
      }
#ifdef GC_PROFILE_PHASES
      __syncwarp();
      t_comb += clock64() - tc;
#endif
    }
#ifdef GC_PROFILE_PHASES
    long long t1 = clock64();
#endif
    const bool need = L.want_derivs;
    const bool alive = L.phase != PH_DONE || need;
    if (blockDim.x > 32) {
      // lock step: every GC_LOCKSTEP_EVERY-th round (a power of two) the warps of the CTA meet here
      if ((wphase & (GC_LOCKSTEP_EVERY - 1)) == 0) {
        if (!__syncthreads_or(alive)) break;
      }
    } else {
      if (!__any_sync(0xffffffffu, alive)) break;
    }
    if (need) {
#if GC_HOT_IN_SMEM
      derivs(reinterpret_cast<const DevModel*>(L.hot), &L.e, &L.th, L.req_tau, col(L, L.req_in), col(L, L.req_out));
#else
      derivs(L.M, &L.e, &L.th, L.req_tau, col(L, L.req_in), col(L, L.req_out));
#endif
      L.want_derivs = false;
      L.n_derivs++;
      L.sum_n += L.e.neq;
    }
#ifdef GC_PROFILE_PHASES
    __syncwarp();
    if (tid == 0) {
      long long t2 = clock64();
      pw[0] += t_adv; pw[1] += t_comb; pw[2] += (t1 - t0) - t_adv - t_comb; pw[3] += t2 - t1; pw[4] += 1;
    }
    if ((threadIdx.x & 31) == 0) {
      long long t2 = clock64();
      atomicAdd(&counters[5], (unsigned long long)t_adv);
      atomicAdd(&counters[6], (unsigned long long)t_comb);
      atomicAdd(&counters[7], (unsigned long long)(t2 - t1));        // barrier + derivs
      atomicAdd(&counters[3], (unsigned long long)((t1 - t0) - t_adv - t_comb));  // flush + loop overhead
    }
#endif
  }
#ifdef GC_PROFILE_PHASES
  if (tid == 0)
    printf("[warp0] iterations %lld advance %.0f combine %.0f flush+loop %.0f derivs %.0f cycles/iteration\n", pw[4],
           (double)pw[0] / pw[4], (double)pw[1] / pw[4], (double)pw[2] / pw[4], (double)pw[3] / pw[4]);
#endif
  if (tid < nitems && !have) status[tid] = 0;
  if (have) {
    status[tid] = L.status;
    atomicAdd(&counters[0], (unsigned long long)L.n_derivs);
    atomicAdd(&counters[1], (unsigned long long)L.sum_n);
    atomicAdd(&counters[2], (unsigned long long)L.n_out);
    atomicAdd(&counters[4], (unsigned long long)L.n_shared);
  }
  }  // next block of items
}

}  // namespace gc

extern "C" int GC_ENTRY(gc_batch_lane_stride)() { return (int)GC_LANE_STRIDE; }

extern "C" void GC_ENTRY(gc_upload_constants_batch)() {
  double denl[260], polfac[260];
  for (int j = 0; j < 260; j++) {
    denl[j] = 1.0 / (2 * j + 1);
    polfac[j] = j >= 1 ? (double)((j + 3) * (j - 1)) / (j + 1) : 0.0;
  }
  cudaMemcpyToSymbol(c_denl, denl, sizeof(denl));
  cudaMemcpyToSymbol(c_polfac, polfac, sizeof(polfac));
}

extern "C" cudaError_t GC_ENTRY(gc_launch_evolve_batch)(const DevModel* models, const int2* items, int nitems, double* workspace, int NV,
                                              int* status, unsigned long long* counters, cudaStream_t stream) {
  // CTA size: as many warps as it takes to give every SM one CTA, up to the 7 of a full batch.  The warps of a CTA run
  // derivs() in lock step (one instruction stream per SM, see the main loop); with fewer warps than SMs every warp is its
  // own CTA.  Measured (evolve ms, distinct points, one-warp CTAs -> this rule): 128 models 1988 -> 1852, 256: 2826 -> 1922,
  // 512: 5101 -> 3148.
  static long lat_items = -1;
  if (lat_items < 0) {
    const char* e = getenv("GALCAMB_LATENCY_ITEMS");  // experiment knob: item count up to which one-warp CTAs are used
    lat_items = e ? atol(e) : 0;
  }
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
  }
  const int nwarps = (nitems + 31) / 32;
  int cta_warps = (nwarps + n_sm - 1) / n_sm;
  cta_warps = cta_warps < 1 ? 1 : cta_warps > GC_EVOLVE_THREADS / 32 ? GC_EVOLVE_THREADS / 32 : cta_warps;
  static long smem_pad = -1;
  if (smem_pad < 0) {
    const char* e = getenv("GALCAMB_EVOLVE_SMEM_PAD");  // experiment knob: unused shared memory per CTA (limits the CTAs per SM)
    smem_pad = e ? atol(e) : 0;
  }
  const int threads = nitems <= lat_items ? 32 : 32 * cta_warps;
  const size_t smem = (size_t)threads * GC_LANE_STRIDE + (threads > 32 ? (size_t)smem_pad : 0);  // lane records
  {  // per device and cheap: set on every launch (several devices may be driven from one process)
    cudaError_t e = cudaFuncSetAttribute(gc::evolve_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)((size_t)GC_EVOLVE_THREADS * GC_LANE_STRIDE + (size_t)smem_pad));
    if (e != cudaSuccess) return e;
  }
  const int blocks = (nitems + threads - 1) / threads;
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gc::evolve_batch_kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int grid = blocks < per_sm * n_sm ? blocks : per_sm * n_sm;  // persistent: what is resident at once
  // Experiment (off by default, it LOSES 16 %): the workspace slots of the resident warps are the first grid * warps slots;
  // as much of that range as an access-policy window may name (128 MiB on B200) marked persisting in L2 for this launch.
  static long l2_persist = -1;
  if (l2_persist < 0) {
    const char* e = getenv("GALCAMB_L2_PERSIST");  // experiment knob: 0 = no access-policy window
    l2_persist = e ? atol(e) : 0;  // measured on 1024 distinct points: 4420 ms without the window, 5136 ms with it
  }
  bool window = false;
  if (l2_persist && grid < blocks) {
    int dev = 0, max_win = 0, max_persist = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    if (max_win > 0 && max_persist > 0) {
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
      const size_t used = (size_t)grid * (threads / 32) * GC_NCOL * NV * GC_LANES * sizeof(double);
      cudaStreamAttrValue attr{};
      attr.accessPolicyWindow.base_ptr = workspace;
      attr.accessPolicyWindow.num_bytes = used < (size_t)max_win ? used : (size_t)max_win;
      attr.accessPolicyWindow.hitRatio = 1.0f;
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
      window = cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess;
    }
  }
  gc::evolve_batch_kernel<<<grid, threads, smem, stream>>>(models, items, nitems, workspace, NV, status, counters);
  const cudaError_t rc = cudaGetLastError();
  if (window) {  // the following kernels of the stream run without a window
    cudaStreamAttrValue attr{};
    attr.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &attr);
  }
  return rc;
}
