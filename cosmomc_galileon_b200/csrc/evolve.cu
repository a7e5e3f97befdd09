// K1 "evolve_sources": per-(k, cosmology) integration of the Einstein-Boltzmann + Galileon system with the
// Verner 6(5) pair of dverk, sampling the CMB sources on the output time grid.
//
// Replaces the OpenMP loop camb/cmbmain.f90:201-206 (DoSourcek :662 -> CalcScalarSources :917 ->
// GaugeInterface_EvolveScal camb/equations_galileon.f90:314 -> dverk camb/subroutines.f90:370 -> derivs/output).
//
// Execution model (DESIGN.md "K1"): one lane per (k, cosmology) work item, FP64 throughout.  Each lane runs the
// reference's control flow (output-time loop, approximation switches, adaptive step accept/reject) as an explicit
// state machine that stops whenever it needs a right-hand-side evaluation; the warp then executes ONE converged
// call of derivs() for all of its lanes.  Lanes of a warp therefore share the ~2 kFLOP RHS instruction stream
// even when they sit in different Runge-Kutta stages, at different step sizes or in different output intervals;
// only the short bookkeeping between evaluations diverges.
#include <algorithm>
#include <cstddef>
#include <cstdio>

#include "evolve_device.cuh"

// 4 warps per CTA, 3 CTAs per SM (168 registers); the warps of a CTA run derivs() in lock-step (see the main loop)
#ifndef GC_EVOLVE_THREADS
#define GC_EVOLVE_THREADS 128
#endif
#ifndef GC_EVOLVE_MINBLOCKS
#define GC_EVOLVE_MINBLOCKS 2
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
  PH_OUT_REQ,
  PH_OUT_DONE,
  PH_DONE
};

#ifndef GC_HOT_IN_SMEM
#define GC_HOT_IN_SMEM 1
#endif
#define GC_HOT_BYTES (offsetof(DevModel, adotrad))
static_assert(GC_HOT_BYTES % 8 == 0, "hot block must be a whole number of doubles");

struct Lane {
  EV e;
  const DevModel* M;
  double* ws;  // lane base of the warp workspace
  double* sm;  // latency mode: lane base of the shared-memory copy of column 9 (and 0 if sm_cols = 2), else nullptr
  int sm_cols;
  int NV;
  int kidx;
  int phase, stage, j, ind, after_switch;
  bool evaluated;
  double tau, tauend, target;
  // dverk communication (c(13), c(14), c(17), c(18), c(19), c(20), c(21), c(23))
  double hmin, hmag, xtrial, htrial, est, c20, c21, nfail;
  // pending RHS request
  double req_tau;
  int req_in, req_out;
  int status;
  // work counters
  unsigned long long n_derivs, sum_n, n_out;
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

// Latency mode (few work items, one-warp CTAs): y (column 0) and the stage input w9 (column 9) -- the two columns
// derivs() reads with dependent loads -- live in shared memory instead (L.sm != nullptr).
__device__ __forceinline__ double* col(const Lane& L, int c) {
  if (L.sm) {
    if (c == 9) return L.sm;
    if (c == 0 && L.sm_cols > 1) return L.sm + (size_t)L.NV * GC_LANES;
  }
  return L.ws + (size_t)c * L.NV * GC_LANES;
}

// Verner 6(5) stage tableau of camb/subroutines.f90:930-993: w9 = y + temp*(sum_j A[s][j]*w_j), evaluated left to
// right with the reference's signs so that every partial sum is bit-identical to the Fortran expression.
__constant__ double c_A[7][7] = {
    {233028180000.0, 0, 0, 0, 0, 0, 0},
    {74569017600.0, 298276070400.0, 0, 0, 0, 0, 0},
    {1165140900000.0, -3728450880000.0, 3495422700000.0, 0, 0, 0, 0},
    {-3604654659375.0, 12816549900000.0, -9284716546875.0, 1237962206250.0, 0, 0, 0},
    {3355605792000.0, -11185352640000.0, 9172628850000.0, -427218330000.0, 482505408000.0, 0, 0},
    {-770204740536.0, 2311639545600.0, -1322092233000.0, -453006781920.0, 326875481856.0, 0, 0},
    {2845924389000.0, -9754668000000.0, 7897110375000.0, -192082660000.0, 400298976000.0, 0, 201586000000.0}};
__constant__ int c_Aused[7][7] = {{1, 0, 0, 0, 0, 0, 0}, {1, 1, 0, 0, 0, 0, 0}, {1, 1, 1, 0, 0, 0, 0}, {1, 1, 1, 1, 0, 0, 0},
                                  {1, 1, 1, 1, 1, 0, 0}, {1, 1, 1, 1, 1, 0, 0}, {1, 1, 1, 1, 1, 0, 1}};


// w9 = y + temp * sum_j A[row][j] * w_(j+1), left to right as in camb/subroutines.f90:930-993
template <int NC>
__device__ __forceinline__ void stage_input_t(const double* __restrict__ y, double* __restrict__ w9, const double* const* kc,
                                              const double* cf, int n, double temp) {
  int i = 1;
  for (; i + 3 <= n; i += 4) {
    double yv[4], kv[NC][4];
#pragma unroll
    for (int u = 0; u < 4; u++) yv[u] = AT(y, i + u);
#pragma unroll
    for (int c = 0; c < NC; c++)
#pragma unroll
      for (int u = 0; u < 4; u++) kv[c][u] = AT(kc[c], i + u);
#pragma unroll
    for (int u = 0; u < 4; u++) {
      double acc = kv[0][u] * cf[0];
#pragma unroll
      for (int c = 1; c < NC; c++) acc = acc + kv[c][u] * cf[c];
      AT(w9, i + u) = yv[u] + temp * acc;
    }
  }
  for (; i <= n; i++) {
    double acc = AT(kc[0], i) * cf[0];
#pragma unroll
    for (int c = 1; c < NC; c++) acc = acc + AT(kc[c], i) * cf[c];
    AT(w9, i) = AT(y, i) + temp * acc;
  }
}

__device__ __forceinline__ void stage_input(Lane& L, int row, int n, double temp) {
  const double* y = col(L, 0);
  double* w9 = col(L, 9);
  const double* kc[6];
  double cf[6];
  int nc = 0;
#pragma unroll
  for (int jj = 0; jj < 7; jj++)
    if (c_Aused[row][jj]) {
      kc[nc] = col(L, jj + 1);
      cf[nc] = c_A[row][jj];
      nc++;
    }
  switch (nc) {
    case 2: stage_input_t<2>(y, w9, kc, cf, n, temp); break;
    case 3: stage_input_t<3>(y, w9, kc, cf, n, temp); break;
    case 4: stage_input_t<4>(y, w9, kc, cf, n, temp); break;
    case 5: stage_input_t<5>(y, w9, kc, cf, n, temp); break;
    default: stage_input_t<6>(y, w9, kc, cf, n, temp); break;
  }
}

// The stage derivatives w2..w8 and the stage input w9 are dead once a trial step has been accepted or rejected;
// they will be rewritten before they are read again.  Dropping their (dirty) lines from L2 without write-back keeps
// the dead 70 % of the workspace from being streamed to HBM and from competing with live lines for L2 capacity.
// One 256-byte row per element = two 128-byte lines; lanes 0 and 16 issue the discard for "their" line.
#ifndef GC_DISCARD_DEAD
#define GC_DISCARD_DEAD 1
#endif
// A line holds the same element of 16 lanes, so it may only be dropped when all 16 lanes of that half-warp are at
// this very point of their trial step together (true in lock-step batches; otherwise the lines are simply kept).
__device__ __forceinline__ void discard_rows(const double* colp, int n) {
#if GC_DISCARD_DEAD
  const unsigned grp = (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu;
  const bool together = (__activemask() & grp) == grp;
  if (together && (threadIdx.x & 15) == 0) {
    for (int i = 1; i <= n; i++) asm volatile("discard.global.L2 [%0], 128;" ::"l"(&AT(colp, i)) : "memory");
  }
#endif
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
__device__ __forceinline__ void advance(Lane& L) {
  const DevModel& M = *L.M;
  EV& e = L.e;
  const int S = M.SourceNum;
  const double tol = M.tol1;
  for (;;) {
    switch (L.phase) {
      case PH_NEXT_OUTPUT: {
        L.j++;
        if (L.j > M.nt) {
          L.phase = PH_DONE;
          return;
        }
        const double tauend = M.ts[(size_t)(L.j - 1) * 8];
        if ((e.q * tauend > M.max_etak_scalar && tauend > M.taurend) && !M.WantLateTime) {
          // camb/cmbmain.f90:1034-1036
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
        do_switch(L);
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
        if (L.ind != 6) {
          L.req_tau = L.tau;
          L.req_in = 0;
          L.req_out = 1;
          L.evaluated = true;
          L.phase = PH_DV_K1;
          return;
        }
        L.evaluated = false;
        L.phase = PH_DV_K1;
        break;
      }
      case PH_DV_K1: {  // step-size selection, camb/subroutines.f90:800-925, then stage 2 input
        const int n = e.neq;
        const double* y = col(L, 0);
        const double x = L.tau, xend = L.target;
        double temp = 0.0;
        for (int i = 1; i <= n; i++) temp = fmax(temp, fabs(AT(y, i)));
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
        temp = L.htrial / 1398169080000.0;
        const double* w1 = col(L, 1);
        double* w9 = col(L, 9);
        for (int i = 1; i <= n; i++) AT(w9, i) = AT(y, i) + temp * AT(w1, i) * 233028180000.0;
        L.req_tau = x + L.htrial / 6.0;
        L.req_in = 9;
        L.req_out = 2;
        L.stage = 2;
        L.phase = PH_DV_STAGE;
        return;
      }
      case PH_DV_STAGE: {
        const int n = e.neq;
        const int s = L.stage;  // the stage whose derivative has just been written (2..8)
        const double* y = col(L, 0);
        double* w9 = col(L, 9);
        const double x = L.tau;
        const double temp = L.htrial / 1398169080000.0;
        if (s < 8) {
          // input of stage s+1 : row s-1 of the tableau.  One specialised, 4-way batched loop per stage so that
          // the (1 + #columns) x 4 independent loads of a batch are in flight together (the columns live in
          // L2/HBM; a one-element-at-a-time loop stalls on every element).
          stage_input(L, s - 1, n, temp);
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
          L.req_in = 9;
          L.req_out = s + 1;
          L.stage = s + 1;
          return;
        }
        // s == 8: solution, error estimate, accept/reject (camb/subroutines.f90:995-1110)
        const double *w1 = col(L, 1), *w3 = col(L, 3), *w4 = col(L, 4), *w5 = col(L, 5), *w6 = col(L, 6), *w7 = col(L, 7),
                     *w8 = col(L, 8);
        double errn = 0.0;
        {
          int i = 1;
          for (; i + 1 <= n; i += 2) {  // 2-way batched: 2 x 8 independent loads in flight
            double v[2][8];
#pragma unroll
            for (int u = 0; u < 2; u++) {
              v[u][0] = AT(y, i + u); v[u][1] = AT(w1, i + u); v[u][2] = AT(w3, i + u); v[u][3] = AT(w4, i + u);
              v[u][4] = AT(w5, i + u); v[u][5] = AT(w6, i + u); v[u][6] = AT(w7, i + u); v[u][7] = AT(w8, i + u);
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
              AT(w9, i + u) = v[u][0] + temp * (v[u][1] * 104862681000.0 + v[u][2] * 545186250000.0 + v[u][3] * 446637345000.0 +
                                               v[u][4] * 188806464000.0 + v[u][6] * 15076875000.0 + v[u][7] * 97599465000.0);
              const double w2 = (v[u][1] * 8738556750.0 + v[u][2] * 9735468750.0 - v[u][3] * 9709507500.0 + v[u][4] * 8582112000.0 +
                                 v[u][5] * 95329710000.0 - v[u][6] * 15076875000.0 - v[u][7] * 97599465000.0) /
                                1398169080000.0;
              errn = fmax(errn, fabs(w2) / fmax(1.0, fabs(v[u][0])));
            }
          }
          for (; i <= n; i++) {
            AT(w9, i) = AT(y, i) + temp * (AT(w1, i) * 104862681000.0 + AT(w3, i) * 545186250000.0 + AT(w4, i) * 446637345000.0 +
                                           AT(w5, i) * 188806464000.0 + AT(w7, i) * 15076875000.0 + AT(w8, i) * 97599465000.0);
            const double w2 = (AT(w1, i) * 8738556750.0 + AT(w3, i) * 9735468750.0 - AT(w4, i) * 9709507500.0 +
                               AT(w5, i) * 8582112000.0 + AT(w6, i) * 95329710000.0 - AT(w7, i) * 15076875000.0 -
                               AT(w8, i) * 97599465000.0) /
                              1398169080000.0;
            errn = fmax(errn, fabs(w2) / fmax(1.0, fabs(AT(y, i))));
          }
        }
        L.est = errn * L.hmag * 1.0;
        L.ind = 5;
        if (L.est > tol) L.ind = 6;
        for (int c = 2; c <= 8; c++) discard_rows(col(L, c), n);
        if (L.ind != 6) {
          L.tau = L.xtrial;
          double* yw = col(L, 0);
          for (int i = 1; i <= n; i++) AT(yw, i) = AT(w9, i);
          discard_rows(w9, n);
          discard_rows(col(L, 1), n);  // k1 is re-evaluated at the start of the next step
          L.nfail = 0.0;
          if (L.tau == L.target) {
            L.ind = 3;
            L.c20 = L.target;
            L.c21 = 1.0;
            L.phase = L.after_switch ? PH_DO_SWITCH : PH_OUT_REQ;
            break;
          }
        } else {
          discard_rows(w9, n);
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
      case PH_OUT_REQ: {  // camb/equations_galileon.f90:1296-1297
        double* yp = col(L, 2);
        for (int i = 1; i <= e.nvar; i++) AT(yp, i) = 0;
        L.req_tau = L.tau;
        L.req_in = 0;
        L.req_out = 2;
        L.phase = PH_OUT_DONE;
        return;
      }
      case PH_OUT_DONE: {
        double src[3] = {0, 0, 0};
        output_sources(L.M, &L.e, col(L, 0), col(L, 2), M.ts + (size_t)(L.j - 1) * 8, L.tau, src);
        for (int s = 0; s < S; s++) M.Src[((size_t)(L.j - 1) * S + s) * M.nk + L.kidx] = src[s];
        L.n_out++;
        L.phase = PH_NEXT_OUTPUT;
        break;
      }
      default:
        return;
    }
  }
}

// items[i] = (model index, k index); workspace = nwarps * GC_NCOL * NV * 32 doubles.
__global__ void __launch_bounds__(GC_EVOLVE_THREADS, GC_EVOLVE_MINBLOCKS) evolve_sources_kernel(const DevModel* __restrict__ models, const int2* __restrict__ items,
                                                             int nitems, double* __restrict__ workspace, int NV,
                                                             int* __restrict__ status,
                                                             unsigned long long* __restrict__ counters, int state_in_smem) {
  // The per-lane control state (layout descriptor EV, dverk communication, pending request) lives in SHARED memory:
  // as a local variable it is a 1.2 KB stack frame per thread (470 KB per SM at 12 warps), which thrashes L1 and
  // turned every field access into a long-scoreboard stall.  Stride = odd multiple of 8 B -> conflict-free.
  extern __shared__ unsigned char gc_smem_raw[];
  Lane& L = *reinterpret_cast<Lane*>(gc_smem_raw + (size_t)threadIdx.x * GC_LANE_STRIDE);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int gwarp = tid >> 5;
  const bool have = tid < nitems && items[tid < nitems ? tid : 0].x >= 0;  // padding items carry model = -1
  L.phase = PH_DONE;
  L.status = 0;
  L.n_derivs = L.sum_n = L.n_out = 0;
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
    L.NV = NV;
    L.ws = workspace + (size_t)gwarp * GC_NCOL * NV * GC_LANES + lane;
    L.sm = state_in_smem ? reinterpret_cast<double*>(gc_smem_raw + (size_t)blockDim.x * GC_LANE_STRIDE) +
                               (size_t)(threadIdx.x >> 5) * state_in_smem * NV * GC_LANES + lane
                         : nullptr;
    L.sm_cols = state_in_smem;
    // DoSourcek, camb/cmbmain.f90:662-689
    L.e.q = M.k[it.y];
    L.e.pig = 0;
    L.e.pigdot = 0;
    for (int i = 0; i < GC_MAX_NU; i++) {
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
      for (int s = 0; s < M.SourceNum; s++) M.Src[(size_t)s * M.nk + L.kidx] = 0;
      L.phase = PH_NEXT_OUTPUT;
    }
  }
  // All warps of the CTA run the RHS evaluation together: the kernel's hot loop is ~8k straight-line
  // instructions (far beyond the 32 KB L1.5 instruction cache), so warps that reach derivs() at the same time
  // share one instruction stream from L2 instead of fetching it once per warp.
  for (;;) {
    if (L.phase != PH_DONE) advance(L);
    const bool need = L.phase != PH_DONE;
    if (blockDim.x > 32) {
      if (!__syncthreads_or(need)) break;
    } else {
      if (!__any_sync(0xffffffffu, need)) break;
    }
    if (need) {
      #if GC_HOT_IN_SMEM
      derivs(reinterpret_cast<const DevModel*>(L.hot), &L.e, L.req_tau, col(L, L.req_in), col(L, L.req_out));
#else
      derivs(L.M, &L.e, L.req_tau, col(L, L.req_in), col(L, L.req_out));
#endif
      L.n_derivs++;
      L.sum_n += L.e.neq;
    }
  }
  if (tid < nitems && !have) status[tid] = 0;
  if (have) {
    status[tid] = L.status;
    atomicAdd(&counters[0], L.n_derivs);
    atomicAdd(&counters[1], L.sum_n);
    atomicAdd(&counters[2], L.n_out);
  }
}

}  // namespace gc

extern "C" void gc_upload_constants() {
  double denl[260], polfac[260];
  for (int j = 0; j < 260; j++) {
    denl[j] = 1.0 / (2 * j + 1);
    polfac[j] = j >= 1 ? (double)((j + 3) * (j - 1)) / (j + 1) : 0.0;
  }
  cudaMemcpyToSymbol(c_denl, denl, sizeof(denl));
  cudaMemcpyToSymbol(c_polfac, polfac, sizeof(polfac));
}

extern "C" cudaError_t gc_launch_evolve(const DevModel* models, const int2* items, int nitems, double* workspace, int NV,
                                        int* status, unsigned long long* counters, cudaStream_t stream) {
  // Few work items (a single model = ~200 items; 128 models = 4 warps per k): one warp per CTA so that every warp
  // gets (half) an SM to itself and no warp waits for another at the lock-step barrier.  The run time is then the
  // critical path of one low-k item, so the latency-critical columns move into shared memory.
  static long lat_items = -1;
  if (lat_items < 0) {
    const char* e = getenv("GALCAMB_LATENCY_ITEMS");  // experiment knob: item count up to which latency mode is used
    lat_items = e ? atol(e) : 160000;  // measured crossover with the lock-step batch mode: ~800 models
  }
  const bool latency_mode = nitems <= lat_items;
  const int threads = latency_mode ? 32 : GC_EVOLVE_THREADS;
  size_t smem = (size_t)threads * GC_LANE_STRIDE;  // lane records
  int state_in_smem = 0;
  if (latency_mode) {
    // up to 2 warps per SM: both y and w9 in shared memory (82 KB per warp); beyond that only w9 (54 KB, 4 warps/SM)
    static int cols_env = -2;
    if (cols_env == -2) {
      const char* e = getenv("GALCAMB_LATENCY_COLS");
      cols_env = e ? atoi(e) : -1;
    }
    const int cols = cols_env >= 0 ? cols_env : ((nitems + 31) / 32 <= 2 * 148 ? 2 : 1);
    const size_t extra = (size_t)cols * NV * GC_LANES * sizeof(double);  // column 9 (and 0) of the warp
    if (cols > 0 && smem + extra <= 113 * 1024) {                        // >= two one-warp CTAs per SM
      smem += extra;
      state_in_smem = cols;
    }
  }
  static size_t configured = 0;
  const size_t want = std::max(smem, (size_t)GC_EVOLVE_THREADS * GC_LANE_STRIDE);
  if (want > configured) {
    cudaError_t e = cudaFuncSetAttribute(gc::evolve_sources_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
    if (e != cudaSuccess) return e;
    configured = want;
  }
  const int blocks = (nitems + threads - 1) / threads;
  gc::evolve_sources_kernel<<<blocks, threads, smem, stream>>>(models, items, nitems, workspace, NV, status, counters, state_in_smem);
  return cudaGetLastError();
}
