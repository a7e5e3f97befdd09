// K1 "evolve_sources": per-(k, cosmology) integration of the Einstein-Boltzmann + Galileon system with the
// Verner 6(5) pair of dverk, sampling the CMB sources on the output time grid.
//
// Replaces the OpenMP loop camb/cmbmain.f90:201-206 (DoSourcek :662 -> CalcScalarSources :917 ->
// GaugeInterface_EvolveScal camb/equations_galileon.f90:314 -> dverk camb/subroutines.f90:370 -> derivs/output).
//
// Execution model (DESIGN.md "K1"):
//   * one work item = one (k, cosmology) pair = one OCTET of eight lanes; a warp holds four items, normally the same
//     k index of four cosmologies; persistent CTAs, every warp pulls groups of four items from a queue that the
//     host sorts expensive-first (the reference's SCHEDULE(DYNAMIC) over k, camb/cmbmain.f90:201);
//   * the whole state of an item lives ON CHIP: y, the stage input and the derivative columns k1..k8 of dverk are
//     nine shared-memory columns (k8 reuses the slot of k2, which is dead by then); nothing of the integrator state
//     ever touches L2/HBM;
//   * a trial step is ONE converged pass of the warp, in three phases:
//       A  the scale factor obeys its own closed equation a' = a^2 H(a) (camb/equations_galileon.f90:2221-2230), so its
//          eight Runge-Kutta stage values are integrated first, alone (one spline lookup per stage, rows prefetched
//          into shared memory);
//       B  everything in the right-hand side that depends on the background point (tau_s, a_s) only -- thermo
//          lookup, massive-neutrino background, and the ~1 kFLOP of Galileon coefficient polynomials, returned as the
//          coefficients of linear forms (galileon_fast.cuh) -- is then evaluated for the EIGHT STAGES SIDE BY SIDE,
//          lane s of the octet taking stage s: full SIMD efficiency, one exposed table latency per trial step;
//       C  the eight stages in order: stage combination (all lanes, shared memory), the perturbation-dependent scalar
//          part of derivs on the lane that holds the stage's background record, the multipole hierarchies across the
//          eight lanes through a per-equation descriptor table, then the error estimate by an octet reduction;
//   * between steps lane 0 of each octet runs the reference's control flow (output-time loop, approximation
//     switches, dverk's step-size control and accept/reject) as a scalar state machine on the item record.
// The y' that output() evaluates at an output time (camb/equations_galileon.f90:1297) is the k1 of the next trial step
// (same tau, y and layout), so the sources are formed right after stage 1 of that step; an explicit single
// evaluation remains where a switch or the end of the run follows the output directly.
#include <algorithm>
#include <cstddef>
#include <cstdio>

// -DGC_WITH_TRANSFER compiles this file a second time with the matter-transfer taps of CP%WantTransfer
// (camb/cmbmain.f90:1017-1068, :1151-1201) in the control flow, in its own namespace and under entry-point names ending
// in _tr: the default kernels stay exactly as they are without the taps (the extra state-machine phases cost registers
// in the hot loop: +14 % on a 1024-point batch when compiled in unconditionally).
#if defined(GC_WITH_TRANSFER)
#define gc gc_tr
#define GC_ENTRY(name) name##_tr
#elif defined(GC_EV_NU)
#define gc gc_nu1
#define GC_ENTRY(name) name##_nu1
#else
#define GC_ENTRY(name) name
#endif

#include "evolve_device.cuh"

#define GC_OCT 8    // lanes per work item
#define GC_IPW 4    // work items per warp
#define GC_NCOLS 9  // y / stage input (2, swapped on acceptance), k1 .. k7 (k8 -> slot of k2)
#define GC_LTAB 132  // multipoles whose 1/(2l+1) and (l+3)(l-1)/(l+1) sit in shared memory (larger l: constant memory)
#define GC_LTAB_BYTES (2 * GC_LTAB * 8)

namespace gc {

// -DGC_PROFILE_PHASES: warp-cycles per phase of the trial step, accumulated into counters[8..15]
#ifdef GC_PROFILE_PHASES
#define GC_TICK(slot)                                         \
  do {                                                        \
    const long long t_now_ = clock64();                       \
    prof[slot] += t_now_ - t_last;                            \
    t_last = t_now_;                                          \
  } while (0)
#else
#define GC_TICK(slot) \
  do {                \
  } while (0)
#endif

// PH_TF_CHECK = label 101 of camb/cmbmain.f90:1049 (is a transfer redshift due before the next time step?),
// PH_TF_ARRIVED = at (or within 1e-5 of) a transfer time: outtransf
enum Phase : int { PH_NEXT_OUTPUT = 0, PH_EVOLVE, PH_DO_SWITCH, PH_DV_ENTER, PH_DV_STEP, PH_DV_FINAL, PH_OUT_MARK, PH_TF_CHECK, PH_TF_ARRIVED, PH_DONE };  // the two PH_TF phases: -DGC_WITH_TRANSFER only

// per-equation descriptor: type | l << 4 | (iq-1) << 12 | nu_i << 16 ; type 0 = written by the scalar part
enum EqType : unsigned {
  EQ_SCALAR = 0, EQ_PHOT_MID, EQ_PHOT_LAST, EQ_POL_MID, EQ_POL_LAST, EQ_NU_MID, EQ_NU_LAST, EQ_MNU_0, EQ_MNU_MID, EQ_MNU_2,
  EQ_MNU_LAST, EQ_MNU_2T, EQ_PERT_0, EQ_PERT_MID, EQ_PERT_LAST
};

// scalars the scalar part of a stage hands to the hierarchy lanes
enum { SC_Z = 0, SC_SIGMA, SC_OPAC, SC_COTH, SC_A, SC_QR, SC_N };

struct Item {
  EV e;
  const DevModel* M;
  int kidx, slot;
  int phase, j, ind, after_switch, status;
  int req;  // 0 = idle/finished, 1 = trial step, 2 = single evaluation (explicit y' of a pending output)
  int ysel;  // which of the two state columns holds y (the other one is the stage input / trial solution)
  int ynorm_valid, out_pending, out_j;
#ifdef GC_WITH_TRANSFER
  // matter transfer functions: next transfer redshift (0-based), kind of the stop being integrated to (0 none, 1 in the
  // time-step loop: taken if within 1e-5, 2 before the loop / transfer-only wavenumber: always taken)
  int itf, tf_stop;
#endif
  double tau, tauend, target;
  // dverk communication (c(13), c(14), c(17), c(18), c(19), c(20), c(21), c(23))
  double hmin, hmag, xtrial, htrial, est, c20, c21, nfail;
  double errn, ynorm_new, ynorm;
  double sc[SC_N];
  int win_i0;
  double galwin[GC_OCT * 6];  // eight consecutive rows of the Galileon background spline (+ 1/(a_{i+1} - a_i)), per trial step
  // massive-neutrino velocity factors of the current stage: v = q/E and E/q per (eigenstate, momentum node)
  double vv[GC_EV_NU * GC_MAX_NQ], ivv[GC_EV_NU * GC_MAX_NQ];
  double nub[3 * GC_EV_NU];  // rhonu, pnu, w = pnu/rhonu of the current stage per eigenstate (dynamic-index copy of BG)
  // work counters: n_derivs / sum_n count what the reference algorithm evaluates
  unsigned n_derivs, sum_n, n_out, n_shared, n_exec, n_trial;
};

__host__ __device__ inline size_t item_stride_bytes(int NVP) {
  size_t b = ((sizeof(Item) + 7) / 8) * 8 + (((size_t)NVP * 4 + 7) / 8) * 8 + (size_t)GC_NCOLS * NVP * 8;
  // consecutive items 64 bytes apart modulo 128: the two octets of a half-warp hit disjoint banks
  b = (b + 63) / 64 * 64;
  if ((b / 64) % 2 == 0) b += 64;
  return b;
}

struct ItemMem {
  Item* it;
  unsigned* desc;
  double* cols;
  int NVP;
  __device__ __forceinline__ double* col(int c) const { return cols + (size_t)c * NVP; }
  __device__ __forceinline__ double* Y() const { return col(it->ysel); }
  __device__ __forceinline__ double* W() const { return col(it->ysel ^ 1); }
  __device__ __forceinline__ double* K(int s) const { return col(2 + (s == 7 ? 1 : s)); }  // s = 0..7 -> k1..k8
};

// ------------------------------------------------------------------------------------------------ tableau
// camb/subroutines.f90:930-1019, written row by row with the reference's signs and summation order
template <int R>
__device__ __forceinline__ double row_val(double y, double temp, const double* k) {
  if (R == 0) return y + temp * k[0] * 233028180000.0;
  if (R == 1) return y + temp * (k[0] * 74569017600.0 + k[1] * 298276070400.0);
  if (R == 2) return y + temp * (k[0] * 1165140900000.0 - k[1] * 3728450880000.0 + k[2] * 3495422700000.0);
  if (R == 3)
    return y + temp * (-k[0] * 3604654659375.0 + k[1] * 12816549900000.0 - k[2] * 9284716546875.0 + k[3] * 1237962206250.0);
  if (R == 4)
    return y + temp * (k[0] * 3355605792000.0 - k[1] * 11185352640000.0 + k[2] * 9172628850000.0 - k[3] * 427218330000.0 +
                       k[4] * 482505408000.0);
  if (R == 5)
    return y + temp * (-k[0] * 770204740536.0 + k[1] * 2311639545600.0 - k[2] * 1322092233000.0 - k[3] * 453006781920.0 +
                       k[4] * 326875481856.0);
  if (R == 6)
    return y + temp * (k[0] * 2845924389000.0 - k[1] * 9754668000000.0 + k[2] * 7897110375000.0 - k[3] * 192082660000.0 +
                       k[4] * 400298976000.0 + k[6] * 201586000000.0);
  return y + temp * (k[0] * 104862681000.0 + k[2] * 545186250000.0 + k[3] * 446637345000.0 + k[4] * 188806464000.0 +
                     k[6] * 15076875000.0 + k[7] * 97599465000.0);
}
__device__ __forceinline__ double err_val(const double* k) {
  return (k[0] * 8738556750.0 + k[2] * 9735468750.0 - k[3] * 9709507500.0 + k[4] * 8582112000.0 + k[5] * 95329710000.0 -
          k[6] * 15076875000.0 - k[7] * 97599465000.0) /
         1398169080000.0;
}

// stage input of stage R+2 for the elements i = i0, i0+8, ... <= n of this lane
template <int R>
__device__ __forceinline__ void combine_row(const ItemMem& m, int i0, int n, double temp) {
  const double* __restrict__ y = m.Y();
  double* __restrict__ w = m.W();
  const double* kc[7];
#pragma unroll
  for (int c = 0; c < 7; c++) kc[c] = m.col(2 + c);
#pragma unroll 2
  for (int i = i0; i <= n; i += GC_OCT) {
    double k[8];
#pragma unroll
    for (int c = 0; c < 7; c++) k[c] = (c <= R + 0 && !(R == 6 && c == 5)) ? kc[c][i - 1] : 0.0;
    k[7] = 0;
    w[i - 1] = row_val<R>(y[i - 1], temp, k);
  }
}

// ------------------------------------------------------------------------------------------------ background
// out-of-line copies of the table lookups (one copy each in the kernel image; the hot loop has to stay small)
// (results by value: handing out addresses of register-resident records would push them into local memory)
__device__ __noinline__ double2 nu_background_ni(const DevModel* M, double am) {
  double2 r;
  Nu_background(*M, am, r.x, r.y);
  return r;
}
__device__ __noinline__ double2 gal_HX_ni(const DevModel* M, double a) {
  double2 r;
  gal_HX(*M, a, r.x, r.y);
  return r;
}

// Galileon background spline with the rows of the current trial step staged in shared memory: win[r*6 + 0..4] =
// {a, hbar, c_h, x, c_x} of row win_i0 + r, win[r*6 + 5] = 1/(a_{r+1} - a_r).  The interval is found by comparing
// against the staged nodes -- the bracket test that also decides in gal_HX (evolve_device.cuh) -- and the spline is
// evaluated with the same arithmetic; a point outside the window takes the global-memory path.
__device__ __forceinline__ void gal_HX_win(const DevModel& M, double a, const double* __restrict__ win, int i0, double& hbar,
                                           double& xgal) {
  const int n = M.ngal;
  int r = -1;
  if (i0 >= 0 && a >= win[0]) {
    r = 0;
#pragma unroll
    for (int q = 1; q < GC_OCT - 1; q++)
      if (a >= win[q * 6]) r = q;
    if (a >= win[(GC_OCT - 1) * 6] && i0 + GC_OCT - 2 != n - 2) r = -1;  // beyond the window (the last interval is open-ended)
  }
  if (r < 0) {
    const double2 hx = gal_HX_ni(&M, a);
    hbar = hx.x;
    xgal = hx.y;
    return;
  }
  const double* r0 = win + r * 6;
  const double* r1 = r0 + 6;
  const double dx = r1[0] - r0[0], delx = a - r0[0];
  const double idx = r0[5], third = 1.0 / 3.0;
  {
    const double dy = r1[1] - r0[1];
    const double b = dy * idx - dx * (r1[2] + 2.0 * r0[2]) * third;
    const double dd = (r1[2] - r0[2]) * idx * third;
    hbar = r0[1] + delx * (b + delx * (r0[2] + delx * dd));
  }
  {
    const double dy = r1[3] - r0[3];
    const double b = dy * idx - dx * (r1[4] + 2.0 * r0[4]) * third;
    const double dd = (r1[4] - r0[4]) * idx * third;
    xgal = a * (r0[3] + delx * (b + delx * (r0[4] + delx * dd)));
  }
}

// 1/(a dtauda(a)), camb/equations_galileon.f90:107-156, with whatever the lookup produced on the way
__device__ __forceinline__ double adotoa_at(const DevModel& M, bool gal, double a, const double* win, int win_i0, double& hbar,
                                            double& xgal, double* rhonu, double* pnu) {
  const double a2 = a * a;
  double grhoa2;
  if (gal) gal_HX_win(M, a, win, win_i0, hbar, xgal);
  if (gal && a >= 1e-10) {
    const double a4 = a2 * a2;
    grhoa2 = a4 * M.grhom * (hbar * hbar);
  } else {
    grhoa2 = M.grhok * a2 + (M.grhoc + M.grhob) * a + M.grhog + M.grhornomass;
    grhoa2 = grhoa2 + M.grhov * (a2 * a2);
    if (M.Num_Nu_massive != 0) {
#pragma unroll
      for (int nu_i = 0; nu_i < GC_EV_NU; nu_i++)
        if (nu_i < M.n_eig) {
          const double2 rp = nu_background_ni(&M, a * M.nu_masses[nu_i]);
          rhonu[nu_i] = rp.x;
          pnu[nu_i] = rp.y;
          grhoa2 = grhoa2 + rhonu[nu_i] * M.grhormass[nu_i];
        }
    }
  }
  return sqrt(grhoa2 * (1.0 / 3.0)) * (1.0 / a);
}

// background record of one right-hand-side evaluation: everything derivs (camb/equations_galileon.f90:2098-2589)
// computes from (tau, a) alone.  Lives in the registers of the lane that owns the stage.
struct BG {
  double tau, a, adotoa, ka;  // ka = y'(1) = adotoa * a
  double ia, iadotoa, cothxor;  // 1/a, 1/adotoa, 1/tau: the divisions leave the serial part of the stage
  double cs2, opacity, dopacity;
  double grhob_t, grhoc_t, grhor_t, grhog_t, gpres;
  double photbar, pb43;
  double iop, ipb;  // tight coupling: 1/opacity, 1/(1 + pb43)
  double rhonu[GC_EV_NU], pnu[GC_EV_NU];
  double chi_p, chi_d, chi_g, chi_e, q_d, q_p, q_g;  // Chigal, qgal as linear forms (galileon_fast.cuh)
  GalPhi2Lin D;                                       // dphisecond
  GalPiLin P;                                         // Pigal (tight coupling only)
};

__device__ __forceinline__ void bg_eval(const DevModel& M, const EV& e, bool tight, double hbar, double xgal, bool have_nu,
                                        BG& b) {
  const bool gal = M.use_galileon != 0;
  const double a = b.a, a2 = a * a;
  const double ia = 1.0 / a, ia2 = ia * ia;
  b.ia = ia;
  b.iadotoa = 1.0 / b.adotoa;
  b.cothxor = 1.0 / b.tau;
  b.grhob_t = M.grhob * ia;
  b.grhoc_t = M.grhoc * ia;
  b.grhor_t = M.grhornomass * ia2;
  b.grhog_t = M.grhog * ia2;
  b.photbar = b.grhog_t / b.grhob_t;
  b.pb43 = 4.0 / 3 * b.photbar;
  if (!have_nu) {
#pragma unroll
    for (int i = 0; i < GC_EV_NU; i++)
      if (i < M.n_eig) {
        const double2 rp = nu_background_ni(&M, a * M.nu_masses[i]);
        b.rhonu[i] = rp.x;
        b.pnu[i] = rp.y;
      }
  }
  double grhov_t = 0;
  double gpresgal = 0;
  if (gal) {
    double lam_nu = 0;
#pragma unroll
    for (int i = 0; i < GC_EV_NU; i++)
      if (i < M.n_eig) lam_nu += M.grhormass[i] * b.pnu[i];
    const GalConst GC{M.c2, M.c3, M.c4, M.c5, M.cG, M.h0, M.orad};
    GalFast F;
    gal_fast_background(GC, a, hbar, xgal, lam_nu, e.k, F);
    gpresgal = F.gpresgal;
    b.chi_p = F.chi_p; b.chi_d = F.chi_d; b.chi_g = F.chi_g; b.chi_e = F.chi_e;
    b.q_d = F.q_d; b.q_p = F.q_p; b.q_g = F.q_g;
    gal_fast_dphisecond_lin(GC, F, b.D);
    if (tight)
      gal_fast_Pigal_lin(GC, F, b.P);
    else
      b.P = GalPiLin{0, 0, 0, 0, 0};
  } else {
    grhov_t = M.grhov * a2;
    b.chi_p = b.chi_d = b.chi_g = b.chi_e = b.q_d = b.q_p = b.q_g = 0;
    b.D = GalPhi2Lin{0, 0, 0, 0, 0, 0};
    b.P = GalPiLin{0, 0, 0, 0, 0};
  }
  b.dopacity = 0;
  thermo(M, b.tau, b.cs2, b.opacity, tight ? &b.dopacity : nullptr);
  b.iop = 0;
  b.ipb = 0;
  if (tight) {
    b.iop = 1.0 / b.opacity;
    b.ipb = 1.0 / (1 + b.pb43);
  }
  double gpres = 0;
  gpres = gpres + (b.grhog_t + b.grhor_t) / 3;
  if (gal)
    gpres = gpres + gpresgal;
  else
    gpres = gpres + grhov_t * (-1.0);
  if (M.Num_Nu_massive > 0) {
#pragma unroll
    for (int i = 0; i < GC_EV_NU; i++)
      if (i < M.n_eig) gpres = gpres + (M.grhormass[i] * ia2) * b.pnu[i];
  }
  b.gpres = gpres;
}

// Nu_Integrate_L012 without the pressure moments (camb/equations_galileon.f90:1058-1107, the call of MassiveNuVars
// :1232): density and momentum sums over the momentum nodes, velocity factors from the stage's table
__device__ __noinline__ void nu_sums_L01(const DevModel& M, const EV& e, const double* __restrict__ y, int nu_i,
                                         const double* __restrict__ vv, const double* __restrict__ ivv, double& drhonu,
                                         double& fnu) {
  drhonu = 0;
  fnu = 0;
  int ind = e.nu_ix[nu_i];
  for (int iq = 1; iq <= e.nq[nu_i]; iq++) {
    const double ker = M.nu_int_kernel[iq - 1];
    drhonu = drhonu + ker * AT(y, ind) * ivv[nu_i * GC_MAX_NQ + iq - 1];
    fnu = fnu + ker * AT(y, ind + 1);
    ind = ind + e.lmaxnu_tau[nu_i] + 1;
  }
  ind = e.nu_pert_ix;
  for (int iq = e.nq[nu_i] + 1; iq <= M.nqmax; iq++) {
    const double r = M.nu_masses[nu_i] * M.inv_nu_q[iq - 1];
    const double pert_scale = (r * r) / 2;
    const double ker = M.nu_int_kernel[iq - 1];
    const double tmp = ker * (AT(y, e.r_ix) + pert_scale * AT(y, ind));
    drhonu = drhonu + tmp * ivv[nu_i * GC_MAX_NQ + iq - 1];
    fnu = fnu + ker * (AT(y, e.r_ix + 1) + pert_scale * AT(y, ind + 1));
  }
}

// ------------------------------------------------------------------------------------------------ derivs, scalar part
// camb/equations_galileon.f90:2139-2429, :2470-2487, :2516-2521, :2533-2547: constraints, baryons, photon / neutrino
// l <= 2, tight-coupling slip, the Galileon field equation, the massive-neutrino fluid approximation.  Run by the one
// lane that holds the stage's BG; reads the stage input `ay`, writes the low-order derivatives into `ayp` and the
// scalars the hierarchy lanes need into sc[].  Writes e.pig / e.pigdot during tight coupling exactly as the reference
// does (its value at a switch is the one left by the LAST call).
__device__ __forceinline__ void rhs_scalar(const DevModel* __restrict__ Mp, EV* __restrict__ ep, const BG& b,
                                        const double* __restrict__ ay, double* __restrict__ ayp, double* __restrict__ sc,
                                        double* __restrict__ vv, double* __restrict__ ivv, double* __restrict__ nub) {
  const DevModel& M = *Mp;
  EV& e = *ep;
  const bool gal = M.use_galileon != 0;
  const double k = e.k, ik = e.ik, ik2 = e.ik2;
  const double a = b.a;
  const double ia = b.ia, ia2 = ia * ia;
  // velocity factors of the massive-neutrino momentum nodes at this a (equations_galileon.f90:1075-1078, :2551-2553):
  // for the momentum sums below and, through shared memory, for the hierarchy lanes
  if (M.Num_Nu_massive > 0) {
    const int nqm = M.nqmax;
#pragma unroll 1
    for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
      const double am = a * M.nu_masses[nu_i];
#pragma unroll 1
      for (int iq = 0; iq < nqm; iq++) {
        const double aq = am * M.inv_nu_q[iq];
        const double e2 = 1.0 + aq * aq;
        const double v = rsqrt(e2);
        vv[nu_i * GC_MAX_NQ + iq] = v;
        ivv[nu_i * GC_MAX_NQ + iq] = e2 * v;
      }
    }
  }
  const double etak = AT(ay, 2), clxc = AT(ay, 3), clxb = AT(ay, 4), vb = AT(ay, 5);
  double dphi = 0, dphiprime = 0;
  if (gal) {
    dphi = AT(ay, e.w_ix);
    dphiprime = AT(ay, e.w_ix + 1);
  }
  double y_r0 = 0, y_r1 = 0, y_r2 = 0, y_g0 = 0, y_g1 = 0, y_g2 = 0, y_g3 = 0, y_e2 = 0, y_e3 = 0, y_r3 = 0;
  if (!e.no_nu) {
    y_r0 = AT(ay, e.r_ix);
    y_r1 = AT(ay, e.r_ix + 1);
    y_r2 = AT(ay, e.r_ix + 2);
    if (!e.high_ktau) y_r3 = AT(ay, e.r_ix + 3);
  }
  if (!e.no_phot) {
    y_g0 = AT(ay, e.g_ix);
    y_g1 = AT(ay, e.g_ix + 1);
    if (!e.TightCoupling) {
      y_g2 = AT(ay, e.g_ix + 2);
      y_g3 = AT(ay, e.g_ix + 3);
      y_e2 = AT(ay, e.polind + 2);
      y_e3 = AT(ay, e.polind + 3);
    }
  }
  const double grhob_t = b.grhob_t, grhoc_t = b.grhoc_t, grhor_t = b.grhor_t, grhog_t = b.grhog_t;
  const double gpres = b.gpres;
  double dgrho_matter = grhob_t * clxb + grhoc_t * clxc;
  double dgq = grhob_t * vb;
  if (M.Num_Nu_massive > 0) {  // MassiveNuVars, :1209-1252
#pragma unroll
    for (int i = 0; i < GC_EV_NU; i++) {  // BG lives in registers: a copy the rolled loops below can index
      nub[i] = b.rhonu[i];
      nub[GC_EV_NU + i] = b.pnu[i];
    }
#pragma unroll 1
    for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
      const double rhonu = nub[nu_i], pnu = nub[GC_EV_NU + nu_i];
      const double grhormass_t = M.grhormass[nu_i] * ia2;
      const double irho = 1.0 / rhonu;
      double clxnu, qnu;
      if (e.MassiveNuApprox[nu_i]) {
        clxnu = AT(ay, e.nu_ix[nu_i]);
        qnu = AT(ay, e.nu_ix[nu_i] + 2);
      } else {
        nu_sums_L01(M, e, ay, nu_i, vv, ivv, clxnu, qnu);
        qnu = qnu * irho;
        clxnu = clxnu * irho;
      }
      const double grhonu_t = grhormass_t * rhonu;
      dgrho_matter = dgrho_matter + grhonu_t * clxnu;
      dgq = dgq + grhonu_t * qnu;
      nub[2 * GC_EV_NU + nu_i] = pnu * irho;
    }
  }
  const double adotoa = b.adotoa;
  const double iadotoa = b.iadotoa;
  const double cothxor = b.cothxor;
  const double opacity = b.opacity, cs2 = b.cs2, dopacity = b.dopacity;
  double dgrho = dgrho_matter;
  double z = 0, dz, clxr, qr, pir, clxg, qg, pig = 0, dgrhogal_t;
  if (e.no_nu) {
    if (gal) {
      dgrhogal_t = b.chi_p * dphiprime + b.chi_d * dphi + b.chi_g * dgrho + b.chi_e * etak;
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
        dgrhogal_t = b.chi_p * dphiprime + b.chi_d * dphi + b.chi_g * dgrho + b.chi_e * etak;
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
  const double photbar = b.photbar;
  const double pb43 = b.pb43;
  AT(ayp, 1) = b.ka;
  if (gal) {
    dgrhogal_t = b.chi_p * dphiprime + b.chi_d * dphi + b.chi_g * dgrho + b.chi_e * etak;
    dgrho = dgrho + dgrhogal_t;
    dgq = dgq + (b.q_d * dphi + b.q_p * dphiprime + b.q_g * dgq);
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
    const double iop = b.iop, iop2 = iop * iop, ipb = b.ipb;
    const double c3245 = 32.0 / 45.0;
    pig = c3245 * iop * k * (sigma + vb);
    const double slip = -(2 * adotoa * ipb + dopacity * iop) * (vb - 3.0 / 4 * qg) +
                        (-adotdota * vb - k / 2 * adotoa * clxg + k * (cs2 * clxbdot - clxgdot / 4)) * (iop * ipb);
    double dgs = grhog_t * pig + grhor_t * pir;
    if (gal) dgs = dgs + (b.P.d * dphi + b.P.pi * dgs + b.P.g * dgrho + b.P.q * dgq + b.P.e * etak);
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
      const double polter = pig / 10 + 9.0 / 15 * y_e2;
      pigdot = denlk(e, 2) * qg - denlk2(e, 2) * y_g3 - opacity * (pig - polter) + 8.0 / 15.0 * k * sigma;
      AT(ayp, e.g_ix + 2) = pigdot;
      AT(ayp, e.polind + 2) = -opacity * (y_e2 - polter) - k / 3.0 * y_e3;
    }
  }
  // clxrdot is read uninitialised by the reference when no_nu_multpoles (:2517 vs :2482); pinned to 0 like
  // the oracle (the value of its defining expression under qr = -4z/3).
  double clxrdot = 0;
  if (!e.no_nu) {
    clxrdot = -k * (4.0 / 3.0 * z + qr);
    AT(ayp, e.r_ix) = clxrdot;
    AT(ayp, e.r_ix + 1) = denlk(e, 1) * clxr - denlk2(e, 1) * pir;
    if (e.high_ktau)
      AT(ayp, e.r_ix + 2) = -3 * pir * cothxor - clxrdot;
    else
      AT(ayp, e.r_ix + 2) = denlk(e, 2) * qr - denlk2(e, 2) * y_r3 + 8.0 / 15.0 * k * sigma;
  }
  if (gal) {
    const double dotdeltaf = grhob_t * (clxbdot - 3 * adotoa * clxb) + grhoc_t * (clxcdot - 3 * adotoa * clxc) +
                             grhor_t * (clxrdot - 4 * adotoa * clxr) + grhog_t * (clxgdot - 4 * adotoa * clxg);
    AT(ayp, e.w_ix) = dphiprime;
    AT(ayp, e.w_ix + 1) =
        b.D.p * dphiprime + b.D.d * dphi + b.D.g * dgrho + b.D.q * dgq + b.D.f * dotdeltaf + b.D.e * etak;
  }
  sc[SC_Z] = z;
  sc[SC_SIGMA] = sigma;
  sc[SC_OPAC] = opacity;
  sc[SC_COTH] = cothxor;
  sc[SC_A] = a;
  sc[SC_QR] = qr;
  if (M.Num_Nu_massive == 0) return;
#pragma unroll 1
  for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
    if (e.MassiveNuApprox[nu_i]) {
      const double ia3 = ia * ia2;
      const double G11_t = e.G11[nu_i] * ia3;
      const double G30_t = e.G30[nu_i] * ia3;
      const int o = e.nu_ix[nu_i];
      const double w = nub[2 * GC_EV_NU + nu_i];
      AT(ayp, o) = -k * z * (w + 1) + 3 * adotoa * (w * AT(ay, o) - AT(ay, o + 1)) - k * AT(ay, o + 2);
      AT(ayp, o + 1) = (3 * w - 2) * adotoa * AT(ay, o + 1) - 5.0 / 3 * k * z * w - k / 3 * G11_t;
      AT(ayp, o + 2) = (3 * w - 1) * adotoa * AT(ay, o + 2) - k * (2.0 / 3 * 1.0 * AT(ay, o + 3) - AT(ay, o + 1));
      AT(ayp, o + 3) = (3 * w - 2) * adotoa * AT(ay, o + 3) + 2 * w * k * sigma - k / 5 * (3 * 1.0 * G30_t - 2 * G11_t);
    }
  }
}

// ------------------------------------------------------------------------------------------------ derivs, hierarchies
// camb/equations_galileon.f90:2436-2466 (photons, polarisation), :2488-2513 (massless neutrinos), :2548-2587 (massive
// neutrinos per momentum node and their relativistic-correction multipoles): equation i of the state vector by lane
// (i-1) mod 8, each with the reference's expression for its block.  s_denl / s_polfac: the l tables
// (camb/equations_galileon.f90:516-527) in shared memory -- the lanes of a warp need them at different l.
struct LTab {
  const double* denl;
  const double* polfac;
  int n;
};
__device__ __forceinline__ double t_denl(const LTab& T, int l) { return l < T.n ? T.denl[l] : c_denl[l]; }
__device__ __forceinline__ double t_polfac(const LTab& T, int l) { return l < T.n ? T.polfac[l] : c_polfac[l]; }

__device__ __forceinline__ double hier_eq(const DevModel& M, const EV& e, const LTab& T, unsigned d, int i, int n,
                                          const double* __restrict__ ay, const double* __restrict__ sc,
                                          const double* __restrict__ vvt) {
  const double k = e.k;
  const unsigned t = d & 15u;
  const int l = (int)((d >> 4) & 255u);
  const double ym = AT(ay, i - 1), y0 = AT(ay, i);
  const double yp = (i < n) ? AT(ay, i + 1) : 0.0;
  const double dl = t_denl(T, l);
  const double dk = dl * k * l, dk2 = dl * k * 1.0 * (l + 1);  // denlk(l), denlk2(l)
  const double opacity = sc[SC_OPAC], cothxor = sc[SC_COTH];
  double v;
  if (t >= EQ_MNU_0 && t <= EQ_MNU_2T) {
    const int iq = (int)((d >> 12) & 15u), nu_i = (int)((d >> 16) & 15u);
    const double vv = vvt[nu_i * GC_MAX_NQ + iq];
    if (t == EQ_MNU_0)
      v = -k * (4.0 / 3.0 * sc[SC_Z] + vv * yp);
    else if (t == EQ_MNU_MID)
      v = vv * (dk * ym - dk2 * yp);
    else if (t == EQ_MNU_2)
      v = vv * (dk * ym - dk2 * yp) + k * 8.0 / 15.0 * sc[SC_SIGMA];
    else if (t == EQ_MNU_LAST)
      v = k * vv * ym - (l + 1) * cothxor * y0;
    else  // lmaxnu_tau == 2: y'(2) = -y'(0) - 3 cothxor y(2), with y'(0) = -k (4/3 z + v y(1))
      v = -(-k * (4.0 / 3.0 * sc[SC_Z] + vv * ym)) - 3 * cothxor * y0;
  } else if (t == EQ_PHOT_MID) {
    v = (dk * ym - dk2 * yp) - opacity * y0;
  } else if (t == EQ_PHOT_LAST) {
    v = k * ym - (l + 1) * cothxor * y0 - opacity * y0;
  } else if (t == EQ_POL_MID) {
    v = -opacity * y0 + (dk * ym - (t_polfac(T, l) * k * 1.0 * dl) * yp);
  } else if (t == EQ_POL_LAST) {
    v = -opacity * y0 + k * e.poltruncfac * ym - (l + 3) * cothxor * y0;
  } else if (t == EQ_NU_MID) {
    v = (dk * ym - dk2 * yp);
  } else if (t == EQ_NU_LAST) {
    v = k * ym - (l + 1) * cothxor * y0;
  } else {
    const double a = sc[SC_A], a2 = a * a;
    if (t == EQ_PERT_0) {
      v = +k * a2 * sc[SC_QR] - k * yp;
    } else if (t == EQ_PERT_MID) {
      const int ind2 = e.r_ix + l;
      v = -a2 * (dk * AT(ay, ind2 - 1) - dk2 * AT(ay, ind2 + 1)) + (dk * ym - dk2 * yp);
    } else {  // EQ_PERT_LAST
      const int ind2 = e.r_ix + l;
      v = k * (ym - a2 * AT(ay, ind2 - 1)) - (l + 1) * cothxor * y0;
    }
  }
  return v;
}

__device__ __forceinline__ void rhs_hier(const DevModel& M, const EV& e, const LTab& T, const unsigned* __restrict__ desc,
                                         const double* __restrict__ ay, double* __restrict__ ayp, const double* __restrict__ sc,
                                         const double* __restrict__ vvt, int n, int sub) {
  // two equations of this lane per pass: their loads and arithmetic are independent and overlap
  for (int i = sub + 1; i <= n; i += 2 * GC_OCT) {
    const int i2 = i + GC_OCT;
    const unsigned d1 = desc[i - 1];
    const unsigned d2 = (i2 <= n) ? desc[i2 - 1] : 0u;
    double v1 = 0, v2 = 0;
    if (d1 & 15u) v1 = hier_eq(M, e, T, d1, i, n, ay, sc, vvt);
    if (d2 & 15u) v2 = hier_eq(M, e, T, d2, i2, n, ay, sc, vvt);
    if (d1 & 15u) AT(ayp, i) = v1;
    if (d2 & 15u) AT(ayp, i2) = v2;
  }
}

// the descriptor table of the current layout (after GetNumEqns / every switch)
__device__ inline void build_desc(const DevModel& M, const EV& e, unsigned* desc) {
  for (int i = 0; i < e.neq; i++) desc[i] = EQ_SCALAR;
  if (!e.no_phot && !e.TightCoupling) {
    for (int l = 3; l <= e.lmaxg - 1; l++) desc[e.g_ix + l - 1] = EQ_PHOT_MID | (l << 4);
    desc[e.g_ix + e.lmaxg - 1] = EQ_PHOT_LAST | (e.lmaxg << 4);
    for (int l = 3; l <= e.lmaxgpol - 1; l++) desc[e.polind + l - 1] = EQ_POL_MID | (l << 4);
    desc[e.polind + e.lmaxgpol - 1] = EQ_POL_LAST | (e.lmaxgpol << 4);
  }
  if (!e.no_nu && !e.high_ktau) {
    for (int l = 3; l <= e.lmaxnr - 1; l++) desc[e.r_ix + l - 1] = EQ_NU_MID | (l << 4);
    desc[e.r_ix + e.lmaxnr - 1] = EQ_NU_LAST | (e.lmaxnr << 4);
  }
  if (M.Num_Nu_massive == 0) return;
  for (int nu_i = 0; nu_i < M.n_eig; nu_i++) {
    if (e.MassiveNuApprox[nu_i]) continue;
    int ind = e.nu_ix[nu_i];
    const int lm = e.lmaxnu_tau[nu_i];
    for (int iq = 0; iq < e.nq[nu_i]; iq++) {
      const unsigned tag = ((unsigned)iq << 12) | ((unsigned)nu_i << 16);
      desc[ind - 1] = EQ_MNU_0 | tag;
      desc[ind] = EQ_MNU_MID | (1 << 4) | tag;
      if (lm == 2) {
        desc[ind + 1] = EQ_MNU_2T | (2 << 4) | tag;
      } else {
        desc[ind + 1] = EQ_MNU_2 | (2 << 4) | tag;
        for (int l = 3; l <= lm - 1; l++) desc[ind + l - 1] = EQ_MNU_MID | (l << 4) | tag;
        desc[ind + lm - 1] = EQ_MNU_LAST | (lm << 4) | tag;
      }
      ind += lm + 1;
    }
  }
  if (e.has_nu_rel) {
    const int ind = e.nu_pert_ix;
    desc[ind - 1] = EQ_PERT_0;
    for (int l = 1; l <= e.lmaxnu_pert - 1; l++) desc[ind + l - 1] = EQ_PERT_MID | (l << 4);
    desc[ind + e.lmaxnu_pert - 1] = EQ_PERT_LAST | (e.lmaxnu_pert << 4);
  }
}

// ------------------------------------------------------------------------------------------------ control flow
struct Switches {
  double ktau, no_nu, no_phot, nu_massless, nu_nonrel, nu_massive, next;
};

// camb/equations_galileon.f90:328-364
__device__ inline void compute_switches(const DevModel& M, const EV& e, Switches& s) {
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

// the switch bodies of camb/equations_galileon.f90:370-463 (y in the state column, scratch = the other one)
__device__ __noinline__ void do_switch(const ItemMem m) {
  Item& L = *m.it;
  const DevModel& M = *L.M;
  EV& e = L.e;
  Switches s;
  compute_switches(M, e, s);
  const double next_switch = s.next;
  const double noSwitch = M.tau0 + 1;
  double* y = m.Y();
  double* yout = m.W();
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
  build_desc(M, e, m.desc);
}

// Run the item's control flow (lane 0 of its octet) until it needs the warp: a trial step (req = 1), a single
// evaluation of y' for a pending output (req = 2), or nothing any more (req = 0, phase = PH_DONE).
__device__ __noinline__ void advance(const ItemMem m) {
  Item& L = *m.it;
  const DevModel& M = *L.M;
  EV& e = L.e;
  const int S = M.SourceNum;
  const double tol = M.tol1;
  L.req = 0;
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
        if ((stops || zero) && L.out_pending) {  // no further step from this state: evaluate y' for the output now
          L.req = 2;
          return;
        }
        if (stops) {
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
        compute_switches(M, e, s);
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
        if (L.out_pending) {  // the next k1 will be taken with the new layout: evaluate y' for the output now
          L.req = 2;
          return;
        }
        do_switch(m);
        L.ynorm_valid = 0;
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
        L.phase = PH_DV_STEP;
        break;
      }
      case PH_DV_STEP: {  // step-size selection, camb/subroutines.f90:800-925 (independent of k1, which stage 1 evaluates)
        const int n = e.neq;
        const double x = L.tau, xend = L.target;
        double temp = L.ynorm;
        if (!L.ynorm_valid) {  // first step after initial() or a switch; otherwise the final row keeps it current
          const double* y = m.Y();
          temp = 0.0;
          for (int i = 1; i <= n; i++) temp = fmax(temp, fabs(AT(y, i)));
          L.ynorm = temp;
          L.ynorm_valid = 1;
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
        // what the reference evaluates in this trial: k1 unless the previous one was rejected (:1095-1110), k2..k8
        const unsigned nev = (L.ind != 6) ? 8u : 7u;
        L.n_derivs += nev;
        L.sum_n += nev * (unsigned)n;
        if (++L.n_trial > 50000000u) {  // runaway guard: no physical mode needs this many trial steps
          L.status = 105;
          L.phase = PH_DONE;
          return;
        }
        L.req = 1;
        L.phase = PH_DV_FINAL;
        return;
      }
      case PH_DV_FINAL: {  // accept/reject, camb/subroutines.f90:1021-1110
        L.est = L.errn * L.hmag * 1.0;
        L.ind = 5;
        if (L.est > tol) L.ind = 6;
        if (L.ind != 6) {
          L.tau = L.xtrial;
          L.ysel ^= 1;  // y <- trial solution: the two state columns trade places
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
        L.phase = PH_DV_STEP;
        break;
      }
      case PH_OUT_MARK: {  // an output time has been reached (camb/cmbmain.f90:1038 -> output)
        L.out_pending = 1;
        L.out_j = L.j;
        L.n_out++;
        L.n_derivs++;  // the derivs call of output(), camb/equations_galileon.f90:1297
        L.sum_n += (unsigned)e.neq;
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
          transfer_outputs(L.M, &e, m.Y(), L.tau, M.Transfer + ((size_t)L.itf * (M.nk + M.nk_tr) + L.kidx) * GC_TRANSFER_MAX);
          L.n_derivs++;  // the derivs call of outtransf, camb/equations_galileon.f90:2090
          L.sum_n += (unsigned)e.neq;
          L.itf++;
          if (kind == 1 && L.j < M.nt && L.itf < M.ntf && M.ts[(size_t)L.j * 8] > M.tautf[L.itf]) again = true;
        }
        L.phase = again ? PH_TF_CHECK : PH_NEXT_OUTPUT;
        break;
      }
#endif
      default:
        return;
    }
  }
}

// DoSourcek / GetTauStart / initial, camb/cmbmain.f90:634-689 (lane 0 of the octet)
__device__ __noinline__ void init_item(const ItemMem m, const DevModel* models, int model, int kidx, int slot, int NVP) {
  Item& L = *m.it;
  L.M = models + model;
  const DevModel& M = *L.M;
  L.kidx = kidx;
  L.slot = slot;
  L.status = 0;
  L.n_derivs = L.sum_n = L.n_out = L.n_shared = L.n_exec = L.n_trial = 0;
  L.req = 0;
  L.ysel = 0;
  L.ynorm_valid = 0;
  L.out_pending = 0;
  L.out_j = 0;
  L.ynorm = L.ynorm_new = L.errn = 0;
  L.win_i0 = -1000;
#ifdef GC_WITH_TRANSFER
  L.itf = 0;
  L.tf_stop = 0;
  // k indices beyond the source grid are the transfer-only wavenumbers (camb/cmbmain.f90:1164-1177)
  L.e.TransferOnly = kidx >= M.nk;
  L.e.q = L.e.TransferOnly ? M.k_tr[kidx - M.nk] : M.k[kidx];
#else
  L.e.q = M.k[kidx];
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
  if (L.e.nvar > NVP) {
    L.status = 103;
    L.phase = PH_DONE;
    return;
  }
  // GetTauStart, camb/cmbmain.f90:634-660
  double taustart = 0.001 / L.e.q;
  taustart = fmin(taustart, 0.1);
  if (M.Num_Nu_massive > 0) {
    double mx = 0;
    for (int i = 0; i < M.n_eig; i++) mx = fmax(mx, M.nu_masses[i]);
    taustart = fmin(taustart, 1.e-3 / mx / M.adotrad);
  }
  // w = 0 (camb/cmbmain.f90:946)
  for (int i = 0; i < GC_NCOLS * NVP; i++) m.cols[i] = 0;
  initial(M, L.e, m.Y(), taustart);
  build_desc(M, L.e, m.desc);
  L.tau = taustart;
  L.ind = 1;
  L.j = 1;
  L.c20 = 0;
  L.c21 = 0;
  L.nfail = 0;
  L.hmag = 0;
  L.est = 0;
  L.hmin = 0;
  L.htrial = 0;
  // Src(:,:,1) = 0 (never written by the time loop, camb/cmbmain.f90:1031)
#ifdef GC_WITH_TRANSFER
  if (!L.e.TransferOnly)
#endif
    for (int s = 0; s < M.SourceNum; s++) M.Src[(size_t)s * M.nk + kidx] = 0;
  L.phase = PH_NEXT_OUTPUT;
}

// ------------------------------------------------------------------------------------------------ one trial step
// One converged pass of the warp: a trial step (req = 1) or a single right-hand-side evaluation (req = 2) for every
// octet that asked for one.  See the header of this file for the three phases.
__device__ __forceinline__ void warp_step(const ItemMem& m, const LTab& T, int sub
#ifdef GC_PROFILE_PHASES
                                          , long long* prof, long long& t_last
#endif
) {
  Item& L = *m.it;
  const int req = L.req;
  const bool active = req != 0, full = req == 1;
  const DevModel& M = *L.M;  // model constants: read-only global, the same lines for the whole octet (L1 resident)
  const bool gal = active && M.use_galileon != 0;
  const int n = active ? L.e.neq : 0;
  const double x = L.tau, h = full ? L.htrial : 0.0;
  const double temp = h / 1398169080000.0;
  const double* __restrict__ Y = m.Y();
  double* __restrict__ W = m.W();
  const bool tight = L.e.TightCoupling;
  // ---- stage the rows of the background spline this step can touch (a grows monotonically through the stages)
  {
    int i0 = L.win_i0;
    bool reload = false;
    if (gal) {
      const double a0 = Y[0];
      // keep the window while the step starts inside its first rows; otherwise restart it one row below the guess
      if (!(i0 >= 0 && a0 >= L.galwin[0] && a0 < L.galwin[3 * 6])) {
        i0 = (int)(log(a0 / M.gal_a0) * M.gal_lnq_inv) - 1;
        i0 = max(0, min(i0, M.ngal - GC_OCT));
        reload = true;
      }
    }
    __syncwarp();
    if (reload) {
      const double* r = M.gal + (size_t)(i0 + sub) * 5;
#pragma unroll
      for (int c = 0; c < 5; c++) L.galwin[sub * 6 + c] = GC_LDG(r + c);
    }
    __syncwarp();
    if (reload) {
      if (sub < GC_OCT - 1) L.galwin[sub * 6 + 5] = 1.0 / (L.galwin[(sub + 1) * 6] - L.galwin[sub * 6]);
      if (sub == 0) L.win_i0 = i0;
    }
    __syncwarp();
  }
  GC_TICK(0);
  // ---- phase A: the scale factor through the eight stages (every lane of the octet, redundantly)
  BG b;
  double hbar_own = 0, xgal_own = 0;
  double ka[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
    const double y1 = Y[0];
    const int win_i0 = L.win_i0;
    const int nst = full ? 8 : 1;
#pragma unroll 1
    for (int s = 0; s < nst; s++) {
      double as;
      switch (s) {  // stage input of the scale factor: the same tableau rows as every other element
        case 0: as = y1; break;
        case 1: as = row_val<0>(y1, temp, ka); break;
        case 2: as = row_val<1>(y1, temp, ka); break;
        case 3: as = row_val<2>(y1, temp, ka); break;
        case 4: as = row_val<3>(y1, temp, ka); break;
        case 5: as = row_val<4>(y1, temp, ka); break;
        case 6: as = row_val<5>(y1, temp, ka); break;
        default: as = row_val<6>(y1, temp, ka); break;
      }
      double hb = 0, xg = 0, rn[GC_EV_NU] = {0}, pn[GC_EV_NU] = {0};
      const double ad = adotoa_at(M, gal, as, L.galwin, win_i0, hb, xg, rn, pn);
      const double kav = ad * as;
      switch (s) {
        case 0: ka[0] = kav; break;
        case 1: ka[1] = kav; break;
        case 2: ka[2] = kav; break;
        case 3: ka[3] = kav; break;
        case 4: ka[4] = kav; break;
        case 5: ka[5] = kav; break;
        case 6: ka[6] = kav; break;
        default: ka[7] = kav; break;
      }
      if (sub == s) {
        b.a = as;
        b.adotoa = ad;
        b.ka = kav;
        hbar_own = hb;
        xgal_own = xg;
#pragma unroll
        for (int i = 0; i < GC_EV_NU; i++) {
          b.rhonu[i] = rn[i];
          b.pnu[i] = pn[i];
        }
      }
    }
  }
  GC_TICK(1);
  // ---- phase B: the background record of stage `sub`, eight stages side by side
  {
    double ts;
    switch (sub) {
      case 0: ts = x; break;
      case 1: ts = x + h / 6.0; break;
      case 2: ts = x + h * (4.0 / 15.0); break;
      case 3: ts = x + h * (2.0 / 3.0); break;
      case 4: ts = x + h * (5.0 / 6.0); break;
      case 5: ts = x + h; break;
      case 6: ts = x + h / 15.0; break;
      default: ts = x + h; break;
    }
    b.tau = ts;
    if (active && (sub == 0 || full)) bg_eval(M, L.e, tight, hbar_own, xgal_own, !(gal && b.a >= 1e-10) && M.Num_Nu_massive != 0, b);
  }
  GC_TICK(2);
  // ---- phase C: the stages in order
  const int i0 = sub + 1;  // first element of this lane
#pragma unroll 1
  for (int s = 0; s < 8; s++) {
    const bool on = active && (s == 0 || full);
    const double* __restrict__ in = (s == 0) ? Y : W;
    if (on && s > 0) {
      const int i1 = (sub == 0) ? i0 + GC_OCT : i0;  // element 1 (the scale factor) comes from phase A
      switch (s) {
        case 1: combine_row<0>(m, i1, n, temp); break;
        case 2: combine_row<1>(m, i1, n, temp); break;
        case 3: combine_row<2>(m, i1, n, temp); break;
        case 4: combine_row<3>(m, i1, n, temp); break;
        case 5: combine_row<4>(m, i1, n, temp); break;
        case 6: combine_row<5>(m, i1, n, temp); break;
        default: combine_row<6>(m, i1, n, temp); break;
      }
      if (sub == s) W[0] = b.a;
    }
    __syncwarp();
    GC_TICK(3);
    double* __restrict__ Kc = m.K(s);
    if (on && sub == s) rhs_scalar(&M, &L.e, b, in, Kc, L.sc, L.vv, L.ivv, L.nub);
    __syncwarp();
    GC_TICK(4);
    if (on) rhs_hier(M, L.e, T, m.desc, in, Kc, L.sc, L.vv, n, sub);
    __syncwarp();
    GC_TICK(5);
    if (s == 0) {
      // k1 at (tau, y) is the y' that output() evaluates (camb/equations_galileon.f90:1297): form the sources now
      if (active && L.out_pending && sub == 0) {
        const DevModel& Mg = *L.M;
        const int S = Mg.SourceNum;
        double src[3] = {0, 0, 0};
        output_sources(L.M, &L.e, Y, Kc, Mg.ts + (size_t)(L.out_j - 1) * 8, Mg.ts[(size_t)(L.out_j - 1) * 8], src);
        for (int q = 0; q < S; q++) Mg.Src[((size_t)(L.out_j - 1) * S + q) * Mg.nk + L.kidx] = src[q];
        L.out_pending = 0;
        if (full) L.n_shared++;
      }
      __syncwarp();
      GC_TICK(6);
    }
    if (!__any_sync(0xffffffffu, full)) break;  // the warp only had single evaluations to do
  }
  // ---- trial solution and error estimate (camb/subroutines.f90:995-1019), octet reduction of the two norms
  double errn = 0.0, ynew = 0.0;
  if (full) {
    const double* kc[8];
#pragma unroll
    for (int c = 0; c < 8; c++) kc[c] = m.K(c);
    for (int i = i0; i <= n; i += GC_OCT) {
      double k[8];
      if (i == 1) {
#pragma unroll
        for (int c = 0; c < 8; c++) k[c] = ka[c];
      } else {
#pragma unroll
        for (int c = 0; c < 8; c++) k[c] = (c == 1) ? 0.0 : kc[c][i - 1];
      }
      const double yv = Y[i - 1];
      const double v = row_val<7>(yv, temp, k);
      const double w2 = err_val(k);
      W[i - 1] = v;
      errn = fmax(errn, fabs(w2) / fmax(1.0, fabs(yv)));
      ynew = fmax(ynew, fabs(v));
    }
  }
#pragma unroll
  for (int o = 1; o < GC_OCT; o <<= 1) {
    errn = fmax(errn, __shfl_xor_sync(0xffffffffu, errn, o));
    ynew = fmax(ynew, __shfl_xor_sync(0xffffffffu, ynew, o));
  }
  if (active && sub == 0) {
    if (full) {
      L.errn = errn;
      L.ynorm_new = ynew;
      L.n_exec += 8;
    } else {
      L.n_exec += 1;
    }
  }
  __syncwarp();
  GC_TICK(7);
}

// groups[g*4 + o] = (model index, k index) of the item octet o of group g works on, model = -1: none.
__global__ void __launch_bounds__(256, 1)
    evolve_sources_kernel(const DevModel* __restrict__ models, const int2* __restrict__ groups, int ngroups, int NVP,
                          unsigned* __restrict__ queue, int* __restrict__ status, unsigned long long* __restrict__ counters) {
  extern __shared__ __align__(16) unsigned char gc_smem_raw[];
  const int lane = threadIdx.x & 31, sub = lane & (GC_OCT - 1), oct = lane >> 3;
  const int warp = threadIdx.x >> 5;
  // l tables first (camb/equations_galileon.f90:516-527), then the item slots
  LTab T;
  {
    double* t = reinterpret_cast<double*>(gc_smem_raw);
    for (int l = threadIdx.x; l < GC_LTAB; l += blockDim.x) {
      t[l] = c_denl[l];
      t[GC_LTAB + l] = c_polfac[l];
    }
    T.denl = t;
    T.polfac = t + GC_LTAB;
    T.n = GC_LTAB;
  }
  __syncthreads();
  const size_t stride = item_stride_bytes(NVP);
  unsigned char* base = gc_smem_raw + GC_LTAB_BYTES + (size_t)(warp * GC_IPW + oct) * stride;
  ItemMem m;
  m.it = reinterpret_cast<Item*>(base);
  m.desc = reinterpret_cast<unsigned*>(base + ((sizeof(Item) + 7) / 8) * 8);
  m.cols = reinterpret_cast<double*>(base + ((sizeof(Item) + 7) / 8) * 8 + (((size_t)NVP * 4 + 7) / 8) * 8);
  m.NVP = NVP;
  Item& L = *m.it;
  for (;;) {
    unsigned g = 0;
    if (lane == 0) g = atomicAdd(queue, 1u);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= (unsigned)ngroups) break;
    const int2 it = groups[(size_t)g * GC_IPW + oct];
    const bool have = it.x >= 0;
    if (sub == 0) {
      L.req = 0;
      L.phase = PH_DONE;
      L.status = 0;
      L.e.neq = 0;
      L.e.TightCoupling = false;
#ifdef GC_WITH_TRANSFER
      L.e.TransferOnly = false;
      L.itf = 0;
      L.tf_stop = 0;
#endif
      L.ysel = 0;
      L.tau = 1.0;
      L.htrial = 0.0;
      L.out_pending = 0;
      L.win_i0 = -1000;
      L.n_derivs = L.sum_n = L.n_out = L.n_shared = L.n_exec = L.n_trial = 0;
      L.M = models;  // an idle octet still reads a few model fields in the converged code
      if (have) init_item(m, models, it.x, it.y, (int)g * GC_IPW + oct, NVP);
    }
    __syncwarp();
#ifdef GC_PROFILE_PHASES
    long long prof[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_last = clock64();
#endif
    for (;;) {
      if (sub == 0 && L.phase != PH_DONE) advance(m);
      __syncwarp();
#ifdef GC_PROFILE_PHASES
      GC_TICK(8);
#endif
      if (!__any_sync(0xffffffffu, L.req != 0)) break;
#ifdef GC_PROFILE_PHASES
      warp_step(m, T, sub, prof, t_last);
#else
      warp_step(m, T, sub);
#endif
    }
#ifdef GC_PROFILE_PHASES
    if (lane == 0)
      for (int i = 0; i < 9; i++) atomicAdd(&counters[8 + i], (unsigned long long)prof[i]);
#endif
    if (sub == 0) {
      status[(size_t)g * GC_IPW + oct] = have ? L.status : 0;
      if (have) {
        atomicAdd(&counters[0], (unsigned long long)L.n_derivs);
        atomicAdd(&counters[1], (unsigned long long)L.sum_n);
        atomicAdd(&counters[2], (unsigned long long)L.n_out);
        atomicAdd(&counters[4], (unsigned long long)L.n_shared);
        atomicAdd(&counters[5], (unsigned long long)L.n_exec);
        atomicAdd(&counters[6], (unsigned long long)L.n_trial);
      }
    }
    __syncwarp();
  }
}

}  // namespace gc

extern "C" void GC_ENTRY(gc_upload_constants)() {
  double denl[260], polfac[260];
  for (int j = 0; j < 260; j++) {
    denl[j] = 1.0 / (2 * j + 1);
    polfac[j] = j >= 1 ? (double)((j + 3) * (j - 1)) / (j + 1) : 0.0;
  }
  cudaMemcpyToSymbol(c_denl, denl, sizeof(denl));
  cudaMemcpyToSymbol(c_polfac, polfac, sizeof(polfac));
}

// shared memory one work item needs for state vectors of up to NVP equations
extern "C" size_t GC_ENTRY(gc_evolve_item_bytes)(int NVP) { return gc::item_stride_bytes(NVP); }

// warps per CTA that fit the 227 KB of one SM (one persistent CTA per SM)
extern "C" int GC_ENTRY(gc_evolve_warps_per_cta)(int NVP) {
  const size_t per_warp = gc::item_stride_bytes(NVP) * GC_IPW;
  int w = (int)((227 * 1024 - GC_LTAB_BYTES) / per_warp);
  return std::max(0, std::min(w, 8));
}

extern "C" cudaError_t GC_ENTRY(gc_launch_evolve)(const DevModel* models, const int2* groups, int ngroups, int NVP, int num_sms,
                                        unsigned* queue, int* status, unsigned long long* counters, cudaStream_t stream) {
  int warps = GC_ENTRY(gc_evolve_warps_per_cta)(NVP);
  if (warps < 1) return cudaErrorInvalidValue;
  // few groups: spread them one warp per CTA over the SMs (latency mode) before stacking warps
  const int ctas = std::min(num_sms, std::max(1, ngroups));
  warps = std::max(1, std::min(warps, (ngroups + ctas - 1) / ctas));
  const size_t smem = GC_LTAB_BYTES + gc::item_stride_bytes(NVP) * GC_IPW * warps;
  cudaError_t e = cudaFuncSetAttribute(gc::evolve_sources_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(queue, 0, sizeof(unsigned), stream);
  if (e != cudaSuccess) return e;
  gc::evolve_sources_kernel<<<ctas, warps * 32, smem, stream>>>(models, groups, ngroups, NVP, queue, status, counters);
  return cudaGetLastError();
}
