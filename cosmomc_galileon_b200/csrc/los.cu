// K0 bessel_table, K2 spline_k, K3 los_integrate, K4 scalar_cls.
//
//   K0  GenerateBessels / bjl                     camb/bessels.f90:50-275
//   K2  InitSourceInterpolation + spline          camb/cmbmain.f90:1267-1278, camb/subroutines.f90:253-295
//   K3  InterpolateSources + DoFlatIntegration    camb/cmbmain.f90:1355-1434, :1447-1597
//   K4  CalcScalCls + InterpolateCls              camb/cmbmain.f90:2167-2267, :2450-2496, camb/modules.f90:1012-1089
#include <cstdio>

#include "model.h"

#define GC_SP(x) ((double)(x##f))

__constant__ GcRegions c_bess;  // BessRanges (cosmology independent), camb/bessels.f90:68-76

namespace gc {

// camb/utils.F90:81-111 ; returns the 1-based index like the reference
__device__ __forceinline__ int Ranges_IndexOf(const GcRegions& R, double tau) {
  for (int i = 0; i < R.count; i++) {
    const GcRegion& A = R.R[i];
    if (tau < A.High && tau >= A.Low) {
      if (A.IsLog) return A.start_index + (int)(log(tau / A.Low) / A.delta);
      return A.start_index + (int)((tau - A.Low) / A.delta);
    }
  }
  return R.npoints;  // tau >= Highest
}

// ------------------------------------------------------------------------------------------------ K0
// camb/bessels.f90:132-275
__device__ double bjl(int L, double X) {
  const double LN2 = 0.6931471805599453094, ONEMLN2 = 0.30685281944005469058277, PID2 = 1.5707963267948966192313217,
               PID4 = 0.78539816339744830961566084582, ROOTPI12 = 21.269446210866192327578,
               GAMMA1 = 2.6789385347077476336556, GAMMA2 = 1.3541179394264004169452;
  const double AX = fabs(X), AX2 = AX * AX;
  double JL;
  if (L < 7) {
    const double s = sin(AX), c = cos(AX);
    if (L == 0)
      JL = AX < 1e-1 ? 1.0 - AX2 / 6.0 * (1.0 - AX2 / 20.0) : s / AX;
    else if (L == 1)
      JL = AX < 2e-1 ? AX / 3.0 * (1.0 - AX2 / 10.0 * (1.0 - AX2 / 28.0)) : (s / AX - c) / AX;
    else if (L == 2)
      JL = AX < 3e-1 ? AX2 / 15.0 * (1.0 - AX2 / 14.0 * (1.0 - AX2 / 36.0)) : (-3.0 * c / AX - s * (1.0 - 3.0 / AX2)) / AX;
    else if (L == 3)
      JL = AX < 4e-1 ? AX * AX2 / 105.0 * (1.0 - AX2 / 18.0 * (1.0 - AX2 / 44.0))
                     : (c * (1.0 - 15.0 / AX2) - s * (6.0 - 15.0 / AX2) / AX) / AX;
    else if (L == 4)
      JL = AX < 6e-1 ? (AX2 * AX2) / 945.0 * (1.0 - AX2 / 22.0 * (1.0 - AX2 / 52.0))
                     : (s * (1.0 - (45.0 - 105.0 / AX2) / AX2) + c * (10.0 - 105.0 / AX2) / AX) / AX;
    else if (L == 5)
      JL = AX < 1.0 ? (AX2 * AX2) * AX / 10395.0 * (1.0 - AX2 / 26.0 * (1.0 - AX2 / 60.0))
                    : (s * (15.0 - (420.0 - 945.0 / AX2) / AX2) / AX - c * (1.0 - (105.0 - 945.0 / AX2) / AX2)) / AX;
    else
      JL = AX < 1.0 ? (AX2 * AX2 * AX2) / 135135.0 * (1.0 - AX2 / 30.0 * (1.0 - AX2 / 68.0))
                    : (s * (-1.0 + (210.0 - (4725.0 - 10395.0 / AX2) / AX2) / AX2) +
                       c * (-21.0 + (1260.0 - 10395.0 / AX2) / AX2) / AX) /
                          AX;
  } else {
    const double NU = 0.5 + L, NU2 = NU * NU;
    if (AX < 1e-40) {
      JL = 0.0;
    } else if ((AX2 / L) < 5e-1) {
      JL = exp(L * log(AX / NU) - LN2 + NU * ONEMLN2 - (1.0 - (1.0 - 3.5 / NU2) / NU2 / 30.0) / 12.0 / NU) / NU *
           (1.0 - AX2 / (4.0 * NU + 4.0) * (1.0 - AX2 / (8.0 * NU + 16.0) * (1.0 - AX2 / (12.0 * NU + 36.0))));
    } else if (((double)L * (double)L / AX) < 5e-1) {
      const double BETA = AX - PID2 * (L + 1);
      JL = (cos(BETA) * (1.0 - (NU2 - 0.25) * (NU2 - 2.25) / 8.0 / AX2 * (1.0 - (NU2 - 6.25) * (NU2 - 12.25) / 48.0 / AX2)) -
            sin(BETA) * (NU2 - 0.25) / 2.0 / AX *
                (1.0 - (NU2 - 2.25) * (NU2 - 6.25) / 24.0 / AX2 * (1.0 - (NU2 - 12.25) * (NU2 - 20.25) / 80.0 / AX2))) /
           AX;
    } else {
      const double L3 = pow(NU, GC_SP(0.325));
      if (AX < NU - GC_SP(1.31) * L3) {
        const double COSB = NU / AX, SX = sqrt(NU2 - AX2), COTB = NU / SX, SECB = AX / NU;
        const double BETA = log(COSB + SX / AX);
        const double COT3B = COTB * COTB * COTB, COT6B = COT3B * COT3B, SEC2B = SECB * SECB;
        const double EXPTERM =
            ((2.0 + 3.0 * SEC2B) * COT3B / 24.0 -
             ((4.0 + SEC2B) * SEC2B * COT6B / 16.0 +
              ((16.0 - (1512.0 + (3654.0 + 375.0 * SEC2B) * SEC2B) * SEC2B) * COT3B / 5760.0 +
               (32.0 + (288.0 + (232.0 + 13.0 * SEC2B) * SEC2B) * SEC2B) * SEC2B * COT6B / 128.0 / NU) *
                  COT6B / NU) /
                 NU) /
            NU;
        JL = sqrt(COTB * COSB) / (2.0 * NU) * exp(-NU * BETA + NU / COTB - EXPTERM);
      } else if (AX > NU + GC_SP(1.48) * L3) {
        const double COSB = NU / AX, SX = sqrt(AX2 - NU2), COTB = NU / SX, SECB = AX / NU;
        const double BETA = acos(COSB);
        const double COT3B = COTB * COTB * COTB, COT6B = COT3B * COT3B, SEC2B = SECB * SECB;
        const double TRIGARG = NU / COTB - NU * BETA - PID4 -
                               ((2.0 + 3.0 * SEC2B) * COT3B / 24.0 +
                                (16.0 - (1512.0 + (3654.0 + 375.0 * SEC2B) * SEC2B) * SEC2B) * COT3B * COT6B / 5760.0 / NU2) /
                                   NU;
        const double EXPTERM = ((4.0 + SEC2B) * SEC2B * COT6B / 16.0 -
                                (32.0 + (288.0 + (232.0 + 13.0 * SEC2B) * SEC2B) * SEC2B) * SEC2B * (COT6B * COT6B) / 128.0 / NU2) /
                               NU2;
        JL = sqrt(COTB * COSB) / NU * exp(-EXPTERM) * cos(TRIGARG);
      } else {
        const double BETA = AX - NU, BETA2 = BETA * BETA, SX = 6.0 / AX, SX2 = SX * SX;
        const double SECB = pow(SX, 0.3333333333333333), SEC2B = SECB * SECB;
        JL = (GAMMA1 * SECB + BETA * GAMMA2 * SEC2B - (BETA2 / 18.0 - 1.0 / 45.0) * BETA * SX * SECB * GAMMA1 -
              ((BETA2 - 1.0) * BETA2 / 36.0 + 1.0 / 420.0) * SX * SEC2B * GAMMA2 +
              (((BETA2 / 1620.0 - 7.0 / 3240.0) * BETA2 + 1.0 / 648.0) * BETA2 - 1.0 / 8100.0) * SX2 * SECB * GAMMA1 +
              (((BETA2 / 4536.0 - 1.0 / 810.0) * BETA2 + 19.0 / 11340.0) * BETA2 - 13.0 / 28350.0) * BETA * SX2 * SEC2B * GAMMA2 -
              ((((BETA2 / 349920.0 - 1.0 / 29160.0) * BETA2 + 71.0 / 583200.0) * BETA2 - 121.0 / 874800.0) * BETA2 +
               7939.0 / 224532000.0) *
                  BETA * SX2 * SX * SECB * GAMMA1) *
             sqrt(SX) / ROOTPI12;
      }
    }
  }
  if (X < 0 && (L % 2) != 0) JL = -JL;
  return JL;
}

// bess[(j*nx + i)] = {ajl, ajlpr}; this kernel fills .x (ajl), camb/bessels.f90:89-111
__global__ void bessel_fill_kernel(double2* __restrict__ bess, const double* __restrict__ xs, const int* __restrict__ ls,
                                   int nx, int nl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= nx || j >= nl) return;
  const int l = ls[j];
  const double x = xs[i];
  double xlim = 0.05 * l;
  xlim = fmax(xlim, 35.0);
  xlim = l - xlim;
  double v = 0;
  if (x > xlim) {
    if ((l == 3 && x <= GC_SP(0.2)) || (l > 3 && x < 0.5) || (l > 5 && x < 1.0))
      v = 0;
    else
      v = bjl(l, x);
  }
  bess[(size_t)j * nx + i].x = v;
}

// natural cubic spline second derivatives (camb/subroutines.f90:253-295) of column j; one thread per l,
// u[] scratch in global memory ([nl][nx]).
__global__ void bessel_spline_kernel(double2* __restrict__ bess, const double* __restrict__ xs, double* __restrict__ u,
                                     int nx, int nl) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nl) return;
  double2* b = bess + (size_t)j * nx;
  double* uu = u + (size_t)j * nx;
  double d1r = (b[1].x - b[0].x) / (xs[1] - xs[0]);
  b[0].y = 0.0;
  uu[0] = 0.0;
  for (int i = 1; i < nx - 1; i++) {
    const double d1l = d1r;
    d1r = (b[i + 1].x - b[i].x) / (xs[i + 1] - xs[i]);
    const double xxdiv = 1.0 / (xs[i + 1] - xs[i - 1]);
    const double sig = (xs[i] - xs[i - 1]) * xxdiv;
    const double xp = 1.0 / (sig * b[i - 1].y + 2.0);
    b[i].y = (sig - 1.0) * xp;
    uu[i] = (6.0 * (d1r - d1l) * xxdiv - sig * uu[i - 1]) * xp;
  }
  b[nx - 1].y = (0.0 - 0.0 * uu[nx - 2]) / (0.0 * b[nx - 2].y + 1.0);
  for (int i = nx - 2; i >= 0; i--) b[i].y = b[i].y * b[i + 1].y + uu[i];
}

// ------------------------------------------------------------------------------------------------ K2
// One thread per (model, time step, source): natural spline along k.  Src[t][s][k] is k-contiguous.
#define GC_MAX_NK 768
__global__ void spline_k_kernel(const DevModel* __restrict__ models, int nmodels) {
  const int m = blockIdx.y;
  const DevModel& M = models[m];
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M.nt * M.SourceNum) return;
  const int n = M.nk;
  const double* x = M.k;
  const double* y = M.Src + (size_t)row * n;
  double* d2 = M.ddSrc + (size_t)row * n;
  double u[GC_MAX_NK];
  double d1r = (y[1] - y[0]) / (x[1] - x[0]);
  d2[0] = 0.0;
  u[0] = 0.0;
  for (int i = 1; i < n - 1; i++) {
    const double d1l = d1r;
    d1r = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
    const double xxdiv = 1.0 / (x[i + 1] - x[i - 1]);
    const double sig = (x[i] - x[i - 1]) * xxdiv;
    const double xp = 1.0 / (sig * d2[i - 1] + 2.0);
    d2[i] = (sig - 1.0) * xp;
    u[i] = (6.0 * (d1r - d1l) * xxdiv - sig * u[i - 1]) * xp;
  }
  d2[n - 1] = (0.0 - 0.0 * u[n - 2]) / (0.0 * d2[n - 2] + 1.0);
  for (int i = n - 2; i >= 0; i--) d2[i] = d2[i] * d2[i + 1] + u[i];
}

// ------------------------------------------------------------------------------------------------ K3
// One CTA per (model, integration q).  Shared memory holds this q's per-time-step operands
// (Source_q(n,1:S), aa, fac, dtau, bes_index); each warp then owns multipoles j = warp, warp+W, ... and its lanes
// stride over the time window, reading the interleaved {ajl, ajlpr} Bessel rows and reducing with shuffles in
// a fixed order (deterministic).  Returns N_iter through a global counter.
__global__ void __launch_bounds__(256) los_kernel(const DevModel* __restrict__ models, const double2* __restrict__ bess,
                                                  const double* __restrict__ xs, const int* __restrict__ ls, int nx, int nl,
                                                  unsigned long long* __restrict__ n_iter_out, int q_lo) {
  extern __shared__ double smem[];
  const int m = blockIdx.y;
  const DevModel& M = models[m];
  const int q_ix = blockIdx.x + q_lo;  // 0-based; q_lo > 0: this device integrates a slice of the q grid
  if (q_ix >= M.nq) return;
  const int nt = M.nt, S = M.SourceNum, nk = M.nk;
  // shared layout (1-based time index n -> slot n-1)
  double* sq = smem;                // [S][nt]
  double* aa = sq + (size_t)S * nt;  // [nt]
  double* fac = aa + nt;             // [nt]
  double* dtau = fac + nt;           // [nt]
  int* bes_index = (int*)(dtau + nt);
  __shared__ int s_step;
  const double q = M.q[q_ix];
  const double* kq = M.k;
  // InterpolateSources: bracket (camb/cmbmain.f90:1366-1372), evaluated redundantly by every thread
  int klo = 1;
  {
    int lo = 1, hi = nk - 1;  // smallest klo in [1, nk-1] with q <= k(klo+1), else nk-1
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (q > kq[mid])  // k(mid+1) in 1-based == kq[mid]
        lo = mid + 1;
      else
        hi = mid;
    }
    klo = lo;
  }
  const int khi = klo + 1;
  const double ho = kq[khi - 1] - kq[klo - 1];
  const double a0 = (kq[khi - 1] - q) / ho, b0 = (q - kq[klo - 1]) / ho;
  const double ho2o6 = (ho * ho) / 6;
  const double a03 = (a0 * a0 * a0 - a0), b03 = (b0 * b0 * b0 - b0);
  if (threadIdx.x == 0) s_step = 2;
  __syncthreads();
  int mystep = 2;
  for (int i = threadIdx.x + 1; i <= nt; i += blockDim.x) {
    const double tau_i = M.ts[(size_t)(i - 1) * 8];
    dtau[i - 1] = M.ts[(size_t)(i - 1) * 8 + 1];
    if (i == 1) {
      for (int s = 0; s < S; s++) sq[(size_t)s * nt] = 0;
      continue;
    }
    const double xf = q * (M.tau0 - tau_i);
    if ((M.WantLateTime || q * tau_i < M.max_etak_scalar) && xf > 1.e-8) {
      mystep = i;
      for (int s = 0; s < S; s++) {
        const size_t off = ((size_t)(i - 1) * S + s) * nk;
        sq[(size_t)s * nt + i - 1] = a0 * M.Src[off + klo - 1] + b0 * M.Src[off + khi - 1] +
                                     (a03 * M.ddSrc[off + klo - 1] + b03 * M.ddSrc[off + khi - 1]) * ho2o6;
      }
    } else {
      for (int s = 0; s < S; s++) sq[(size_t)s * nt + i - 1] = 0;
    }
  }
  atomicMax(&s_step, mystep);
  __syncthreads();
  const int SourceSteps = s_step;
  // per-time-step Bessel operands, camb/cmbmain.f90:1516-1524
  for (int j = threadIdx.x + 1; j <= SourceSteps; j += blockDim.x) {
    const double xf = fabs(q * (M.tau0 - M.ts[(size_t)(j - 1) * 8]));
    const int bes_ix = Ranges_IndexOf(c_bess, xf);  // 1-based
    const double p0 = xs[bes_ix - 1], p1 = xs[bes_ix];  // BessRanges%points(bes_ix), (bes_ix+1)
    double f = p1 - p0;
    const double av = (p1 - xf) / f;
    f = (f * f) * av / 6;
    bes_index[j - 1] = bes_ix;
    aa[j - 1] = av;
    fac[j - 1] = f;
  }
  __syncthreads();
  // DoSourceIntegration, camb/cmbmain.f90:1447-1481 (flat)
  const double pi = 3.14159265358979323846264338328;
  int llmax = (int)lround(q * M.tau0);  // nu*chi0 with r = 1
  if (llmax < 15)
    llmax = 17;
  else
    llmax = (int)lround(q * (M.tau0 + 6 * pi / q));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const double tp2 = M.ts[8];  // TimeSteps%points(2)
  unsigned long long iters = 0;
  for (int j = warp; j < nl; j += nwarps) {
    const int l = ls[j];
    if (l > llmax) break;
    double xlim = 0.05 * l;
    xlim = fmax(xlim, 35.0);
    xlim = l - xlim;
    const double xlmax1 = 80 * l * M.AccuracyBoost;
    double tmin = M.tau0 - xlmax1 / q;
    tmin = fmax(tp2, tmin);
    double tmax = M.tau0 - xlim / q;
    tmax = fmin(M.tau0, tmax);
    tmin = fmax(tp2, tmin);
    if (tmax < tp2) break;
    const int n_lo = Ranges_IndexOf(M.TimeSteps, tmin);
    const int n_hi = min(SourceSteps, Ranges_IndexOf(M.TimeSteps, tmax));
    const double2* bj = bess + (size_t)j * nx - 1;  // 1-based x index
    double s1 = 0, s2 = 0, s3 = 0;
    bool DoInt = true;
    if (S > 2) {
      double qmax_int = max(850, l) * 3 * M.AccuracyBoost / M.tau0;
      if (M.HighAccuracyDefault) qmax_int = qmax_int * GC_SP(1.2);
      DoInt = q < qmax_int;
    }
    if (DoInt) {
      for (int n = n_lo + lane; n <= n_hi; n += 32) {
        const double a2 = aa[n - 1];
        const int bes_ix = bes_index[n - 1];
        const double2 b0v = bj[bes_ix], b1v = bj[bes_ix + 1];
        double J_l = a2 * b0v.x + (1 - a2) * (b1v.x - ((a2 + 1) * b0v.y + (2 - a2) * b1v.y) * fac[n - 1]);
        J_l = J_l * dtau[n - 1];
        s1 = s1 + sq[n - 1] * J_l;
        s2 = s2 + sq[(size_t)nt + n - 1] * J_l;
        if (S > 2) s3 = s3 + sq[(size_t)2 * nt + n - 1] * J_l;
      }
      if (lane == 0 && n_hi >= n_lo) iters += (unsigned long long)(n_hi - n_lo + 1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        s3 += __shfl_xor_sync(0xffffffffu, s3, o);
      }
    }
    if (S > 2) {
      const bool UseLimber = l > 400 * sqrt(M.AccuracyBoost);
      if (!DoInt || UseLimber) {  // Limber substitution for the lensing source, camb/cmbmain.f90:1580-1590
        double xf = M.tau0 - (l + 0.5) / q;
        if (xf < M.TimeSteps.Highest && xf > M.TimeSteps.Lowest) {
          const int n = Ranges_IndexOf(M.TimeSteps, xf);
          const double t0 = M.ts[(size_t)(n - 1) * 8], t1 = M.ts[(size_t)n * 8];
          xf = (xf - t0) / (t1 - t0);
          s3 = (sq[(size_t)2 * nt + n - 1] * (1 - xf) + xf * sq[(size_t)2 * nt + n]) * sqrt(pi / 2 / (l + 0.5)) / q;
        } else {
          s3 = 0;
        }
      }
    }
    if (lane == 0) {
      double* D = M.Delta + ((size_t)q_ix * nl + j) * S;
      D[0] = s1;
      D[1] = s2;
      if (S > 2) D[2] = s3;
    }
  }
  if (lane == 0 && iters) atomicAdd(n_iter_out, iters);
}

// zero Delta (l > llmax entries are never touched by los_kernel; reference: Init_ClTransfer camb/modules.f90:1175)
__global__ void zero_delta_kernel(const DevModel* __restrict__ models, int nl) {
  const DevModel& M = models[blockIdx.y];
  const size_t n = (size_t)M.SourceNum * nl * M.nq;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) M.Delta[i] = 0;
}

// ------------------------------------------------------------------------------------------------ K4
// camb/cmbmain.f90:2167-2267: one CTA per (model, l sample); fixed-order tree reduction over q.
__global__ void __launch_bounds__(256) cl_kernel(const DevModel* __restrict__ models, const int* __restrict__ ls, int nl) {
  const int m = blockIdx.y, j = blockIdx.x;
  const DevModel& M = models[m];
  const int S = M.SourceNum, nq = M.nq;
  const int ntype = S > 2 ? 6 : 3;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int q_ix = threadIdx.x; q_ix < nq; q_ix += blockDim.x) {
    const double ks = M.q[q_ix];
    const double dlnk = M.dq[q_ix] / ks;
    const double lnrat = log(ks / M.k_0_scalar);  // ScalarPower, camb/power_tilt.f90:149-157
    const double apowers = M.ScalarPowerAmp * exp(lnrat * (M.an - 1 + lnrat * (M.n_run / 2 + M.n_runrun / 6 * lnrat)));
    const double* D = M.Delta + ((size_t)q_ix * nl + j) * S;
    acc[0] += apowers * (D[0] * D[0]) * dlnk;
    acc[1] += apowers * (D[1] * D[1]) * dlnk;
    acc[2] += apowers * D[0] * D[1] * dlnk;
    if (S > 2) {
      acc[3] += apowers * (D[2] * D[2]) * dlnk;
      acc[4] += apowers * D[2] * D[0] * dlnk;
      acc[5] += apowers * D[2] * D[1] * dlnk;
    }
  }
  __shared__ double red[6][256];
  for (int t = 0; t < 6; t++) red[t][threadIdx.x] = acc[t];
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int t = 0; t < ntype; t++) red[t][threadIdx.x] += red[t][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double pi = 3.14159265358979323846264338328, twopi = 2 * pi, fourpi = 4 * pi;
    const double ell = (double)ls[j];
    const double ctnorm = (ell * ell - 1) * (ell + 2) * ell;
    const double dbletmp = (ell * (ell + 1)) / twopi * fourpi;
    M.iCl[0 * nl + j] = red[0][0] * dbletmp;
    M.iCl[1 * nl + j] = red[1][0] * dbletmp * ctnorm;
    M.iCl[2 * nl + j] = red[2][0] * dbletmp * sqrt(ctnorm);
    if (S > 2) {
      M.iCl[3 * nl + j] = M.ALens * red[3][0] * fourpi * (ell * ell * ell * ell);  // camb/cmbmain.f90:2254-2259
      M.iCl[4 * nl + j] = sqrt(M.ALens) * red[4][0] * fourpi * (ell * ell * ell);
      M.iCl[5 * nl + j] = sqrt(M.ALens) * red[5][0] * fourpi * (ell * ell * ell) * sqrt(ctnorm);
    }
  }
}

// camb/modules.f90:1012-1089: one CTA per (model, type).  tmpl = [3][lmax_tmpl-1] or NULL.
#define GC_MAX_NL 512
__global__ void __launch_bounds__(256) cl_interp_kernel(const DevModel* __restrict__ models, const int* __restrict__ ls, int nl,
                                                        int lmax, const double* __restrict__ tmpl, int lmax_tmpl) {
  const int m = blockIdx.y, type = blockIdx.x;
  const DevModel& M = models[m];
  __shared__ double v[GC_MAX_NL], d2[GC_MAX_NL], u[GC_MAX_NL];
  const int nL = lmax - 1;
  const bool use_t = tmpl != nullptr && type < 3;
  int max_ind = nl;  // samples used (1-based count)
  if (use_t)
    while (ls[max_ind - 1] > lmax_tmpl) max_ind--;
  const double* tt = use_t ? tmpl + (size_t)type * (lmax_tmpl - 1) : nullptr;
  for (int i = threadIdx.x; i < nl; i += blockDim.x) {
    double val = M.iCl[(size_t)type * nl + i];
    if (use_t && i < max_ind) val = val - tt[ls[i] - 2];
    v[i] = val;
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // spline(xl, v, max_ind, 1e30, 1e30, d2)
    const int n = max_ind;
    double d1r = (v[1] - v[0]) / (double)(ls[1] - ls[0]);
    d2[0] = 0.0;
    u[0] = 0.0;
    for (int i = 1; i < n - 1; i++) {
      const double d1l = d1r;
      d1r = (v[i + 1] - v[i]) / ((double)ls[i + 1] - (double)ls[i]);
      const double xxdiv = 1.0 / ((double)ls[i + 1] - (double)ls[i - 1]);
      const double sig = ((double)ls[i] - (double)ls[i - 1]) * xxdiv;
      const double xp = 1.0 / (sig * d2[i - 1] + 2.0);
      d2[i] = (sig - 1.0) * xp;
      u[i] = (6.0 * (d1r - d1l) * xxdiv - sig * u[i - 1]) * xp;
    }
    d2[n - 1] = (0.0 - 0.0 * u[n - 2]) / (0.0 * d2[n - 2] + 1.0);
    for (int i = n - 2; i >= 0; i--) d2[i] = d2[i] * d2[i + 1] + u[i];
  }
  __syncthreads();
  const int ltop = ls[max_ind - 1];
  double* out = M.Cl + (size_t)type * nL;
  for (int il = 2 + threadIdx.x; il <= ltop; il += blockDim.x) {
    // llo = number of samples with l < il, at least 1 (sequential rule of camb/modules.f90:1030-1033)
    int lo = 0, hi = max_ind;  // count of ls[i] < il
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (ls[mid] < il)
        lo = mid + 1;
      else
        hi = mid;
    }
    int llo = lo < 1 ? 1 : lo;
    if (llo > max_ind - 1) llo = max_ind - 1;
    const int lhi = llo + 1;
    const double ho = ls[lhi - 1] - ls[llo - 1];
    const double a0 = (ls[lhi - 1] - il) / ho;
    const double b0 = (il - ls[llo - 1]) / ho;
    double val = a0 * v[llo - 1] + b0 * v[lhi - 1] + ((a0 * a0 * a0 - a0) * d2[llo - 1] + (b0 * b0 * b0 - b0) * d2[lhi - 1]) * (ho * ho) / 6;
    if (use_t) val = val + tt[il - 2];
    out[il - 2] = val;
  }
  // (maxdelta < max_ind tail of InterpolateClArrTemplated needs lmax > 8000: not reachable for lmax <= 8000)
}

}  // namespace gc

// ------------------------------------------------------------------------------------------------ launchers
extern "C" cudaError_t gc_set_bess_regions(const GcRegions* r) { return cudaMemcpyToSymbol(c_bess, r, sizeof(GcRegions)); }

extern "C" cudaError_t gc_launch_bessels(double2* bess, const double* xs, const int* ls, double* scratch, int nx, int nl,
                                         cudaStream_t st) {
  dim3 g((nx + 127) / 128, nl);
  gc::bessel_fill_kernel<<<g, 128, 0, st>>>(bess, xs, ls, nx, nl);
  gc::bessel_spline_kernel<<<(nl + 31) / 32, 32, 0, st>>>(bess, xs, scratch, nx, nl);
  return cudaGetLastError();
}

extern "C" cudaError_t gc_launch_spline_k(const DevModel* models, int nmodels, int max_rows, cudaStream_t st) {
  dim3 g((max_rows + 63) / 64, nmodels);
  gc::spline_k_kernel<<<g, 64, 0, st>>>(models, nmodels);
  return cudaGetLastError();
}

extern "C" cudaError_t gc_launch_los(const DevModel* models, int nmodels, int max_nq, int max_nt, int S, const double2* bess,
                                     const double* xs, const int* ls, int nx, int nl, unsigned long long* n_iter, int q_lo,
                                     cudaStream_t st) {
  const size_t smem = ((size_t)S * max_nt + 3 * (size_t)max_nt) * sizeof(double) + (size_t)max_nt * sizeof(int);
  if (smem > 48 * 1024) {  // per device and cheap: set on every launch (several devices may be driven from one process)
    cudaError_t e = cudaFuncSetAttribute(gc::los_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 g(max_nq, nmodels);  // max_nq = number of q values to integrate from q_lo on
  gc::los_kernel<<<g, 256, smem, st>>>(models, bess, xs, ls, nx, nl, n_iter, q_lo);
  return cudaGetLastError();
}

extern "C" cudaError_t gc_launch_zero_delta(const DevModel* models, int nmodels, int nl, cudaStream_t st) {
  dim3 g(64, nmodels);
  gc::zero_delta_kernel<<<g, 256, 0, st>>>(models, nl);
  return cudaGetLastError();
}

// what = 1: the sums over q (iCl), 2: the interpolation in l, 3: both
extern "C" cudaError_t gc_launch_cls(const DevModel* models, int nmodels, const int* ls, int nl, int ntype, int lmax,
                                     const double* tmpl, int lmax_tmpl, int what, cudaStream_t st) {
  if (what & 1) {
    dim3 g(nl, nmodels);
    gc::cl_kernel<<<g, 256, 0, st>>>(models, ls, nl);
  }
  if (what & 2) {
    dim3 g2(ntype, nmodels);
    gc::cl_interp_kernel<<<g2, 256, 0, st>>>(models, ls, nl, lmax, tmpl, lmax_tmpl);
  }
  return cudaGetLastError();
}
