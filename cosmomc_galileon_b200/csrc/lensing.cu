// Lensed C_l (TT, EE, BB, TE) from the unlensed scalar C_l and the lensing potential by the curved-sky correlation
// function method: CorrFuncFullSkyImpl, camb/lensing.f90:107-528 (lensing_method = 1, AccurateBB = F, no tensors).
//
// Decomposition: one lane per angle theta_i (the reference's OMP loop, :288), 128 angles per CTA, grid = (angle
// blocks, models).  Every lane runs the Legendre/Wigner-d recurrences in l three times, all lanes in lock step:
//   pass 1  sigma^2(theta), C_gl,2(theta)                         (:312-321)
//   pass 2  the correlation functions at the sampled l            (:343-460)
//   pass 3  the Gauss-Legendre-like sum over theta for each l     (:478-491), reduced over the CTA's angles
// The recurrences are recomputed instead of stored (3 x ~60 flops per (l, theta) against 7 arrays of lmax doubles per
// lane); the only memory traffic is the per-model C_l input, broadcast to all lanes, and the per-block partial sums.
#include <cuda_runtime.h>

#include "model.h"

namespace gc {

struct LensGeom {
  int lmax;         // last l of the sums (extrapolated with the template beyond Max_l)
  int Max_l;        // last computed l
  int lmax_lensed;  // last lensed l
  int npoints, interp_fac, apod_w3;  // apod_w3 = 3 * apodize_point_width
  int nblk;         // angle blocks per model
  double dtheta;
};

constexpr int LENS_THREADS = 128;
constexpr int LENS_LCHUNK = 32;

// l-only factors shared by all models: {1/l, 1/(l(l+1)), 1/((l+2)(l-1)), 1/lrootfacs}, {lrootfacs, theta_cut, sqrt(l), 0}
__global__ void lens_ltab_kernel(double4* __restrict__ ta, double4* __restrict__ tb, int lmax) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l > lmax + 4) return;
  const double lf = (double)l * (l + 1), lf2 = (double)(l + 2) * (l - 1);
  const double lr = sqrt(lf * lf2);
  ta[l] = make_double4(l > 0 ? 1.0 / l : 0.0, l > 0 ? 1.0 / lf : 0.0, l > 1 ? 1.0 / lf2 : 0.0, l > 1 ? 1.0 / lr : 0.0);
  tb[l] = make_double4(lr, l > 1 ? 0.244949 / sqrt(3.0 * lf - 8.0) : 0.0, sqrt((double)l), 0.0);
}

// Cphil3, CTT, CEE, CTE (:221-257) as cin[m][l] = (Cphil3, CTT, CEE, CTE); status[m] (zeroed by the launcher) is set
// if the spectrum (:230-235) or the high-L template (:243-249) is not realistically normalised
__global__ void lens_prep_kernel(const DevModel* __restrict__ models, LensGeom G, const double* __restrict__ tmpl,
                                 const double* __restrict__ tmpl_pp, int lmax_tmpl, double4* __restrict__ cin,
                                 int* __restrict__ status) {
  const int m = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l > G.lmax) return;
  const double pi = 3.14159265358979323846264338328;
  double4 out = make_double4(0, 0, 0, 0);
  if (l >= 2) {
    const DevModel& M = models[m];
    const int nL = G.Max_l - 1;
    const double* Cl = M.Cl;
    auto row = [&](int ll) {
      const double fac = (2 * ll + 1) / (4 * pi) * 2 * pi / (ll * (ll + 1));
      return make_double4(Cl[3 * (size_t)nL + ll - 2] * (2 * ll + 1) * (ll + 1) / ((double)ll * (double)ll * (double)ll) / (4 * pi),
                          Cl[ll - 2] * fac, Cl[(size_t)nL + ll - 2] * fac, Cl[2 * (size_t)nL + ll - 2] * fac);
    };
    if (l <= G.Max_l) {
      out = row(l);
      if (l == 10 && out.x > (double)1e-7f) atomicOr(&status[m], 1);
    } else {
      const int nT = lmax_tmpl - 1;
      const int L = G.Max_l;
      const double4 e = row(L);
      double sc = (2 * L + 1) / (4 * pi) * 2 * pi / (L * (L + 1));
      const double fac2 = e.y / (sc * tmpl[L - 2]);
      const double fac = e.x / (sc * tmpl_pp[L - 2]);
      sc = (2 * l + 1) / (4 * pi) * 2 * pi / (l * (l + 1));
      out = make_double4(tmpl_pp[l - 2] * fac * sc, tmpl[l - 2] * fac2 * sc, tmpl[(size_t)nT + l - 2] * fac2 * sc,
                         tmpl[2 * (size_t)nT + l - 2] * fac2 * sc);
      if (l == G.Max_l + 1 && out.x > (double)1e-7f) atomicOr(&status[m], 1);  // :243-249
    }
    if (M.ALens_fid > 0) {  // :252-257: the lensing potential of the template itself, scaled
      const double sc = (2 * l + 1) / (4 * pi) * 2 * pi / (l * (l + 1));
      out.x = sc * tmpl_pp[l - 2] * M.ALens_fid;
    }
  }
  cin[(size_t)m * (G.lmax + 1) + l] = out;
}

struct Legendre {
  double x, inv_sin2, pmm, pmmp1, P, dP;
  __device__ __forceinline__ void start(double x_, double sinth) {
    x = x_;
    inv_sin2 = 1.0 / (sinth * sinth);
    pmm = 1;
    pmmp1 = x_;
  }
  __device__ __forceinline__ void step(int l, double inv_l) {  // :313-316
    P = ((2 * l - 1) * x * pmmp1 - (l - 1) * pmm) * inv_l;
    dP = l * (pmmp1 - x * P) * inv_sin2;
    pmm = pmmp1;
    pmmp1 = P;
  }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(LENS_THREADS) lens_corr_kernel(LensGeom G, const double4* __restrict__ ta,
                                                                 const double4* __restrict__ tb,
                                                                 const double4* __restrict__ cin,
                                                                 const float* __restrict__ apod,
                                                                 double* __restrict__ partial) {
  const int m = blockIdx.y, blk = blockIdx.x;
  const int i_raw = blk * LENS_THREADS + threadIdx.x + 1;
  const bool live = i_raw <= G.npoints - 1;
  const int i = live ? i_raw : G.npoints - 1;
  const double4* C = cin + (size_t)m * (G.lmax + 1);
  const double theta = i * G.dtheta;
  double sinth, x;
  sincos(theta, &sinth, &x);
  const double fac1 = 1 - x, fac2 = 1 + x, facr = fac1 / fac2, inv_facr = fac2 / fac1;
  const double inv_fac1 = 1.0 / fac1, inv_fac2 = 1.0 / fac2;
  Legendre L;

  // ---- pass 1 ----
  double sigmasq = 0, Cg2 = 0;
  L.start(x, sinth);
  for (int l = 2; l <= G.lmax; l++) {
    const double4 a = ta[l];
    L.step(l, a.x);
    const double d11 = fac1 * L.dP * a.y + L.P;
    const double dm11 = fac2 * L.dP * a.y - L.P;
    const double cphi = C[l].x;
    sigmasq += (1 - d11) * cphi;
    Cg2 += dm11 * cphi;
  }

  // ---- pass 2 ----
  double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  const double Cg2sq = Cg2 * Cg2;
  const double th2 = theta * theta;
  const double sinfac = 4 / sinth, inv_sinth = 1.0 / sinth;
  L.start(x, sinth);
  const int half = G.interp_fac / 2;
  for (int l = 2; l <= G.lmax; l++) {
    const double4 a = ta[l];
    L.step(l, a.x);
    if (!(l <= 15 || ((l - 15) % G.interp_fac) == half)) continue;
    const double4 b = tb[l];
    const double llp1 = (double)l * (l + 1), lf2 = (double)(l + 2) * (l - 1);
    const double P = L.P, dP = L.dP;
    const double d11 = fac1 * dP * a.y + P;
    const double dm11 = fac2 * dP * a.y - P;
    const double d22 = (((4 * x - 8) * inv_fac2 + llp1) * P + 4 * facr * (fac2 + (x - 2) * a.y) * dP) * a.z;
    double d2m2;
    if (theta > b.y)
      d2m2 = ((llp1 - (4 * x + 8) * inv_fac1) * P + 4 * inv_facr * (-fac1 + (x + 2) * a.y) * dP) * a.z;
    else
      d2m2 = llp1 * lf2 * (th2 * th2) * (1.0 / 384.0 - (3.0 * llp1 - 8.0) / 23040.0 * th2);
    const double d20 = (2 * x * dP - llp1 * P) * a.w;
    const double rootllp1 = b.z * tb[l + 1].z;
    const double rootfac1 = tb[l + 2].z * tb[l - 1].z;
    const double rootfac2 = tb[l + 3].z * tb[l - 2].z;
    const double inv_rf1 = 1.0 / rootfac1;
    const double d1m2 = sinth * inv_rf1 * (dP - 2 * inv_fac1 * dm11);
    const double d12 = sinth * inv_rf1 * (dP - 2 * inv_fac2 * d11);
    double d1m3 = 0, d2m3 = 0, d3m3 = 0, d13 = 0;
    if (l >= 3) {
      const double inv_rf2 = 1.0 / rootfac2;
      d1m3 = (-(x + 0.5) * d1m2 * sinfac - lf2 * dm11 * inv_rf1) * inv_rf2;
      d2m3 = (-fac2 * d2m2 * sinfac - rootfac1 * d1m2) * inv_rf2;
      d3m3 = (-(x + 1.5) * d2m3 * sinfac - rootfac1 * d1m3) * inv_rf2;
      d13 = ((x - 0.5) * d12 * sinfac - lf2 * d11 * inv_rf1) * inv_rf2;
    }
    double d04 = 0, d2m4 = 0, d4m4 = 0, rootfac3 = 0;
    if (l >= 4) {
      rootfac3 = tb[l - 3].z * tb[l + 4].z;
      d04 = ((-llp1 + (18 * x * x + 6) * L.inv_sin2) * d20 - 6 * x * lf2 * dP * a.w) / (rootfac2 * rootfac3);
      d2m4 = (-(6 * x + 4) * d2m3 * inv_sinth - rootfac2 * d2m2) / rootfac3;
      d4m4 = (-7 / 5.0 * (llp1 - 6) * d2m2 + 12 / 5.0 * (-llp1 + (9 * x + 26) * inv_fac1) * d3m3) / (llp1 - 12);
    }
    const double X000 = exp(-llp1 * sigmasq / 4);
    const double X022 = X000 * (1 + sigmasq);
    const double X220 = b.x / 4 * X000;
    const double X121 = -0.5 * rootfac1 * X000;
    const double X132 = -0.5 * rootfac2 * X000;
    const double X242 = 0.25 * rootfac2 * rootfac3 * X022;
    const double dX000 = -llp1 / 4 * X000;
    const double dX022 = (1 - llp1 / 4) * X022;
    const double f1 = dX000 * dX000, f3 = X220 * X220;
    const double f2 = (Cg2 * dX022) * (Cg2 * dX022) + (X022 * X022 - 1);
    const double4 c = C[l];
    const double tTT = c.y * (((X000 * X000 - 1) + Cg2sq * f1) * P + Cg2sq * f3 * d2m2 + 8 * a.y * f1 * Cg2 * dm11);
    const double tE1 = c.z * (2 * Cg2 * X121 * X132 * d13 + f2 * d22 + Cg2sq * X242 * X220 * d04);
    const double tE2 = c.z * ((f3 * P + X242 * X242 * d4m4) * Cg2sq / 2 + Cg2 * (X121 * X121 * dm11 + X132 * X132 * d3m3) +
                              f2 * d2m2);
    const double tTE = c.w * ((X000 * X022 - 1) * d20 + 2 * dX000 * Cg2 * (X121 * d11 + X132 * d1m3) / rootllp1 +
                              Cg2sq * (X220 / 2 * d2m4 * X242 + (f3 / 2 + dX022 * dX000) * d20));
    if (l <= 15) {  // j = l - 1 <= 14 : the unweighted head of the sum (:463)
      s1[0] += tTT; s1[1] += tE1; s1[2] += tE2; s1[3] += tTE;
    } else {
      s2[0] += tTT; s2[1] += tE1; s2[2] += tE2; s2[3] += tTE;
    }
  }
  double corr[4];
  {
    double w = live ? 1.0 : 0.0;
    const int ia = i - G.npoints + G.apod_w3;  // :467-470
    if (ia > 0) w *= (double)apod[ia];
#pragma unroll
    for (int t = 0; t < 4; t++) corr[t] = (s1[t] + G.interp_fac * s2[t]) * w;
  }

  // ---- pass 3 ----
  __shared__ double wp[LENS_LCHUNK][LENS_THREADS / 32][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double halfsinth = sinth / 2;
  double* out = partial + ((size_t)m * G.nblk + blk) * 4 * (G.lmax_lensed + 1);
  L.start(x, sinth);
  for (int l0 = 2; l0 <= G.lmax_lensed; l0 += LENS_LCHUNK) {
    const int l1 = min(l0 + LENS_LCHUNK - 1, G.lmax_lensed);
    for (int l = l0; l <= l1; l++) {
      const double4 a = ta[l];
      L.step(l, a.x);
      const double llp1 = (double)l * (l + 1), lf2 = (double)(l + 2) * (l - 1);
      const double P = L.P, dP = L.dP;
      const double d22 = (((4 * x - 8) * inv_fac2 + llp1) * P + 4 * facr * (fac2 + (x - 2) * a.y) * dP) * a.z;
      double d2m2;
      if (theta > tb[l].y)
        d2m2 = ((llp1 - (4 * x + 8) * inv_fac1) * P + 4 * inv_facr * (-fac1 + (x + 2) * a.y) * dP) * a.z;
      else
        d2m2 = llp1 * lf2 * (th2 * th2) * (1.0 / 384.0 - (3.0 * llp1 - 8.0) / 23040.0 * th2);
      const double d20 = (2 * x * dP - llp1 * P) * a.w;
      const double T2 = corr[1] * d22, T4 = corr[2] * d2m2;
      const double v0 = warp_sum(corr[0] * P * sinth);
      const double v1 = warp_sum((T2 + T4) * halfsinth);
      const double v2 = warp_sum((T2 - T4) * halfsinth);
      const double v3 = warp_sum(corr[3] * d20 * sinth);
      if (lane == 0) {
        wp[l - l0][warp][0] = v0;
        wp[l - l0][warp][1] = v1;
        wp[l - l0][warp][2] = v2;
        wp[l - l0][warp][3] = v3;
      }
    }
    __syncthreads();
    {
      const int dl = threadIdx.x >> 2, t = threadIdx.x & 3;  // 32 l x 4 types = 128 threads
      if (l0 + dl <= l1) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < LENS_THREADS / 32; w++) s += wp[dl][w][t];
        out[(size_t)t * (G.lmax_lensed + 1) + l0 + dl] = s;
      }
    }
    __syncthreads();
  }
}

// :497-508 : sum the angle blocks, scale, add the unlensed spectrum.  lensed[m][type TT,EE,BB,TE][l-2]
__global__ void lens_finish_kernel(const DevModel* __restrict__ models, LensGeom G, const double* __restrict__ partial,
                                   const int* __restrict__ status, double* __restrict__ lensed) {
  const int m = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x + 2;
  if (l > G.lmax_lensed) return;
  const double pi = 3.14159265358979323846264338328;
  const int nLl = G.lmax_lensed - 1, nL = G.Max_l - 1;
  const double fac = l * (l + 1) / (2 * pi) * G.dtheta * 2 * pi;
  const double* Cl = models[m].Cl;
  const bool bad = status[m] != 0;
#pragma unroll
  for (int t = 0; t < 4; t++) {
    double s = 0;
    for (int b = 0; b < G.nblk; b++) s += partial[(((size_t)m * G.nblk + b) * 4 + t) * (G.lmax_lensed + 1) + l];
    double u = 0;
    if (t == 0) u = Cl[l - 2];
    if (t == 1) u = Cl[(size_t)nL + l - 2];
    if (t == 3) u = Cl[2 * (size_t)nL + l - 2];
    lensed[((size_t)m * 4 + t) * nLl + (l - 2)] = bad ? nan("") : s * fac + u;
  }
}

}  // namespace gc

extern "C" cudaError_t gc_launch_lens_ltab(double4* ta, double4* tb, int lmax, cudaStream_t st) {
  gc::lens_ltab_kernel<<<(lmax + 5 + 255) / 256, 256, 0, st>>>(ta, tb, lmax);
  return cudaGetLastError();
}

extern "C" cudaError_t gc_launch_lensing(const DevModel* models, int nmodels, const int* geom, double dtheta,
                                         const double4* ta, const double4* tb, const double* tmpl, const double* tmpl_pp,
                                         int lmax_tmpl, double4* cin, const float* apod, double* partial, int* status,
                                         double* lensed, cudaStream_t st) {
  gc::LensGeom G;
  G.lmax = geom[0];
  G.Max_l = geom[1];
  G.lmax_lensed = geom[2];
  G.npoints = geom[3];
  G.interp_fac = geom[4];
  G.apod_w3 = geom[5];
  G.nblk = geom[6];
  G.dtheta = dtheta;
  cudaMemsetAsync(status, 0, (size_t)nmodels * sizeof(int), st);
  gc::lens_prep_kernel<<<dim3((G.lmax + 256) / 256, nmodels), 256, 0, st>>>(models, G, tmpl, tmpl_pp, lmax_tmpl, cin, status);
  gc::lens_corr_kernel<<<dim3(G.nblk, nmodels), gc::LENS_THREADS, 0, st>>>(G, ta, tb, cin, apod, partial);
  gc::lens_finish_kernel<<<dim3((G.lmax_lensed - 1 + 127) / 128, nmodels), 128, 0, st>>>(models, G, partial, status, lensed);
  return cudaGetLastError();
}
