// TEST: host-side check of cosmomc_galileon_b200/csrc/galileon_fast.cuh (the restructured Galileon terms the
// CUDA kernel uses) against the oracle's reference-order evaluation of camb/galileon.cc, on 20000 random points.
// Prints "name worst_rel_err" lines; tests/test_galileon_fast.py asserts on them.
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <dlfcn.h>
#include "../cosmomc_galileon_b200/csrc/galileon_fast.cuh"
extern "C" {
void* orc_new(); int orc_set(void*, const char*, double); void orc_set_physical(void*, double,double,double,double);
int orc_stage_background(void*); double orc_get(void*, const char*); double orc_gal_H(void*, double); double orc_gal_X(void*, double);
void orc_gal_terms(void*, const double*, double*); void orc_nu_background(void*, double, double*, double*);
double orc_nu_dp(void*, double, double, double);
}
int main() {
  void* h = orc_new();
  orc_set_physical(h, 0.0221743, 0.1177806, 0.00064, 78.07227);
  orc_set(h, "use_galileon", 1); orc_set(h, "c2", -8.438179); orc_set(h, "c3", -3.366555); orc_set(h, "c4", -1.03411); orc_set(h, "cG", 0.001188243);
  orc_set(h, "use_spline_template", 0);
  if (orc_stage_background(h)) { printf("bg fail\n"); return 1; }
  gc::GalConst G{orc_get(h,"gal_c2"), orc_get(h,"gal_c3"), orc_get(h,"gal_c4"), orc_get(h,"gal_c5"), orc_get(h,"gal_cG"), orc_get(h,"gal_h0"), orc_get(h,"gal_orad")};
  const double grhormass = orc_get(h, "grhormass1"), numass = orc_get(h, "nu_masses1");
  double worst[9] = {0};
  srand(1);
  for (int t = 0; t < 20000; t++) {
    double u = rand() / (double)RAND_MAX;
    double a = std::exp(std::log(5e-7) + u * (std::log(1.0) - std::log(5e-7)));
    auto rnd = []() { return (rand() / (double)RAND_MAX - 0.5) * 2; };
    double k = std::exp(std::log(1e-5) + rand() / (double)RAND_MAX * std::log(0.5 / 1e-5));
    double in[12] = {a, rnd() * 1e-3, rnd() * 1e-4, rnd() * 1e-5, rnd(), rnd() * 10, rnd() * 0.1, k, rnd() * 1e-6, rnd(), 1e-3, 3e-4};
    double ref[9];
    orc_gal_terms(h, in, ref);
    double hbar = orc_gal_H(h, a), X = orc_gal_X(h, a), rho, p;
    orc_nu_background(h, a * numass, &rho, &p);
    gc::GalFast F;
    gc::GalDyn B;
    gc::gal_fast_background(G, a, hbar, X, grhormass * p, k, F, &B);
    gc::GalPiDotLin L;
    const double Hc = a * hbar;
    gc::gal_fast_pigalprime_lin(G, F, B, grhormass * (orc_nu_dp(h, a * numass, Hc, p) - 4 * p * Hc), in[10] + in[11], L);
    const double pigalprime = L.p * in[6] + L.d * in[5] + L.pid * in[9] + L.g * in[1] + L.q * in[2] + L.pi * in[3] + L.e * in[4];
    double got[9] = {F.dh, F.dx, F.grhogal, F.gpresgal, gc::gal_fast_Chigal(F, in[1], in[4], in[5], in[6]), gc::gal_fast_qgal(F, in[2], in[5], in[6]),
                     gc::gal_fast_Pigal(G, F, in[1], in[2], in[3], in[4], in[5]), gc::gal_fast_dphisecond(G, F, in[1], in[2], in[4], in[5], in[6], in[8]), pigalprime};
    for (int i = 0; i < 9; i++) {
      double e = std::fabs(got[i] - ref[i]) / (std::fabs(ref[i]) + 1e-300);
      if (ref[i] == 0 && got[i] == 0) e = 0;
      if (e > worst[i]) worst[i] = e;
    }
  }
  const char* names[9] = {"dh", "dx", "grhogal", "gpresgal", "Chigal", "qgal", "Pigal", "dphisecond", "pigalprime"};
  for (int i = 0; i < 9; i++) printf("%-10s worst rel err %.3e\n", names[i], worst[i]);
  return 0;
}
