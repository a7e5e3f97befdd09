// TEST INFRASTRUCTURE.  Glue that lets the reference's own camb/galileon.cc run outside CAMB:
//   * the three Fortran call-backs it imports (galileon.cc:96-99, module MassiveNu of camb/modules.f90:1721-1850)
//     are served from the oracle's massive-neutrino tables of one oracle Model;
//   * ref_set_globals() writes the file-scope parameter globals (galileon.cc:80-92) so that the closed-form terms
//     can be evaluated without running arrays_;
//   * ref_set_background() hands a background table to the reference's spline globals (galileon.cc:75-77)
//     the way arrays_ does (:666-682).
// Built by oracle/Makefile target `ref` into oracle/_ref/libgalref.so together with the UNMODIFIED reference source,
// compiled where it lies under /root/reference (never copied into this repo).
#include <gsl/gsl_spline.h>

#include "../orc.hpp"

extern double om, orad, h0, c2, c3, c4, c5, cG, c0, q;
extern int tracker, nb;
extern gsl_interp_accel* acc;
extern gsl_spline* spline_h;
extern gsl_spline* spline_x;

static const orc::Model* g_nu_model = nullptr;

extern "C" {

void massivenu_mp_nu_rho_(double* am, double* rhonu) { orc::Nu_rho(*g_nu_model, *am, *rhonu); }
void massivenu_mp_nu_background_(double* am, double* rhonu, double* pnu) {
  orc::Nu_background(*g_nu_model, *am, *rhonu, *pnu);
}
double massivenu_mp_nu_dp_(double* am, double* adotoa, double* pnu) { return orc::Nu_dp(*g_nu_model, *am, *adotoa, *pnu); }

void ref_set_nu_model(void* oracle_model) { g_nu_model = (const orc::Model*)oracle_model; }

void ref_set_globals(double om_, double orad_, double h0_, double c2_, double c3_, double c4_, double c5_, double cG_) {
  om = om_; orad = orad_; h0 = h0_; c2 = c2_; c3 = c3_; c4 = c4_; c5 = c5_; cG = cG_;
}
void ref_get_globals(double* out /*[9]*/) {
  out[0] = om; out[1] = orad; out[2] = h0; out[3] = c2; out[4] = c3; out[5] = c4; out[6] = c5; out[7] = cG; out[8] = q;
}
void ref_set_background(const double* a, const double* h, const double* x, int n) {
  gsl_spline_free(spline_h);
  gsl_spline_free(spline_x);
  spline_h = gsl_spline_alloc(gsl_interp_cspline, n);
  spline_x = gsl_spline_alloc(gsl_interp_cspline, n);
  gsl_spline_init(spline_h, a, h, n);
  gsl_spline_init(spline_x, a, x, n);
}
// the reference passes dh, dx by C++ reference (galileon.cc:746); a pointer-only wrapper for ctypes
void GetdHdX_(double*, double*, double*, double&, double&, double*, double*, int*);
void ref_GetdHdX(double* a, double* hc, double* xc, double* dh, double* dx, double* grhormass, double* nu_masses, int* n) {
  GetdHdX_(a, hc, xc, *dh, *dx, grhormass, nu_masses, n);
}
}
