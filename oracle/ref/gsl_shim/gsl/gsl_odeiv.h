// TEST INFRASTRUCTURE -- stand-in for the part of GSL's legacy ODE API (<gsl/gsl_odeiv.h>) that
// camb/galileon.cc:344-363 uses: gsl_odeiv_step_rkf45, gsl_odeiv_control_y_new, gsl_odeiv_evolve_apply.
// Written from the published algorithm (Fehlberg 4(5) tableau; GSL manual "Adaptive Step-size Control":
// D_i = eps_abs + eps_rel*(a_y|y_i| + a_dydt h|y'_i|), decrease by 0.9 r^(-1/q) (floor 1/5) when r > 1.1,
// increase by 0.9 r^(-1/(q+1)) (cap 5) when r < 0.5, q = 5; "Evolution": a failed step is retried with the
// reduced h from the saved state, the last step is clipped to land exactly on t1).  Not copied from GSL.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <vector>

#include "gsl_errno.h"

struct gsl_odeiv_system {
  int (*function)(double t, const double y[], double dydt[], void* params);
  int (*jacobian)(double, const double[], double*, double[], void*);
  size_t dimension;
  void* params;
};
struct gsl_odeiv_step_type { int order; };
static const gsl_odeiv_step_type gsl_odeiv_step_rkf45_obj = {5};
static const gsl_odeiv_step_type* const gsl_odeiv_step_rkf45 = &gsl_odeiv_step_rkf45_obj;
struct gsl_odeiv_step {
  size_t dim;
  int order;
  std::vector<double> k1, k2, k3, k4, k5, k6, y0, ytmp;
};
struct gsl_odeiv_control { double eps_abs, eps_rel, a_y, a_dydt; };
struct gsl_odeiv_evolve {
  size_t dim;
  std::vector<double> y0, yerr, dydt_in, dydt_out;
  double last_step;
  unsigned long count, failed_steps;
};

inline gsl_odeiv_step* gsl_odeiv_step_alloc(const gsl_odeiv_step_type* T, size_t dim) {
  gsl_odeiv_step* s = new gsl_odeiv_step;
  s->dim = dim;
  s->order = T->order;
  for (auto* v : {&s->k1, &s->k2, &s->k3, &s->k4, &s->k5, &s->k6, &s->y0, &s->ytmp}) v->assign(dim, 0.0);
  return s;
}
inline void gsl_odeiv_step_free(gsl_odeiv_step* s) { delete s; }
inline gsl_odeiv_control* gsl_odeiv_control_y_new(double eps_abs, double eps_rel) {
  return new gsl_odeiv_control{eps_abs, eps_rel, 1.0, 0.0};
}
inline void gsl_odeiv_control_free(gsl_odeiv_control* c) { delete c; }
inline gsl_odeiv_evolve* gsl_odeiv_evolve_alloc(size_t dim) {
  gsl_odeiv_evolve* e = new gsl_odeiv_evolve;
  e->dim = dim;
  for (auto* v : {&e->y0, &e->yerr, &e->dydt_in, &e->dydt_out}) v->assign(dim, 0.0);
  e->last_step = 0;
  e->count = e->failed_steps = 0;
  return e;
}
inline void gsl_odeiv_evolve_free(gsl_odeiv_evolve* e) { delete e; }

// one Fehlberg 4(5) step: y <- 5th-order solution, yerr = difference to the embedded 4th-order one
inline int gsl_shim_rkf45_apply(gsl_odeiv_step* s, double t, double h, double y[], double yerr[], const double dydt_in[],
                                double dydt_out[], const gsl_odeiv_system* sys) {
  const size_t n = s->dim;
  static const double ah[] = {1.0 / 4.0, 3.0 / 8.0, 12.0 / 13.0, 1.0, 1.0 / 2.0};
  static const double b3[] = {3.0 / 32.0, 9.0 / 32.0};
  static const double b4[] = {1932.0 / 2197.0, -7200.0 / 2197.0, 7296.0 / 2197.0};
  static const double b5[] = {8341.0 / 4104.0, -32832.0 / 4104.0, 29440.0 / 4104.0, -845.0 / 4104.0};
  static const double b6[] = {-6080.0 / 20520.0, 41040.0 / 20520.0, -28352.0 / 20520.0, 9295.0 / 20520.0, -5643.0 / 20520.0};
  static const double c1 = 902880.0 / 7618050.0, c3 = 3953664.0 / 7618050.0, c4 = 3855735.0 / 7618050.0,
                      c5 = -1371249.0 / 7618050.0, c6 = 277020.0 / 7618050.0;
  static const double ec[] = {0.0, 1.0 / 360.0, 0.0, -128.0 / 4275.0, -2197.0 / 75240.0, 1.0 / 50.0, 2.0 / 55.0};
  double *k1 = s->k1.data(), *k2 = s->k2.data(), *k3 = s->k3.data(), *k4 = s->k4.data(), *k5 = s->k5.data(),
         *k6 = s->k6.data(), *ytmp = s->ytmp.data(), *y0 = s->y0.data();
  for (size_t i = 0; i < n; i++) y0[i] = y[i];
  if (dydt_in) {
    for (size_t i = 0; i < n; i++) k1[i] = dydt_in[i];
  } else if (sys->function(t, y, k1, sys->params) != GSL_SUCCESS) return GSL_EBADFUNC;
  for (size_t i = 0; i < n; i++) ytmp[i] = y[i] + ah[0] * h * k1[i];
  if (sys->function(t + ah[0] * h, ytmp, k2, sys->params) != GSL_SUCCESS) return GSL_EBADFUNC;
  for (size_t i = 0; i < n; i++) ytmp[i] = y[i] + h * (b3[0] * k1[i] + b3[1] * k2[i]);
  if (sys->function(t + ah[1] * h, ytmp, k3, sys->params) != GSL_SUCCESS) return GSL_EBADFUNC;
  for (size_t i = 0; i < n; i++) ytmp[i] = y[i] + h * (b4[0] * k1[i] + b4[1] * k2[i] + b4[2] * k3[i]);
  if (sys->function(t + ah[2] * h, ytmp, k4, sys->params) != GSL_SUCCESS) return GSL_EBADFUNC;
  for (size_t i = 0; i < n; i++) ytmp[i] = y[i] + h * (b5[0] * k1[i] + b5[1] * k2[i] + b5[2] * k3[i] + b5[3] * k4[i]);
  if (sys->function(t + ah[3] * h, ytmp, k5, sys->params) != GSL_SUCCESS) return GSL_EBADFUNC;
  for (size_t i = 0; i < n; i++)
    ytmp[i] = y[i] + h * (b6[0] * k1[i] + b6[1] * k2[i] + b6[2] * k3[i] + b6[3] * k4[i] + b6[4] * k5[i]);
  if (sys->function(t + ah[4] * h, ytmp, k6, sys->params) != GSL_SUCCESS) return GSL_EBADFUNC;
  for (size_t i = 0; i < n; i++) {
    const double d_i = c1 * k1[i] + c3 * k3[i] + c4 * k4[i] + c5 * k5[i] + c6 * k6[i];
    y[i] += h * d_i;
  }
  if (dydt_out) {
    if (sys->function(t + h, y, dydt_out, sys->params) != GSL_SUCCESS) {
      for (size_t i = 0; i < n; i++) y[i] = y0[i];
      return GSL_EBADFUNC;
    }
  }
  for (size_t i = 0; i < n; i++)
    yerr[i] = h * (ec[1] * k1[i] + ec[3] * k3[i] + ec[4] * k4[i] + ec[5] * k5[i] + ec[6] * k6[i]);
  return GSL_SUCCESS;
}

enum { GSL_ODEIV_HADJ_INC = 1, GSL_ODEIV_HADJ_NIL = 0, GSL_ODEIV_HADJ_DEC = -1 };

inline int gsl_shim_hadjust(const gsl_odeiv_control* c, int ord, size_t dim, const double y[], const double yerr[],
                            const double yp[], double* h) {
  const double S = 0.9, h_old = *h;
  double rmax = DBL_MIN;
  for (size_t i = 0; i < dim; i++) {
    const double D0 = c->eps_rel * (c->a_y * std::fabs(y[i]) + c->a_dydt * std::fabs(h_old * yp[i])) + c->eps_abs;
    const double r = std::fabs(yerr[i]) / std::fabs(D0);
    rmax = r > rmax ? r : rmax;
  }
  if (rmax > 1.1) {
    double r = S / std::pow(rmax, 1.0 / ord);
    if (r < 0.2) r = 0.2;
    *h = r * h_old;
    return GSL_ODEIV_HADJ_DEC;
  } else if (rmax < 0.5) {
    double r = S / std::pow(rmax, 1.0 / (ord + 1.0));
    if (r > 5.0) r = 5.0;
    if (r < 1.0) r = 1.0;
    *h = r * h_old;
    return GSL_ODEIV_HADJ_INC;
  }
  return GSL_ODEIV_HADJ_NIL;
}

inline int gsl_odeiv_evolve_apply(gsl_odeiv_evolve* e, gsl_odeiv_control* con, gsl_odeiv_step* step,
                                  const gsl_odeiv_system* dydt, double* t, double t1, double* h, double y[]) {
  const double t0 = *t;
  double h0 = *h;
  const double dt = t1 - t0;
  if ((dt < 0.0 && h0 > 0.0) || (dt > 0.0 && h0 < 0.0)) return GSL_EINVAL;
  for (size_t i = 0; i < e->dim; i++) e->y0[i] = y[i];
  if (dydt->function(t0, y, e->dydt_in.data(), dydt->params) != GSL_SUCCESS) return GSL_EBADFUNC;
  for (;;) {
    int final_step = 0;
    if ((dt >= 0.0 && h0 > dt) || (dt < 0.0 && h0 < dt)) {
      h0 = dt;
      final_step = 1;
    }
    const int st = gsl_shim_rkf45_apply(step, t0, h0, y, e->yerr.data(), e->dydt_in.data(), e->dydt_out.data(), dydt);
    if (st != GSL_SUCCESS) {
      *h = h0;
      *t = t0;
      return st;
    }
    e->count++;
    e->last_step = h0;
    *t = final_step ? t1 : t0 + h0;
    const double h_old = h0;
    const int adj = gsl_shim_hadjust(con, step->order, e->dim, y, e->yerr.data(), e->dydt_out.data(), &h0);
    if (adj == GSL_ODEIV_HADJ_DEC) {
      // the step is repeated with the reduced h -- unless h cannot shrink any more
      const double t_curr = *t, t_next = *t + h0;
      if (std::fabs(h0) < std::fabs(h_old) && t_next != t_curr) {
        for (size_t i = 0; i < e->dim; i++) y[i] = e->y0[i];
        e->failed_steps++;
        continue;
      }
      h0 = h_old;
    }
    break;
  }
  *h = h0;
  return GSL_SUCCESS;
}
