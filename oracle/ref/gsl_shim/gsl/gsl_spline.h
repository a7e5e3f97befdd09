// TEST INFRASTRUCTURE -- stand-in for the part of <gsl/gsl_spline.h> that camb/galileon.cc uses
// (galileon.cc:75-77, 507-509, 666-682, 706, 734, 1282-1284): gsl_interp_cspline = NATURAL cubic spline,
// evaluated with a NULL/unused accelerator (plain bisection for the interval).
// Written from GSL's documented algorithm (gsl-doc "Interpolation": cspline with natural boundary conditions,
// c_i = y''(x_i)/2 from the symmetric tridiagonal system, b_i and d_i from c_i), not copied from GSL.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <vector>

struct gsl_interp_accel { size_t cache; };
struct gsl_interp_type { int id; };
static const gsl_interp_type gsl_interp_cspline_obj = {1};
static const gsl_interp_type* const gsl_interp_cspline = &gsl_interp_cspline_obj;

struct gsl_spline {
  size_t n;
  std::vector<double> x, y, c;
};

inline gsl_interp_accel* gsl_interp_accel_alloc() { return new gsl_interp_accel{0}; }
inline void gsl_interp_accel_free(gsl_interp_accel* a) { delete a; }
inline gsl_spline* gsl_spline_alloc(const gsl_interp_type*, size_t n) {
  gsl_spline* s = new gsl_spline;
  s->n = n;
  return s;
}
inline void gsl_spline_free(gsl_spline* s) { delete s; }

inline int gsl_spline_init(gsl_spline* s, const double* xa, const double* ya, size_t n) {
  s->n = n;
  s->x.assign(xa, xa + n);
  s->y.assign(ya, ya + n);
  s->c.assign(n, 0.0);
  if (n < 3) return 0;
  // interior unknowns c_1..c_{n-2}; natural ends c_0 = c_{n-1} = 0
  const size_t m = n - 2;
  std::vector<double> diag(m), off(m), rhs(m);
  for (size_t i = 0; i < m; i++) {
    const double h_i = xa[i + 1] - xa[i], h_ip1 = xa[i + 2] - xa[i + 1];
    const double yd_i = ya[i + 1] - ya[i], yd_ip1 = ya[i + 2] - ya[i + 1];
    off[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    rhs[i] = 3.0 * (yd_ip1 / h_ip1 - yd_i / h_i);
  }
  // symmetric tridiagonal solve (Cholesky-like LDL^T, as gsl_linalg_solve_symm_tridiag does)
  std::vector<double> gamma(m), alpha(m), cc(m), z(m);
  alpha[0] = diag[0];
  gamma[0] = off[0] / alpha[0];
  for (size_t i = 1; i + 1 < m; i++) {
    alpha[i] = diag[i] - off[i - 1] * gamma[i - 1];
    gamma[i] = off[i] / alpha[i];
  }
  if (m > 1) alpha[m - 1] = diag[m - 1] - off[m - 2] * gamma[m - 2];
  z[0] = rhs[0];
  for (size_t i = 1; i < m; i++) z[i] = rhs[i] - gamma[i - 1] * z[i - 1];
  for (size_t i = 0; i < m; i++) cc[i] = z[i] / alpha[i];
  s->c[m] = cc[m - 1];
  if (m >= 2)
    for (size_t i = m - 2, j = 0; j <= m - 2; j++, i--) s->c[i + 1] = cc[i] - gamma[i] * s->c[i + 2];
  return 0;
}

inline double gsl_spline_eval(const gsl_spline* s, double x, gsl_interp_accel*) {
  const size_t n = s->n;
  size_t lo = 0, hi = n - 1;  // gsl_interp_bsearch(xa, x, 0, n-1)
  while (hi > lo + 1) {
    const size_t i = (hi + lo) / 2;
    if (s->x[i] > x) hi = i; else lo = i;
  }
  const double x_lo = s->x[lo], x_hi = s->x[lo + 1], dx = x_hi - x_lo;
  const double y_lo = s->y[lo], y_hi = s->y[lo + 1], dy = y_hi - y_lo, delx = x - x_lo;
  const double c_i = s->c[lo], c_ip1 = s->c[lo + 1];
  const double b_i = (dy / dx) - dx * (c_ip1 + 2.0 * c_i) / 3.0;
  const double d_i = (c_ip1 - c_i) / (3.0 * dx);
  return y_lo + delx * (b_i + delx * (c_i + delx * d_i));
}
