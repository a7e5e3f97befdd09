// TEST INFRASTRUCTURE (oracle) -- see orc.hpp header.
// Numerical primitives of camb/subroutines.f90 and the Ranges module of camb/utils.F90.
#include "orc.hpp"

namespace orc {

// camb/subroutines.f90:36-48
void splini(double* g, int n) {
  // g is 1-based: g[1..n]
  g[1] = 0.0;
  for (int i = 2; i <= n; i++) g[i] = 1 / (4.0 - g[i - 1]);
}

// camb/subroutines.f90:6-34   (y, dy, g 1-based)
void splder(const double* y, double* dy, int n, const double* g) {
  std::vector<double> f(n + 1);
  int n1 = n - 1;
  f[1] = (-10.0 * y[1] + 15.0 * y[2] - 6.0 * y[3] + y[4]) / 6.0;
  f[n] = (10.0 * y[n] - 15.0 * y[n1] + 6.0 * y[n - 2] - y[n - 3]) / 6.0;
  for (int i = 2; i <= n1; i++) f[i] = g[i] * (3.0 * (y[i + 1] - y[i - 1]) - f[i - 1]);
  dy[n] = f[n];
  for (int i = n1; i >= 1; i--) dy[i] = f[i] - g[i] * dy[i + 1];
}

// camb/subroutines.f90:253-295  (x, y, d2 1-based)
void spline(const double* x, const double* y, int n, double d11, double d1n, double* d2) {
  std::vector<double> u(n + 1);
  double xp, qn, sig, un, xxdiv, d1l, d1r;
  d1r = (y[2] - y[1]) / (x[2] - x[1]);
  if (d11 > .99e30) {
    d2[1] = 0.0;
    u[1] = 0.0;
  } else {
    d2[1] = -0.5;
    u[1] = (3.0 / (x[2] - x[1])) * (d1r - d11);
  }
  for (int i = 2; i <= n - 1; i++) {
    d1l = d1r;
    d1r = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
    xxdiv = 1.0 / (x[i + 1] - x[i - 1]);
    sig = (x[i] - x[i - 1]) * xxdiv;
    xp = 1.0 / (sig * d2[i - 1] + 2.0);
    d2[i] = (sig - 1.0) * xp;
    u[i] = (6.0 * (d1r - d1l) * xxdiv - sig * u[i - 1]) * xp;
  }
  d1l = d1r;
  if (d1n > .99e30) {
    qn = 0.0;
    un = 0.0;
  } else {
    qn = 0.5;
    un = (3.0 / (x[n] - x[n - 1])) * (d1n - d1l);
  }
  d2[n] = (un - qn * u[n - 1]) / (qn * d2[n - 1] + 1.0);
  for (int i = n - 1; i >= 1; i--) d2[i] = d2[i] * d2[i + 1] + u[i];
}

// camb/subroutines.f90:342-364  (y 1-based)
void splint(const double* y, double* z, int n) {
  int n1 = n - 1;
  double dy1 = 0.0;
  double dyn = (11.0 * y[n] - 18.0 * y[n1] + 9.0 * y[n - 2] - 2.0 * y[n - 3]) / 6.0;
  double zz = 0.5 * (y[1] + y[n]) + (dy1 - dyn) / 12.0;
  double s = 0;
  for (int i = 2; i <= n1; i++) s += y[i];
  *z = zz + s;
}

// camb/subroutines.f90:117-176
double rombint(scalar_fn f, void* ctx, double a, double b, double tol) {
  const int MAXITER = 20, MAXJ = 5;
  double g[MAXJ + 2];
  double h = 0.5 * (b - a);
  double gmax = h * (f(ctx, a) + f(ctx, b));
  g[1] = gmax;
  int nint = 1;
  double error = 1.0e20;
  int i = 0;
  double g0 = 0, g1, fourj;
  for (;;) {
    i = i + 1;
    if (i > MAXITER || (i > 5 && std::fabs(error) < tol)) break;
    g0 = 0.0;
    for (int k = 1; k <= nint; k++) g0 = g0 + f(ctx, a + (k + k - 1) * h);
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
    if (std::fabs(g0) > tol)
      error = 1.0 - gmax / g0;
    else
      error = gmax;
    gmax = g0;
    g[jmax + 1] = g0;
  }
  if (i > MAXITER && std::fabs(error) > tol)
    fprintf(stderr, "Warning: Rombint failed to converge; integral, error, tol: %g %g %g\n", g0, error, tol);
  return g0;
}

// camb/subroutines.f90:52-114
double rombint2(scalar_fn f, void* ctx, double a, double b, double tol, int maxit, int minsteps) {
  const int MAXJ = 5;
  double g[MAXJ + 2];
  double h = 0.5 * (b - a);
  double gmax = h * (f(ctx, a) + f(ctx, b));
  g[1] = gmax;
  int nint = 1;
  double error = 1.0e20;
  int i = 0;
  double g0 = 0, g1, fourj;
  for (;;) {
    i = i + 1;
    // Fortran precedence: (i>maxit .or. (i>5 .and. |error|<tol)) .and. nint>minsteps
    // -- .and. binds tighter than .or.: i>maxit .or. ((i>5 .and. |error|<tol) .and. nint>minsteps)
    if (i > maxit || ((i > 5 && std::fabs(error) < tol) && nint > minsteps)) break;
    g0 = 0.0;
    for (int k = 1; k <= nint; k++) g0 = g0 + f(ctx, a + (k + k - 1) * h);
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
    if (std::fabs(g0) > tol)
      error = 1.0 - gmax / g0;
    else
      error = gmax;
    gmax = g0;
    g[jmax + 1] = g0;
  }
  if (i > maxit && std::fabs(error) > tol)
    fprintf(stderr, "Warning: Rombint2 failed to converge; integral, error, tol: %g %g %g\n", g0, error, tol);
  return g0;
}

// camb/subroutines.f90:370-1128.  Verner 6(5) pair.  Only the option set the reference
// uses is restated: c(1..9)=0 on ind=1 (default error control, weights 1/max(1,|y|)),
// ind=2 keeps user options, ind=3 continuation.  Interrupts (c(8), c(9)) are never set by
// the reference callers so ind=4,5,6 re-entry is not reachable.
void dverk(void* ctx, int n, deriv_fn fcn, double& x, double* y, double xend, double tol, int& ind, double* c,
           int nw, double* w) {
#define W(k, j) w[((j)-1) * (size_t)nw + (k)]
  // y is 0-based here (y[0..n-1]); W(k,j) with k 0-based.
  auto fail = [&]() {
    fprintf(stderr, "Error in dverk, x = %g xend= %g\n", x, xend);
    std::abort();
  };
  if (ind < 1 || ind > 6) fail();
  if (ind == 3) {
    // label 45 (:846 region of the original listing)
    if (c[21] != 0.0 && (x != c[20] || xend == c[20])) fail();
    c[21] = 0.0;
  } else {
    if (ind > 3) fail();
    if (n > nw || tol <= 0.0) fail();
    if (ind == 2) {
      for (int k = 1; k <= 9; k++) c[k] = std::fabs(c[k]);
    } else {
      for (int k = 1; k <= 9; k++) c[k] = 0.0;
    }
    c[10] = std::ldexp(1.0, -56);
    c[11] = 1.e-35;
    c[20] = x;
    for (int k = 21; k <= 24; k++) c[k] = 0.0;
  }
  double temp;
  for (;;) {  // 99999
    if (!(c[7] == 0.0 || c[24] < c[7])) {
      ind = -1;
      return;
    }
    if (ind != 6) {
      fcn(ctx, n, x, y, &W(0, 1));
      c[24] = c[24] + 1.0;
    }
    // hmin
    c[13] = c[3];
    if (c[3] == 0.0) {
      temp = 0.0;
      if (c[1] == 1.0) {
        for (int k = 0; k < n; k++) temp = std::fmax(temp, std::fabs(y[k]));
        c[12] = temp;
      } else if (c[1] == 2.0) {
        c[12] = 1.0;
      } else if (c[1] == 3.0) {
        for (int k = 0; k < n; k++) temp = std::fmax(temp, std::fabs(y[k]) / c[2]);
        c[12] = std::fmin(temp, 1.0);
      } else if (c[1] == 4.0 || c[1] == 5.0) {
        fail();  // per-component floors not used by the reference callers
      } else {
        for (int k = 0; k < n; k++) temp = std::fmax(temp, std::fabs(y[k]));
        c[12] = std::fmin(temp, 1.0);
      }
      c[13] = 10.0 * std::fmax(c[11], c[10] * std::fmax(c[12] / tol, std::fabs(x)));
    }
    // scale, hmax
    c[15] = c[5];
    if (c[5] == 0.0) c[15] = 1.0;
    if (c[6] != 0.0 && c[5] != 0.0) c[16] = std::fmin(c[6], 2.0 / c[5]);
    if (c[6] != 0.0 && c[5] == 0.0) c[16] = c[6];
    if (c[6] == 0.0 && c[5] != 0.0) c[16] = 2.0 / c[5];
    if (c[6] == 0.0 && c[5] == 0.0) c[16] = 2.0;
    if (c[13] > c[16]) {
      ind = -2;
      return;
    }
    // hmag
    if (ind <= 2) {
      c[14] = c[4];
      if (c[4] == 0.0) c[14] = c[16] * std::pow(tol, 1.0 / 6.0);
    } else if (c[23] <= 1.0) {
      temp = 2.0 * c[14];
      const double r = 2.0 / .9;
      if (tol < (r * r * r * r * r * r) * c[19]) temp = .9 * std::pow(tol / c[19], 1.0 / 6.0) * c[14];
      c[14] = std::fmax(temp, .5 * c[14]);
    } else {
      c[14] = .5 * c[14];
    }
    c[14] = std::fmin(c[14], c[16]);
    c[14] = std::fmax(c[14], c[13]);
    if (c[8] != 0.0) fail();  // interrupt 1 never requested
    // 1111: xtrial, htrial
    if (c[14] >= std::fabs(xend - x)) {
      c[14] = std::fabs(xend - x);
      c[17] = xend;
    } else {
      c[14] = std::fmin(c[14], .5 * std::fabs(xend - x));
      c[17] = x + std::copysign(c[14], xend - x);
    }
    c[18] = c[17] - x;
    temp = c[18] / 1398169080000.0;
    for (int k = 0; k < n; k++) W(k, 9) = y[k] + temp * W(k, 1) * 233028180000.0;
    fcn(ctx, n, x + c[18] / 6.0, &W(0, 9), &W(0, 2));
    for (int k = 0; k < n; k++) W(k, 9) = y[k] + temp * (W(k, 1) * 74569017600.0 + W(k, 2) * 298276070400.0);
    fcn(ctx, n, x + c[18] * (4.0 / 15.0), &W(0, 9), &W(0, 3));
    for (int k = 0; k < n; k++)
      W(k, 9) = y[k] + temp * (W(k, 1) * 1165140900000.0 - W(k, 2) * 3728450880000.0 + W(k, 3) * 3495422700000.0);
    fcn(ctx, n, x + c[18] * (2.0 / 3.0), &W(0, 9), &W(0, 4));
    for (int k = 0; k < n; k++)
      W(k, 9) = y[k] + temp * (-W(k, 1) * 3604654659375.0 + W(k, 2) * 12816549900000.0 -
                               W(k, 3) * 9284716546875.0 + W(k, 4) * 1237962206250.0);
    fcn(ctx, n, x + c[18] * (5.0 / 6.0), &W(0, 9), &W(0, 5));
    for (int k = 0; k < n; k++)
      W(k, 9) = y[k] + temp * (W(k, 1) * 3355605792000.0 - W(k, 2) * 11185352640000.0 +
                               W(k, 3) * 9172628850000.0 - W(k, 4) * 427218330000.0 + W(k, 5) * 482505408000.0);
    fcn(ctx, n, x + c[18], &W(0, 9), &W(0, 6));
    for (int k = 0; k < n; k++)
      W(k, 9) = y[k] + temp * (-W(k, 1) * 770204740536.0 + W(k, 2) * 2311639545600.0 -
                               W(k, 3) * 1322092233000.0 - W(k, 4) * 453006781920.0 + W(k, 5) * 326875481856.0);
    fcn(ctx, n, x + c[18] / 15.0, &W(0, 9), &W(0, 7));
    for (int k = 0; k < n; k++)
      W(k, 9) = y[k] + temp * (W(k, 1) * 2845924389000.0 - W(k, 2) * 9754668000000.0 +
                               W(k, 3) * 7897110375000.0 - W(k, 4) * 192082660000.0 + W(k, 5) * 400298976000.0 +
                               W(k, 7) * 201586000000.0);
    fcn(ctx, n, x + c[18], &W(0, 9), &W(0, 8));
    for (int k = 0; k < n; k++)
      W(k, 9) = y[k] + temp * (W(k, 1) * 104862681000.0 + W(k, 3) * 545186250000.0 + W(k, 4) * 446637345000.0 +
                               W(k, 5) * 188806464000.0 + W(k, 7) * 15076875000.0 + W(k, 8) * 97599465000.0);
    c[24] = c[24] + 7.0;
    for (int k = 0; k < n; k++)
      W(k, 2) = (W(k, 1) * 8738556750.0 + W(k, 3) * 9735468750.0 - W(k, 4) * 9709507500.0 +
                 W(k, 5) * 8582112000.0 + W(k, 6) * 95329710000.0 - W(k, 7) * 15076875000.0 -
                 W(k, 8) * 97599465000.0) /
                1398169080000.0;
    temp = 0.0;
    if (c[1] == 1.0) {
      for (int k = 0; k < n; k++) temp = std::fmax(temp, std::fabs(W(k, 2)));
    } else if (c[1] == 2.0) {
      for (int k = 0; k < n; k++) temp = std::fmax(temp, std::fabs(W(k, 2) / y[k]));
    } else if (c[1] == 3.0) {
      for (int k = 0; k < n; k++) temp = std::fmax(temp, std::fabs(W(k, 2)) / std::fmax(c[2], std::fabs(y[k])));
    } else {
      for (int k = 0; k < n; k++) temp = std::fmax(temp, std::fabs(W(k, 2)) / std::fmax(1.0, std::fabs(y[k])));
    }
    c[19] = temp * c[14] * c[15];
    ind = 5;
    if (c[19] > tol) ind = 6;
    if (c[9] != 0.0) fail();  // interrupt 2 never requested
    if (ind != 6) {
      x = c[17];
      for (int k = 0; k < n; k++) y[k] = W(k, 9);
      c[22] = c[22] + 1.0;
      c[23] = 0.0;
      if (x == xend) {
        ind = 3;
        c[20] = xend;
        c[21] = 1.0;
        return;
      }
    } else {
      c[23] = c[23] + 1.0;
      if (!(c[14] > c[13])) {
        ind = -3;
        return;
      }
    }
  }
#undef W
}

// ------------------------------------------------------------------ Ranges
// camb/utils.F90:38-67
void Ranges_Init(Regions& R) {
  R.points.clear();
  R.dpoints.clear();
  R.count = 0;
  R.npoints = 0;
  R.has_dpoints = false;
}

// camb/utils.F90:81-111
int Ranges_IndexOf(const Regions& Reg, double tau) {
  for (int i = 1; i <= Reg.count; i++) {
    const Region& A = Reg.R[i];
    if (tau < A.High && tau >= A.Low) {
      if (A.IsLog)
        return A.start_index + (int)(std::log(tau / A.Low) / A.delta);
      else
        return A.start_index + (int)((tau - A.Low) / A.delta);
    }
  }
  if (tau >= Reg.Highest) return Reg.npoints;
  fprintf(stderr, "Ranges_IndexOf: value out of range\n");
  std::abort();
}

// camb/utils.F90:114-149
void Ranges_GetArray(Regions& Reg, bool want_dpoints) {
  Reg.has_dpoints = want_dpoints;
  Reg.points.assign(Reg.npoints + 1, 0.0);
  int ix = 0;
  for (int i = 1; i <= Reg.count; i++) {
    const Region& A = Reg.R[i];
    for (int j = 0; j <= A.steps - 1; j++) {
      ix++;
      if (A.IsLog)
        Reg.points[ix] = A.Low * std::exp(j * A.delta);
      else
        Reg.points[ix] = A.Low + A.delta * j;
    }
  }
  ix++;
  Reg.points[ix] = Reg.Highest;
  if (ix != Reg.npoints) {
    fprintf(stderr, "Ranges_GetArray: ERROR\n");
    std::abort();
  }
  if (Reg.has_dpoints) Ranges_Getdpoints(Reg, true);
}

// camb/utils.F90:152-176
void Ranges_Getdpoints(Regions& Reg, bool half_ends) {
  int n = Reg.npoints;
  Reg.dpoints.assign(n + 1, 0.0);
  for (int i = 2; i <= n - 1; i++) Reg.dpoints[i] = (Reg.points[i + 1] - Reg.points[i - 1]) / 2;
  if (half_ends) {
    Reg.dpoints[1] = (Reg.points[2] - Reg.points[1]) / 2;
    Reg.dpoints[n] = (Reg.points[n] - Reg.points[n - 1]) / 2;
  } else {
    Reg.dpoints[1] = (Reg.points[2] - Reg.points[1]);
    Reg.dpoints[n] = (Reg.points[n] - Reg.points[n - 1]);
  }
}

static const double RangeTol = 0.1;

// camb/utils.F90:179-203
void Ranges_Add_delta(Regions& Reg, double t_start, double t_end, double t_approx_delta, bool IsLog) {
  if (t_end <= t_start) {
    fprintf(stderr, "Ranges_Add_delta: end must be larger than start\n");
    std::abort();
  }
  if (t_approx_delta <= 0) {
    fprintf(stderr, "Ranges_Add_delta: delta must be > 0\n");
    std::abort();
  }
  int n;
  if (IsLog)
    n = std::max(1, (int)(std::log(t_end / t_start) / t_approx_delta + 1.0 - RangeTol));
  else
    n = std::max(1, (int)((t_end - t_start) / t_approx_delta + 1.0 - RangeTol));
  Ranges_Add(Reg, t_start, t_end, n, IsLog);
}

// camb/utils.F90:206-466
void Ranges_Add(Regions& Reg, double t_start, double t_end, int nstep, bool WantLog) {
  const int Max_Ranges = 100;
  static thread_local Region NewRegions[Max_Ranges + 1];
  double EndPoints[2 * Max_Ranges + 2];
  double RequestDelta[Max_Ranges + 1];
  double delta;
  if (WantLog)
    delta = std::log(t_end / t_start) / nstep;
  else
    delta = (t_end - t_start) / nstep;
  if (t_end <= t_start || nstep <= 0 || Reg.count >= Max_Ranges) {
    fprintf(stderr, "Ranges_Add: bad arguments\n");
    std::abort();
  }
  for (int i = 1; i <= Reg.count; i++) NewRegions[i] = Reg.R[i];
  int nreg = Reg.count + 1;
  {
    Region& A = NewRegions[nreg];
    A = Region();
    A.Low = t_start;
    A.High = t_end;
    A.delta = delta;
    A.steps = nstep;
    A.IsLog = WantLog;
  }
  // Get end points in order (:258-302)
  int ix = 0, ixin;
  for (int i = 1; i <= nreg; i++) {
    const Region& A = NewRegions[i];
    if (ix == 0) {
      ix = 1;
      EndPoints[ix] = A.Low;
      ix = 2;
      EndPoints[ix] = A.High;
    } else {
      ixin = ix;
      for (int j = 1; j <= ixin; j++) {
        if (A.Low < EndPoints[j]) {
          for (int m = ix; m >= j; m--) EndPoints[m + 1] = EndPoints[m];
          EndPoints[j] = A.Low;
          ix = ix + 1;
          break;
        }
      }
      if (ixin == ix) {
        ix = ix + 1;
        EndPoints[ix] = A.Low;
        ix = ix + 1;
        EndPoints[ix] = A.High;
      } else {
        ixin = ix;
        for (int j = 1; j <= ixin; j++) {
          if (A.High < EndPoints[j]) {
            for (int m = ix; m >= j; m--) EndPoints[m + 1] = EndPoints[m];
            EndPoints[j] = A.High;
            ix = ix + 1;
            break;
          }
        }
        if (ixin == ix) {
          ix = ix + 1;
          EndPoints[ix] = A.High;
        }
      }
    }
  }
  // remove duplicate points (:305-313)
  ixin = ix;
  ix = 1;
  for (int i = 2; i <= ixin; i++) {
    if (EndPoints[i] != EndPoints[ix]) {
      ix = ix + 1;
      EndPoints[ix] = EndPoints[i];
    }
  }
  Reg.Lowest = EndPoints[1];
  Reg.Highest = EndPoints[ix];
  Reg.count = 0;
  double max_delta = Reg.Highest - Reg.Lowest;
  double Diff, min_log_step, max_log_step, min_request, max_request;
  for (int i = 1; i <= ix - 1; i++) {
    Region& A = Reg.R[i];
    A.Low = EndPoints[i];
    A.High = EndPoints[i + 1];
    delta = max_delta;
    A.IsLog = false;
    for (int j = 1; j <= nreg; j++) {
      const Region& N = NewRegions[j];
      if (A.Low >= N.Low && A.Low < N.High) {
        if (N.IsLog) {
          if (A.IsLog) {
            delta = std::fmin(delta, N.delta);
          } else {
            min_log_step = A.Low * (std::exp(N.delta) - 1);
            if (min_log_step < delta) {
              max_log_step = A.High * (1 - std::exp(-N.delta));
              if (delta < max_log_step) {
                delta = min_log_step;
              } else {
                A.IsLog = true;
                delta = N.delta;
              }
            }
          }
        } else {
          if (A.IsLog) {
            max_log_step = A.High * (1 - std::exp(-delta));
            if (N.delta < max_log_step) {
              min_log_step = A.Low * (std::exp(delta) - 1);
              if (min_log_step < N.delta) {
                A.IsLog = false;
                delta = min_log_step;
              } else {
                delta = -std::log(1 - N.delta / A.High);
              }
            }
          } else {
            delta = std::fmin(delta, N.delta);
          }
        }
      }
    }
    if (A.IsLog)
      Diff = std::log(A.High / A.Low);
    else
      Diff = A.High - A.Low;
    if (delta >= Diff) {
      A.delta = Diff;
      A.steps = 1;
    } else {
      A.steps = std::max(1, (int)(Diff / delta + 1.0 - RangeTol));
      A.delta = Diff / A.steps;
    }
    Reg.count = Reg.count + 1;
    RequestDelta[Reg.count] = delta;
    if (A.IsLog) {
      if (A.steps == 1) {
        A.delta_min = A.High - A.Low;
        A.delta_max = A.delta_min;
      } else {
        A.delta_min = A.Low * (std::exp(A.delta) - 1);
        A.delta_max = A.High * (1 - std::exp(-A.delta));
      }
    } else {
      A.delta_max = A.delta;
      A.delta_min = A.delta;
    }
  }
  // Get rid of tiny regions (:390-441)
  ix = Reg.count;
  for (int i = ix; i >= 1; i--) {
    Region& A = Reg.R[i];
    if (A.steps == 1) {
      Diff = A.High - A.Low;
      if (A.IsLog) {
        min_request = A.Low * (std::exp(RequestDelta[i]) - 1);
        max_request = A.High * (1 - std::exp(-RequestDelta[i]));
      } else {
        min_request = RequestDelta[i];
        max_request = min_request;
      }
      if (i != Reg.count) {
        Region& L = Reg.R[i + 1];
        if (RequestDelta[i] >= A.delta && Diff <= L.delta_min && L.delta_min <= max_request) {
          L.Low = A.Low;
          if (Diff > L.delta_min * RangeTol) L.steps = L.steps + 1;
          if (L.IsLog)
            L.delta = std::log(L.High / L.Low) / L.steps;
          else
            L.delta = (L.High - L.Low) / L.steps;
          for (int m = i; m <= Reg.count - 1; m++) Reg.R[m] = Reg.R[m + 1];
          Reg.count = Reg.count - 1;
          continue;
        }
      }
      if (i != 1) {
        Region& L = Reg.R[i - 1];
        if (RequestDelta[i] >= A.delta && Diff <= L.delta_max && L.delta_max <= min_request) {
          L.High = A.High;
          if (Diff > L.delta_max * RangeTol) L.steps = L.steps + 1;
          if (L.IsLog)
            L.delta = std::log(L.High / L.Low) / L.steps;
          else
            L.delta = (L.High - L.Low) / L.steps;
          for (int m = i; m <= Reg.count - 1; m++) Reg.R[m] = Reg.R[m + 1];
          Reg.count = Reg.count - 1;
        }
      }
    }
  }
  // start indices (:445-463)
  int nsteps = 1;
  for (int i = 1; i <= Reg.count; i++) {
    Region& A = Reg.R[i];
    A.start_index = nsteps;
    nsteps = nsteps + A.steps;
    if (A.IsLog) {
      if (A.steps == 1) {
        A.delta_min = A.High - A.Low;
        A.delta_max = A.delta_min;
      } else {
        A.delta_min = A.Low * (std::exp(A.delta) - 1);
        A.delta_max = A.High * (1 - std::exp(-A.delta));
      }
    } else {
      A.delta_max = A.delta;
      A.delta_min = A.delta;
    }
  }
  Reg.npoints = nsteps;
}

}  // namespace orc
