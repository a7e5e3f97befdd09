// Host pre-stage of the scalar path: everything the reference does between CAMBParams_Set and the k loop, for a BATCH
// of parameter points, one host thread per point:
//   CAMBParams_Set                 camb/modules.f90:260-484       densities, massive-neutrino set-up, tau0
//   init_massive_nu / Nu_init      camb/modules.f90:1595-1680     (tables shared by all points, built once)
//   init_background -> arrays_     camb/equations_galileon.f90:63-104, camb/galileon.cc:503-688 (csrc/galileon_host.cu)
//   Reionization_Init              camb/reionization.f90:153-208  z_re from the optical depth (bisection + Romberg)
//   InitVars                       camb/cmbmain.f90:722-795
//   Recombination_init (RECFAST)   camb/recfast.f90:460-722       3-variable ODE with dverk on 10 000 redshift steps
//   inithermo / SetTimeSteps       camb/modules.f90:2787-3146     20 000-node thermo tables, visibility, output times
//   GaugeInterface_Init            camb/equations_galileon.f90:510-553  neutrino switch times
//   SetkValuesForSources / ForInt  camb/cmbmain.f90:797-852, :1281-1353  the two k grids (Ranges, camb/utils.F90)
//   initlval                       camb/modules.f90:851-1008      l sampling
// Each point is sequential (stiff 3-variable ODE, 20 000-step recurrences, a 30 000-node background integration), so
// the batch is spread over host threads; the tables land in the galcamb_model records that galcamb_set_model_ /
// galcamb_batch_cls_ take.  Plain C++ (no device code); arrays are 1-based like the reference.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/galcamb.h"
#include "pack.h"
#include "model.h"

int gc_galileon_background(double omegar, double omegam, double h0, double c2, double c3in, double c4in, double cG, int neig,
                           const double* grhormass, const double* nu_masses, const double* nu_tab, double dlnam, double* coef,
                           std::vector<double>& a, std::vector<double>& h, std::vector<double>& x);

struct GcBackgroundIn {  // csrc/galileon_host.cu
  double omegar, omegam, h0, c2, c3in, c4in, cG;
  int neig;
  double grhormass[GC_MAX_NU], nu_masses[GC_MAX_NU];
};
int gc_galileon_background_device(int device, const GcBackgroundIn* in, int n, const double* nu_tab, size_t nu_rows, double dlnam,
                                  std::vector<double>& a_out, const double** h_out, const double** x_out, double* coef_out,
                                  int* status_out);

namespace pre {

#define SP(x) ((double)(x##f))  // Fortran default-real literals keep their single-precision rounding

// ------------------------------------------------------------------------------------------------ constants
// camb/constants.f90:11-63
const double const_pi = 3.1415926535897932384626433832795;
const double c_light = 2.99792458e8, h_P = 6.62606896e-34, G_newton = 6.6738e-11, sigma_thomson = 6.6524616e-29;
const double sigma_boltz = 5.6704e-8, k_B = 1.3806504e-23, m_p = 1.672621637e-27, m_H = 1.673575e-27, m_e = 9.10938215e-31;
const double mass_ratio_He_H = 3.9715, Mpc = 3.085678e22;
const double MPC_in_sec = Mpc / c_light;
const double barssc0 = k_B / m_p / (c_light * c_light);
const double kappa = 8.0 * const_pi * G_newton;
const double pi_aml = 3.14159265358979323846264338328;  // camb/utils.F90:957
static double a_rad() {
  const double pi5 = const_pi * const_pi * const_pi * const_pi * const_pi;
  const double kb4 = k_B * k_B * k_B * k_B;
  return 8.0 * pi5 * kb4 / 15 / (c_light * c_light * c_light) / (h_P * h_P * h_P);
}
static double Compton_CT() { return MPC_in_sec * (8.0 / 3.0) * (sigma_thomson / (m_e * c_light)) * a_rad(); }
const double base_tol = 1.0e-4;
const int NTHERMO = 20000, NRHOPN = 2000, MAXNU = GALCAMB_MAX_NU;

// ------------------------------------------------------------------------------------------------ numerics
// camb/subroutines.f90:6-48: derivative of a function tabulated on a uniform grid (splini + splder)
static void splder(const double* y, double* dy, int n) {
  std::vector<double> g(n + 1), f(n + 1);
  g[1] = 0.0;
  for (int i = 2; i <= n; i++) g[i] = 1 / (4.0 - g[i - 1]);
  const int n1 = n - 1;
  f[1] = (-10.0 * y[1] + 15.0 * y[2] - 6.0 * y[3] + y[4]) / 6.0;
  f[n] = (10.0 * y[n] - 15.0 * y[n1] + 6.0 * y[n - 2] - y[n - 3]) / 6.0;
  for (int i = 2; i <= n1; i++) f[i] = g[i] * (3.0 * (y[i + 1] - y[i - 1]) - f[i - 1]);
  dy[n] = f[n];
  for (int i = n1; i >= 1; i--) dy[i] = f[i] - g[i] * dy[i + 1];
}

// camb/subroutines.f90:253-295 with both end derivatives "natural" (1e40)
static void spline_natural(const double* x, const double* y, int n, double* d2) {
  std::vector<double> u(n + 1);
  double d1r = (y[2] - y[1]) / (x[2] - x[1]);
  d2[1] = 0.0;
  u[1] = 0.0;
  for (int i = 2; i <= n - 1; i++) {
    const double d1l = d1r;
    d1r = (y[i + 1] - y[i]) / (x[i + 1] - x[i]);
    const double xxdiv = 1.0 / (x[i + 1] - x[i - 1]);
    const double sig = (x[i] - x[i - 1]) * xxdiv;
    const double xp = 1.0 / (sig * d2[i - 1] + 2.0);
    d2[i] = (sig - 1.0) * xp;
    u[i] = (6.0 * (d1r - d1l) * xxdiv - sig * u[i - 1]) * xp;
  }
  d2[n] = (0.0 - 0.0 * u[n - 1]) / (0.0 * d2[n - 1] + 1.0);
  for (int i = n - 1; i >= 1; i--) d2[i] = d2[i] * d2[i + 1] + u[i];
}

// camb/subroutines.f90:342-364
static double splint(const double* y, int n) {
  const double dyn = (11.0 * y[n] - 18.0 * y[n - 1] + 9.0 * y[n - 2] - 2.0 * y[n - 3]) / 6.0;
  double s = 0;
  for (int i = 2; i <= n - 1; i++) s += y[i];
  return 0.5 * (y[1] + y[n]) + (0.0 - dyn) / 12.0 + s;
}

// camb/subroutines.f90:52-176: Romberg integration; minsteps < 0 = rombint (:117), otherwise rombint2 (:52)
template <class F>
static double romberg(F&& f, double a, double b, double tol, int maxit = 20, int minsteps = -1) {
  const int MAXJ = 5;
  double g[MAXJ + 2];
  double h = 0.5 * (b - a);
  double gmax = h * (f(a) + f(b));
  g[1] = gmax;
  int nint = 1, i = 0;
  double error = 1.0e20, g0 = 0;
  for (;;) {
    i++;
    const bool conv = i > 5 && std::fabs(error) < tol;
    if (minsteps < 0 ? (i > maxit || conv) : (i > maxit || (conv && nint > minsteps))) break;
    g0 = 0.0;
    for (int k = 1; k <= nint; k++) g0 = g0 + f(a + (k + k - 1) * h);
    g0 = 0.5 * g[1] + h * g0;
    h = 0.5 * h;
    nint = nint + nint;
    const int jmax = i < MAXJ ? i : MAXJ;
    double fourj = 1.0;
    for (int j = 1; j <= jmax; j++) {
      fourj = 4.0 * fourj;
      const double g1 = g0 + (g0 - g[j]) / (fourj - 1.0);
      g[j] = g0;
      g0 = g1;
    }
    error = std::fabs(g0) > tol ? 1.0 - gmax / g0 : gmax;
    gmax = g0;
    g[jmax + 1] = g0;
  }
  return g0;
}

// camb/subroutines.f90:370-1128: dverk with the default option set (c(1..9) = 0), the only one RECFAST uses.
// State between calls (ind = 3 continuation) in the struct.  Returns ind (3 = reached xend, < 0 = failure).
template <int N>
struct Dverk {
  double hmag = 0, est = 0, nfail = 0, c20 = 0, c21 = 0;
  double w[9][N];
  template <class F>
  int run(F&& fcn, double& x, double* y, double xend, double tol, int ind) {
    if (ind == 3) {
      if (c21 != 0.0 && (x != c20 || xend == c20)) return -9;
      c21 = 0.0;
    } else {
      c20 = x;
      c21 = 0.0;
      nfail = 0.0;
    }
    for (;;) {
      if (ind != 6) fcn(x, y, w[0]);
      double temp = 0.0;
      for (int k = 0; k < N; k++) temp = std::fmax(temp, std::fabs(y[k]));
      const double c12 = std::fmin(temp, 1.0);
      const double hmin = 10.0 * std::fmax(1.e-35, std::ldexp(1.0, -56) * std::fmax(c12 / tol, std::fabs(x)));
      const double hmax = 2.0;
      if (hmin > hmax) return -2;
      if (ind <= 2) {
        hmag = hmax * std::pow(tol, 1.0 / 6.0);
      } else if (nfail <= 1.0) {
        temp = 2.0 * hmag;
        const double r = 2.0 / .9;
        if (tol < (r * r * r * r * r * r) * est) temp = .9 * std::pow(tol / est, 1.0 / 6.0) * hmag;
        hmag = std::fmax(temp, .5 * hmag);
      } else {
        hmag = .5 * hmag;
      }
      hmag = std::fmin(hmag, hmax);
      hmag = std::fmax(hmag, hmin);
      double xtrial;
      if (hmag >= std::fabs(xend - x)) {
        hmag = std::fabs(xend - x);
        xtrial = xend;
      } else {
        hmag = std::fmin(hmag, .5 * std::fabs(xend - x));
        xtrial = x + std::copysign(hmag, xend - x);
      }
      const double ht = xtrial - x;
      temp = ht / 1398169080000.0;
      double s[N];
      for (int k = 0; k < N; k++) s[k] = y[k] + temp * w[0][k] * 233028180000.0;
      fcn(x + ht / 6.0, s, w[1]);
      for (int k = 0; k < N; k++) s[k] = y[k] + temp * (w[0][k] * 74569017600.0 + w[1][k] * 298276070400.0);
      fcn(x + ht * (4.0 / 15.0), s, w[2]);
      for (int k = 0; k < N; k++)
        s[k] = y[k] + temp * (w[0][k] * 1165140900000.0 - w[1][k] * 3728450880000.0 + w[2][k] * 3495422700000.0);
      fcn(x + ht * (2.0 / 3.0), s, w[3]);
      for (int k = 0; k < N; k++)
        s[k] = y[k] + temp * (-w[0][k] * 3604654659375.0 + w[1][k] * 12816549900000.0 - w[2][k] * 9284716546875.0 +
                              w[3][k] * 1237962206250.0);
      fcn(x + ht * (5.0 / 6.0), s, w[4]);
      for (int k = 0; k < N; k++)
        s[k] = y[k] + temp * (w[0][k] * 3355605792000.0 - w[1][k] * 11185352640000.0 + w[2][k] * 9172628850000.0 -
                              w[3][k] * 427218330000.0 + w[4][k] * 482505408000.0);
      fcn(x + ht, s, w[5]);
      for (int k = 0; k < N; k++)
        s[k] = y[k] + temp * (-w[0][k] * 770204740536.0 + w[1][k] * 2311639545600.0 - w[2][k] * 1322092233000.0 -
                              w[3][k] * 453006781920.0 + w[4][k] * 326875481856.0);
      fcn(x + ht / 15.0, s, w[6]);
      for (int k = 0; k < N; k++)
        s[k] = y[k] + temp * (w[0][k] * 2845924389000.0 - w[1][k] * 9754668000000.0 + w[2][k] * 7897110375000.0 -
                              w[3][k] * 192082660000.0 + w[4][k] * 400298976000.0 + w[6][k] * 201586000000.0);
      fcn(x + ht, s, w[7]);
      for (int k = 0; k < N; k++)
        s[k] = y[k] + temp * (w[0][k] * 104862681000.0 + w[2][k] * 545186250000.0 + w[3][k] * 446637345000.0 +
                              w[4][k] * 188806464000.0 + w[6][k] * 15076875000.0 + w[7][k] * 97599465000.0);
      temp = 0.0;
      for (int k = 0; k < N; k++) {
        const double e = (w[0][k] * 8738556750.0 + w[2][k] * 9735468750.0 - w[3][k] * 9709507500.0 + w[4][k] * 8582112000.0 +
                          w[5][k] * 95329710000.0 - w[6][k] * 15076875000.0 - w[7][k] * 97599465000.0) /
                         1398169080000.0;
        temp = std::fmax(temp, std::fabs(e) / std::fmax(1.0, std::fabs(y[k])));
      }
      est = temp * hmag * 1.0;
      ind = est > tol ? 6 : 5;
      if (ind != 6) {
        x = xtrial;
        for (int k = 0; k < N; k++) y[k] = s[k];
        nfail = 0.0;
        if (x == xend) {
          c20 = xend;
          c21 = 1.0;
          return 3;
        }
      } else {
        nfail = nfail + 1.0;
        if (!(hmag > hmin)) return -3;
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------ Ranges
// camb/utils.F90:8-466.  One region = [Low, High) sampled in `steps` linear or logarithmic steps.
struct Region {
  int start_index = 0, steps = 0;
  bool IsLog = false;
  double Low = 0, High = 0, delta = 0, delta_max = 0, delta_min = 0;
};
struct Ranges {
  int count = 0, npoints = 0;
  double Lowest = 0, Highest = 0;
  Region R[101];
  std::vector<double> points, dpoints;  // 1-based
  void add(double t_start, double t_end, int nstep, bool WantLog);
  void add_delta(double t_start, double t_end, double t_approx_delta, bool IsLog = false) {  // :179-203
    const double RangeTol = 0.1;
    const int n = IsLog ? std::max(1, (int)(std::log(t_end / t_start) / t_approx_delta + 1.0 - RangeTol))
                        : std::max(1, (int)((t_end - t_start) / t_approx_delta + 1.0 - RangeTol));
    add(t_start, t_end, n, IsLog);
  }
  void get_array(bool want_dpoints) {  // :114-176
    points.assign(npoints + 1, 0.0);
    int ix = 0;
    for (int i = 1; i <= count; i++)
      for (int j = 0; j <= R[i].steps - 1; j++) points[++ix] = R[i].IsLog ? R[i].Low * std::exp(j * R[i].delta) : R[i].Low + R[i].delta * j;
    points[++ix] = Highest;
    if (!want_dpoints) return;
    const int n = npoints;
    dpoints.assign(n + 1, 0.0);
    for (int i = 2; i <= n - 1; i++) dpoints[i] = (points[i + 1] - points[i - 1]) / 2;
    dpoints[1] = (points[2] - points[1]) / 2;
    dpoints[n] = (points[n] - points[n - 1]) / 2;
  }
};

static void set_log_bounds(Region& A) {
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

// camb/utils.F90:206-466: merge a new region into the set -- sorted break points, finest requested spacing per
// piece (linear/log arbitration), absorption of one-step slivers into a neighbour, start indices
void Ranges::add(double t_start, double t_end, int nstep, bool WantLog) {
  const double RangeTol = 0.1;
  Region New[102];
  double End[204], Request[102];
  for (int i = 1; i <= count; i++) New[i] = R[i];
  const int nreg = count + 1;
  New[nreg] = Region();
  New[nreg].Low = t_start;
  New[nreg].High = t_end;
  New[nreg].delta = WantLog ? std::log(t_end / t_start) / nstep : (t_end - t_start) / nstep;
  New[nreg].steps = nstep;
  New[nreg].IsLog = WantLog;
  // break points in ascending order, in the reference's insertion order (:258-302), then without duplicates
  int ix = 0;
  auto insert_before_larger = [&](double v) {  // returns true if v went in front of a larger entry
    for (int j = 1; j <= ix; j++)
      if (v < End[j]) {
        for (int m = ix; m >= j; m--) End[m + 1] = End[m];
        End[j] = v;
        ix++;
        return true;
      }
    return false;
  };
  for (int i = 1; i <= nreg; i++) {
    const Region& A = New[i];
    if (ix == 0) {
      End[1] = A.Low;
      End[2] = A.High;
      ix = 2;
    } else if (!insert_before_larger(A.Low)) {
      End[++ix] = A.Low;
      End[++ix] = A.High;
    } else if (!insert_before_larger(A.High)) {
      End[++ix] = A.High;
    }
  }
  int nend = 1;
  for (int i = 2; i <= ix; i++)
    if (End[i] != End[nend]) End[++nend] = End[i];
  Lowest = End[1];
  Highest = End[nend];
  count = 0;
  const double max_delta = Highest - Lowest;
  for (int i = 1; i <= nend - 1; i++) {
    Region& A = R[i];
    A = Region();
    A.Low = End[i];
    A.High = End[i + 1];
    double delta = max_delta;
    for (int j = 1; j <= nreg; j++) {
      const Region& N = New[j];
      if (!(A.Low >= N.Low && A.Low < N.High)) continue;
      if (N.IsLog) {
        if (A.IsLog) {
          delta = std::fmin(delta, N.delta);
        } else {
          const double min_log_step = A.Low * (std::exp(N.delta) - 1);
          if (min_log_step < delta) {
            const double max_log_step = A.High * (1 - std::exp(-N.delta));
            if (delta < max_log_step) {
              delta = min_log_step;
            } else {
              A.IsLog = true;
              delta = N.delta;
            }
          }
        }
      } else if (A.IsLog) {
        const double max_log_step = A.High * (1 - std::exp(-delta));
        if (N.delta < max_log_step) {
          const double min_log_step = A.Low * (std::exp(delta) - 1);
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
    const double Diff = A.IsLog ? std::log(A.High / A.Low) : A.High - A.Low;
    if (delta >= Diff) {
      A.delta = Diff;
      A.steps = 1;
    } else {
      A.steps = std::max(1, (int)(Diff / delta + 1.0 - RangeTol));
      A.delta = Diff / A.steps;
    }
    Request[++count] = delta;
    set_log_bounds(A);
  }
  // one-step slivers join a neighbour whose spacing covers them (:390-441)
  for (int i = count; i >= 1; i--) {
    Region& A = R[i];
    if (A.steps != 1) continue;
    const double Diff = A.High - A.Low;
    const double min_request = A.IsLog ? A.Low * (std::exp(Request[i]) - 1) : Request[i];
    const double max_request = A.IsLog ? A.High * (1 - std::exp(-Request[i])) : Request[i];
    auto refit = [](Region& L) { L.delta = L.IsLog ? std::log(L.High / L.Low) / L.steps : (L.High - L.Low) / L.steps; };
    auto drop = [&](int at) {
      for (int m = at; m <= count - 1; m++) R[m] = R[m + 1];
      count--;
    };
    if (i != count) {
      Region& L = R[i + 1];
      if (Request[i] >= A.delta && Diff <= L.delta_min && L.delta_min <= max_request) {
        L.Low = A.Low;
        if (Diff > L.delta_min * RangeTol) L.steps++;
        refit(L);
        drop(i);
        continue;
      }
    }
    if (i != 1) {
      Region& L = R[i - 1];
      if (Request[i] >= A.delta && Diff <= L.delta_max && L.delta_max <= min_request) {
        L.High = A.High;
        if (Diff > L.delta_max * RangeTol) L.steps++;
        refit(L);
        drop(i);
      }
    }
  }
  int nsteps = 1;
  for (int i = 1; i <= count; i++) {
    R[i].start_index = nsteps;
    nsteps += R[i].steps;
    set_log_bounds(R[i]);
  }
  npoints = nsteps;
}

// ------------------------------------------------------------------------------------------------ massive neutrinos
// camb/modules.f90:1544-1716: log(rho), log(p) and their derivatives on the 2000-node log(am) grid.  The same for every
// cosmology -> built once per process.
const double nu_const = 7.0 / 120 * (const_pi * const_pi * const_pi * const_pi);
const double nu_const2 = 5.0 / 7 / (const_pi * const_pi);
const double zeta3 = 1.2020569031595942853997, zeta5 = 1.0369277551433699263313, zeta7 = 1.0083492773819228268397;
const double am_min = 0.01, am_max = 600.0;
const double am_minp = am_min * SP(1.1), am_maxp = am_max * SP(0.9);

struct NuTables {
  double dlnam = 0;
  std::vector<double> r1, p1, dr1, dp1, ddr1, ddp1;  // 1-based
  std::vector<double> rows;                           // [2000][6] = {r1, dr1, p1, dp1, ddr1, ddp1} (galileon_host / device layout)
};
static NuTables g_nu;
static std::once_flag g_nu_once;

static void nu_rho_pres(double am, double& rhonu, double& pnu) {  // :1683-1716
  const double qmax = 30.0;
  const int nq = 100;
  double dum1[nq + 2], dum2[nq + 2];
  const double adq = qmax / nq;
  dum1[1] = dum2[1] = 0;
  for (int i = 1; i <= nq; i++) {
    const double q = i * adq, aq = am / q;
    const double v = 1.0 / std::sqrt(1.0 + aq * aq);
    const double aqdn = adq * q * q * q / (std::exp(q) + 1.0);
    dum1[i + 1] = aqdn / v;
    dum2[i + 1] = aqdn * v;
  }
  rhonu = (splint(dum1, nq + 1) + dum1[nq + 1] / adq) / nu_const;
  pnu = (splint(dum2, nq + 1) + dum2[nq + 1] / adq) / nu_const / 3.0;
}

static void build_nu_tables() {
  NuTables& T = g_nu;
  for (auto* v : {&T.r1, &T.p1, &T.dr1, &T.dp1, &T.ddr1, &T.ddp1}) v->assign(NRHOPN + 1, 0.0);
  T.dlnam = -(std::log(am_min / am_max)) / (NRHOPN - 1);
  for (int i = 1; i <= NRHOPN; i++) {
    double rhonu, pnu;
    nu_rho_pres(am_min * std::exp((i - 1) * T.dlnam), rhonu, pnu);
    T.r1[i] = std::log(rhonu);
    T.p1[i] = std::log(pnu);
  }
  splder(T.r1.data(), T.dr1.data(), NRHOPN);
  splder(T.p1.data(), T.dp1.data(), NRHOPN);
  splder(T.dr1.data(), T.ddr1.data(), NRHOPN);
  splder(T.dp1.data(), T.ddp1.data(), NRHOPN);
  T.rows.resize((size_t)NRHOPN * 6);
  for (int i = 0; i < NRHOPN; i++) {
    double* r = &T.rows[(size_t)i * 6];
    r[0] = T.r1[i + 1]; r[1] = T.dr1[i + 1]; r[2] = T.p1[i + 1]; r[3] = T.dp1[i + 1]; r[4] = T.ddr1[i + 1]; r[5] = T.ddp1[i + 1];
  }
}

static inline double hermite4(const std::vector<double>& f, const std::vector<double>& df, int i, double d) {
  return f[i] + d * (df[i] + d * (3.0 * (f[i + 1] - f[i]) - 2.0 * df[i] - df[i + 1] + d * (df[i] + df[i + 1] + 2.0 * (f[i] - f[i + 1]))));
}

static double Nu_rho(double am) {  // :1757-1786
  if (am <= am_minp) return 1.0 + nu_const2 * am * am;
  if (am >= am_maxp) return 3 / (2 * nu_const) * (zeta3 * am + (15 * zeta5) / 2 / am);
  double d = std::log(am / am_min) / g_nu.dlnam + 1.0;
  const int i = (int)d;
  d = d - i;
  return std::exp(hermite4(g_nu.r1, g_nu.dr1, i, d));
}

// ------------------------------------------------------------------------------------------------ one model
struct Model {
  galcamb_params P;
  int status = 0;
  std::string err;
  // CAMBParams_Set
  double Max_eta_k = 0, omegak = 0, tau0 = 0, chi0 = 0, scale = 1;
  double grhom = 0, grhog = 0, grhor = 0, grhob = 0, grhoc = 0, grhov = 0, grhornomass = 0, grhok = 0, adotrad = 0;
  double akthom = 0, Nnow = 0;
  int Nu_mass_eigenstates = 0, Num_Nu_massive = 0, nqmax = 0;
  double Nu_mass_degeneracies[MAXNU + 1] = {0}, grhormass[MAXNU + 1] = {0}, nu_masses[MAXNU + 1] = {0};
  double nu_q[GALCAMB_MAX_NQ + 1] = {0}, nu_int_kernel[GALCAMB_MAX_NQ + 1] = {0};
  // Galileon background (ascending a; natural-spline coefficients c = y''/2 for hbar only: dtauda needs no x)
  double gal_coef[5] = {0};
  double gal_h0 = 0, gal_orad = 0;
  std::vector<double> gal_a, gal_h, gal_x, gal_ch;
  double gal_lnq_inv = 0;
  // reionization (camb/reionization.f90:79-86)
  bool reionization = true;
  double reion_fHe = 0, reion_tau_start = 0, reion_tau_complete = 0, reion_fraction = 0, reion_redshift = 0;
  double WindowVarMid = 0, WindowVarDelta = 0;
  // cmbmain module state
  double qmin = 0, qmax = 0, dtaurec = 0, max_etak_scalar = 0, maximum_qeta = 0;
  double taurst = 0, taurend = 0, tau_maxvis = 0;
  // recfast
  std::vector<double> zrec, xrec, dxrec;
  // thermo
  std::vector<double> cs2, dcs2, dotmu, ddotmu, dddotmu;  // 0-based copies for the ABI (filled last)
  double tauminn = 0, dlntau = 0, tight_tau = 0, matter_verydom_tau = 0;
  Ranges TimeSteps, Evolve_q, q_int;
  std::vector<double> vis, dvis, ddvis, expmmu, opac, dopac;  // 1-based over TimeSteps
  // GaugeInterface
  double epsw = 0;
  double nu_tau_notmassless[GALCAMB_MAX_NQ + 1][MAXNU + 1] = {{0}}, nu_tau_nonrelativistic[MAXNU + 1] = {0}, nu_tau_massive[MAXNU + 1] = {0};
  std::vector<int> ls;  // l samples, 0-based
  int kmaxfile = 0;
  // matter transfer functions: output times (ascending) and the wavenumbers beyond the source grid
  std::vector<double> tautf, k_transfer;
  double prof_bg_ms = 0, prof_rec_ms = 0;  // GALCAMB_PRESTAGE_PROFILE
  // Galileon background of a batch integrated on the device (prestage_into): 0 = integrate here on the host, 1 = stop in
  // params_set() once the inputs of the background are known, 2 = gal_a / gal_h / gal_x / gal_coef / bg_status are given
  int bg_mode = 0, bg_status = 0;
  // the tables in the device layout (csrc/pack.h), packed by the pre-stage thread that built the point
  std::vector<double> packed;
  gcpack::Info pinfo;

  int fail(int id, const char* msg) {
    status = id;
    err = msg;
    return id;
  }
  // hbar(a): natural cubic spline of the tabulated background (camb/galileon.cc:719-744, gsl_interp_cspline); the
  // interval comes from the log-uniform spacing of the nodes, corrected against the nodes themselves
  double gal_hbar(double a) const {
    const int n = (int)gal_a.size();
    int i = (int)(std::log(a / gal_a[0]) * gal_lnq_inv);
    i = std::max(0, std::min(i, n - 2));
    while (i > 0 && a < gal_a[i]) i--;
    while (i < n - 2 && a >= gal_a[i + 1]) i++;
    const double dx = gal_a[i + 1] - gal_a[i], dy = gal_h[i + 1] - gal_h[i], delx = a - gal_a[i];
    const double b = dy / dx - dx * (gal_ch[i + 1] + 2.0 * gal_ch[i]) / 3.0;
    const double d = (gal_ch[i + 1] - gal_ch[i]) / (3.0 * dx);
    return gal_h[i] + delx * (b + delx * (gal_ch[i] + delx * d));
  }
  // camb/equations_galileon.f90:107-156
  double dtauda(double a) const {
    const double a2 = a * a;
    double grhoa2;
    if (P.use_galileon && a >= 1e-10) {
      const double hbar = gal_hbar(a);
      grhoa2 = (a2 * a2) * grhom * (hbar * hbar);
    } else {
      grhoa2 = grhok * a2 + (grhoc + grhob) * a + grhog + grhornomass;
      grhoa2 = grhoa2 + grhov * (a2 * a2);
      if (Num_Nu_massive != 0)
        for (int nu_i = 1; nu_i <= Nu_mass_eigenstates; nu_i++) grhoa2 = grhoa2 + Nu_rho(a * nu_masses[nu_i]) * grhormass[nu_i];
    }
    return std::sqrt(3 / grhoa2);
  }
  double DeltaTime(double a1, double a2, double in_tol = -1) const {  // camb/modules.f90:557-580
    const double atol = in_tol > 0 ? in_tol : base_tol / 1000 / std::exp(P.AccuracyBoost - 1);
    return romberg([this](double a) { return dtauda(a); }, a1, a2, atol);
  }
  double DeltaTimeMaxed(double a1, double a2, double tol = -1) const {  // camb/equations_galileon.f90:471-481
    if (a1 > 1.0) return 0;
    if (a2 > 1.0) return DeltaTime(a1, 1.01, tol);
    return DeltaTime(a1, a2, tol);
  }
  int params_set();
  double reion_xe(double a, double xe_recomb, bool have_start) const;
  void reion_set_zre();
  int reion_init();
  int recombination();
  double recomb_xe(double a) const;
  int init_vars();
  void gauge_init();
  void k_grids();
  void transfer_grid();
  void l_samples();
  int run();
};

// camb/modules.f90:260-484 (scalar, flat) + init_massive_nu (:1856) + init_background (equations_galileon.f90:63-104)
int Model::params_set() {
  Max_eta_k = P.Max_eta_k;
  if (P.HighAccuracyDefault) Max_eta_k = std::fmax(std::min(P.Max_l, 3000) * 2.5, P.Max_eta_k);  // camb/camb.f90:130
  Num_Nu_massive = P.Num_Nu_massive;
  Nu_mass_eigenstates = P.Nu_mass_eigenstates;
  if (Nu_mass_eigenstates < 0 || Nu_mass_eigenstates > MAXNU) return fail(5, "Nu_mass_eigenstates out of range");
  double Num_Nu_massless = P.Num_Nu_massless;
  int mass_numbers[MAXNU];
  for (int i = 0; i < MAXNU; i++) {
    Nu_mass_degeneracies[i + 1] = P.Nu_mass_degeneracies[i];
    mass_numbers[i] = P.Nu_mass_numbers[i];
  }
  if (P.omegan == 0 && Num_Nu_massive != 0) {
    if (P.share_delta_neff) {
      Num_Nu_massless += Num_Nu_massive;
    } else {
      double s = 0;
      for (int i = 1; i <= Nu_mass_eigenstates; i++) s += Nu_mass_degeneracies[i];
      Num_Nu_massless += s;
    }
    Num_Nu_massive = 0;
    for (int i = 0; i < MAXNU; i++) mass_numbers[i] = 0;
  }
  double nu_massless_degeneracy = Num_Nu_massless;
  if (Num_Nu_massive > 0) {
    if (Nu_mass_eigenstates == 0) return fail(5, "Have Num_nu_massive>0 but no nu_mass_eigenstates");
    if (Nu_mass_eigenstates == 1 && mass_numbers[0] == 0) mass_numbers[0] = Num_Nu_massive;
    bool allzero = true;
    for (int i = 0; i < Nu_mass_eigenstates; i++) allzero = allzero && mass_numbers[i] == 0;
    if (allzero)
      for (int i = 0; i < MAXNU; i++) mass_numbers[i] = 1;
    if (P.share_delta_neff) {
      const double fractional_number = Num_Nu_massless + Num_Nu_massive;
      const int actual_massless = (int)(Num_Nu_massless + 1e-6);
      const double neff_i = fractional_number / (actual_massless + Num_Nu_massive);
      nu_massless_degeneracy = neff_i * actual_massless;
      for (int i = 1; i <= Nu_mass_eigenstates; i++) Nu_mass_degeneracies[i] = mass_numbers[i - 1] * neff_i;
    }
  } else {
    Nu_mass_eigenstates = 0;
  }
  omegak = 1 - (P.omegab + P.omegac + P.omegav + P.omegan);
  if (std::fabs(omegak) > 5e-7) return fail(5, "only the flat path is implemented");
  const double c = c_light;
  grhom = 3 * (P.H0 * P.H0) / (c * c) * (1000.0 * 1000.0);
  grhog = kappa / (c * c) * 4 * sigma_boltz / (c * c * c) * (P.TCMB * P.TCMB * P.TCMB * P.TCMB) * (Mpc * Mpc);
  grhor = 7.0 / 8 * std::pow(4.0 / 11, 4.0 / 3) * grhog;
  grhornomass = grhor * nu_massless_degeneracy;
  double sum_rm = 0;
  for (int nu_i = 1; nu_i <= Nu_mass_eigenstates; nu_i++) {
    grhormass[nu_i] = grhor * Nu_mass_degeneracies[nu_i];
    sum_rm += grhormass[nu_i];
  }
  grhoc = grhom * P.omegac;
  grhob = grhom * P.omegab;
  grhov = grhom * P.omegav;
  grhok = grhom * omegak;
  adotrad = std::sqrt((grhog + grhornomass + sum_rm) / 3);
  Nnow = P.omegab * (1 - P.YHe) * grhom * (c * c) / kappa / m_H / (Mpc * Mpc);
  akthom = sigma_thomson * Nnow * Mpc;
  // Nu_init, camb/modules.f90:1595-1680 (the tables themselves are process-wide)
  nqmax = 3;
  if (P.AccuracyBoost > 1) nqmax = 4;
  if (P.AccuracyBoost > 2) nqmax = 5;
  if (P.AccuracyBoost > 3) return fail(5, "AccuracyBoost > 3 needs more than 5 momentum nodes (not supported)");
  if (P.omegan != 0) {
    std::call_once(g_nu_once, build_nu_tables);
    for (int i = 1; i <= Nu_mass_eigenstates; i++)
      nu_masses[i] = nu_const / (1.5 * zeta3) * grhom / grhor * P.omegan * P.Nu_mass_fractions[i - 1] / Nu_mass_degeneracies[i];
    static const double q3[3] = {SP(0.913201), SP(3.37517), SP(7.79184)}, w3[3] = {SP(0.0687359), SP(3.31435), SP(2.29911)};
    static const double q4[4] = {SP(0.7), SP(2.62814), SP(5.90428), SP(12.0)}, w4[4] = {SP(0.0200251), SP(1.84539), SP(3.52736), SP(0.289427)};
    static const double q5[5] = {SP(0.583165), SP(2.0), SP(4.0), SP(7.26582), SP(13.0)},
                        w5[5] = {SP(0.0081201), SP(0.689407), SP(2.8063), SP(2.05156), SP(0.126817)};
    const double* q = nqmax == 3 ? q3 : nqmax == 4 ? q4 : q5;
    const double* w = nqmax == 3 ? w3 : nqmax == 4 ? w4 : w5;
    for (int i = 1; i <= nqmax; i++) {
      nu_q[i] = q[i - 1];
      nu_int_kernel[i] = w[i - 1] / nu_const;
    }
  }
  if (P.use_galileon) {
    const double omegar = (grhog + grhornomass) / grhom, omegam = P.omegab + P.omegac, hub = P.H0 / c * 1000;
    gal_h0 = hub;
    gal_orad = omegar;
    if (bg_mode == 1) return 0;
    const auto tb0 = std::chrono::steady_clock::now();
    const int st = bg_mode == 2 ? bg_status
                                : gc_galileon_background(omegar, omegam, hub, P.c2, P.c3, P.c4, P.cG, Nu_mass_eigenstates, grhormass + 1,
                                                         nu_masses + 1, g_nu.rows.empty() ? nullptr : g_nu.rows.data(), g_nu.dlnam,
                                                         gal_coef, gal_a, gal_h, gal_x);
    prof_bg_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb0).count();
    if (st != 0) return fail(6, "Didn't compute background successfully");
    // natural-spline coefficients of hbar (c = y''/2, gsl_interp_cspline): spline_natural gives y''
    const int n = (int)gal_a.size();
    std::vector<double> d2(n + 1);
    spline_natural(gal_a.data() - 1, gal_h.data() - 1, n, d2.data());
    gal_ch.resize(n);
    for (int i = 0; i < n; i++) gal_ch[i] = 0.5 * d2[i + 1];
    gal_lnq_inv = (n - 2) / std::log(gal_a[n - 2] / gal_a[0]);
  }
  tau0 = DeltaTime(0.0, 1.0);
  if (int rc = reion_init()) return rc;
  chi0 = tau0;
  scale = 1;
  return 0;
}

// camb/reionization.f90:64-98
double Model::reion_xe(double a, double xe_recomb, bool have_start) const {
  const double zexp = 1.5;
  const double xstart = have_start ? xe_recomb : 0.0;
  double xod = (WindowVarMid - 1.0 / std::pow(a, zexp)) / WindowVarDelta;
  double tgh = xod > 100 ? 1.0 : std::tanh(xod);
  double xe = (reion_fraction - xstart) * (tgh + 1.0) / 2.0 + xstart;
  if (a > (1 / (1 + P.helium_redshiftstart))) {
    xod = (1 + P.helium_redshift - 1.0 / a) / P.helium_delta_redshift;
    tgh = xod > 100 ? 1.0 : std::tanh(xod);
    xe = xe + reion_fHe * (tgh + 1.0) / 2.0;
  }
  return xe;
}

void Model::reion_set_zre() {  // :141-149
  const double zexp = 1.5;
  WindowVarMid = std::pow(1.0 + reion_redshift, zexp);
  WindowVarDelta = zexp * std::pow(1.0 + reion_redshift, zexp - 1.0) * P.delta_redshift;
}

int Model::reion_init() {  // :153-208, :261-367
  const double maxz = 50.0, tol = 1e-5;
  reion_fHe = P.YHe / (mass_ratio_He_H * (1.0 - P.YHe));
  reion_tau_start = reion_tau_complete = tau0;
  reionization = P.Reionization != 0;
  reion_fraction = P.reion_fraction;
  reion_redshift = P.reion_redshift;
  if (reionization && ((P.use_optical_depth && P.optical_depth < SP(0.001)) || (!P.use_optical_depth && P.reion_redshift < SP(0.001))))
    reionization = false;
  if (!reionization) return 0;
  if (reion_fraction == -1.0) reion_fraction = 1.0 + reion_fHe;
  if (P.use_optical_depth) {
    double try_b = 0, try_t = maxz, tau = 0;
    for (int i = 1;; i++) {
      reion_redshift = (try_t + try_b) / 2;
      reion_set_zre();
      tau = romberg([this](double z) { const double a = 1.0 / (1.0 + z); return reion_xe(a, 0, false) * akthom * dtauda(a); }, 0.0, maxz,
                    tol, 20, (int)std::lround(maxz / P.delta_redshift * 5));
      if (tau > P.optical_depth) try_t = reion_redshift; else try_b = reion_redshift;
      if (std::fabs(try_b - try_t) < SP(2e-3) / 1.0) break;
      if (i > 100) return fail(1, "Reionization_zreFromOptDepth: failed to converge");
    }
    if (std::fabs(tau - P.optical_depth) > SP(0.002)) return fail(1, "Reionization_zreFromOptDepth: Did not converge to optical depth");
  }
  reion_set_zre();
  const double astart = 1.0 / (1.0 + reion_redshift + P.delta_redshift * 8);
  auto f = [this](double a) { return dtauda(a); };
  reion_tau_start = std::fmax(0.05, romberg(f, 0.0, astart, 1e-3));
  reion_tau_complete = std::fmin(tau0, reion_tau_start + romberg(f, astart, 1.0 / (1.0 + std::fmax(0.0, reion_redshift - P.delta_redshift * 8)), 1e-3));
  return 0;
}

// ------------------------------------------------------------------------------------------------ RECFAST 1.5.2
namespace rec {
const double Lambda = 8.2245809, Lambda_He = 51.3;
const double L_H_ion = 1.096787737e7, L_H_alpha = 8.225916453e6, L_He1_ion = 1.98310772e7, L_He2_ion = 4.389088863e7;
const double L_He_2s = 1.66277434e7, L_He_2p = 1.71134891e7;
const double A2P_s = 1.798287e9, A2P_t = 177.58, L_He_2Pt = 1.690871466e7, L_He_2St = 1.5985597526e7;
const double L_He2St_ion = 3.8454693845e6, sigma_He_2Ps = 1.436289e-22, sigma_He_2Pt = 1.484872e-22;
const double bigH = 100.0e3 / Mpc, not4 = mass_ratio_He_H;
const double zinitial = 1e4, zfinal = 0.0;
const int Nz = 10000;
const double delta_z = (zinitial - zfinal) / Nz;
const double AGauss1 = -0.14, AGauss2 = 0.079, zGauss1 = 7.28, zGauss2 = 6.73, wGauss1 = 0.18, wGauss2 = 0.33;
const double Pi = pi_aml, G = G_newton, C = c_light;

struct Data {  // RECDATA, camb/recfast.f90:200-236 / :480-560
  double Lalpha, Lalpha_He, DeltaB, CDB, DeltaB_He, CDB_He, CB1, CB1_He1, CB1_He2, CR, CK, CK_He, CL, CL_He, CT, Bfact, H_frac,
      fu, mu_H, mu_T, Tnow, HO, OmegaK, OmegaT, z_eq, Nnow, fHe;
};

// the three Saha branches of camb/recfast.f90:725-776 / :620-680
static double saha_He2(const Data& R, double z) {
  const double rhs = std::exp(1.5 * std::log(R.CR * R.Tnow / (1.0 + z)) - R.CB1_He2 / (R.Tnow * (1.0 + z))) / R.Nnow * 1.0;
  return 0.5 * (std::sqrt((rhs - 1.0 - R.fHe) * (rhs - 1.0 - R.fHe) + 4.0 * (1.0 + 2.0 * R.fHe) * rhs) - (rhs - 1.0 - R.fHe));
}
static double saha_He1(const Data& R, double z) {
  const double rhs = std::exp(1.5 * std::log(R.CR * R.Tnow / (1.0 + z)) - R.CB1_He1 / (R.Tnow * (1.0 + z))) / R.Nnow * 4.0;
  return 0.5 * (std::sqrt((rhs - 1.0) * (rhs - 1.0) + 4.0 * (1.0 + R.fHe) * rhs) - (rhs - 1.0));
}
static double saha_H(const Data& R, double z) {
  const double rhs = std::exp(1.5 * std::log(R.CR * R.Tnow / (1.0 + z)) - R.CB1 / (R.Tnow * (1.0 + z))) / R.Nnow;
  return 0.5 * (std::sqrt(rhs * rhs + 4.0 * rhs) - rhs);
}
}  // namespace rec

// camb/recfast.f90:460-722 with ION (:778-1012) as the right-hand side
int Model::recombination() {
  using namespace rec;
  Data R;
  zrec.assign(Nz + 1, 0);
  xrec.assign(Nz + 1, 0);
  dxrec.assign(Nz + 1, 0);
  R.Tnow = P.TCMB;
  R.OmegaT = P.omegac + P.omegab;
  R.OmegaK = 1.0 - R.OmegaT - P.omegav;
  R.HO = P.H0 / 100.0 * bigH;
  const double Yp = P.YHe;
  R.mu_H = 1.0 / (1.0 - Yp);
  R.mu_T = not4 / (not4 - (not4 - 1.0) * Yp);
  R.fHe = Yp / (not4 * (1.0 - Yp));
  R.Nnow = 3.0 * R.HO * R.HO * P.omegab / (8.0 * Pi * G * R.mu_H * m_H);
  const double fnu = (21.0 / 8.0) * std::pow(4.0 / 11.0, 4.0 / 3.0);
  const double Tnow4 = R.Tnow * R.Tnow * R.Tnow * R.Tnow;
  R.z_eq = (3.0 * ((R.HO * C) * (R.HO * C)) / (8.0 * Pi * G * a_rad() * (1.0 + fnu) * Tnow4)) * (P.omegab + P.omegac);
  R.z_eq = R.z_eq - 1.0;
  R.Lalpha = 1.0 / L_H_alpha;
  R.Lalpha_He = 1.0 / L_He_2p;
  R.DeltaB = h_P * C * (L_H_ion - L_H_alpha);
  R.CDB = R.DeltaB / k_B;
  R.DeltaB_He = h_P * C * (L_He1_ion - L_He_2s);
  R.CDB_He = R.DeltaB_He / k_B;
  R.CB1 = h_P * C * L_H_ion / k_B;
  R.CB1_He1 = h_P * C * L_He1_ion / k_B;
  R.CB1_He2 = h_P * C * L_He2_ion / k_B;
  R.CR = 2.0 * Pi * (m_e / h_P) * (k_B / h_P);
  R.CK = (R.Lalpha * R.Lalpha * R.Lalpha) / (8.0 * Pi);
  R.CK_He = (R.Lalpha_He * R.Lalpha_He * R.Lalpha_He) / (8.0 * Pi);
  R.CL = C * h_P / (k_B * R.Lalpha);
  R.CL_He = C * h_P / (k_B / L_He_2s);
  R.CT = Compton_CT() / MPC_in_sec;
  R.Bfact = h_P * C * (L_He_2p - L_He_2s) / k_B;
  R.H_frac = 1e-3;
  R.fu = P.RECFAST_fudge;
  const int Hswitch = P.RECFAST_Hswitch, Heswitch = P.RECFAST_Heswitch;
  const double fudge_He = P.RECFAST_fudge_He;
  // temperature-independent pieces of ION
  const double a_PPB = 4.309, b_PPB = -0.6166, c_PPB = 0.6703, d_PPB = 0.5300;
  const double a_VF = std::pow(10.0, -16.744), b_VF = 0.711, T_0 = std::pow(10.0, 0.477121), T_1 = std::pow(10.0, 5.114);
  const double a_trip = std::pow(10.0, -16.306), b_trip = 0.761;
  auto ION = [&](double z, const double* y, double* f) {
    const double fHe = R.fHe, Nnow = R.Nnow, Tnow = R.Tnow;
    const double x_H = y[0], x_He = y[1], x = x_H + fHe * x_He, Tmat = y[2];
    const double zp1 = 1.0 + z, zp13 = zp1 * zp1 * zp1;
    const double n = Nnow * zp13, n_He = fHe * Nnow * zp13, Trad = Tnow * zp1;
    const double Hz = 1 / dtauda(1 / zp1) * (zp1 * zp1) / MPC_in_sec;
    // Only the terms an equation reads are formed: y'(x_H) is zero above x_H = 0.99 and the helium equation is zero below
    // x_He = 1e-15 (camb/recfast.f90:1039-1073), so the hydrogen rates are skipped during helium recombination and the
    // helium rates afterwards -- same expressions, same results, about half of the transcendental calls.
    const bool needH = !(x_H > SP(0.99)), needHe = !(x_He < SP(1.e-15));
    const double crt15 = (needH || needHe) ? std::pow(R.CR * Tmat, 1.5) : 0.0;
    double Rdown = 0, Rup = 0, K = 0;
    if (needH) {
      Rdown = 1.e-19 * a_PPB * std::pow(Tmat / 1.e4, b_PPB) / (1.0 + c_PPB * std::pow(Tmat / 1.e4, d_PPB));
      Rup = Rdown * crt15 * std::exp(-R.CDB / Tmat);
      if (!(x_H > 0.985)) {
        if (!Hswitch) {
          K = R.CK / Hz;
        } else {
          const double lz = std::log(zp1);
          const double g1 = (lz - zGauss1) / wGauss1, g2 = (lz - zGauss2) / wGauss2;
          K = R.CK / Hz * (1.0 + AGauss1 * std::exp(-(g1 * g1)) + AGauss2 * std::exp(-(g2 * g2)));
        }
      }
    }
    double Rdown_He = 0, Rup_He = 0, He_Boltz = 0, Rdown_trip = 0, Rup_trip = 0, K_He = 0, CfHe_t = 0;
    const int Heflag = ((x_He < 5.e-9) || (x_He > 0.98)) ? 0 : Heswitch;
    if (needHe) {
      const double sq_0 = std::sqrt(Tmat / T_0), sq_1 = std::sqrt(Tmat / T_1);
      Rdown_He = a_VF / (sq_0 * std::pow(1.0 + sq_0, 1.0 - b_VF));
      Rdown_He = Rdown_He / std::pow(1.0 + sq_1, 1.0 + b_VF);
      Rup_He = Rdown_He * crt15 * std::exp(-R.CDB_He / Tmat);
      Rup_He = 4.0 * Rup_He;
      He_Boltz = (R.Bfact / Tmat) > 680.0 ? std::exp(680.0) : std::exp(R.Bfact / Tmat);
      if (Heflag >= 3) {  // the triplet rates are read only then (:1004-1030, :1068-1073)
        Rdown_trip = a_trip / (sq_0 * std::pow(1.0 + sq_0, 1.0 - b_trip));
        Rdown_trip = Rdown_trip / std::pow(1.0 + sq_1, 1.0 + b_trip);
        Rup_trip = Rdown_trip * std::exp(-h_P * C * L_He2St_ion / (k_B * Tmat));
        Rup_trip = Rup_trip * crt15 * (4.0 / 3.0);
      }
      if (Heflag == 0) {
        K_He = R.CK_He / Hz;
      } else {
        const double tauHe_s = A2P_s * R.CK_He * 3.0 * n_He * (1.0 - x_He) / Hz;
        const double pHe_s = (1.0 - std::exp(-tauHe_s)) / tauHe_s;
        K_He = 1.0 / (A2P_s * pHe_s * 3.0 * n_He * (1.0 - x_He));
        if (((Heflag == 2) || (Heflag >= 5)) && x_H < 0.9999999) {
          double Doppler = 2.0 * k_B * Tmat / (m_H * not4 * C * C);
          Doppler = C * L_He_2p * std::sqrt(Doppler);
          const double gamma_2Ps = 3.0 * A2P_s * fHe * (1.0 - x_He) * C * C /
                                   (std::sqrt(Pi) * sigma_He_2Ps * 8.0 * Pi * Doppler * (1.0 - x_H)) / ((C * L_He_2p) * (C * L_He_2p));
          const double AHcon = A2P_s / (1.0 + 0.36 * std::pow(gamma_2Ps, fudge_He));
          K_He = 1.0 / ((A2P_s * pHe_s + AHcon) * 3.0 * n_He * (1.0 - x_He));
        }
        if (Heflag >= 3) {
          double tauHe_t = A2P_t * n_He * (1.0 - x_He) * 3.0;
          tauHe_t = tauHe_t / (8.0 * Pi * Hz * (L_He_2Pt * L_He_2Pt * L_He_2Pt));
          const double pHe_t = (1.0 - std::exp(-tauHe_t)) / tauHe_t;
          const double CL_PSt = h_P * C * (L_He_2Pt - L_He_2St) / k_B;
          if ((Heflag == 3) || (Heflag == 5) || (x_H > 0.99999)) {
            CfHe_t = A2P_t * pHe_t * std::exp(-CL_PSt / Tmat);
            CfHe_t = CfHe_t / (Rup_trip + CfHe_t);
          } else {
            double Doppler = 2.0 * k_B * Tmat / (m_H * not4 * C * C);
            Doppler = C * L_He_2Pt * std::sqrt(Doppler);
            const double gamma_2Pt = 3.0 * A2P_t * fHe * (1.0 - x_He) * C * C /
                                     (std::sqrt(Pi) * sigma_He_2Pt * 8.0 * Pi * Doppler * (1.0 - x_H)) / ((C * L_He_2Pt) * (C * L_He_2Pt));
            const double AHcon = A2P_t / (1.0 + 0.66 * std::pow(gamma_2Pt, 0.9)) / 3.0;
            CfHe_t = (A2P_t * pHe_t + AHcon) * std::exp(-CL_PSt / Tmat);
            CfHe_t = CfHe_t / (Rup_trip + CfHe_t);
          }
        }
      }
    }
    const double Trad4 = Trad * Trad * Trad * Trad;
    const double timeTh = (1.0 / (R.CT * Trad4)) * (1.0 + x + fHe) / x;
    const double timeH = 2. / (3. * R.HO * std::pow(zp1, 1.5));
    if (x_H > SP(0.99)) {
      f[0] = 0.0;
    } else if (x_H > 0.985) {
      f[0] = (x * x_H * n * Rdown - Rup * (1.0 - x_H) * std::exp(-R.CL / Tmat)) / (Hz * zp1);
    } else {
      f[0] = ((x * x_H * n * Rdown - Rup * (1.0 - x_H) * std::exp(-R.CL / Tmat)) * (1.0 + K * Lambda * n * (1.0 - x_H))) /
             (Hz * zp1 * (1.0 / R.fu + K * Lambda * n * (1.0 - x_H) / R.fu + K * Rup * n * (1.0 - x_H)));
    }
    if (x_He < SP(1.e-15)) {
      f[1] = 0.0;
    } else {
      f[1] = ((x * x_He * n * Rdown_He - Rup_He * (1 - x_He) * std::exp(-R.CL_He / Tmat)) *
              (1 + K_He * Lambda_He * n_He * (1.0 - x_He) * He_Boltz)) /
             (Hz * zp1 * (1 + K_He * (Lambda_He + Rup_He) * n_He * (1.0 - x_He) * He_Boltz));
      if (Heflag >= 3)
        f[1] = f[1] + (x * x_He * n * Rdown_trip - (1.0 - x_He) * 3.0 * Rup_trip * std::exp(-h_P * C * L_He_2St / (k_B * Tmat))) *
                          CfHe_t / (Hz * zp1);
    }
    if (timeTh < R.H_frac * timeH) {
      const double dHdz = (R.HO * R.HO / 2.0 / Hz) * (4.0 * zp13 / (1.0 + R.z_eq) * R.OmegaT + 3.0 * R.OmegaT * (zp1 * zp1) + 2.0 * R.OmegaK * zp1);
      const double epsilon = Hz * (1.0 + x + fHe) / (R.CT * (Trad * Trad * Trad) * x);
      f[2] = Tnow + epsilon * ((1.0 + fHe) / (1.0 + fHe + x)) * ((f[0] + fHe * f[1]) / x) - epsilon * dHdz / Hz + 3.0 * epsilon / zp1;
    } else {
      f[2] = R.CT * Trad4 * x / (1.0 + x + fHe) * (Tmat - Trad) / (Hz * zp1) + 2.0 * Tmat / zp1;
    }
  };
  double z = zinitial;
  double y[3] = {1.0, 1.0, R.Tnow * (1.0 + z)};
  Dverk<3> dv;
  int ind = 1;
  const double tol = 1e-5;
  for (int i = 1; i <= Nz; i++) {
    double zstart = zinitial - (double)(i - 1) * delta_z;
    const double zend = zinitial - (double)i * delta_z;
    z = zend;
    double x0;
    if (zend > 8000.0) {
      x0 = 1.0 + 2.0 * R.fHe;
      y[0] = 1.0; y[1] = 1.0; y[2] = R.Tnow * (1.0 + z);
    } else if (z > 5000.0) {
      x0 = saha_He2(R, z);
      y[0] = 1.0; y[1] = 1.0; y[2] = R.Tnow * (1.0 + z);
    } else if (z > 3500.0) {
      x0 = 1.0 + R.fHe * 1.0;
      y[0] = 1.0; y[1] = 1.0; y[2] = R.Tnow * (1.0 + z);
    } else if (y[1] > SP(0.99)) {
      x0 = saha_He1(R, z);
      y[0] = 1.0; y[1] = (x0 - 1.0) / R.fHe; y[2] = R.Tnow * (1.0 + z);
    } else if (y[0] > 0.99) {
      const double x_H0 = saha_H(R, z);
      ind = dv.run(ION, zstart, y, zend, tol, ind);
      if (ind < 0) return fail(2, "RECFAST: dverk failed");
      y[0] = x_H0;
      x0 = y[0] + R.fHe * y[1];
    } else {
      ind = dv.run(ION, zstart, y, zend, tol, ind);
      if (ind < 0) return fail(2, "RECFAST: dverk failed");
      x0 = y[0] + R.fHe * y[1];
    }
    zrec[i] = zend;
    xrec[i] = x0;
  }
  spline_natural(zrec.data(), xrec.data(), Nz, dxrec.data());
  return 0;
}

double Model::recomb_xe(double a) const {  // camb/recfast.f90:434-458
  using namespace rec;
  const double z = 1 / a - 1;
  if (z >= zrec[1]) return xrec[1];
  if (z <= zrec[Nz]) return xrec[Nz];
  const double zst = (zinitial - z) / delta_z;
  const int ihi = (int)zst, ilo = ihi + 1;
  const double az = zst - (int)zst, bz = 1 - az;
  return az * xrec[ilo] + bz * xrec[ihi] + ((az * az * az - az) * dxrec[ilo] + (bz * bz * bz - bz) * dxrec[ihi]) / 6.0;
}

// ------------------------------------------------------------------------------------------------ InitVars + inithermo
static inline double herm(const std::vector<double>& f, const std::vector<double>& df, int i, double d) {
  return f[i] + d * (df[i] + d * (3 * (f[i + 1] - f[i]) - 2 * df[i] - df[i + 1] + d * (df[i] + df[i + 1] + 2 * (f[i] - f[i + 1]))));
}

int Model::init_vars() {
  // camb/cmbmain.f90:722-795
  maximum_qeta = Max_eta_k;
  qmax = maximum_qeta / tau0;
  qmin = 0.1 / tau0 / P.AccuracyBoost;
  dtaurec = 4 / qmax / P.AccuracyBoost;
  max_etak_scalar = P.AccuracyBoost * std::fmax(1700.0, maximum_qeta) / 20;
  if (maximum_qeta < 3500 && P.AccuracyBoost < 2) max_etak_scalar = max_etak_scalar * 1.5;
  // GetTauStart(maxq), camb/cmbmain.f90:634-660, maxq of :757-762 (WantCls)
  const double maxq = P.WantTransfer ? std::fmax(qmax, P.transfer_kmax) : qmax;
  double taumin = std::fmin(0.001 / maxq, 0.1);
  if (Num_Nu_massive > 0) {
    double mx = 0;
    for (int i = 1; i <= Nu_mass_eigenstates; i++) mx = std::fmax(mx, nu_masses[i]);
    taumin = std::fmin(taumin, 1.e-3 / mx / adotrad);
  }
  // inithermo(taumin, tau0), camb/modules.f90:2787-3103
  const auto tr0 = std::chrono::steady_clock::now();
  if (int rc = recombination()) return rc;
  prof_rec_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr0).count();
  const int nt = NTHERMO;
  std::vector<double> tb(nt + 1), c2(nt + 1), xe(nt + 1), dc2(nt + 1), dm(nt + 1), ddm(nt + 1), sdm(nt + 1), emmu(nt + 1), demmu(nt + 1),
      dddm(nt + 1), ddddm(nt + 1);
  const double Maxtau = tau0;
  int ncount = 0;
  const double thomc0 = Compton_CT() * (P.TCMB * P.TCMB * P.TCMB * P.TCMB);
  tauminn = 0.05 * taumin;
  dlntau = std::log(tau0 / tauminn) / (nt - 1);
  tight_tau = 0;
  matter_verydom_tau = 0;
  const double a_verydom = P.AccuracyBoost * 5 * (grhog + grhornomass) / (grhoc + grhob);
  double tau01 = tauminn, adot0 = adotrad, a0 = adotrad * tauminn;
  tb[1] = P.TCMB / a0;
  xe[1] = 1.0 + 0.25 * P.YHe / (1.0 - P.YHe) * (0.0 + 2 * 1.0);
  double barssc = barssc0 * (1.0 - 0.75 * P.YHe + (1.0 - P.YHe) * xe[1]);
  c2[1] = 4.0 / 3.0 * barssc * tb[1];
  dm[1] = xe[1] * akthom / (a0 * a0);
  sdm[1] = 0;
  for (int i = 2; i <= nt; i++) {
    const double tau = tauminn * std::exp((i - 1) * dlntau);
    const double dtau = tau - tau01;
    double a = a0 + adot0 * dtau;
    const double a2 = a * a;  // the reference keeps the a^2 of this first-order estimate for dotmu (:2893, :2935)
    const double adot = 1 / dtauda(a);
    if (matter_verydom_tau == 0 && a > a_verydom) matter_verydom_tau = tau;
    a = a0 + 2.0 * dtau / (1.0 / adot0 + 1.0 / adot);
    const double tg0 = P.TCMB / a0, ahalf = 0.5 * (a0 + a), adothalf = 0.5 * (adot0 + adot);
    const double fe = (1.0 - P.YHe) * xe[i - 1] / (1.0 - 0.75 * P.YHe + (1.0 - P.YHe) * xe[i - 1]);
    const double thomc = thomc0 * fe / adothalf / (ahalf * ahalf * ahalf);
    const double etc = std::exp(-thomc * (a - a0));
    const double a2t = a0 * a0 * (tb[i - 1] - tg0) * etc - P.TCMB / thomc * (1.0 - etc);
    tb[i] = P.TCMB / a + a2t / (a * a);
    if (reionization && tau > reion_tau_start) {
      if (ncount == 0) ncount = i - 1;
      xe[i] = reion_xe(a, xe[ncount], true);
    } else {
      xe[i] = recomb_xe(a);
    }
    const double dtbdla = -2.0 * tb[i] - thomc * adothalf / adot * (a * tb[i] - P.TCMB);
    barssc = barssc0 * (1.0 - 0.75 * P.YHe + (1.0 - P.YHe) * xe[i]);
    c2[i] = barssc * tb[i] * (1 - dtbdla / tb[i] / 3.0);
    dm[i] = xe[i] * akthom / a2;
    if (tight_tau == 0 && 1 / (tau * dm[i]) > SP(0.005)) tight_tau = tau;
    sdm[i] = tau < SP(0.001) ? 0 : sdm[i - 1] + 2.0 * dtau / (1.0 / dm[i] + 1.0 / dm[i - 1]);
    a0 = a;
    tau01 = tau;
    adot0 = adot;
  }
  for (int j1 = 1; j1 <= nt; j1++) emmu[j1] = (sdm[j1] - sdm[nt] < -69) ? 1.e-30 : std::exp(sdm[j1] - sdm[nt]);
  int iv = 0;
  double vfi = 0.0, maxvis = 0;
  const double cf1 = ncount == 0 ? 1.0 : std::exp(sdm[nt] - sdm[ncount]);
  const int ns = ncount == 0 ? nt : ncount;
  for (int j1 = 1; j1 <= ns; j1++) {
    const double v = emmu[j1] * dm[j1];
    const double tau = tauminn * std::exp((j1 - 1) * dlntau);
    vfi = vfi + v * cf1 * dlntau * tau;
    if ((iv == 0) && (vfi > 1.0e-7 / P.AccuracyBoost)) {
      taurst = 9.0 / 10.0 * tau;
      iv = 1;
    } else if (iv == 1) {
      if (v > maxvis) {
        maxvis = v;
        tau_maxvis = tau;
      }
      if (vfi > SP(0.995)) {
        taurend = tau;
        iv = 2;
        break;
      }
    }
  }
  if (iv != 2) return fail(1, "inithermo: failed to find end of recombination");
  dtaurec = std::fmin(dtaurec, taurst / 40) / P.AccuracyBoost;
  if (reionization) taurend = std::fmin(taurend, reion_tau_start);
  splder(c2.data(), dc2.data(), nt);
  splder(dm.data(), ddm.data(), nt);
  splder(ddm.data(), dddm.data(), nt);
  splder(dddm.data(), ddddm.data(), nt);
  splder(emmu.data(), demmu.data(), nt);
  // SetTimeSteps, camb/modules.f90:3106-3146
  TimeSteps = Ranges();
  TimeSteps.add_delta(taurst, taurend, dtaurec);
  TimeSteps.add_delta(taurend, tau0, Maxtau / 500.0 / P.AccuracyBoost);
  if (reionization) TimeSteps.add(reion_tau_start, reion_tau_complete, (int)(50 * P.AccuracyBoost), false);
  TimeSteps.get_array(true);
  const int nstep = TimeSteps.npoints;
  for (auto* v : {&vis, &dvis, &ddvis, &expmmu, &opac, &dopac}) v->assign(nstep + 1, 0.0);
  for (int j2 = 1; j2 <= nstep; j2++) {  // DoThermoSpline, :3159-3202
    const double tau = TimeSteps.points[j2];
    double d = std::log(tau / tauminn) / dlntau + 1.0;
    const int i = (int)d;
    d = d - i;
    double ddopac;
    if (i < nt) {
      opac[j2] = herm(dm, ddm, i, d);
      dopac[j2] = herm(ddm, dddm, i, d) / (tau * dlntau);
      ddopac = (herm(dddm, ddddm, i, d) - (dlntau * dlntau) * tau * dopac[j2]) / ((tau * dlntau) * (tau * dlntau));
      expmmu[j2] = herm(emmu, demmu, i, d);
    } else {
      opac[j2] = dm[nt];
      dopac[j2] = ddm[nt];
      ddopac = dddm[nt];
      expmmu[j2] = emmu[nt];
    }
    const double op = opac[j2];
    vis[j2] = op * expmmu[j2];
    dvis[j2] = expmmu[j2] * (op * op + dopac[j2]);
    ddvis[j2] = expmmu[j2] * (op * op * op + 3 * op * dopac[j2] + ddopac);
  }
  cs2.assign(c2.begin() + 1, c2.end());
  dcs2.assign(dc2.begin() + 1, dc2.end());
  dotmu.assign(dm.begin() + 1, dm.end());
  ddotmu.assign(ddm.begin() + 1, ddm.end());
  dddotmu.assign(dddm.begin() + 1, dddm.end());
  gauge_init();
  // times of the transfer-function outputs, camb/cmbmain.f90:783-793 (TimeOfz, camb/modules.f90:574-580)
  tautf.clear();
  if (P.WantTransfer) {
    if (P.transfer_high_precision) return fail(5, "transfer_high_precision is not supported");
    if (P.transfer_num_redshifts < 1 || P.transfer_num_redshifts > GALCAMB_MAX_TRANSFER_Z)
      return fail(5, "transfer_num_redshifts out of range");
    for (int itf = 0; itf < P.transfer_num_redshifts; itf++) {
      tautf.push_back(std::fmin(DeltaTime(0.0, 1.0 / (P.transfer_redshifts[itf] + 1.0)), tau0));
      if (itf > 0 && tautf[itf] <= tautf[itf - 1]) return fail(5, "Transfer redshifts not set or out of order");
    }
  }
  return 0;
}

// camb/equations_galileon.f90:510-553
void Model::gauge_init() {
  epsw = 100 / tau0;
  for (int nu_i = 1; nu_i <= Nu_mass_eigenstates; nu_i++) {
    const double nu_mass = std::fmax(0.1, nu_masses[nu_i]);
    const double a_mass = 1.e-1 / nu_mass / (double)(float)P.lAccuracyBoost;
    double time = DeltaTime(0.0, nu_q[1] * a_mass);
    nu_tau_notmassless[1][nu_i] = time;
    for (int j = 2; j <= nqmax; j++) {
      time = time + DeltaTimeMaxed(nu_q[j - 1] * a_mass, nu_q[j] * a_mass, 0.01);
      nu_tau_notmassless[j][nu_i] = time;
    }
    const double a_nonrel = 2.5 / nu_mass * P.AccuracyBoost;
    nu_tau_nonrelativistic[nu_i] = DeltaTimeMaxed(0.0, a_nonrel);
    const double a_massive = 17.0 / nu_mass * P.AccuracyBoost;
    nu_tau_massive[nu_i] = nu_tau_nonrelativistic[nu_i] + DeltaTimeMaxed(a_nonrel, a_massive);
  }
}

// SetkValuesForSources camb/cmbmain.f90:797-852 and SetkValuesForInt :1281-1353
void Model::k_grids() {
  const double B = P.AccuracyBoost;
  double dlnk0 = (reionization && P.AccuratePolarization) ? 2.0 / 10 / B : 5.0 / 10 / B;
  if (P.AccurateReionization) dlnk0 = dlnk0 / 2;
  const double dkn1 = 0.6 / taurst / B;
  double dkn2 = 0.9 / taurst / B;
  if (P.HighAccuracyDefault) dkn2 = dkn2 / SP(1.2);
  const double qmax_log = dkn1 / dlnk0;
  const double q_switch = (double)(2 * 6.3f) / taurst;
  double q_cmb = 2 * 3000 / chi0 * B;
  if (P.Max_l > 5000 && P.AccuratePolarization) q_cmb = q_cmb * SP(1.4);
  double dksmooth = q_cmb / 2 / (B * B);
  dksmooth = dksmooth / 6;
  Evolve_q = Ranges();
  Evolve_q.add_delta(qmin, qmax_log, dlnk0, true);
  Evolve_q.add_delta(qmax_log, std::fmin(qmax, q_switch), dkn1);
  if (qmax > q_switch) {
    Evolve_q.add_delta(q_switch, std::fmin(q_cmb, qmax), dkn2);
    if (qmax > q_cmb) {
      dksmooth = std::log(1 + dksmooth / q_cmb);
      Evolve_q.add_delta(q_cmb, qmax, dksmooth, true);
    }
  }
  Evolve_q.get_array(false);
  // integration grid (max_bessels_etak = maximum_qeta, camb/cmbmain.f90:255)
  const double qmax_int = std::fmin(qmax, maximum_qeta / tau0);
  const int lognum = (int)std::lround(10 * B);
  const double dlnk1 = 1.0 / lognum;
  const int no = (int)std::lround(600 * B);
  const double dk0 = 1.8 / chi0 / B;
  double dk = 3.0 / chi0 / B;
  if (P.HighAccuracyDefault) dk = dk / SP(1.6);
  const double k_max_log = lognum * dk0, k_max_0 = no * dk0, dk2 = SP(0.04) / B;
  q_int = Ranges();
  q_int.add_delta(qmin, k_max_log, dlnk1, true);
  q_int.add_delta(k_max_log, std::fmin(qmax_int, k_max_0), dk0);
  if (qmax_int > k_max_0) {
    const double max_k_dk = std::max(3000, 2 * P.Max_l) / tau0;
    q_int.add_delta(k_max_0, std::fmin(qmax_int, max_k_dk), dk);
    if (qmax_int > max_k_dk) q_int.add_delta(max_k_dk, qmax_int, dk2);
  }
  q_int.get_array(true);
  kmaxfile = (int)std::fmin(maximum_qeta, Max_eta_k) + 1;  // camb/bessels.f90:44, :58
}

// InitTransfer, camb/cmbmain.f90:508-632 (flat, WantCls): the transfer wavenumber grid; kept are the values above the top
// of the source grid, which CalcScalarSources does not visit (:597-623)
void Model::transfer_grid() {
  k_transfer.clear();
  if (!P.WantTransfer) return;
  const double kmax = P.transfer_kmax;
  std::vector<double> qt;  // 0-based here
  if (P.transfer_k_per_logint == 0) {
    const double boost = P.AccuracyBoost;
    const double q_switch_lowk1 = SP(0.7) / taurst, dlog_lowk1 = 2 * boost;
    const double q_switch_lowk = 8 / taurst;
    double dlog_lowk = 8 * boost;
    if (P.HighAccuracyDefault) dlog_lowk = dlog_lowk * 2.5;
    const double q_switch_osc = std::fmin(kmax, 30 / taurst);
    double d_osc = 200 * boost;
    if (P.HighAccuracyDefault) d_osc = d_osc * SP(1.8);
    const double q_switch_highk = std::fmin(kmax, (P.HighAccuracyDefault ? 90 : 60) / taurst);
    const double dlog_osc = 17 * boost, dlog_highk = 3 * boost, amin = 5e-5;
    int nq = (int)(std::log(q_switch_lowk1 / amin) * dlog_lowk1) + 1;
    for (int i = 1; i <= nq; i++) qt.push_back(amin * std::exp((i - 1) / dlog_lowk1));
    double last = qt.back();
    nq = (int)(std::log(q_switch_lowk / last) * dlog_lowk) + 1;
    for (int i = 1; i <= nq; i++) qt.push_back(last * std::exp(i / dlog_lowk));
    last = qt.back();
    nq = (int)((q_switch_osc - last) * d_osc) + 1;
    for (int i = 1; i <= nq; i++) qt.push_back(last + i / d_osc);
    last = qt.back();
    if (kmax > last) {
      nq = (int)(std::log(q_switch_highk / last) * dlog_osc) + 1;
      for (int i = 1; i <= nq; i++) qt.push_back(last * std::exp(i / dlog_osc));
      last = qt.back();
    }
    if (kmax > last) {
      nq = (int)(std::log(kmax / last) * dlog_highk) + 1;
      for (int i = 1; i <= nq; i++) qt.push_back(last * std::exp(i / dlog_highk));
    }
  } else {
    const int n = (int)((std::log(kmax) - std::log(qmin)) * P.transfer_k_per_logint) + 1;
    for (int i = 1; i <= n; i++) qt.push_back(qmin * (double)std::exp((float)i / (float)P.transfer_k_per_logint));  // default-real exp
  }
  for (size_t i = 0; i < qt.size(); i++)
    if (qt[i] > Evolve_q.Highest + 1e-4) {
      k_transfer.assign(qt.begin() + i, qt.end());
      break;
    }
}

// camb/modules.f90:851-1008 (Log_lvalues = F)
void Model::l_samples() {
  const int max_l = P.Max_l, lmin = 2;
  std::vector<int> v(8002, 0);
  int lind = 0, lvar, step, top, bot;
  const double Ascale = scale / P.lSampleBoost;
  auto done = [&]() { ls.assign(v.begin() + 1, v.begin() + 1 + lind); };
  if (P.lSampleBoost >= 50) {
    for (lvar = lmin; lvar <= max_l; lvar++) v[++lind] = lvar;
    return done();
  }
  for (lvar = lmin; lvar <= 10; lvar++) v[++lind] = lvar;
  if (P.AccurateReionization) {
    for (lvar = 11; lvar <= 37; lvar += (P.lSampleBoost > 1 ? 1 : 2)) v[++lind] = lvar;
    step = std::max((int)std::lround(5 * Ascale), 2);
    bot = 40;
    top = bot + step * 10;
  } else {
    if (P.lSampleBoost > 1) {
      for (lvar = 11; lvar <= 15; lvar++) v[++lind] = lvar;
    } else {
      v[++lind] = 12;
      v[++lind] = 15;
    }
    step = std::max((int)std::lround(10 * Ascale), 3);
    bot = 15 + std::max(step / 2, 2);
    top = bot + step * 7;
  }
  for (lvar = bot; lvar <= top; lvar += step) v[++lind] = lvar;
  step = std::max((int)std::lround(20 * Ascale), 4);
  bot = v[lind] + step;
  top = bot + step * 2;
  for (lvar = bot; lvar <= top; lvar += step) v[++lind] = lvar;
  auto clip = [&]() {
    for (lvar = lind; lvar >= 1; lvar--)
      if (v[lvar] <= max_l) break;
    lind = lvar;
    if (v[lind] < max_l) v[++lind] = max_l;
  };
  if (v[lind] >= max_l) {
    clip();
    return done();
  }
  step = std::max((int)std::lround(25 * Ascale), 4);
  bot = v[lind] + step;
  top = bot + step;
  for (lvar = bot; lvar <= top; lvar += step) v[++lind] = lvar;
  if (v[lind] >= max_l) {
    clip();
    return done();
  }
  step = (P.HighAccuracyDefault && !P.use_spline_template) ? std::max((int)std::lround(42 * Ascale), 7) : std::max((int)std::lround(50 * Ascale), 7);
  bot = v[lind] + step;
  top = std::min(5000, max_l);
  for (lvar = bot; lvar <= top; lvar += step) v[++lind] = lvar;
  if (max_l > 5000) {
    step = std::max((int)std::lround(400 * Ascale), 50);
    lvar = v[lind];
    for (;;) {
      lvar = lvar + step;
      if (lvar > max_l) break;
      v[++lind] = lvar;
      step = (int)std::lround(step * SP(1.5));
    }
  }
  if (v[lind] != max_l) v[++lind] = max_l;
  done();
}

int Model::run() {
  status = 0;
  err.clear();
  static const bool prof = getenv("GALCAMB_PRESTAGE_PROFILE") != nullptr;  // per-phase wall clock of every point to stderr
  const auto t0 = std::chrono::steady_clock::now();
  if (int rc = params_set()) return rc;
  const auto t1 = std::chrono::steady_clock::now();
  l_samples();
  if (int rc = init_vars()) return rc;
  const auto t2 = std::chrono::steady_clock::now();
  k_grids();
  transfer_grid();
  if (prof) {
    const auto t3 = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "[prestage] params_set %.2f ms (background %.2f) init_vars %.2f ms (recfast %.2f) grids %.2f ms\n", ms(t0, t1), prof_bg_ms,
            ms(t1, t2), prof_rec_ms, ms(t2, t3));
  }
  return 0;
}

// the batch held by the library between galcamb_prestage_ and the next call
static std::vector<Model> g_batch;
static std::string g_err;
// second buffer + worker of the pipelined form (galcamb_prestage_async_ / _wait_)
static std::vector<Model> g_next;
static std::string g_next_err;
static std::vector<galcamb_params> g_async_params;
static std::thread g_async;
static int g_async_rc = 0;
static int g_host_threads = 0;
static int g_bg_device = -1;  // galcamb_set_prestage_device_: device that integrates the Galileon backgrounds of a batch (-1: host)
static std::mutex g_bg_mutex;
// cs2 table address of every point of the current batch -> its index (galcamb_model views carry no back pointer)
static std::unordered_map<const double*, int> g_by_table;
static void index_current_batch() {
  g_by_table.clear();
  for (size_t i = 0; i < g_batch.size(); i++)
    if (g_batch[i].status == 0 && !g_batch[i].packed.empty()) g_by_table[g_batch[i].cs2.data()] = (int)i;
}

static void fill_view(const Model& M, galcamb_model* m) {
  std::memset(m, 0, sizeof(*m));
  const galcamb_params& P = M.P;
  m->grhom = M.grhom; m->grhog = M.grhog; m->grhor = M.grhor; m->grhob = M.grhob; m->grhoc = M.grhoc; m->grhov = M.grhov;
  m->grhornomass = M.grhornomass; m->grhok = M.grhok; m->adotrad = M.adotrad; m->tau0 = M.tau0; m->taurst = M.taurst;
  m->taurend = M.taurend; m->tau_maxvis = M.tau_maxvis; m->tight_tau = M.tight_tau; m->matter_verydom_tau = M.matter_verydom_tau;
  m->epsw = M.epsw; m->max_etak_scalar = M.max_etak_scalar; m->AccuracyBoost = P.AccuracyBoost; m->lAccuracyBoost = P.lAccuracyBoost;
  m->use_galileon = P.use_galileon; m->DoLateRadTruncation = P.DoLateRadTruncation; m->HighAccuracyDefault = P.HighAccuracyDefault;
  m->AccuratePolarization = P.AccuratePolarization; m->AccurateReionization = P.AccurateReionization;
  m->WantLateTime = P.DoLensing; m->MassiveNuMethod = P.MassiveNuMethod;
  m->Nu_mass_eigenstates = M.Nu_mass_eigenstates; m->Num_Nu_massive = M.Num_Nu_massive; m->nqmax = M.nqmax;
  for (int i = 0; i < MAXNU; i++) {
    m->grhormass[i] = M.grhormass[i + 1];
    m->nu_masses[i] = M.nu_masses[i + 1];
    m->nu_tau_nonrelativistic[i] = M.nu_tau_nonrelativistic[i + 1];
    m->nu_tau_massive[i] = M.nu_tau_massive[i + 1];
    for (int q = 0; q < GALCAMB_MAX_NQ; q++) m->nu_tau_notmassless[q][i] = M.nu_tau_notmassless[q + 1][i + 1];
  }
  for (int q = 0; q < GALCAMB_MAX_NQ; q++) {
    m->nu_q[q] = M.nu_q[q + 1];
    m->nu_int_kernel[q] = M.nu_int_kernel[q + 1];
  }
  if (M.Nu_mass_eigenstates > 0 && !g_nu.r1.empty()) {
    m->dlnam = g_nu.dlnam;
    m->nu_r1 = g_nu.r1.data() + 1; m->nu_p1 = g_nu.p1.data() + 1; m->nu_dr1 = g_nu.dr1.data() + 1; m->nu_dp1 = g_nu.dp1.data() + 1;
    m->nu_ddr1 = g_nu.ddr1.data() + 1; m->nu_ddp1 = g_nu.ddp1.data() + 1;
  }
  m->nthermo = NTHERMO; m->tauminn = M.tauminn; m->dlntau = M.dlntau;
  m->cs2 = M.cs2.data(); m->dcs2 = M.dcs2.data(); m->dotmu = M.dotmu.data(); m->ddotmu = M.ddotmu.data(); m->dddotmu = M.dddotmu.data();
  m->nt = M.TimeSteps.npoints;
  m->tau = M.TimeSteps.points.data() + 1; m->dtau = M.TimeSteps.dpoints.data() + 1;
  m->vis = M.vis.data() + 1; m->dvis = M.dvis.data() + 1; m->ddvis = M.ddvis.data() + 1; m->expmmu = M.expmmu.data() + 1;
  m->opac = M.opac.data() + 1; m->dopac = M.dopac.data() + 1;
  m->n_tau_regions = M.TimeSteps.count;
  for (int i = 0; i < M.TimeSteps.count && i < GALCAMB_MAX_REGIONS; i++) {
    const Region& A = M.TimeSteps.R[i + 1];
    m->tau_regions[i].Low = A.Low; m->tau_regions[i].High = A.High; m->tau_regions[i].delta = A.delta;
    m->tau_regions[i].start_index = A.start_index; m->tau_regions[i].IsLog = A.IsLog;
  }
  if (P.use_galileon) {
    m->c2 = M.gal_coef[0]; m->c3 = M.gal_coef[1]; m->c4 = M.gal_coef[2]; m->c5 = M.gal_coef[3]; m->cG = M.gal_coef[4];
    m->h0 = M.gal_h0; m->orad = M.gal_orad;
    m->ngal = (int)M.gal_a.size() - 1;  // without the extrapolated node (the device library appends its own)
    m->gal_a = M.gal_a.data(); m->gal_h = M.gal_h.data(); m->gal_x = M.gal_x.data();
  }
  m->nk = M.Evolve_q.npoints; m->k = M.Evolve_q.points.data() + 1;
  m->nq = M.q_int.npoints; m->q = M.q_int.points.data() + 1; m->dq = M.q_int.dpoints.data() + 1;
  m->ScalarPowerAmp = P.ScalarPowerAmp; m->an = P.an; m->n_run = P.n_run; m->n_runrun = P.n_runrun; m->k_0_scalar = P.k_0_scalar;
  m->lmax = P.Max_l;
  m->WantTransfer = P.WantTransfer ? 1 : 0;
  m->transfer_num_redshifts = (int)M.tautf.size();
  for (size_t i = 0; i < M.tautf.size() && i < GALCAMB_MAX_TRANSFER_Z; i++) m->tautf[i] = M.tautf[i];
  m->nk_transfer = (int)M.k_transfer.size();
  m->k_transfer = M.k_transfer.empty() ? nullptr : M.k_transfer.data();
  m->H0 = P.H0;
  m->ALens = P.ALens > 0 ? P.ALens : 1.0;
  m->ALens_Fiducial = P.ALens_Fiducial;
}

}  // namespace pre

extern "C" {

void galcamb_default_params_(galcamb_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->H0 = 70; p->TCMB = 2.7255; p->YHe = 0.24; p->Num_Nu_massless = 2.046; p->Num_Nu_massive = 1; p->Nu_mass_eigenstates = 1;
  p->share_delta_neff = 1; p->Nu_mass_fractions[0] = 1; p->Nu_mass_numbers[0] = 1;
  p->ScalarPowerAmp = 2.1e-9; p->an = 0.96; p->k_0_scalar = 0.05;
  p->Reionization = 1; p->use_optical_depth = 1; p->optical_depth = 0.09; p->reion_redshift = 11; p->delta_redshift = 1.5;
  p->reion_fraction = -1; p->helium_redshift = 3.5; p->helium_delta_redshift = 0.5; p->helium_redshiftstart = 5.0;
  p->RECFAST_fudge = 1.14 - (1.14 - (1.105 + 0.02)); p->RECFAST_fudge_He = 0.86; p->RECFAST_Heswitch = 6; p->RECFAST_Hswitch = 1;
  p->Max_l = 2500; p->Max_eta_k = 5000; p->AccuratePolarization = 1; p->AccurateReionization = 1; p->MassiveNuMethod = 1;
  p->DoLateRadTruncation = 1; p->HighAccuracyDefault = 1; p->AccuracyBoost = 1; p->lSampleBoost = 1; p->lAccuracyBoost = 1;
  p->use_spline_template = 1;
  p->WantTransfer = 0; p->transfer_kmax = 0.9; p->transfer_num_redshifts = 1; p->transfer_redshifts[0] = 0;
  p->ALens = 1; p->ALens_Fiducial = 0;
}

int galcamb_sizeof_params_(void) { return (int)sizeof(galcamb_params); }

// the work of galcamb_prestage_ on a given buffer
static int prestage_into(std::vector<pre::Model>& batch, std::string& err, const galcamb_params* params, int n, int nthreads) {
  using namespace pre;
  batch.clear();
  batch.resize(n);
  for (int i = 0; i < n; i++) batch[i].P = params[i];
  int nt = nthreads > 0 ? nthreads : g_host_threads > 0 ? g_host_threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min(nt, n));
  // Galileon backgrounds of the whole batch in one device launch (csrc/galileon_host.cu), next to whatever the device is
  // running for the previous batch; the host threads below then start from the finished tables.
  if (g_bg_device >= 0 && n > 1) {
    std::vector<GcBackgroundIn> in;
    std::vector<int> who;
    for (int i = 0; i < n; i++) {
      Model& M = batch[i];
      if (!M.P.use_galileon) continue;
      M.bg_mode = 1;
      const int rc = M.params_set();  // up to the inputs of the background
      M.bg_mode = 0;
      if (rc != 0) continue;          // the full run below reports the error
      GcBackgroundIn b{};
      b.omegar = M.gal_orad;
      b.omegam = M.P.omegab + M.P.omegac;
      b.h0 = M.gal_h0;
      b.c2 = M.P.c2; b.c3in = M.P.c3; b.c4in = M.P.c4; b.cG = M.P.cG;
      b.neig = M.Nu_mass_eigenstates;
      if (b.neig > GC_MAX_NU) continue;
      for (int j = 0; j < b.neig; j++) {
        b.grhormass[j] = M.grhormass[j + 1];
        b.nu_masses[j] = M.nu_masses[j + 1];
      }
      in.push_back(b);
      who.push_back(i);
    }
    if (!in.empty()) {
      std::lock_guard<std::mutex> lock(g_bg_mutex);
      std::vector<double> a, coef(in.size() * 5);
      std::vector<int> st(in.size());
      const double *h = nullptr, *x = nullptr;
      const int rc = gc_galileon_background_device(g_bg_device, in.data(), (int)in.size(), g_nu.rows.empty() ? nullptr : g_nu.rows.data(),
                                                   g_nu.rows.size() / 6, g_nu.dlnam, a, &h, &x, coef.data(), st.data());
      if (rc == 0) {
        const size_t per = a.size();
        for (size_t m = 0; m < who.size(); m++) {
          Model& M = batch[who[m]];
          M.bg_mode = 2;
          M.bg_status = st[m];
          if (st[m] == 0) {
            M.gal_a = a;
            M.gal_h.assign(h + m * per, h + (m + 1) * per);
            M.gal_x.assign(x + m * per, x + (m + 1) * per);
            for (int c = 0; c < 5; c++) M.gal_coef[c] = coef[m * 5 + c];
          }
        }
      }  // on a CUDA failure the host threads integrate as usual
    }
  }
  // the shared neutrino tables are built by whichever thread gets there first (std::call_once)
  std::atomic<int> next{0};
  auto work = [&]() {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n) break;
      batch[i].run();
      if (batch[i].status == 0) {  // device layout of the tables, ready for galcamb_set_model_ / galcamb_batch_upload_
        galcamb_model v;
        fill_view(batch[i], &v);
        const int nk_tr = v.WantTransfer ? v.nk_transfer : 0;
        gcpack::sizes(&v, nk_tr, batch[i].pinfo);
        batch[i].packed.resize(batch[i].pinfo.total);
        gcpack::tables(&v, nk_tr, batch[i].packed.data(), batch[i].pinfo);
      }
      batch[i].zrec = std::vector<double>();  // the recombination history is only needed up to inithermo
      batch[i].xrec = std::vector<double>();
      batch[i].dxrec = std::vector<double>();
      batch[i].gal_ch = std::vector<double>();
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; t++) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  int rc = 0;
  for (int i = 0; i < n; i++)
    if (batch[i].status && !rc) {
      rc = batch[i].status;
      err = "point " + std::to_string(i) + ": " + batch[i].err;
    }
  return rc;
}

int galcamb_prestage_(const galcamb_params* params, const int* n, const int* nthreads) {
  if (*n < 1) {
    pre::g_err = "n < 1";
    return 5;
  }
  const int rc = prestage_into(pre::g_batch, pre::g_err, params, *n, *nthreads);
  pre::index_current_batch();
  return rc;
}

// Pipelined form: galcamb_prestage_async_ starts the same work on a second buffer and returns at once -- the device
// stages of the CURRENT batch (galcamb_prestaged_to_cls_) run meanwhile; galcamb_prestage_wait_ joins it and makes the
// new batch the current one.  The parameter array is copied, the caller may reuse it.
int galcamb_prestage_async_(const galcamb_params* params, const int* n, const int* nthreads) {
  using namespace pre;
  if (*n < 1) return 5;
  if (g_async.joinable()) return 5;  // one batch in flight at a time
  g_async_params.assign(params, params + *n);
  const int nth = *nthreads;
  g_async = std::thread([nth]() {
    g_async_rc = prestage_into(g_next, g_next_err, g_async_params.data(), (int)g_async_params.size(), nth);
  });
  return 0;
}

int galcamb_prestage_wait_(void) {
  using namespace pre;
  if (!g_async.joinable()) return 5;
  g_async.join();
  g_batch.swap(g_next);
  g_next.clear();
  index_current_batch();
  g_err = g_next_err;
  return g_async_rc;
}

// the packed tables of a model view handed out by galcamb_prestage_model_ (csrc/capi.cu pack_model), or nullptr
const double* gc_prestage_prepacked(const galcamb_model* m, size_t* n, gcpack::Info* info) {
  const auto it = pre::g_by_table.find(m->cs2);
  if (it == pre::g_by_table.end()) return nullptr;
  const pre::Model& M = pre::g_batch[it->second];
  *n = M.packed.size();
  *info = M.pinfo;
  return M.packed.data();
}

int galcamb_prestage_count_(void) { return (int)pre::g_batch.size(); }

// host threads the library may use for the pre-stage and for packing model tables (0 = all cores): with one process per
// GPU on a shared host each rank passes cores / ranks
void galcamb_set_host_threads_(const int* n) { pre::g_host_threads = *n; }
// device >= 0: the Galileon backgrounds of a batch are integrated on that device (one lane per cosmology, one launch per
// batch, next to the per-k kernel of the previous batch); -1 (default): on the host threads
int galcamb_set_prestage_device_(const int* device) {
  pre::g_bg_device = *device;
  return 0;
}
int galcamb_get_host_threads_(void) { return pre::g_host_threads; }

int galcamb_prestage_status_(const int* i, int* status) {
  if (*i < 0 || *i >= (int)pre::g_batch.size()) return 5;
  *status = pre::g_batch[*i].status;
  return 0;
}

const char* galcamb_prestage_error_(void) { return pre::g_err.c_str(); }

int galcamb_prestage_model_(const int* i, galcamb_model* m) {
  if (*i < 0 || *i >= (int)pre::g_batch.size()) return 5;
  if (pre::g_batch[*i].status) return pre::g_batch[*i].status;
  pre::fill_view(pre::g_batch[*i], m);
  return 0;
}

int galcamb_prestage_lsamples_(const int* i, int* nl, int* ls, int* kmaxfile) {
  if (*i < 0 || *i >= (int)pre::g_batch.size()) return 5;
  const pre::Model& M = pre::g_batch[*i];
  *nl = (int)M.ls.size();
  *kmaxfile = M.kmaxfile;
  if (ls)
    for (size_t j = 0; j < M.ls.size(); j++) ls[j] = M.ls[j];
  return 0;
}

void galcamb_prestage_free_(void) {
  pre::g_by_table.clear();
  pre::g_batch.clear();
  pre::g_batch.shrink_to_fit();
}

}  // extern "C"
