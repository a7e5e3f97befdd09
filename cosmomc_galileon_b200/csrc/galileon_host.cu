// Host-side replacement of camb/galileon.cc -- the reference's own C ABI (camb/interface.f90:7-92): background
// integration of the Galileon field (arrays_), table accessors (GetX_, GetH_), and the closed-form background /
// perturbation terms the Fortran equations call per RHS evaluation.  First "next" row of SURVEY.md section 8f (the
// Galileon part of the host pre-stage).  Differences from the reference, all at the boundary:
//   * no GSL: the Fehlberg 4(5) stepper with gsl_odeiv_control_y_new(1e-18, 1e-16) semantics and the natural cubic
//     spline (= gsl_interp_cspline) are written out here (camb/galileon.cc:344-363, :666-682);
//   * no call-backs into Fortran (MassiveNu::Nu_rho/Nu_background/Nu_dp, camb/galileon.cc:96-99): the host passes
//     its massive-neutrino tables once with galileon_set_nu_tables_();
//   * errors come back as the integer status of arrays_ (camb/equations_galileon.f90:97-100); an out-of-range `a` in
//     GetX_/GetH_ returns NaN instead of exit() (camb/galileon.cc:698-701).
// The closed forms are the same functions the device code evaluates (evolve_device.cuh, __host__ __device__).
#include <cmath>
#include <cstdio>
#include <vector>

#include "evolve_device.cuh"
#include "../../include/galileon_abi.h"

namespace {

using gc::Nu_background;

// one background problem: the constants of the ODE + the memory of the two "no sudden jump" stability tests
// what the background problem reads of a model (the field names of DevModel, so that the neutrino look-ups of
// evolve_device.cuh take either): small enough to stay in registers when the integrator runs as a device lane
struct BgModel {
  double c2, c3, c4, c5, cG, h0, orad;
  double dlnam, inv_dlnam;
  double grhormass[GC_MAX_NU], nu_masses[GC_MAX_NU];
  const double* nu_tab;  // [n][6] rows {r1, dr1, p1, dp1, ddr1, ddp1}; host or device memory
  int n_eig;
};

struct Background {
  BgModel M{};
  double noghost_previous = 0, cs2_previous = 0;
  GC_HOSTDEV void nu_pressures(double a, double* rhonu, double* pnu) const {
    for (int i = 0; i < M.n_eig; i++) gc::Nu_background(M, a * M.nu_masses[i], rhonu[i], pnu[i]);
  }
  // the pressure alone (same branches and arithmetic as Nu_background, camb/modules.f90:1721-1754, without the density's
  // exponential): all the right-hand side and the stability conditions read
  GC_HOSTDEV double nu_pressure(double am) const {
    if (am <= GC_AM_MINP) return (2 - (1.0 + GC_NU_CONST2 * am * am)) / 3.0;
    if (am >= GC_AM_MAXP) return 900.0 / 120.0 / GC_NU_CONST * (GC_ZETA5 - 63.0 / 4 * GC_ZETA7 / (am * am)) / am;
    double d = log(am * 100.0) * M.inv_dlnam + 1.0;
    const int i = (int)d;
    d = d - i;
    const double* r0 = M.nu_tab + (size_t)(i - 1) * 6;
    const double* r1 = r0 + 6;
    return exp(gc::herm(r0[2], r0[3], r1[2], r1[3], d));
  }
  GC_HOSTDEV void rhs(double a, const double y[3], double f[3]) const;
  GC_HOSTDEV int stability(double a, const double y[3]);
  GC_HOSTDEV void evolve_apply(double& t, double t1, double& h, double y[3]) const;
  // the node loop of calcHubbleGalileon on given node positions; hub / xg = NB + 1 values, element i at [i * stride]
  GC_HOSTDEV int run_nodes(double om, bool reject_negative_density, const double* grid, double* hub, double* xg, long stride);
  int run(double om, bool reject_negative_density, std::vector<double>& grid, std::vector<double>& hub, std::vector<double>& xg);
};

struct HostGal : Background {
  double om = 0;
  std::vector<double> nu_tab;            // [n][6] = {r1, dr1, p1, dp1, ddr1, ddp1}
  std::vector<double> a, h, x, ch, cx;   // nb + 2 nodes, ascending a, + natural-spline coefficients
  bool ready = false;
};
HostGal G;

constexpr int NB = 29999;  // camb/galileon.cc:72 (30 000 nodes)

void nu_pressures(double a, double* rhonu, double* pnu) { G.nu_pressures(a, rhonu, pnu); }

// d(hbar, dpi/da, pi)/da, camb/galileon.cc:254-302
GC_HOSTDEV void Background::rhs(double a, const double y[3], double f[3]) const {
  const double c2 = M.c2, c3 = M.c3, c4 = M.c4, c5 = M.c5, cG = M.cG, h0 = M.h0, orad = M.orad;
  const double prod = a * y[0] * y[1];
  const double prod2 = prod * prod, prod3 = prod * prod2, prod4 = prod * prod3, prod5 = prod * prod4;
  const double h = y[0], h2 = h * h, h3 = h * h2, h4 = h2 * h2, a2 = a * a;
  const double alpha = c2 / 6.0 * prod - 3 * c3 * h * prod2 + 15 * c4 * h2 * prod3 - 17.5 * c5 * h3 * prod4 - 3.0 * cG * h2 * prod;
  const double gamma = c2 / 3.0 * h * prod - c3 * h2 * prod2 + 2.5 * c5 * h4 * prod4 - 2.0 * cG * h3 * prod;
  const double beta = c2 / 6.0 * h2 - 2 * c3 * h3 * prod + 9 * c4 * h4 * prod2 - 10 * c5 * h4 * h * prod3 - cG * h4;
  const double sigma = 2.0 * h + 2.0 * c3 * prod3 - 15.0 * c4 * h * prod4 + 21.0 * c5 * h2 * prod5 + 6.0 * cG * h * prod2;
  double lambda = 3.0 * h2 + orad / (a2 * a2) + c2 / 2.0 * prod2 - 2.0 * c3 * h * prod3 + 7.5 * c4 * h2 * prod4 -
                  9.0 * c5 * h3 * prod5 - cG * h2 * prod2;
  for (int i = 0; i < M.n_eig; i++) lambda += M.grhormass[i] * nu_pressure(a * M.nu_masses[i]) / (a2 * a2 * h0 * h0);
  const double omega = 2 * c3 * h2 * prod2 - 12 * c4 * h3 * prod3 + 15 * c5 * h4 * prod4 + 4.0 * cG * h3 * prod;
  const double denom = sigma * beta - alpha * omega;
  f[0] = (omega * gamma - lambda * beta) / (a * denom);
  f[1] = (alpha * lambda - sigma * gamma) / (a2 * denom) - 2.0 * y[1] / a;
  f[2] = y[1];
}

// ghost / Laplace stability of scalar and tensor perturbations at one node, camb/galileon.cc:130-245
// returns 0 or the number of the failed condition (1 no-ghost, 2 c_s^2, 3 tensor no-ghost, 4 c_T^2)
GC_HOSTDEV int Background::stability(double a, const double y[3]) {
  const double c2 = M.c2, c3 = M.c3, c4 = M.c4, c5 = M.c5, cG = M.cG, h0 = M.h0, orad = M.orad;
  const double xgal = a * y[1];
  const double prod = xgal * y[0];
  const double prod2 = prod * prod, prod3 = prod * prod2, prod4 = prod * prod3, prod5 = prod * prod4;
  const double h = y[0], h2 = h * h, h3 = h * h2, h4 = h2 * h2;
  const double alpha = c2 / 6.0 * prod - 3 * c3 * h * prod2 + 15 * c4 * h2 * prod3 - 17.5 * c5 * h3 * prod4 - 3.0 * cG * h2 * prod;
  const double gamma = c2 / 3.0 * h * prod - c3 * h2 * prod2 + 2.5 * c5 * h4 * prod4 - 2.0 * cG * h3 * prod;
  const double beta = -2 * c3 * h3 * prod + c2 / 6.0 * h2 + 9 * c4 * h4 * prod2 - 10 * c5 * h4 * h * prod3 - cG * h4;
  const double sigma = 2.0 * h + 2.0 * c3 * prod3 - 15.0 * c4 * h * prod4 + 21.0 * c5 * h2 * prod5 + 6.0 * cG * h * prod2;
  double lambda = 3.0 * h2 + orad / (a * a * a * a) - 2.0 * c3 * h * prod3 + c2 / 2.0 * prod2 + 7.5 * c4 * h2 * prod4 -
                  9.0 * c5 * h3 * prod5 - cG * h2 * prod2;
  const double omega = 2 * c3 * h2 * prod2 - 12 * c4 * h3 * prod3 + 15 * c5 * h4 * prod4 + 4.0 * cG * h3 * prod;
  const double denom = sigma * beta - alpha * omega;
  for (int i = 0; i < M.n_eig; i++) lambda += M.grhormass[i] * nu_pressure(a * M.nu_masses[i]) / (a * a * a * a * h0 * h0);
  const double dhdlna = (omega * gamma - lambda * beta) / denom;
  const double dxdlna = (alpha * lambda - sigma * gamma) / denom - xgal;
  const double k1 = -6 * c4 * h * prod2 * (dhdlna * xgal + h * dxdlna + prod / 3.0) +
                    c5 * h2 * prod3 * (12 * h * dxdlna + 15 * dhdlna * xgal + 3 * prod) +
                    2.0 * cG * (dhdlna * prod + h2 * dxdlna + h * prod);
  const double k2 = -0.5 * c2 + 6 * c3 * h * prod - 27 * c4 * h2 * prod2 + 30 * c5 * h3 * prod3 + 3.0 * cG * h2;
  const double k3 = -1.0 - 0.5 * c4 * prod4 - 3 * c5 * h * prod4 * (h * dxdlna + dhdlna * xgal) + cG * prod2;
  const double k4 = -2.0 + 3.0 * c4 * prod4 - 6 * c5 * h * prod5 - 2.0 * cG * prod2;
  const double k5 = 2 * c3 * prod2 - 12 * c4 * h * prod3 + 15.0 * c5 * h2 * prod4 + 4.0 * cG * h * prod;
  const double k6 = 0.5 * c2 - 2.0 * c3 * (h2 * dxdlna + dhdlna * prod + 2 * h * prod) +
                    c4 * (12 * h3 * prod * dxdlna + 18 * h * prod2 * dhdlna + 13 * h2 * prod2) -
                    c5 * (18 * h2 * h2 * prod2 * dxdlna + 30 * h2 * prod3 * dhdlna + 12 * h3 * prod3) -
                    cG * (2.0 * h * dhdlna + 3.0 * h2);
  const double noghost = k2 + 1.5 * k5 * k5 / k4;
  const double cs2 = (4 * k1 * k4 * k5 - 2 * k3 * k5 * k5 - 2 * k4 * k4 * k6) / (k4 * (2 * k4 * k2 + 3 * k5 * k5));
  if (a == 1) {
    noghost_previous = noghost;
    cs2_previous = cs2;
  }
  const double noghostt = 0.5 - 0.75 * c4 * prod4 + 1.5 * c5 * prod5 * h + 0.5 * cG * prod2;
  const double ct2 = (0.5 + 0.25 * c4 * prod4 + 1.5 * c5 * prod4 * h * (h * dxdlna + dhdlna * xgal) - 0.5 * cG * prod2) / noghostt;
  if (noghost > -1.0 && noghost > -1e-8) return 1;
  if (cs2 < 0) return 2;
  if (noghostt < 0) return 3;
  if (ct2 < 0) return 4;
  if (noghost > -1.0 && noghost < -1e-8 && fabs((noghost - noghost_previous) / noghost) > 0.25) return 1;
  noghost_previous = noghost;
  if (fabs((cs2 - cs2_previous) / cs2_previous) > 1) return 2;
  cs2_previous = cs2;
  return 0;
}

// Fehlberg 4(5) embedded pair (published tableau); the derivative at the start of the step is handed in.  (GSL's rkf45 also
// evaluates the derivative at the end of every trial step for its caller; the y-error controller used here never reads
// it, so it is not computed: one right-hand side in seven less, bit-identical results.)
struct Rkf45 {
  const Background& B;
  GC_HOSTDEV void step(double t, double h, double y[3], double yerr[3], const double k1[3]) const {
    const double c[] = {1.0 / 4.0, 3.0 / 8.0, 12.0 / 13.0, 1.0, 1.0 / 2.0};
    // rows of the tableau (row s - 1 = the coefficients of stage s + 1, padded with zeros that are never read)
    const double rows[5][5] = {{1.0 / 4.0, 0, 0, 0, 0},
                               {3.0 / 32.0, 9.0 / 32.0, 0, 0, 0},
                               {1932.0 / 2197.0, -7200.0 / 2197.0, 7296.0 / 2197.0, 0, 0},
                               {8341.0 / 4104.0, -32832.0 / 4104.0, 29440.0 / 4104.0, -845.0 / 4104.0, 0},
                               {-6080.0 / 20520.0, 41040.0 / 20520.0, -28352.0 / 20520.0, 9295.0 / 20520.0, -5643.0 / 20520.0}};
    const double b[] = {902880.0 / 7618050.0, 0.0, 3953664.0 / 7618050.0, 3855735.0 / 7618050.0, -1371249.0 / 7618050.0,
                        277020.0 / 7618050.0};
    const double e[] = {1.0 / 360.0, 0.0, -128.0 / 4275.0, -2197.0 / 75240.0, 1.0 / 50.0, 2.0 / 55.0};
    double k[6][3], yt[3];
    for (int i = 0; i < 3; i++) k[0][i] = k1[i];
#pragma unroll
    for (int s = 1; s < 6; s++) {
      for (int i = 0; i < 3; i++) {
        double acc = rows[s - 1][0] * k[0][i];
#pragma unroll
        for (int j = 1; j < s; j++) acc += rows[s - 1][j] * k[j][i];
        yt[i] = y[i] + (s == 1 ? rows[0][0] * h * k[0][i] : h * acc);
      }
      B.rhs(t + c[s - 1] * h, yt, k[s]);
    }
    for (int i = 0; i < 3; i++) {
      const double d = b[0] * k[0][i] + b[2] * k[2][i] + b[3] * k[3][i] + b[4] * k[4][i] + b[5] * k[5][i];
      y[i] += h * d;
    }
    for (int i = 0; i < 3; i++)
      yerr[i] = h * (e[0] * k[0][i] + e[2] * k[2][i] + e[3] * k[3][i] + e[4] * k[4][i] + e[5] * k[5][i]);
  }
};

// one accepted step towards t1 with the standard y-error controller (eps_abs 1e-18, eps_rel 1e-16, a_y = 1, a_dydt = 0)
GC_HOSTDEV void Background::evolve_apply(double& t, double t1, double& h, double y[3]) const {
  const double eps_abs = 1e-18, eps_rel = 1e-16, S = 0.9, ord = 5.0;
  const Rkf45 stepper{*this};
  const double t0 = t, dt = t1 - t0;
  double h0 = h;
  const double y0[3] = {y[0], y[1], y[2]};
  double yerr[3], dydt_in[3];
  rhs(t0, y, dydt_in);
  for (;;) {
    const bool final_step = (dt >= 0.0 && h0 > dt) || (dt < 0.0 && h0 < dt);
    if (final_step) h0 = dt;
    stepper.step(t0, h0, y, yerr, dydt_in);
    t = final_step ? t1 : t0 + h0;
    const double h_old = h0;
    double rmax = 2.2250738585072014e-308;
    for (int i = 0; i < 3; i++) rmax = fmax(rmax, fabs(yerr[i]) / fabs(eps_rel * fabs(y[i]) + eps_abs));
    if (rmax > 1.1) {  // shrink and retry from (t0, y0) unless the step cannot shrink any more
      h0 = h_old * fmax(0.2, S / pow(rmax, 1.0 / ord));
      if (fabs(h0) < fabs(h_old) && t + h0 != t) {
        for (int i = 0; i < 3; i++) y[i] = y0[i];
        continue;
      }
      h0 = h_old;
    } else if (rmax < 0.5) {
      h0 = h_old * fmin(5.0, fmax(1.0, S / pow(rmax, 1.0 / (ord + 1.0))));
    }
    break;
  }
  h = h0;
}

// natural cubic spline, c[i] = y''(x_i)/2 (gsl_interp_cspline's representation)
void spline_coefficients(const std::vector<double>& xa, const std::vector<double>& ya, std::vector<double>& c) {
  const int n = (int)xa.size(), sys = n - 2;
  c.assign(n, 0.0);
  std::vector<double> rhs_(sys), diag(sys), off(sys), al(sys), ga(sys), z(sys);
  for (int i = 0; i < sys; i++) {
    const double h_i = xa[i + 1] - xa[i], h_ip1 = xa[i + 2] - xa[i + 1];
    const double yd_i = ya[i + 1] - ya[i], yd_ip1 = ya[i + 2] - ya[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    off[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    rhs_[i] = 3.0 * (yd_ip1 * g_ip1 - yd_i * g_i);
  }
  al[0] = diag[0];
  ga[0] = off[0] / al[0];
  for (int i = 1; i < sys - 1; i++) {
    al[i] = diag[i] - off[i - 1] * ga[i - 1];
    ga[i] = off[i] / al[i];
  }
  if (sys > 1) al[sys - 1] = diag[sys - 1] - off[sys - 2] * ga[sys - 2];
  z[0] = rhs_[0];
  for (int i = 1; i < sys; i++) z[i] = rhs_[i] - ga[i - 1] * z[i - 1];
  for (int i = 0; i < sys; i++) z[i] = z[i] / al[i];
  c[sys] = z[sys - 1];
  for (int i = sys - 2; i >= 0; i--) c[i + 1] = z[i] - ga[i] * c[i + 2];
}

double spline_eval(const std::vector<double>& xa, const std::vector<double>& ya, const std::vector<double>& c, double x) {
  int lo = 0, hi = (int)xa.size() - 1;
  while (hi > lo + 1) {
    const int mid = (hi + lo) / 2;
    if (xa[mid] > x)
      hi = mid;
    else
      lo = mid;
  }
  const double dx = xa[lo + 1] - xa[lo], dy = ya[lo + 1] - ya[lo], delx = x - xa[lo];
  const double b_i = (dy / dx) - dx * (c[lo + 1] + 2.0 * c[lo]) / 3.0;
  const double d_i = (c[lo + 1] - c[lo]) / (3.0 * dx);
  return ya[lo] + delx * (b_i + delx * (c[lo] + delx * d_i));
}

GC_HOSTDEV int Background::run_nodes(double om, bool reject_negative_density, const double* grid, double* hub, double* xg,
                                     long stride) {
  double y[3] = {1.0, 1.0, 0.0};
  double a = 1, h = -1e-6;
  double rhonu[GC_MAX_NU], pnu[GC_MAX_NU];
  for (int i = 0; i <= NB; ++i) {
    while (a > grid[i]) evolve_apply(a, grid[i], h, y);
    if (isnan(y[0]) || isnan(y[1]) || isnan(y[2])) return 8;
    if (stability(a, y)) return 7;
    const double h2 = y[0] * y[0];
    const double x1 = a * y[1], x2 = x1 * x1, x3 = x2 * x1, a3 = a * a * a;
    const double OmegaP = (0.5 * M.c2 * x2 - 6.0 * M.c3 * h2 * x3 + 22.5 * M.c4 * h2 * h2 * x2 * x2 -
                           21.0 * M.c5 * h2 * h2 * h2 * x3 * x2 - 9.0 * M.cG * h2 * x2) / 3.0;
    double OmegaM = 1 - OmegaP - M.orad / (a3 * a * h2);
    nu_pressures(grid[i], rhonu, pnu);
    for (int j = 0; j < M.n_eig; j++) OmegaM -= M.grhormass[j] * rhonu[j] / (a3 * a * 3. * h2 * (M.h0 * M.h0));
    const double OmTest = om / (a3 * h2);
    if (reject_negative_density && OmegaP < 0) return 5;                  // camb/theta.cc:263
    if (fabs((OmegaM - OmTest) / OmTest) > 1e-4) return 4;                // Friedmann closure, camb/galileon.cc:420
    hub[i * stride] = y[0];
    xg[i * stride] = y[1];
  }
  return 0;
}

// the node positions (amax q^i, i = 0 .. NB) do not depend on the model: computed once per process
static const std::vector<double>& background_nodes() {
  static const std::vector<double> nodes = [] {
    const double amin = 1e-10, amax = 1.;
    const double q = std::pow(amin / amax, 1. / NB);
    std::vector<double> g(NB + 1);
    for (int i = 0; i <= NB; i++) g[i] = amax * std::pow(q, i);
    return g;
  }();
  return nodes;
}

// calcHubbleGalileon, camb/galileon.cc:332-438 (= calcHubble, camb/theta.cc:179-295, which also rejects a negative
// Galileon density): from a = 1 (hbar = 1, dpi/da = 1) down to 1e-10 on NB + 1 log-spaced nodes, every node checked
int Background::run(double om, bool reject_negative_density, std::vector<double>& grid, std::vector<double>& hub,
                    std::vector<double>& xg) {
  grid = background_nodes();
  hub.assign(NB + 1, 0.0);
  xg.assign(NB + 1, 0.0);
  return run_nodes(om, reject_negative_density, grid.data(), hub.data(), xg.data(), 1);
}

// flatness closure: the highest non-zero coefficient follows from the others (camb/galileon.cc:520-575)
GC_HOSTDEV static void close_coefficients(Background& B, double om, double c3in, double c4in) {
  BgModel& M = B.M;
  const int neig = M.n_eig;
  const double grhom = 3 * (M.h0 * M.h0);
  double rhonu[GC_MAX_NU], pnu[GC_MAX_NU];
  B.nu_pressures(1.0, rhonu, pnu);
  if (c3in == 0 && c4in == 0) {
    M.c3 = 0.5 * (-1. + om + M.orad + M.c2 / 6. - 3 * M.cG);
    for (int i = 0; i < neig; i++) M.c3 += 0.5 * rhonu[i] * M.grhormass[i] / grhom;
    M.c4 = 0;
    M.c5 = 0;
  } else if (c4in == 0) {
    M.c3 = c3in;
    // the reference's right-hand side reads the c4 left over from the previous call (:546); kept
    M.c4 = (1. - om - M.orad - M.c2 / 6. + 2 * M.c3 - 7.5 * M.c4 + 3 * M.cG) / 7.5;
    for (int i = 0; i < neig; i++) M.c4 += -rhonu[i] * M.grhormass[i] / (7.5 * grhom);
    M.c5 = 0;
  } else {
    M.c3 = c3in;
    M.c4 = c4in;
    M.c5 = (-1. + om + M.orad + M.c2 / 6. - 2 * M.c3 + 7.5 * M.c4 - 3 * M.cG) / 7.;
    for (int i = 0; i < neig; i++) M.c5 += rhonu[i] * M.grhormass[i] / (7. * grhom);
  }
}

// ascending-a tables of NB + 2 nodes from the integration grid (descending a), with the linearly extrapolated node
// beyond a = 1 (camb/galileon.cc:655-663)
static void ascending_tables(const std::vector<double>& grid, const std::vector<double>& hub, const std::vector<double>& xg,
                             std::vector<double>& a, std::vector<double>& h, std::vector<double>& x) {
  const int n = NB + 2;
  a.assign(n, 0);
  h.assign(n, 0);
  x.assign(n, 0);
  for (int i = 0; i <= NB; i++) {
    a[i] = grid[NB - i];
    h[i] = hub[NB - i];
    x[i] = xg[NB - i];
  }
  const double q = std::pow(1e-10 / 1., 1. / NB);
  a[NB + 1] = 1. / q;
  h[NB + 1] = (h[NB] - h[NB - 1]) / (a[NB] - a[NB - 1]) * (a[NB + 1] - a[NB - 1]) + h[NB - 1];
  x[NB + 1] = (x[NB] - x[NB - 1]) / (a[NB] - a[NB - 1]) * (a[NB + 1] - a[NB - 1]) + x[NB - 1];
}

// exact integral of the natural spline between two abscissae (what gsl_spline_eval_integ returns)
double spline_integral(const std::vector<double>& xa, const std::vector<double>& ya, const std::vector<double>& c,
                       double a, double b) {
  auto find = [&](double x) {
    int lo = 0, hi = (int)xa.size() - 1;
    while (hi > lo + 1) {
      const int mid = (hi + lo) / 2;
      if (xa[mid] > x)
        hi = mid;
      else
        lo = mid;
    }
    return lo;
  };
  const int ia = find(a), ib = find(b);
  double result = 0.0;
  for (int i = ia; i <= ib; i++) {
    const double x_lo = xa[i], x_hi = xa[i + 1], y_lo = ya[i], dx = x_hi - x_lo, dy = ya[i + 1] - y_lo;
    const double b_i = (dy / dx) - dx * (c[i + 1] + 2.0 * c[i]) / 3.0;
    const double c_i = c[i];
    const double d_i = (c[i + 1] - c[i]) / (3.0 * dx);
    if (i == ia || i == ib) {
      const double x1 = (i == ia) ? a : x_lo, x2 = (i == ib) ? b : x_hi;
      const double r1 = x1 - x_lo, r2 = x2 - x_lo, r12 = r1 + r2, bterm = 0.5 * b_i * r12;
      const double cterm = (1.0 / 3.0) * c_i * (r1 * r1 + r2 * r2 + r1 * r2), dterm = 0.25 * d_i * r12 * (r1 * r1 + r2 * r2);
      result += (r2 - r1) * (y_lo + bterm + cterm + dterm);
    } else {
      result += dx * (y_lo + dx * (0.5 * b_i + dx * (c_i / 3.0 + 0.25 * d_i * dx)));
    }
  }
  return result;
}

}  // namespace

extern "C" {

int galileon_set_nu_tables_(const double* r1, const double* p1, const double* dr1, const double* dp1, const double* ddr1,
                            const double* ddp1, const int* n, const double* dlnam) {
  if (*n < 2) return 5;
  G.nu_tab.resize((size_t)*n * 6);
  for (int i = 0; i < *n; i++) {
    double* row = &G.nu_tab[(size_t)i * 6];
    row[0] = r1[i]; row[1] = dr1[i]; row[2] = p1[i]; row[3] = dp1[i]; row[4] = ddr1[i]; row[5] = ddp1[i];
  }
  G.M.nu_tab = G.nu_tab.data();
  G.M.dlnam = *dlnam;
  G.M.inv_dlnam = 1.0 / *dlnam;
  return 0;
}

// camb/galileon.cc:503-688 with calcHubbleGalileon (:332-438)
int arrays_(double* omegar, double* omegam, double* H0in, double* c2in, double* c3in, double* c4in, double* cGin,
            double* grhormass, double* nu_masses, int* nu_mass_eigenstates) {
  G.ready = false;
  BgModel& M = G.M;
  const int neig = *nu_mass_eigenstates;
  if (neig < 0 || neig > GC_MAX_NU) return 5;
  if (neig > 0 && G.nu_tab.empty()) return 5;  // galileon_set_nu_tables_ first
  M.n_eig = neig;
  for (int i = 0; i < neig; i++) {
    M.grhormass[i] = grhormass[i];
    M.nu_masses[i] = nu_masses[i];
  }
  G.om = *omegam;
  M.orad = *omegar;
  M.h0 = *H0in;
  M.c2 = *c2in;
  M.cG = *cGin;
  close_coefficients(G, G.om, *c3in, *c4in);
  std::vector<double> grid, hub, xg;
  if (int status = G.run(G.om, false, grid, hub, xg)) return status;
  ascending_tables(grid, hub, xg, G.a, G.h, G.x);
  spline_coefficients(G.a, G.h, G.ch);
  spline_coefficients(G.a, G.x, G.cx);
  G.ready = true;
  return 0;
}

}  // extern "C"

// Re-entrant form of arrays_ for the batched pre-stage (csrc/prestage.cu): one Background per call, no file-scope
// state, so host threads integrate different cosmologies side by side.  nu_tab = [n][6] rows {r1, dr1, p1, dp1, ddr1,
// ddp1}.  coef = {c2, c3, c4, c5, cG} after the flatness closure; a/h/x = NB + 2 ascending nodes.
int gc_galileon_background(double omegar, double omegam, double h0, double c2, double c3in, double c4in, double cG, int neig,
                           const double* grhormass, const double* nu_masses, const double* nu_tab, double dlnam, double* coef,
                           std::vector<double>& a, std::vector<double>& h, std::vector<double>& x) {
  if (neig < 0 || neig > GC_MAX_NU) return 5;
  Background B;
  BgModel& M = B.M;
  M.n_eig = neig;
  for (int i = 0; i < neig; i++) {
    M.grhormass[i] = grhormass[i];
    M.nu_masses[i] = nu_masses[i];
  }
  M.nu_tab = nu_tab;
  M.dlnam = dlnam;
  M.inv_dlnam = 1.0 / dlnam;
  M.orad = omegar;
  M.h0 = h0;
  M.c2 = c2;
  M.cG = cG;
  M.c3 = M.c4 = M.c5 = 0;
  close_coefficients(B, omegam, c3in, c4in);
  std::vector<double> grid, hub, xg;
  if (int status = B.run(omegam, false, grid, hub, xg)) return status;
  ascending_tables(grid, hub, xg, a, h, x);
  coef[0] = M.c2; coef[1] = M.c3; coef[2] = M.c4; coef[3] = M.c5; coef[4] = M.cG;
  return 0;
}

extern "C" {

// camb/theta.cc:296-426 -- theta = r_s(a*) / D_A(a*) for the theta <-> H0 solve of CosmoMC
// (source/CosmologyParameterizations.f90:263-271).  Its own background problem: the tables of arrays_ are untouched.
// h0 in km/s/Mpc here.  Returns -1 when the background is rejected, like the reference (:380).
double theta_(double* omegam, double* orad, double* ombh2, double* h0, double* astar, double* c2in, double* c3in, double* c4in,
              double* cGin, double* grhormass, double* nu_masses, int* nu_mass_eigenstates) {
  const int neig = *nu_mass_eigenstates;
  if (neig < 0 || neig > GC_MAX_NU || (neig > 0 && G.nu_tab.empty())) return -1;
  const double cl = 2.99792458e8;
  Background B;
  BgModel& M = B.M;
  M.nu_tab = G.M.nu_tab;
  M.dlnam = G.M.dlnam;
  M.inv_dlnam = G.M.inv_dlnam;
  M.n_eig = neig;
  for (int i = 0; i < neig; i++) {
    M.grhormass[i] = grhormass[i];
    M.nu_masses[i] = nu_masses[i];
  }
  const double om = *omegam;
  M.orad = *orad;
  M.h0 = *h0 * 1000 / cl;
  M.c2 = *c2in;
  M.cG = *cGin;
  const double grhom = 3 * (M.h0 * M.h0);
  double rhonu[GC_MAX_NU], pnu[GC_MAX_NU];
  B.nu_pressures(1.0, rhonu, pnu);
  if (*c3in == 0 && *c4in == 0) {
    M.c3 = 0.5 * (-1. + om + M.orad + M.c2 / 6. - 3 * M.cG);
    for (int i = 0; i < neig; i++) M.c3 += 0.5 * rhonu[i] * M.grhormass[i] / grhom;
    M.c4 = M.c5 = 0;
  } else if (*c4in == 0) {
    M.c3 = *c3in;
    M.c4 = (1. - om - M.orad - M.c2 / 6. + 2 * M.c3 + 3 * M.cG) / 7.5;  // local c4 = 0 on the right-hand side (:329)
    for (int i = 0; i < neig; i++) M.c4 += -rhonu[i] * M.grhormass[i] / (7.5 * grhom);
    M.c5 = 0;
  } else {
    M.c3 = *c3in;
    M.c4 = *c4in;
    M.c5 = (-1. + om + M.orad + M.c2 / 6. - 2 * M.c3 + 7.5 * M.c4 - 3 * M.cG) / 7.;
    for (int i = 0; i < neig; i++) M.c5 += rhonu[i] * M.grhormass[i] / (7. * grhom);
  }
  std::vector<double> grid, hub, xg;
  if (B.run(om, true, grid, hub, xg)) return -1;
  const int n = NB + 1;
  std::vector<double> av(n), rs(n), da(n), crs, cda;
  for (int i = 0; i < n; i++) {
    const double a = grid[NB - i], hb = hub[NB - i];
    av[i] = a;
    rs[i] = cl / (a * a * hb * (*h0) * 1000 * std::sqrt(3 * (1 + 3e4 * a * (*ombh2))));
    da[i] = cl / (a * a * hb * (*h0) * 1000);
  }
  spline_coefficients(av, rs, crs);
  spline_coefficients(av, da, cda);
  return spline_integral(av, rs, crs, 1e-8, *astar) / spline_integral(av, da, cda, *astar, 1.0);
}

void freegal_(void) {
  G.a.clear(); G.h.clear(); G.x.clear(); G.ch.clear(); G.cx.clear();
  G.ready = false;
}

// the tabulated background for galcamb_model.gal_a / gal_h / gal_x (the accessor INTEGRATION.md asks galileon.cc for)
int galgettable_(const double** a, const double** h, const double** x, int* n) {
  if (!G.ready) return 6;
  *a = G.a.data();
  *h = G.h.data();
  *x = G.x.data();
  *n = NB + 1;  // without the extrapolated node (the device library appends its own)
  return 0;
}

int galgetcoefficients_(double* c2, double* c3, double* c4, double* c5, double* cG) {
  *c2 = G.M.c2; *c3 = G.M.c3; *c4 = G.M.c4; *c5 = G.M.c5; *cG = G.M.cG;
  return G.ready ? 0 : 6;
}

double GetX_(double* point) {
  if (!G.ready || !(*point >= 0.0 && *point <= 1.1)) return std::nan("");
  return *point * spline_eval(G.a, G.x, G.cx, *point);
}

double GetH_(double* point) {
  if (!G.ready || !(*point >= 0.0 && *point <= 1.1)) return std::nan("");
  return spline_eval(G.a, G.h, G.ch, *point);
}

// the closed forms below are the ones the device evaluates (galileon_fast.cuh): one background record for given
// (a, H = a*hbar, x, dh, dx), then the term as a linear form in the perturbations
static gc::GalConst host_const() { return gc::GalConst{G.M.c2, G.M.c3, G.M.c4, G.M.c5, G.M.cG, G.M.h0, G.M.orad}; }
static double host_lam_nu(const double* pnu) {
  double lam = 0;
  for (int i = 0; i < G.M.n_eig; i++) lam += G.M.grhormass[i] * pnu[i];
  return lam;
}

void GetdHdX_(double* point, double* hcamb, double* xcamb, double& dh, double& dx, double* grhormass, double* nu_masses,
              int* nu_mass_eigenstates) {
  (void)grhormass; (void)nu_masses; (void)nu_mass_eigenstates;  // the values given to arrays_ are in use
  double rhonu[GC_MAX_NU], pnu[GC_MAX_NU];
  nu_pressures(*point, rhonu, pnu);
  gc::gal_fast_dhdx(host_const(), *point, *hcamb / *point, *xcamb, host_lam_nu(pnu), dh, dx);
}

double ct_(double* point, double* grhormass, double* nu_masses, int* nu_mass_eigenstates) {
  double xgal = GetX_(point), h = GetH_(point), ha = *point * h, dh = 0, dx = 0;
  const double prod = xgal * h, prod2 = prod * prod, prod4 = prod2 * prod2, prod5 = prod4 * prod;
  GetdHdX_(point, &ha, &xgal, dh, dx, grhormass, nu_masses, nu_mass_eigenstates);
  const double c4 = G.M.c4, c5 = G.M.c5, cG = G.M.cG;
  return std::sqrt((0.5 + 0.25 * c4 * prod4 + 1.5 * c5 * prod4 * h * (h * dx + dh * xgal) - 0.5 * cG * prod2) /
                   (0.5 - 0.75 * c4 * prod4 + 1.5 * c5 * prod5 * h + 0.5 * cG * prod2));
}

static void host_fill(double a, double hc, double x, double dh, double dx, double k, gc::GalFast& F) {
  gc::gal_fast_fill(host_const(), a, hc / a, x, dh, dx, k, F);
}

double grhogal_(double* point, double* hcamb, double* xcamb) {
  gc::GalFast F;
  host_fill(*point, *hcamb, *xcamb, 0, 0, 1, F);
  return F.grhogal;
}

double gpresgal_(double* point, double* hcamb, double* xcamb, double* dhcamb, double* dxcamb) {
  gc::GalFast F;
  host_fill(*point, *hcamb, *xcamb, *dhcamb, *dxcamb, 1, F);
  return F.gpresgal;
}

double Chigal_(double* point, double* hcamb, double* xcamb, double* dgrho, double* eta, double* dphi, double* dphiprime,
               double* k) {
  gc::GalFast F;
  host_fill(*point, *hcamb, *xcamb, 0, 0, *k, F);
  return gc::gal_fast_Chigal(F, *dgrho, *eta, *dphi, *dphiprime);
}

double qgal_(double* point, double* hcamb, double* xcamb, double* dgq, double* eta, double* dphi, double* dphiprime, double* k) {
  (void)eta;
  gc::GalFast F;
  host_fill(*point, *hcamb, *xcamb, 0, 0, *k, F);
  return gc::gal_fast_qgal(F, *dgq, *dphi, *dphiprime);
}

double Pigal_(double* point, double* hcamb, double* xcamb, double* dhcamb, double* dxcamb, double* dgrho, double* dgq,
              double* dgpi, double* eta, double* dphi, double* k) {
  gc::GalFast F;
  host_fill(*point, *hcamb, *xcamb, *dhcamb, *dxcamb, *k, F);
  return gc::gal_fast_Pigal(host_const(), F, *dgrho, *dgq, *dgpi, *eta, *dphi);
}

double dphisecond_(double* point, double* hcamb, double* xcamb, double* dhcamb, double* dxcamb, double* dgrho, double* dgq,
                   double* eta, double* dphi, double* dphiprime, double* k, double* deltafprime) {
  gc::GalFast F;
  host_fill(*point, *hcamb, *xcamb, *dhcamb, *dxcamb, *k, F);
  return gc::gal_fast_dphisecond(host_const(), F, *dgrho, *dgq, *eta, *dphi, *dphiprime, *deltafprime);
}

double pigalprime_(double* point, double* hcamb, double* xcamb, double* dhcamb, double* dxcamb, double* dgrho, double* dgq,
                   double* dgpi, double* pidot, double* eta, double* dphi, double* dphiprime, double* k, double* grho,
                   double* gpres, double* grhormass, double* nu_masses, int* nu_mass_eigenstates) {
  (void)grhormass; (void)nu_masses; (void)nu_mass_eigenstates;
  double rhonu[GC_MAX_NU], pnu[GC_MAX_NU];
  nu_pressures(*point, rhonu, pnu);
  const double a = *point, hc = *hcamb;
  // the six polynomials of the background system are re-formed from (a, H, x) as the reference does (:1105-1110);
  // dh and dx are the caller's
  gc::GalDyn B;
  double dh_own, dx_own;
  gc::gal_fast_dhdx(host_const(), a, hc / a, *xcamb, host_lam_nu(pnu), dh_own, dx_own, &B);
  gc::GalFast F;
  host_fill(a, hc, *xcamb, *dhcamb, *dxcamb, *k, F);
  double lam_nu_prime = 0;
  for (int i = 0; i < G.M.n_eig; i++)
    lam_nu_prime += G.M.grhormass[i] * (gc::Nu_dp(G.M, a * G.M.nu_masses[i], hc, pnu[i]) - 4 * pnu[i] * hc);
  gc::GalPiDotLin L;
  gc::gal_fast_pigalprime_lin(host_const(), F, B, lam_nu_prime, *grho + *gpres, L);
  return L.p * *dphiprime + L.d * *dphi + L.pid * *pidot + L.g * *dgrho + L.q * *dgq + L.pi * *dgpi + L.e * *eta;
}

}  // extern "C"


// ------------------------------------------------------------------------------------------------------------------
// The same background integration for a whole batch of cosmologies on the device: one lane per cosmology (the node loop
// is sequential: 30 000 Fehlberg steps, each depending on the one before), every lane running Background::run_nodes --
// the code above, compiled for the device without FMA contraction (this file is built -fmad=false), on the node positions
// computed by the host.  What can differ from the host run is the last bit of log / exp / pow (the neutrino pressure
// look-up and the step-size controller); tests/test_gpu_parity.py::test_device_background_against_host bounds it.
// The launch needs one warp slot and 8 K registers per 32 cosmologies and no shared memory, so it runs NEXT TO the
// per-k kernel of the previous batch (which leaves exactly that much of an SM free) on its own stream.
struct GcBackgroundIn {  // inputs of arrays_ (camb/galileon.cc:503-519) for one cosmology
  double omegar, omegam, h0, c2, c3in, c4in, cG;
  int neig;
  double grhormass[GC_MAX_NU], nu_masses[GC_MAX_NU];
};

namespace {

__global__ void __launch_bounds__(32) gal_background_kernel(const GcBackgroundIn* __restrict__ in, int n, const double* __restrict__ nu_tab,
                                                            double dlnam, const double* __restrict__ nodes, double* __restrict__ h,
                                                            double* __restrict__ x, double* __restrict__ coef, int* __restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const GcBackgroundIn I = in[i];
  if (I.neig < 0 || I.neig > GC_MAX_NU) {
    status[i] = 5;
    return;
  }
  Background B;
  BgModel& M = B.M;
  M.n_eig = I.neig;
  for (int j = 0; j < I.neig; j++) {
    M.grhormass[j] = I.grhormass[j];
    M.nu_masses[j] = I.nu_masses[j];
  }
  M.nu_tab = nu_tab;
  M.dlnam = dlnam;
  M.inv_dlnam = 1.0 / dlnam;
  M.orad = I.omegar;
  M.h0 = I.h0;
  M.c2 = I.c2;
  M.cG = I.cG;
  M.c3 = M.c4 = M.c5 = 0;
  close_coefficients(B, I.omegam, I.c3in, I.c4in);
  // ascending a: node j of the descending integration grid is element NB - j of the model's table (NB + 2 per model)
  double* hm = h + (size_t)i * (NB + 2) + NB;
  double* xm = x + (size_t)i * (NB + 2) + NB;
  status[i] = B.run_nodes(I.omegam, false, nodes, hm, xm, -1);
  coef[i * 5 + 0] = M.c2; coef[i * 5 + 1] = M.c3; coef[i * 5 + 2] = M.c4; coef[i * 5 + 3] = M.c5; coef[i * 5 + 4] = M.cG;
}

struct DeviceBackground {
  int device = -1;
  cudaStream_t stream = nullptr;
  GcBackgroundIn* d_in = nullptr;
  double *d_nu = nullptr, *d_nodes = nullptr, *d_h = nullptr, *d_x = nullptr, *d_coef = nullptr;
  int* d_status = nullptr;
  double *p_h = nullptr, *p_x = nullptr;  // pinned
  size_t cap = 0;
  size_t nu_rows = 0;
};
DeviceBackground DB;

}  // namespace

// n cosmologies at once on `device`.  h_out / x_out: per model NB + 2 ascending values (the extrapolated node beyond a = 1
// included, like gc_galileon_background); a_out: the NB + 2 abscissae (the same for every model); coef_out [n][5];
// status_out [n] as arrays_.  Returns 0 or a CUDA failure as 100.
int gc_galileon_background_device(int device, const GcBackgroundIn* in, int n, const double* nu_tab, size_t nu_rows, double dlnam,
                                  std::vector<double>& a_out, const double** h_out, const double** x_out, double* coef_out,
                                  int* status_out) {
#define GCB(call)                        \
  do {                                   \
    if ((call) != cudaSuccess) return 100; \
  } while (0)
  for (int i = 0; i < n; i++)
    if (in[i].neig > 0 && (nu_tab == nullptr || nu_rows == 0)) return 100;  // no neutrino tables: the caller integrates on the host
  GCB(cudaSetDevice(device));
  const std::vector<double>& nodes = background_nodes();
  if (DB.device != device) {
    if (DB.device >= 0) return 100;  // one device per process for this path
    GCB(cudaStreamCreateWithFlags(&DB.stream, cudaStreamNonBlocking));
    GCB(cudaMalloc(&DB.d_nodes, nodes.size() * sizeof(double)));
    GCB(cudaMemcpyAsync(DB.d_nodes, nodes.data(), nodes.size() * sizeof(double), cudaMemcpyHostToDevice, DB.stream));
    DB.device = device;
  }
  if (nu_rows > 0 && DB.nu_rows != nu_rows) {
    if (DB.d_nu) cudaFree(DB.d_nu);
    GCB(cudaMalloc(&DB.d_nu, nu_rows * 6 * sizeof(double)));
    GCB(cudaMemcpyAsync(DB.d_nu, nu_tab, nu_rows * 6 * sizeof(double), cudaMemcpyHostToDevice, DB.stream));
    DB.nu_rows = nu_rows;
  }
  if ((size_t)n > DB.cap) {
    if (DB.d_in) { cudaFree(DB.d_in); cudaFree(DB.d_h); cudaFree(DB.d_x); cudaFree(DB.d_coef); cudaFree(DB.d_status); cudaFreeHost(DB.p_h); cudaFreeHost(DB.p_x); }
    const size_t per = (size_t)(NB + 2);
    GCB(cudaMalloc(&DB.d_in, n * sizeof(GcBackgroundIn)));
    GCB(cudaMalloc(&DB.d_h, n * per * sizeof(double)));
    GCB(cudaMalloc(&DB.d_x, n * per * sizeof(double)));
    GCB(cudaMalloc(&DB.d_coef, (size_t)n * 5 * sizeof(double)));
    GCB(cudaMalloc(&DB.d_status, n * sizeof(int)));
    GCB(cudaMallocHost(&DB.p_h, n * per * sizeof(double)));
    GCB(cudaMallocHost(&DB.p_x, n * per * sizeof(double)));
    DB.cap = n;
  }
  const size_t per = (size_t)(NB + 2);
  GCB(cudaMemcpyAsync(DB.d_in, in, n * sizeof(GcBackgroundIn), cudaMemcpyHostToDevice, DB.stream));
  gal_background_kernel<<<(n + 31) / 32, 32, 0, DB.stream>>>(DB.d_in, n, DB.d_nu, dlnam, DB.d_nodes, DB.d_h, DB.d_x, DB.d_coef, DB.d_status);
  GCB(cudaGetLastError());
  GCB(cudaMemcpyAsync(DB.p_h, DB.d_h, n * per * sizeof(double), cudaMemcpyDeviceToHost, DB.stream));
  GCB(cudaMemcpyAsync(DB.p_x, DB.d_x, n * per * sizeof(double), cudaMemcpyDeviceToHost, DB.stream));
  GCB(cudaMemcpyAsync(coef_out, DB.d_coef, (size_t)n * 5 * sizeof(double), cudaMemcpyDeviceToHost, DB.stream));
  GCB(cudaMemcpyAsync(status_out, DB.d_status, n * sizeof(int), cudaMemcpyDeviceToHost, DB.stream));
  GCB(cudaStreamSynchronize(DB.stream));
#undef GCB
  // abscissae and the linearly extrapolated node beyond a = 1 (camb/galileon.cc:655-663), as ascending_tables()
  a_out.assign(NB + 2, 0.0);
  for (int i = 0; i <= NB; i++) a_out[i] = nodes[NB - i];
  const double q = std::pow(1e-10 / 1., 1. / NB);
  a_out[NB + 1] = 1. / q;
  for (int m = 0; m < n; m++) {
    if (status_out[m]) continue;
    double* h = DB.p_h + (size_t)m * per;
    double* x = DB.p_x + (size_t)m * per;
    h[NB + 1] = (h[NB] - h[NB - 1]) / (a_out[NB] - a_out[NB - 1]) * (a_out[NB + 1] - a_out[NB - 1]) + h[NB - 1];
    x[NB + 1] = (x[NB] - x[NB - 1]) / (a_out[NB] - a_out[NB - 1]) * (a_out[NB + 1] - a_out[NB - 1]) + x[NB - 1];
  }
  *h_out = DB.p_h;
  *x_out = DB.p_x;
  return 0;
}
