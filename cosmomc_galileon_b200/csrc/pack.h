// Packing of one model's host tables (galcamb_model, SoA like the Fortran arrays) into the device layout of csrc/model.h:
//   th[nthermo][5] | dotmu_pmin[nthermo] | ts[nt][8] | gal[ngal][5] | k[nk] | q[nq] | dq[nq] | k_tr[nk_tr]
// Used by galcamb_set_model_ / galcamb_batch_upload_ (csrc/capi.cu) and, for models the library's own pre-stage builds,
// already by the pre-stage threads (csrc/prestage.cu) -- in the pipelined form that is behind the device work of the
// previous batch.  Host code only.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <vector>

#include "../../include/galcamb.h"

namespace gcpack {

struct Info {
  size_t total = 0, n_th = 0, n_pm = 0, n_ts = 0, n_gal = 0;
  int ngal = 0;
  double gal_a0 = 0, gal_lnq_inv = 0;  // index guess into the background table: i ~ log(a/a0) * gal_lnq_inv
};

// natural cubic spline, c[i] = y''(x_i)/2 (the representation gsl_interp_cspline keeps)
inline void natural_spline_c(const std::vector<double>& xa, const std::vector<double>& ya, std::vector<double>& c) {
  const int n = (int)xa.size();
  const int sys = n - 2;
  c.assign(n, 0.0);
  std::vector<double> g_(sys), diag(sys), off(sys), al(sys), ga(sys), z(sys);
  for (int i = 0; i < sys; i++) {
    const double h_i = xa[i + 1] - xa[i], h_ip1 = xa[i + 2] - xa[i + 1];
    const double yd_i = ya[i + 1] - ya[i], yd_ip1 = ya[i + 2] - ya[i + 1];
    const double g_i = (h_i != 0.0) ? 1.0 / h_i : 0.0, g_ip1 = (h_ip1 != 0.0) ? 1.0 / h_ip1 : 0.0;
    off[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g_[i] = 3.0 * (yd_ip1 * g_ip1 - yd_i * g_i);
  }
  al[0] = diag[0];
  ga[0] = off[0] / al[0];
  for (int i = 1; i < sys - 1; i++) {
    al[i] = diag[i] - off[i - 1] * ga[i - 1];
    ga[i] = off[i] / al[i];
  }
  if (sys > 1) al[sys - 1] = diag[sys - 1] - off[sys - 2] * ga[sys - 2];
  z[0] = g_[0];
  for (int i = 1; i < sys; i++) z[i] = g_[i] - ga[i - 1] * z[i - 1];
  for (int i = 0; i < sys; i++) z[i] = z[i] / al[i];
  c[sys] = z[sys - 1];
  for (int i = sys - 2; i >= 0; i--) c[i + 1] = z[i] - ga[i] * c[i + 2];
}


inline void sizes(const galcamb_model* m, int nk_tr, Info& I) {
  I.ngal = m->use_galileon ? m->ngal + 1 : 0;  // + the linearly extrapolated node of galileon.cc:677-679
  I.n_th = (size_t)m->nthermo * 5;
  I.n_pm = m->nthermo;
  I.n_ts = (size_t)m->nt * 8;
  I.n_gal = (size_t)I.ngal * 5;
  I.total = I.n_th + I.n_pm + I.n_ts + I.n_gal + m->nk + 2 * (size_t)m->nq + (size_t)nk_tr;
}

// p: I.total doubles (sizes() first)
inline void tables(const galcamb_model* m, int nk_tr, double* p, Info& I) {
  for (int i = 0; i < m->nthermo; i++) {
    p[i * 5 + 0] = m->cs2[i]; p[i * 5 + 1] = m->dcs2[i]; p[i * 5 + 2] = m->dotmu[i];
    p[i * 5 + 3] = m->ddotmu[i]; p[i * 5 + 4] = m->dddotmu[i];
  }
  p += I.n_th;
  {
    double mn = m->dotmu[0];
    for (int i = 0; i < m->nthermo; i++) { mn = std::min(mn, m->dotmu[i]); p[i] = mn; }
  }
  p += I.n_pm;
  for (int i = 0; i < m->nt; i++) {
    p[i * 8 + 0] = m->tau[i]; p[i * 8 + 1] = m->dtau[i]; p[i * 8 + 2] = m->vis[i]; p[i * 8 + 3] = m->dvis[i];
    p[i * 8 + 4] = m->ddvis[i]; p[i * 8 + 5] = m->expmmu[i]; p[i * 8 + 6] = m->opac[i]; p[i * 8 + 7] = m->dopac[i];
  }
  p += I.n_ts;
  const int ngal = I.ngal;
  if (ngal) {
    const int nb1 = m->ngal;  // nb+1 tabulated points, ascending a
    std::vector<double> a(ngal), hh(ngal), xx(ngal), ch, cx;
    for (int i = 0; i < nb1; i++) { a[i] = m->gal_a[i]; hh[i] = m->gal_h[i]; xx[i] = m->gal_x[i]; }
    // galileon.cc:677: intvar_interp[nb+1] = 1./q with q the geometric ratio; a[nb] = 1 so q = a[nb-1] exactly
    a[nb1] = 1. / (a[nb1 - 2] / a[nb1 - 1]);
    hh[nb1] = (hh[nb1 - 1] - hh[nb1 - 2]) / (a[nb1 - 1] - a[nb1 - 2]) * (a[nb1] - a[nb1 - 2]) + hh[nb1 - 2];
    xx[nb1] = (xx[nb1 - 1] - xx[nb1 - 2]) / (a[nb1 - 1] - a[nb1 - 2]) * (a[nb1] - a[nb1 - 2]) + xx[nb1 - 2];
    natural_spline_c(a, hh, ch);
    natural_spline_c(a, xx, cx);
    for (int i = 0; i < ngal; i++) {
      p[i * 5 + 0] = a[i]; p[i * 5 + 1] = hh[i]; p[i * 5 + 2] = ch[i]; p[i * 5 + 3] = xx[i]; p[i * 5 + 4] = cx[i];
    }
    I.gal_a0 = a[0];
    I.gal_lnq_inv = (nb1 - 1) / std::log(a[nb1 - 1] / a[0]);
  }
  p += I.n_gal;
  std::copy(m->k, m->k + m->nk, p); p += m->nk;
  std::copy(m->q, m->q + m->nq, p); p += m->nq;
  std::copy(m->dq, m->dq + m->nq, p); p += m->nq;
  if (nk_tr) { std::copy(m->k_transfer, m->k_transfer + nk_tr, p); p += nk_tr; }
}

}  // namespace gcpack
