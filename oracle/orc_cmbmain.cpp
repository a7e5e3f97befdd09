// TEST INFRASTRUCTURE (oracle) -- see orc.hpp header.
// Orchestration of one scalar, flat model: module CAMBmain of camb/cmbmain.f90, the l-sampling of
// camb/modules.f90:851-1089, the flat Bessel tables of camb/bessels.f90:34-275 and the primordial
// power law of camb/power_tilt.f90:149-172.  OpenMP structure mirrors the reference:
// SCHEDULE(DYNAMIC) over source k (cmbmain.f90:201-206), SCHEDULE(STATIC,4) over integration k
// (:263-267), serial CalcScalCls (:2191-2193).
#include <omp.h>

#include <chrono>

#include "orc.hpp"

namespace orc {

#define SP(x) ((double)(x##f))
static inline long nint(double x) { return std::lround(x); }

void CalcScalarSourcesForK(Model& M, int q_ix, Counters& cnt, int& err, std::string& errmsg);
void GetTransferForK(Model& M, int q_ix, Counters& cnt, int& err, std::string& errmsg);
double orc_GetTauStart(const Model& M, double q);

static const double xlimmin = 35.0, xlimfrac = 0.05;  // camb/bessels.f90:14

// camb/modules.f90:851-1008
void initlval(Model& M, int max_l) {
  const Params& P = M.P;
  std::vector<int> ls(4001, 0);
  int lind = 0, lvar, step, top, bot;
  const double Ascale = M.scale / P.lSampleBoost;
  lSamples& lSet = M.lSamp;
  auto finish = [&]() {
    lSet.l0 = lind;
    lSet.l.assign(lind + 1, 0);
    for (int i = 1; i <= lind; i++) lSet.l[i] = ls[i];
  };
  if (P.lSampleBoost >= 50) {
    for (lvar = lmin; lvar <= max_l; lvar++) ls[++lind] = lvar;
    finish();
    return;
  }
  for (lvar = lmin; lvar <= 10; lvar++) ls[++lind] = lvar;
  if (P.AccurateReionization) {
    if (P.lSampleBoost > 1) {
      for (lvar = 11; lvar <= 37; lvar += 1) ls[++lind] = lvar;
    } else {
      for (lvar = 11; lvar <= 37; lvar += 2) ls[++lind] = lvar;
    }
    step = std::max((int)nint(5 * Ascale), 2);
    bot = 40;
    top = bot + step * 10;
  } else {
    if (P.lSampleBoost > 1) {
      for (lvar = 11; lvar <= 15; lvar++) ls[++lind] = lvar;
    } else {
      ls[++lind] = 12;
      ls[++lind] = 15;
    }
    step = std::max((int)nint(10 * Ascale), 3);
    bot = 15 + std::max(step / 2, 2);
    top = bot + step * 7;
  }
  for (lvar = bot; lvar <= top; lvar += step) ls[++lind] = lvar;
  // Log_lvalues = .false.
  step = std::max((int)nint(20 * Ascale), 4);
  bot = ls[lind] + step;
  top = bot + step * 2;
  for (lvar = bot; lvar <= top; lvar += step) ls[++lind] = lvar;
  auto clip_to_max = [&]() {
    for (lvar = lind; lvar >= 1; lvar--)
      if (ls[lvar] <= max_l) break;
    lind = lvar;
    if (ls[lind] < max_l) ls[++lind] = max_l;
  };
  if (ls[lind] >= max_l) {
    clip_to_max();
  } else {
    step = std::max((int)nint(25 * Ascale), 4);
    bot = ls[lind] + step;
    top = bot + step;
    for (lvar = bot; lvar <= top; lvar += step) ls[++lind] = lvar;
    if (ls[lind] >= max_l) {
      clip_to_max();
    } else {
      if (P.HighAccuracyDefault && !P.use_spline_template)
        step = std::max((int)nint(42 * Ascale), 7);
      else
        step = std::max((int)nint(50 * Ascale), 7);
      bot = ls[lind] + step;
      top = std::min(5000, max_l);
      for (lvar = bot; lvar <= top; lvar += step) ls[++lind] = lvar;
      if (max_l > 5000) {
        step = std::max((int)nint(400 * Ascale), 50);
        lvar = ls[lind];
        for (;;) {
          lvar = lvar + step;
          if (lvar > max_l) break;
          ls[++lind] = lvar;
          step = (int)nint(step * SP(1.5));
        }
      }
      if (ls[lind] != max_l) ls[++lind] = max_l;
    }
  }
  finish();
}

// camb/cmbmain.f90:722-795
void cmb_InitVars(Model& M) {
  const Params& P = M.P;
  const double initAccuracyBoost = P.AccuracyBoost;
  M.qmax = M.maximum_qeta / M.tau0;
  M.qmin = 0.1 / M.tau0 / initAccuracyBoost;
  M.dtaurec_q = 4 / M.qmax / initAccuracyBoost;
  M.dtaurec = M.dtaurec_q;
  M.max_etak_scalar = initAccuracyBoost * std::fmax(1700.0, M.maximum_qeta) / 20;
  if (M.maximum_qeta < 3500 && P.AccuracyBoost < 2) M.max_etak_scalar = M.max_etak_scalar * 1.5;
  double maxq = M.qmax;  // WantCls (:757-762)
  if (P.WantTransfer) maxq = std::fmax(M.qmax, P.transfer_kmax);
  const double taumin = orc_GetTauStart(M, maxq);
  inithermo(M, taumin, M.tau0);
  if (M.error_flag) return;
  GaugeInterface_Init(M);
  // times of the transfer-function outputs, :783-793  (TimeOfz, camb/modules.f90:574-580)
  M.tautf.assign(1, 0.0);
  if (P.WantTransfer) {
    for (int itf = 1; itf <= P.transfer_num_redshifts; itf++) {
      M.tautf.push_back(std::fmin(DeltaTime(M, 0.0, 1.0 / (P.transfer_redshifts[itf - 1] + 1.0)), M.tau0));
      if (itf > 1 && M.tautf[itf] <= M.tautf[itf - 1]) {
        GlobalError(M, "Transfer redshifts not set or out of order", error_unsupported_params);
        return;
      }
    }
  }
}

// camb/cmbmain.f90:508-632 (flat, WantCls): the transfer wavenumber grid, merged behind the C_l source grid
void InitTransfer(Model& M) {
  const Params& P = M.P;
  std::vector<double> q_transfer(1, 0.0);  // 1-based
  int num_q_trans = 0;
  if (P.transfer_k_per_logint == 0) {
    double boost = P.AccuracyBoost;
    if (P.transfer_high_precision) boost = boost * 1.5;
    const double q_switch_lowk1 = SP(0.7) / M.taurst;
    const double dlog_lowk1 = 2 * boost;
    const double q_switch_lowk = 8 / M.taurst;
    double dlog_lowk = 8 * boost;
    if (P.HighAccuracyDefault) dlog_lowk = dlog_lowk * 2.5;
    const double q_switch_osc = std::fmin(P.transfer_kmax, 30 / M.taurst);
    double d_osc = 200 * boost;
    if (P.HighAccuracyDefault) d_osc = d_osc * SP(1.8);
    double q_switch_highk = std::fmin(P.transfer_kmax, 60 / M.taurst);
    const double dlog_osc = 17 * boost;
    if (P.HighAccuracyDefault) q_switch_highk = std::fmin(P.transfer_kmax, 90 / M.taurst);
    const double dlog_highk = 3 * boost;
    const double amin = 5e-5;
    int nq = (int)((std::log(P.transfer_kmax / amin)) * d_osc) + 1;
    q_transfer.assign(nq + 2, 0.0);
    nq = (int)((std::log(q_switch_lowk1 / amin)) * dlog_lowk1) + 1;
    for (int q_ix = 1; q_ix <= nq; q_ix++) q_transfer[q_ix] = amin * std::exp((q_ix - 1) / dlog_lowk1);
    num_q_trans = nq;
    nq = (int)(std::log(q_switch_lowk / q_transfer[num_q_trans]) * dlog_lowk) + 1;
    for (int q_ix = 1; q_ix <= nq; q_ix++) q_transfer[num_q_trans + q_ix] = q_transfer[num_q_trans] * std::exp(q_ix / dlog_lowk);
    num_q_trans = num_q_trans + nq;
    nq = (int)((q_switch_osc - q_transfer[num_q_trans]) * d_osc) + 1;
    for (int q_ix = 1; q_ix <= nq; q_ix++) q_transfer[num_q_trans + q_ix] = q_transfer[num_q_trans] + q_ix / d_osc;
    num_q_trans = num_q_trans + nq;
    if (P.transfer_kmax > q_transfer[num_q_trans]) {
      nq = (int)(std::log(q_switch_highk / q_transfer[num_q_trans]) * dlog_osc) + 1;
      for (int q_ix = 1; q_ix <= nq; q_ix++) q_transfer[num_q_trans + q_ix] = q_transfer[num_q_trans] * std::exp(q_ix / dlog_osc);
      num_q_trans = num_q_trans + nq;
    }
    if (P.transfer_kmax > q_transfer[num_q_trans]) {
      nq = (int)(std::log(P.transfer_kmax / q_transfer[num_q_trans]) * dlog_highk) + 1;
      for (int q_ix = 1; q_ix <= nq; q_ix++) q_transfer[num_q_trans + q_ix] = q_transfer[num_q_trans] * std::exp(q_ix / dlog_highk);
      num_q_trans = num_q_trans + nq;
    }
  } else {
    num_q_trans = (int)((std::log(P.transfer_kmax) - std::log(M.qmin)) * P.transfer_k_per_logint) + 1;
    q_transfer.assign(num_q_trans + 1, 0.0);
    for (int q_ix = 1; q_ix <= num_q_trans; q_ix++) q_transfer[q_ix] = M.qmin * (double)std::exp((float)q_ix / (float)P.transfer_k_per_logint);  // default-real exp
  }
  // WantCls: keep the source grid, append the transfer wavenumbers above its top (:597-623)
  const int ntodo = num_q_trans, nk = M.Evolve_q.npoints;
  int first_i = ntodo + 1;
  for (int q_ix = 1; q_ix <= ntodo; q_ix++)
    if (q_transfer[q_ix] > M.Evolve_q.Highest + 1e-4) {
      first_i = q_ix;
      break;
    }
  M.num_q_trans = first_i > ntodo ? nk : nk + (ntodo - first_i + 1);
  M.q_trans.assign(M.num_q_trans + 1, 0.0);
  for (int i = 1; i <= nk; i++) M.q_trans[i] = M.Evolve_q.points[i];
  for (int i = nk + 1; i <= M.num_q_trans; i++) M.q_trans[i] = q_transfer[first_i + (i - nk - 1)];
  M.TransferData.assign((size_t)Transfer_max * M.num_q_trans * P.transfer_num_redshifts, 0.0);
  M.sigma_8.assign(P.transfer_num_redshifts, 0.0);
}

// camb/cmbmain.f90:1151-1179: the wavenumbers beyond the C_l source grid (OpenMP over k like the reference)
void TransferOut(Model& M) {
  const int nk = M.Evolve_q.npoints;
  const int nth = M.P.nthreads > 0 ? M.P.nthreads : omp_get_max_threads();
  int err_any = 0;
  std::string err_msg;
  Counters total;
#pragma omp parallel for schedule(dynamic) num_threads(nth)
  for (int q_ix = nk + 1; q_ix <= M.num_q_trans; q_ix++) {
    Counters c;
    int err = 0;
    std::string msg;
    GetTransferForK(M, q_ix, c, err, msg);
#pragma omp critical
    {
      total.n_derivs += c.n_derivs;
      total.sum_n_trial += c.sum_n_trial;
      total.flop_ode += c.flop_ode;
      if (err) {
        err_any = err;
        err_msg = msg;
      }
    }
  }
  M.cnt.n_derivs += total.n_derivs;
  M.cnt.sum_n_trial += total.sum_n_trial / 8.0;
  M.cnt.flop_ode += total.flop_ode;
  if (err_any) GlobalError(M, err_msg.c_str(), err_any);
}

// Transfer_Get_sigmas -> Transfer_Get_SigmaR, camb/modules.f90:2450-2472, :2304-2370: sigma_8 of the total matter
// transfer function (transfer_power_var = Transfer_tot, R = 8 Mpc/h) at every transfer redshift
void Transfer_Get_sigmas(Model& M) {
  const int nz = M.P.transfer_num_redshifts;
  const double radius = 8.0, H = M.P.H0 / 100.0;
  std::vector<double> dsig8(nz, 0.0), dsig8o(nz, 0.0), sig8(nz, 0.0);
  double lnko = 0, dlnk;
  for (int ik = 1; ik <= M.num_q_trans; ik++) {
    const double kh = M.TransferData[(Transfer_kh - 1) + (size_t)Transfer_max * (ik - 1)];
    if (kh == 0) continue;
    const double k = kh * H;
    const double x = kh * radius;
    const double win = 3 * (std::sin(x) - x * std::cos(x)) / (x * x * x);
    const double lnk = std::log(k);
    if (ik == 1)
      dlnk = 0.5;
    else
      dlnk = lnk - lnko;
    const double powers = ScalarPower(M, k);
    for (int iz = 0; iz < nz; iz++) {
      const double t = M.TransferData[(Transfer_tot - 1) + (size_t)Transfer_max * ((ik - 1) + (size_t)M.num_q_trans * iz)];
      dsig8[iz] = ((win * (k * k)) * (win * (k * k))) * powers * (t * t);
      sig8[iz] = sig8[iz] + (dsig8[iz] + dsig8o[iz]) * dlnk / 2;
      dsig8o[iz] = dsig8[iz];
    }
    lnko = lnk;
  }
  for (int iz = 0; iz < nz; iz++) M.sigma_8[iz] = std::sqrt(sig8[iz]);
}

// camb/cmbmain.f90:797-852
void SetkValuesForSources(Model& M) {
  const Params& P = M.P;
  const double SourceAccuracyBoost = P.AccuracyBoost;
  double dlnk0, dkn1, dkn2;
  if (M.reionization && P.AccuratePolarization)
    dlnk0 = 2.0 / 10 / SourceAccuracyBoost;
  else
    dlnk0 = 5.0 / 10 / SourceAccuracyBoost;
  if (P.AccurateReionization) dlnk0 = dlnk0 / 2;
  dkn1 = 0.6 / M.taurst / SourceAccuracyBoost;
  dkn2 = 0.9 / M.taurst / SourceAccuracyBoost;
  if (P.HighAccuracyDefault) dkn2 = dkn2 / SP(1.2);
  const double qmax_log = dkn1 / dlnk0;
  const double q_switch = (double)(2 * 6.3f) / M.taurst;
  const int l_smooth_sample = 3000;
  double q_cmb = 2 * l_smooth_sample / M.chi0 * P.AccuracyBoost;
  if (M.maximum_l > 5000 && P.AccuratePolarization) q_cmb = q_cmb * SP(1.4);
  double dksmooth = q_cmb / 2 / (P.AccuracyBoost * P.AccuracyBoost);
  dksmooth = dksmooth / 6;  // Want_CMB
  Ranges_Init(M.Evolve_q);
  Ranges_Add_delta(M.Evolve_q, M.qmin, qmax_log, dlnk0, true);
  Ranges_Add_delta(M.Evolve_q, qmax_log, std::fmin(M.qmax, q_switch), dkn1);
  if (M.qmax > q_switch) {
    Ranges_Add_delta(M.Evolve_q, q_switch, std::fmin(q_cmb, M.qmax), dkn2);
    if (M.qmax > q_cmb) {
      dksmooth = std::log(1 + dksmooth / q_cmb);
      Ranges_Add_delta(M.Evolve_q, q_cmb, M.qmax, dksmooth, true);
    }
  }
  Ranges_GetArray(M.Evolve_q, false);
}

// hot loop 1, camb/cmbmain.f90:201-206
void CalcAllSources(Model& M) {
  const int nk = M.Evolve_q.npoints, nt = M.TimeSteps.npoints, S = M.SourceNum;
  M.Src.assign((size_t)nk * S * nt, 0.0);
  M.ddSrc.assign((size_t)nk * S * nt, 0.0);
  M.per_k_trials.assign(nk + 1, 0.0);
  Counters total;
  int err_any = 0;
  std::string err_msg;
  const int nth = M.P.nthreads > 0 ? M.P.nthreads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic) num_threads(nth)
  for (int q_ix = 1; q_ix <= nk; q_ix++) {
    int stop;
#pragma omp atomic read
    stop = err_any;
    if (stop) continue;
    Counters c;
    int err = 0;
    std::string msg;
    CalcScalarSourcesForK(M, q_ix, c, err, msg);
    M.per_k_trials[q_ix] = c.n_derivs / 8.0;
#pragma omp critical
    {
      total.n_derivs += c.n_derivs;
      total.n_out += c.n_out;
      total.sum_n_trial += c.sum_n_trial;
      total.flop_ode += c.flop_ode;
      if (err) {
        err_any = err;
        err_msg = msg;
      }
    }
  }
  M.cnt.n_derivs = total.n_derivs;
  M.cnt.n_out = total.n_out;
  M.cnt.n_trial = (total.n_derivs - total.n_out) / 8.0;
  M.cnt.sum_n_trial = total.sum_n_trial / 8.0;
  M.cnt.flop_ode = total.flop_ode;
  if (err_any) GlobalError(M, err_msg.c_str(), err_any);
}

// camb/cmbmain.f90:1267-1278
void InitSourceInterpolation(Model& M) {
  const int nk = M.Evolve_q.npoints, nt = M.TimeSteps.npoints, S = M.SourceNum;
  const int nth = M.P.nthreads > 0 ? M.P.nthreads : omp_get_max_threads();
#pragma omp parallel for num_threads(nth)
  for (int i = 1; i <= nt; i++)
    for (int j = 1; j <= S; j++) {
      size_t off = (size_t)nk * ((j - 1) + (size_t)S * (i - 1));
      spline(M.Evolve_q.points.data(), M.Src.data() + off - 1, nk, spl_large, spl_large, M.ddSrc.data() + off - 1);
    }
}

// camb/bessels.f90:132-275
void bjl(int L, double X, double& JL) {
  const double LN2 = 0.6931471805599453094, ONEMLN2 = 0.30685281944005469058277, PID2 = 1.5707963267948966192313217,
               PID4 = 0.78539816339744830961566084582, ROOTPI12 = 21.269446210866192327578,
               GAMMA1 = 2.6789385347077476336556, GAMMA2 = 1.3541179394264004169452;
  const double AX = std::fabs(X), AX2 = AX * AX;
  if (L < 7) {
    if (L == 0) {
      JL = AX < 1e-1 ? 1.0 - AX2 / 6.0 * (1.0 - AX2 / 20.0) : std::sin(AX) / AX;
    } else if (L == 1) {
      JL = AX < 2e-1 ? AX / 3.0 * (1.0 - AX2 / 10.0 * (1.0 - AX2 / 28.0)) : (std::sin(AX) / AX - std::cos(AX)) / AX;
    } else if (L == 2) {
      JL = AX < 3e-1 ? AX2 / 15.0 * (1.0 - AX2 / 14.0 * (1.0 - AX2 / 36.0))
                     : (-3.0 * std::cos(AX) / AX - std::sin(AX) * (1.0 - 3.0 / AX2)) / AX;
    } else if (L == 3) {
      JL = AX < 4e-1 ? AX * AX2 / 105.0 * (1.0 - AX2 / 18.0 * (1.0 - AX2 / 44.0))
                     : (std::cos(AX) * (1.0 - 15.0 / AX2) - std::sin(AX) * (6.0 - 15.0 / AX2) / AX) / AX;
    } else if (L == 4) {
      JL = AX < 6e-1 ? (AX2 * AX2) / 945.0 * (1.0 - AX2 / 22.0 * (1.0 - AX2 / 52.0))
                     : (std::sin(AX) * (1.0 - (45.0 - 105.0 / AX2) / AX2) + std::cos(AX) * (10.0 - 105.0 / AX2) / AX) / AX;
    } else if (L == 5) {
      JL = AX < 1.0 ? (AX2 * AX2) * AX / 10395.0 * (1.0 - AX2 / 26.0 * (1.0 - AX2 / 60.0))
                    : (std::sin(AX) * (15.0 - (420.0 - 945.0 / AX2) / AX2) / AX -
                       std::cos(AX) * (1.0 - (105.0 - 945.0 / AX2) / AX2)) /
                          AX;
    } else {
      JL = AX < 1.0 ? (AX2 * AX2 * AX2) / 135135.0 * (1.0 - AX2 / 30.0 * (1.0 - AX2 / 68.0))
                    : (std::sin(AX) * (-1.0 + (210.0 - (4725.0 - 10395.0 / AX2) / AX2) / AX2) +
                       std::cos(AX) * (-21.0 + (1260.0 - 10395.0 / AX2) / AX2) / AX) /
                          AX;
    }
  } else {
    const double NU = 0.5 + L, NU2 = NU * NU;
    if (AX < 1e-40) {
      JL = 0.0;
    } else if ((AX2 / L) < 5e-1) {
      JL = std::exp(L * std::log(AX / NU) - LN2 + NU * ONEMLN2 - (1.0 - (1.0 - 3.5 / NU2) / NU2 / 30.0) / 12.0 / NU) / NU *
           (1.0 - AX2 / (4.0 * NU + 4.0) * (1.0 - AX2 / (8.0 * NU + 16.0) * (1.0 - AX2 / (12.0 * NU + 36.0))));
    } else if (((double)L * (double)L / AX) < 5e-1) {
      const double BETA = AX - PID2 * (L + 1);
      JL = (std::cos(BETA) * (1.0 - (NU2 - 0.25) * (NU2 - 2.25) / 8.0 / AX2 * (1.0 - (NU2 - 6.25) * (NU2 - 12.25) / 48.0 / AX2)) -
            std::sin(BETA) * (NU2 - 0.25) / 2.0 / AX *
                (1.0 - (NU2 - 2.25) * (NU2 - 6.25) / 24.0 / AX2 * (1.0 - (NU2 - 12.25) * (NU2 - 20.25) / 80.0 / AX2))) /
           AX;
    } else {
      const double L3 = std::pow(NU, SP(0.325));
      if (AX < NU - SP(1.31) * L3) {
        const double COSB = NU / AX, SX = std::sqrt(NU2 - AX2), COTB = NU / SX, SECB = AX / NU;
        const double BETA = std::log(COSB + SX / AX);
        const double COT3B = COTB * COTB * COTB, COT6B = COT3B * COT3B, SEC2B = SECB * SECB;
        const double EXPTERM =
            ((2.0 + 3.0 * SEC2B) * COT3B / 24.0 -
             ((4.0 + SEC2B) * SEC2B * COT6B / 16.0 +
              ((16.0 - (1512.0 + (3654.0 + 375.0 * SEC2B) * SEC2B) * SEC2B) * COT3B / 5760.0 +
               (32.0 + (288.0 + (232.0 + 13.0 * SEC2B) * SEC2B) * SEC2B) * SEC2B * COT6B / 128.0 / NU) *
                  COT6B / NU) /
                 NU) /
            NU;
        JL = std::sqrt(COTB * COSB) / (2.0 * NU) * std::exp(-NU * BETA + NU / COTB - EXPTERM);
      } else if (AX > NU + SP(1.48) * L3) {
        const double COSB = NU / AX, SX = std::sqrt(AX2 - NU2), COTB = NU / SX, SECB = AX / NU;
        const double BETA = std::acos(COSB);
        const double COT3B = COTB * COTB * COTB, COT6B = COT3B * COT3B, SEC2B = SECB * SECB;
        const double TRIGARG =
            NU / COTB - NU * BETA - PID4 -
            ((2.0 + 3.0 * SEC2B) * COT3B / 24.0 +
             (16.0 - (1512.0 + (3654.0 + 375.0 * SEC2B) * SEC2B) * SEC2B) * COT3B * COT6B / 5760.0 / NU2) /
                NU;
        const double EXPTERM =
            ((4.0 + SEC2B) * SEC2B * COT6B / 16.0 -
             (32.0 + (288.0 + (232.0 + 13.0 * SEC2B) * SEC2B) * SEC2B) * SEC2B * (COT6B * COT6B) / 128.0 / NU2) /
            NU2;
        JL = std::sqrt(COTB * COSB) / NU * std::exp(-EXPTERM) * std::cos(TRIGARG);
      } else {
        const double BETA = AX - NU, BETA2 = BETA * BETA, SX = 6.0 / AX, SX2 = SX * SX;
        const double SECB = std::pow(SX, 0.3333333333333333), SEC2B = SECB * SECB;
        JL = (GAMMA1 * SECB + BETA * GAMMA2 * SEC2B - (BETA2 / 18.0 - 1.0 / 45.0) * BETA * SX * SECB * GAMMA1 -
              ((BETA2 - 1.0) * BETA2 / 36.0 + 1.0 / 420.0) * SX * SEC2B * GAMMA2 +
              (((BETA2 / 1620.0 - 7.0 / 3240.0) * BETA2 + 1.0 / 648.0) * BETA2 - 1.0 / 8100.0) * SX2 * SECB * GAMMA1 +
              (((BETA2 / 4536.0 - 1.0 / 810.0) * BETA2 + 19.0 / 11340.0) * BETA2 - 13.0 / 28350.0) * BETA * SX2 * SEC2B *
                  GAMMA2 -
              ((((BETA2 / 349920.0 - 1.0 / 29160.0) * BETA2 + 71.0 / 583200.0) * BETA2 - 121.0 / 874800.0) * BETA2 +
               7939.0 / 224532000.0) *
                  BETA * SX2 * SX * SECB * GAMMA1) *
             std::sqrt(SX) / ROOTPI12;
      }
    }
  }
  if (X < 0 && (L % 2) != 0) JL = -JL;
}

// camb/bessels.f90:50-120
void GenerateBessels(Model& M) {
  const Params& P = M.P;
  M.kmaxfile = (int)std::fmin(M.Max_eta_k, M.max_bessels_etak) + 1;
  Regions& B = M.BessRanges;
  Ranges_Init(B);
  Ranges_Add_delta(B, 0.0, 1.0, 0.01);
  Ranges_Add_delta(B, 1.0, 5.0, 0.1);
  Ranges_Add_delta(B, 5.0, 25.0, 0.2);
  Ranges_Add_delta(B, 25.0, 150.0, 0.5 / P.AccuracyBoost);
  Ranges_Add_delta(B, 150.0, (double)M.kmaxfile, 0.8 / P.AccuracyBoost);
  Ranges_GetArray(B, false);
  M.num_xx = B.npoints;
  const int nx = M.num_xx;
  const int max_ix = std::min(M.max_bessels_l_index, M.lSamp.l0);
  M.ajl.assign((size_t)nx * max_ix, 0.0);
  M.ajlpr.assign((size_t)nx * max_ix, 0.0);
  const int nth = P.nthreads > 0 ? P.nthreads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic) num_threads(nth)
  for (int j = 1; j <= max_ix; j++) {
    double* aj = M.ajl.data() + (size_t)(j - 1) * nx - 1;  // 1-based view
    double* ajp = M.ajlpr.data() + (size_t)(j - 1) * nx - 1;
    const int l = M.lSamp.l[j];
    for (int i = 1; i <= nx; i++) {
      const double x = B.points[i];
      double xlim = xlimfrac * l;
      xlim = std::fmax(xlim, xlimmin);
      xlim = l - xlim;
      if (x > xlim) {
        if ((l == 3 && x <= SP(0.2)) || (l > 3 && x < 0.5) || (l > 5 && x < 1.0))
          aj[i] = 0;
        else
          bjl(l, x, aj[i]);
      } else {
        aj[i] = 0;
      }
    }
    spline(B.points.data(), aj, nx, spl_large, spl_large, ajp);
  }
}

// camb/cmbmain.f90:1281-1353
void SetkValuesForInt(Model& M) {
  const Params& P = M.P;
  const double qmax_int = std::fmin(M.qmax, M.max_bessels_etak / M.tau0);
  const double IntSampleBoost = P.AccuracyBoost;
  Regions& Q = M.q_int;
  Ranges_Init(Q);
  const int lognum = (int)nint(10 * IntSampleBoost);
  const double dlnk1 = 1.0 / lognum;
  const int no = (int)nint(600 * IntSampleBoost);
  const double dk0 = 1.8 / M.r / M.chi0 / IntSampleBoost;
  double dk = 3.0 / M.r / M.chi0 / IntSampleBoost;
  if (P.HighAccuracyDefault) dk = dk / SP(1.6);
  const double k_max_log = lognum * dk0;
  const double k_max_0 = no * dk0;
  const double dk2 = SP(0.04) / IntSampleBoost;
  Ranges_Add_delta(Q, M.qmin, k_max_log, dlnk1, true);
  Ranges_Add_delta(Q, k_max_log, std::fmin(qmax_int, k_max_0), dk0);
  if (qmax_int > k_max_0) {
    const double max_k_dk = std::max(3000, 2 * M.maximum_l) / M.tau0;
    Ranges_Add_delta(Q, k_max_0, std::fmin(qmax_int, max_k_dk), dk);
    if (qmax_int > max_k_dk) Ranges_Add_delta(Q, max_k_dk, qmax_int, dk2);
  }
  Ranges_GetArray(Q, true);  // Init_ClTransfer, camb/modules.f90:1164-1177
  const int nl = std::min(M.max_bessels_l_index, M.lSamp.l0);
  M.Delta_p_l_k.assign((size_t)M.SourceNum * nl * Q.npoints, 0.0);
}

// camb/cmbmain.f90:483-507 (SourceToTransfers) = InterpolateSources :1355-1434 +
// DoSourceIntegration :1447-1481 + DoFlatIntegration :1500-1597, for one integration q.
static void SourceToTransfers(Model& M, int q_ix, std::vector<double>& Source_q, std::vector<double>& aa,
                              std::vector<double>& fac, std::vector<int>& bes_index, double& n_iter) {
  const Params& P = M.P;
  const int nk = M.Evolve_q.npoints, nt = M.TimeSteps.npoints, S = M.SourceNum;
  const std::vector<double>& kq = M.Evolve_q.points;
  const std::vector<double>& tp = M.TimeSteps.points;
  const double q = M.q_int.points[q_ix];
#define SQ(i, s) Source_q[(size_t)((s)-1) * (nt + 1) + (i)]
  // IntegrationVars_init
  for (int s = 1; s <= S; s++) {
    SQ(1, s) = 0;
    SQ(nt, s) = 0;
    SQ(nt - 1, s) = 0;
  }
  // InterpolateSources
  int klo = 1;
  while ((q > kq[klo + 1]) && (klo < (nk - 1))) klo = klo + 1;
  const int khi = klo + 1;
  const double ho = kq[khi] - kq[klo];
  const double a0 = (kq[khi] - q) / ho, b0 = (q - kq[klo]) / ho;
  const double ho2o6 = (ho * ho) / 6;
  const double a03 = (a0 * a0 * a0 - a0), b03 = (b0 * b0 * b0 - b0);
  int step = 2;
  for (int i = 2; i <= nt; i++) {
    const double xf = q * (M.tau0 - tp[i]);
    if ((M.WantLateTime || q * tp[i] < M.max_etak_scalar) && xf > 1.e-8) {
      step = i;
      for (int s = 1; s <= S; s++) {
        const size_t off = (size_t)nk * ((s - 1) + (size_t)S * (i - 1));
        SQ(i, s) = a0 * M.Src[off + klo - 1] + b0 * M.Src[off + khi - 1] +
                   (a03 * M.ddSrc[off + klo - 1] + b03 * M.ddSrc[off + khi - 1]) * ho2o6;
      }
    } else {
      for (int s = 1; s <= S; s++) SQ(i, s) = 0;
    }
  }
  const int SourceSteps = step;
  // DoSourceIntegration (flat)
  const double nu = q * M.r;
  int llmax = (int)nint(nu * M.chi0);
  if (llmax < 15)
    llmax = 17;
  else
    llmax = (int)nint(nu * (M.tau0 / M.r + 6 * pi_aml / nu));
  // DoFlatIntegration
  const Regions& B = M.BessRanges;
  for (int j = 1; j <= SourceSteps; j++) {
    const double xf = std::fabs(q * (M.tau0 - tp[j]));
    bes_index[j] = Ranges_IndexOf(B, xf);
    const int bes_ix = bes_index[j];
    fac[j] = B.points[bes_ix + 1] - B.points[bes_ix];
    aa[j] = (B.points[bes_ix + 1] - xf) / fac[j];
    fac[j] = (fac[j] * fac[j]) * aa[j] / 6;
  }
  const int nx = M.num_xx;
  const int nl = std::min(M.max_bessels_l_index, M.lSamp.l0);
  for (int j = 1; j <= nl; j++) {
    const int l = M.lSamp.l[j];
    if (l > llmax) return;
    double xlim = xlimfrac * l;
    xlim = std::fmax(xlim, xlimmin);
    xlim = l - xlim;
    const double xlmax1 = 80 * l * P.AccuracyBoost;
    double tmin = M.tau0 - xlmax1 / q;
    tmin = std::fmax(tp[2], tmin);
    double tmax = M.tau0 - xlim / q;
    tmax = std::fmin(M.tau0, tmax);
    tmin = std::fmax(tp[2], tmin);
    if (tmax < tp[2]) break;
    double sums[4] = {0, 0, 0, 0};
    const double* aj = M.ajl.data() + (size_t)(j - 1) * nx - 1;
    const double* ajp = M.ajlpr.data() + (size_t)(j - 1) * nx - 1;
    const int n_lo = Ranges_IndexOf(M.TimeSteps, tmin);
    const int n_hi = std::min(SourceSteps, Ranges_IndexOf(M.TimeSteps, tmax));
    if (S == 2) {
      for (int n = n_lo; n <= n_hi; n++) {
        const double a2 = aa[n];
        const int bes_ix = bes_index[n];
        double J_l = a2 * aj[bes_ix] +
                     (1 - a2) * (aj[bes_ix + 1] - ((a2 + 1) * ajp[bes_ix] + (2 - a2) * ajp[bes_ix + 1]) * fac[n]);
        J_l = J_l * M.TimeSteps.dpoints[n];
        sums[1] = sums[1] + SQ(n, 1) * J_l;
        sums[2] = sums[2] + SQ(n, 2) * J_l;
      }
      if (n_hi >= n_lo) n_iter += (n_hi - n_lo + 1);
    } else {
      double qmax_int = std::max(850, l) * 3 * P.AccuracyBoost / M.tau0;
      if (P.HighAccuracyDefault) qmax_int = qmax_int * SP(1.2);
      const bool DoInt = q < qmax_int;
      if (DoInt) {
        for (int n = n_lo; n <= n_hi; n++) {
          const double a2 = aa[n];
          const int bes_ix = bes_index[n];
          double J_l = a2 * aj[bes_ix] +
                       (1 - a2) * (aj[bes_ix + 1] - ((a2 + 1) * ajp[bes_ix] + (2 - a2) * ajp[bes_ix + 1]) * fac[n]);
          J_l = J_l * M.TimeSteps.dpoints[n];
          sums[1] = sums[1] + SQ(n, 1) * J_l;
          sums[2] = sums[2] + SQ(n, 2) * J_l;
          sums[3] = sums[3] + SQ(n, 3) * J_l;
        }
        if (n_hi >= n_lo) n_iter += (n_hi - n_lo + 1);
      }
      const bool UseLimber = l > 400 * std::pow(P.AccuracyBoost, (double)0.5f);
      if (!DoInt || UseLimber) {
        double xf = M.tau0 - (l + 0.5) / q;
        if (xf < M.TimeSteps.Highest && xf > M.TimeSteps.Lowest) {
          const int n = Ranges_IndexOf(M.TimeSteps, xf);
          xf = (xf - tp[n]) / (tp[n + 1] - tp[n]);
          sums[3] = (SQ(n, 3) * (1 - xf) + xf * SQ(n + 1, 3)) * std::sqrt(pi_aml / 2 / (l + 0.5)) / q;
        } else {
          sums[3] = 0;
        }
      }
    }
    double* D = M.Delta_p_l_k.data() + (size_t)S * ((j - 1) + (size_t)nl * (q_ix - 1));
    for (int s = 1; s <= S; s++) D[s - 1] = D[s - 1] + sums[s];
  }
#undef SQ
}

// hot loop 2, camb/cmbmain.f90:263-267
void IntegrateAll(Model& M) {
  const int nq = M.q_int.npoints, nt = M.TimeSteps.npoints, S = M.SourceNum;
  const int nth = M.P.nthreads > 0 ? M.P.nthreads : omp_get_max_threads();
  double n_iter_total = 0;
#pragma omp parallel num_threads(nth) reduction(+ : n_iter_total)
  {
    std::vector<double> Source_q((size_t)(nt + 1) * S + 1, 0.0), aa(nt + 2), fac(nt + 2);
    std::vector<int> bes_index(nt + 2);
#pragma omp for schedule(static, 4)
    for (int q_ix = 1; q_ix <= nq; q_ix++) SourceToTransfers(M, q_ix, Source_q, aa, fac, bes_index, n_iter_total);
  }
  M.cnt.n_iter = n_iter_total;
}

// camb/power_tilt.f90:149-157
double ScalarPower(const Model& M, double k) {
  const Params& P = M.P;
  const double lnrat = std::log(k / P.k_0_scalar);
  return P.ScalarPowerAmp * std::exp(lnrat * (P.an - 1 + lnrat * (P.n_run / 2 + P.n_runrun / 6 * lnrat)));
}

// camb/cmbmain.f90:2167-2267 (flat; no 2D array; limber_phiphi = 0)
void CalcScalCls(Model& M) {
  const int nq = M.q_int.npoints, l0 = M.lSamp.l0, S = M.SourceNum;
  const int nl = std::min(M.max_bessels_l_index, l0);
  M.iCl_scalar.assign((size_t)l0 * M.C_last, 0.0);
#define ICL(j, t) M.iCl_scalar[(size_t)((t)-1) * l0 + ((j)-1)]
  std::vector<double> dlnks(nq + 1), pows(nq + 1);
  for (int q_ix = 1; q_ix <= nq; q_ix++) {
    const double ks = M.q_int.points[q_ix];
    dlnks[q_ix] = M.q_int.dpoints[q_ix] / M.q_int.points[q_ix];
    pows[q_ix] = ScalarPower(M, ks);
  }
  const double twopi = 2 * pi_aml, fourpi = 4 * pi_aml;
  for (int j = 1; j <= l0; j++) {
    const double ell = (double)M.lSamp.l[j];
    if (j <= nl) {
      for (int q_ix = 1; q_ix <= nq; q_ix++) {
        const double dlnk = dlnks[q_ix], apowers = pows[q_ix];
        const double* D = M.Delta_p_l_k.data() + (size_t)S * ((j - 1) + (size_t)nl * (q_ix - 1));
        ICL(j, 1) = ICL(j, 1) + apowers * (D[0] * D[0]) * dlnk;
        ICL(j, 2) = ICL(j, 2) + apowers * (D[1] * D[1]) * dlnk;
        ICL(j, 3) = ICL(j, 3) + apowers * D[0] * D[1] * dlnk;
        if (S > 2) {
          ICL(j, 4) = ICL(j, 4) + apowers * (D[2] * D[2]) * dlnk;
          ICL(j, 5) = ICL(j, 5) + apowers * D[2] * D[0] * dlnk;
          ICL(j, 6) = ICL(j, 6) + apowers * D[2] * D[1] * dlnk;
        }
      }
    }
    const double ctnorm = (ell * ell - 1) * (ell + 2) * ell;
    const double dbletmp = (ell * (ell + 1)) / twopi * fourpi;
    ICL(j, 1) = ICL(j, 1) * dbletmp;
    ICL(j, 2) = ICL(j, 2) * dbletmp * ctnorm;
    ICL(j, 3) = ICL(j, 3) * dbletmp * std::sqrt(ctnorm);
    if (S > 2) {
      const double ALens = M.P.ALens;  // camb/cmbmain.f90:122
      ICL(j, 4) = ALens * ICL(j, 4) * fourpi * (ell * ell * ell * ell);
      ICL(j, 5) = std::sqrt(ALens) * ICL(j, 5) * fourpi * (ell * ell * ell);
      ICL(j, 6) = std::sqrt(ALens) * ICL(j, 6) * fourpi * (ell * ell * ell) * std::sqrt(ctnorm);
    }
  }
#undef ICL
}

// camb/modules.f90:1012-1044   (iCl 1-based over samples; all_Cl indexed by l-lmin)
static void InterpolateClArr(const lSamples& lSet, const double* iCl, double* all_Cl, int max_ind) {
  std::vector<double> ddCl(lSet.l0 + 1), xl(lSet.l0 + 1);
  for (int i = 1; i <= lSet.l0; i++) xl[i] = (double)lSet.l[i];
  spline(xl.data(), iCl, max_ind, 1.e30, 1.e30, ddCl.data());
  int llo = 1;
  for (int il = lmin; il <= lSet.l[max_ind]; il++) {
    const int xi = il;
    if ((xi > lSet.l[llo + 1]) && (llo < max_ind)) llo = llo + 1;
    const int lhi = llo + 1;
    const double ho = lSet.l[lhi] - lSet.l[llo];
    const double a0 = (lSet.l[lhi] - xi) / ho;
    const double b0 = (xi - lSet.l[llo]) / ho;
    all_Cl[il - lmin] = a0 * iCl[llo] + b0 * iCl[lhi] +
                        ((a0 * a0 * a0 - a0) * ddCl[llo] + (b0 * b0 * b0 - b0) * ddCl[lhi]) * (ho * ho) / 6;
  }
}

// camb/modules.f90:1046-1089
static void InterpolateClArrTemplated(const Model& M, const double* iCl, double* all_Cl, int max_ind,
                                      int template_index) {
  const lSamples& lSet = M.lSamp;
  if (M.P.use_spline_template && template_index <= 3 && !M.highL_template.empty()) {
    int maxdelta = max_ind;
    while (lSet.l[maxdelta] > lmax_extrap_highl) maxdelta = maxdelta - 1;
    const double* tmpl = M.highL_template.data() + (size_t)(lmax_extrap_highl - lmin + 1) * (template_index - 1);
    std::vector<double> DeltaCL(lSet.l0 + 1, 0.0);
    for (int i = 1; i <= maxdelta; i++) DeltaCL[i] = iCl[i] - tmpl[lSet.l[i] - lmin];
    InterpolateClArr(lSet, DeltaCL.data(), all_Cl, maxdelta);
    for (int il = lmin; il <= lSet.l[maxdelta]; il++) all_Cl[il - lmin] = all_Cl[il - lmin] + tmpl[il - lmin];
    if (maxdelta < max_ind) {
      std::vector<double> tmpall(lSet.l[max_ind] - lmin + 1);
      InterpolateClArr(lSet, iCl, tmpall.data(), max_ind);
      for (int il = lSet.l[maxdelta - 2]; il <= lSet.l[max_ind]; il++) all_Cl[il - lmin] = tmpall[il - lmin];
    }
    return;
  }
  InterpolateClArr(lSet, iCl, all_Cl, max_ind);
}

// camb/cmbmain.f90:2450-2496 (scalars)
void InterpolateCls(Model& M) {
  const int l0 = M.lSamp.l0, nL = M.P.Max_l - lmin + 1;
  M.Cl_scalar.assign((size_t)nL * M.C_last, 0.0);
  for (int i = 1; i <= M.C_last; i++)
    InterpolateClArrTemplated(M, M.iCl_scalar.data() + (size_t)(i - 1) * l0 - 1, M.Cl_scalar.data() + (size_t)(i - 1) * nL,
                              l0, i);
}

// camb/modules.f90:1222-1245.  Accepts the reference text format (L + 7 columns) or the
// repo's binary fixture (float64[7999][4] = TT,EE,TE,PP, little-endian; tests/golden).
int load_highL_template(Model& M, const char* path) {
  const int n = lmax_extrap_highl - lmin + 1;
  M.highL_template.assign((size_t)n * 4, 0.0);
  FILE* f = fopen(path, "rb");
  if (!f) return 1;
  const size_t len = strlen(path);
  int rc = 0;
  if (len > 4 && strcmp(path + len - 4, ".f64") == 0) {
    std::vector<double> buf((size_t)n * 4);
    if (fread(buf.data(), sizeof(double), buf.size(), f) != buf.size()) rc = 2;
    for (int l = 0; l < n; l++)
      for (int t = 0; t < 4; t++) M.highL_template[(size_t)t * n + l] = buf[(size_t)l * 4 + t];
  } else {
    int L = 0;
    double a[7];
    while (fscanf(f, "%d %lf %lf %lf %lf %lf %lf %lf", &L, &a[0], &a[1], &a[2], &a[3], &a[4], &a[5], &a[6]) == 8) {
      if (L > lmax_extrap_highl) break;
      M.highL_template[(size_t)0 * n + (L - lmin)] = a[0];
      M.highL_template[(size_t)1 * n + (L - lmin)] = a[1];
      M.highL_template[(size_t)2 * n + (L - lmin)] = a[3];
      M.highL_template[(size_t)3 * n + (L - lmin)] = a[4];
    }
    if (L < lmax_extrap_highl) rc = 2;
  }
  fclose(f);
  if (rc) M.highL_template.clear();
  return rc;
}

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// CAMB_GetResults -> CAMBParams_Set + cmbmain (camb/camb.f90:106-145, camb/cmbmain.f90:131-291).
// stage_times (optional, 8 doubles): params_set, initvars, sources, spline_k, bessels, los, cls, total.
int run_model(Model& M, double* stage_times, bool reuse_bessels) {
  double t0 = now_s(), t = t0, tn;
  double st[8] = {0};
  if (CAMBParams_Set(M)) return M.error_flag;
  tn = now_s(); st[0] = tn - t; t = tn;
  M.WantLateTime = M.P.DoLensing != 0;
  M.maximum_l = M.P.Max_l;
  M.maximum_qeta = M.Max_eta_k;
  initlval(M, M.maximum_l);
  cmb_InitVars(M);
  if (M.error_flag) return M.error_flag;
  tn = now_s(); st[1] = tn - t; t = tn;
  SetkValuesForSources(M);
  if (M.P.WantTransfer) InitTransfer(M);
  if (M.WantLateTime) {
    M.SourceNum = 3;
    M.C_last = 6;
  } else {
    M.SourceNum = 2;
    M.C_last = 3;
  }
  CalcAllSources(M);
  if (M.error_flag) return M.error_flag;
  if (M.P.WantTransfer) {  // camb/cmbmain.f90:217-237
    TransferOut(M);
    if (M.error_flag) return M.error_flag;
    Transfer_Get_sigmas(M);
  }
  tn = now_s(); st[2] = tn - t; t = tn;
  InitSourceInterpolation(M);
  tn = now_s(); st[3] = tn - t; t = tn;
  M.max_bessels_l_index = M.lSamp.l0;
  M.max_bessels_etak = M.maximum_qeta;
  // InitSpherBessels cache test, camb/bessels.f90:34-48
  if (!(reuse_bessels && !M.ajl.empty() && (int)std::fmin(M.max_bessels_etak, M.Max_eta_k) + 1 <= M.kmaxfile &&
        (size_t)M.num_xx * M.lSamp.l0 == M.ajl.size()))
    GenerateBessels(M);
  tn = now_s(); st[4] = tn - t; t = tn;
  SetkValuesForInt(M);
  IntegrateAll(M);
  tn = now_s(); st[5] = tn - t; t = tn;
  CalcScalCls(M);
  InterpolateCls(M);
  tn = now_s(); st[6] = tn - t; t = tn;
  st[7] = tn - t0;
  if (stage_times)
    for (int i = 0; i < 8; i++) stage_times[i] = st[i];
  return 0;
}

int run_all(Model& M) { return run_model(M, nullptr, false); }

}  // namespace orc
