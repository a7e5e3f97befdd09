// C ABI of libgalcamb_b200.so (see include/galcamb.h): context, model upload, stage launches, results.
// Host-side helpers here are product code: table packing into the AoS device layout (model.h), the natural
// cubic spline of the Galileon background (what gsl_spline_init does inside arrays_, camb/galileon.cc:666-682),
// the running minimum used by Thermo_OpacityToTime, the BessRanges grid of camb/bessels.f90:68-76 and the
// work-item ordering of the per-k kernel.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/galcamb.h"
#include "model.h"
#include "pack.h"

// packed tables of a model view handed out by the pre-stage (csrc/prestage.cu), or nullptr
extern "C" const double* gc_prestage_prepacked(const galcamb_model* m, size_t* n, gcpack::Info* info);
extern "C" void gc_upload_constants();
extern "C" cudaError_t gc_launch_evolve(const DevModel*, const int2*, int, int, int, unsigned*, int*, unsigned long long*, cudaStream_t);
extern "C" int gc_evolve_warps_per_cta(int);
extern "C" void gc_upload_constants_batch();
extern "C" cudaError_t gc_launch_evolve_batch(const DevModel*, const int2*, int, double*, int, int*, unsigned long long*, cudaStream_t);
// the same two kernels compiled with the matter-transfer taps (-DGC_WITH_TRANSFER): used when a model of the batch has WantTransfer
extern "C" void gc_upload_constants_tr();
extern "C" cudaError_t gc_launch_evolve_tr(const DevModel*, const int2*, int, int, int, unsigned*, int*, unsigned long long*, cudaStream_t);
extern "C" int gc_evolve_warps_per_cta_tr(int);
extern "C" void gc_upload_constants_batch_tr();
extern "C" cudaError_t gc_launch_evolve_batch_tr(const DevModel*, const int2*, int, double*, int, int*, unsigned long long*, cudaStream_t);
// ... and the lane-per-item kernel with its lane records sized for one massive-neutrino eigenstate (-DGC_EV_NU=1): used
// when no model of the batch has more
extern "C" void gc_upload_constants_batch_nu1();
extern "C" cudaError_t gc_launch_evolve_batch_nu1(const DevModel*, const int2*, int, double*, int, int*, unsigned long long*, cudaStream_t);
extern "C" cudaError_t gc_set_bess_regions(const GcRegions*);
extern "C" cudaError_t gc_launch_bessels(double2*, const double*, const int*, double*, int, int, cudaStream_t);
extern "C" cudaError_t gc_launch_spline_k(const DevModel*, int, int, cudaStream_t);
extern "C" cudaError_t gc_launch_los(const DevModel*, int, int, int, int, const double2*, const double*, const int*, int, int,
                                     unsigned long long*, int, cudaStream_t);
extern "C" cudaError_t gc_launch_zero_delta(const DevModel*, int, int, cudaStream_t);
extern "C" cudaError_t gc_launch_cls(const DevModel*, int, const int*, int, int, int, const double*, int, int, cudaStream_t);
extern "C" cudaError_t gc_launch_lens_ltab(double4*, double4*, int, cudaStream_t);
extern "C" cudaError_t gc_launch_lensing(const DevModel*, int, const int*, double, const double4*, const double4*, const double*,
                                         const double*, int, double4*, const float*, double*, int*, double*, cudaStream_t);

namespace {

struct Slot {
  bool set = false;
  DevModel host{};          // device pointers filled in
  double* d_tables = nullptr;  // one allocation: th | pmin | ts | gal | k | q | dq
  size_t tables_cap = 0;
  double* d_out = nullptr;     // Src | ddSrc | Delta | iCl | Cl
  size_t out_cap = 0;
  int lmax = 0, ntype = 3, status = 0;
  double* staging = nullptr;  // pinned host buffer (packed AoS tables)
  size_t staging_cap = 0;
  size_t total = 0, n_th = 0, n_pm = 0, n_ts = 0, n_gal = 0;
  int m_nk = 0, m_nq = 0, m_lmax = 0, m_nk_tr = 0;
  int ensure_staging(size_t n) {
    if (n <= staging_cap) return 0;
    if (staging) cudaFreeHost(staging);
    staging = nullptr;
    staging_cap = 0;
    if (cudaMallocHost((void**)&staging, n * sizeof(double)) != cudaSuccess) return 1;
    staging_cap = n;
    return 0;
  }
};

struct Ctx {
  bool ready = false;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[10]{};
  cudaEvent_t ev_block = nullptr;  // blocking-sync event: the host thread sleeps through the long K1 launch instead of spinning
  // Bessel tables
  int nl = 0, nx = 0, kmaxfile = 0;
  double bess_boost = 0;
  std::vector<int> ls;
  std::vector<double> xs;
  int* d_ls = nullptr;
  double* d_xs = nullptr;
  double2* d_bess = nullptr;
  // template
  double* d_tmpl = nullptr;
  int lmax_tmpl = 0;
  double* d_tmpl_pp = nullptr;  // lensing potential column of the template
  int lmax_tmpl_pp = 0;
  // lensing work space
  double4 *d_lens_ta = nullptr, *d_lens_tb = nullptr, *d_lens_cin = nullptr;
  float* d_lens_apod = nullptr;
  double *d_lens_partial = nullptr, *d_lensed = nullptr;
  int* d_lens_status = nullptr;
  size_t lens_tab_cap = 0, lens_tabb_cap = 0, lens_cin_cap = 0, lens_partial_cap = 0, lensed_cap = 0, lens_status_cap = 0, lens_apod_cap = 0;
  int lens_tab_lmax = -1, lmax_lensed = 0, lens_nm = 0;
  std::vector<int> lens_status;
  // neutrino tables (cosmology independent; uploaded from the first model that has them)
  double* d_nu_tab = nullptr;
  bool nu_tab_set = false;
  // batch
  std::vector<Slot> slots;  // the slots of the current batch
  std::vector<Slot> spare;  // slots of earlier, larger batches: their device tables / pinned staging are reused, not leaked
  DevModel* d_models = nullptr;
  size_t models_cap = 0;
  int2* d_items = nullptr;
  size_t items_cap = 0;
  int* d_status = nullptr;
  unsigned* d_queue = nullptr;  // work-queue head of the per-k kernel (eight-lane form)
  double* d_ws = nullptr;       // integrator workspace of the lane-per-item form (large batches)
  size_t ws_cap = 0;
  int k1_form = 0;              // which form of K1 the last call used: 0 = eight lanes per item, 1 = lane per item
  // one spectrum shared by several devices (galcamb_multi_single_cls_): this device takes the source wavenumbers
  // kk % kshard_world == kshard_rank, the integration wavenumbers [q_lo, q_hi), and only the sums of the C_l stage
  int kshard_rank = 0, kshard_world = 1;
  int q_lo = 0, q_hi = -1;      // q_hi < 0: the whole grid
  int cls_what = 3;             // gc_launch_cls: 1 = sums over q, 2 = interpolation in l, 3 = both
  unsigned long tmpl_version = 0;
  int num_sms = 0;
  unsigned long long* d_counters = nullptr;  // [0..2] evolve, [3] n_iter
  std::vector<int2> items;
  double ms[8] = {0};
  double host_ms[8] = {0};  // wall clock of the host phases of the last batch call (galcamb_get_host_timings_)
  double counters[4] = {0};
  double evolve_stats[3] = {0};
  std::string err;
};

// One context per device.  A host thread works on the context of the device it last named in galcamb_init_ /
// galcamb_select_device_ (thread-local), so several host threads -- OpenMP threads of a Fortran host, the workers of
// the galcamb_multi_* entry points -- can drive one GPU each from the same process.  Calls on one context are not
// re-entrant (like CAMB's own module state, camb/cmbmain.f90:7-8).
std::mutex g_ctx_mutex;
std::map<int, Ctx*> g_ctxs;
Ctx g_none;  // what a thread sees before it has named a device: never ready
thread_local Ctx* tl_ctx = nullptr;
inline Ctx& cur_ctx() { return tl_ctx ? *tl_ctx : g_none; }
#define g (cur_ctx())

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      g.err = std::string(#call) + ": " + cudaGetErrorString(e_);                                  \
      return 100;                                                                                  \
    }                                                                                              \
  } while (0)

template <class T>
int grow(T*& p, size_t& cap, size_t n) {
  if (n <= cap) return 0;
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
  if (cudaMalloc((void**)&p, n * sizeof(T)) != cudaSuccess) return 1;
  cap = n;
  return 0;
}

std::mutex g_err_mutex;  // pack_model's workers share their caller's context and may fail at the same time
int fail(int code, const char* msg) {
  std::lock_guard<std::mutex> lk(g_err_mutex);
  g.err = msg;
  return code;
}

// BessRanges of camb/bessels.f90:68-76: five adjacent, non-overlapping regions, so the general merge rules of
// Ranges_Add (camb/utils.F90:206-466) reduce to "n = max(1,int(span/delta + 1 - 0.1)), step = span/n" per region.
void build_bess_ranges(int kmaxfile, double boost, GcRegions& R, std::vector<double>& xs) {
  const double lo[5] = {0, 1, 5, 25, 150}, hi[5] = {1, 5, 25, 150, (double)kmaxfile};
  const double dl[5] = {0.01, 0.1, 0.2, 0.5 / boost, 0.8 / boost};
  R.count = 0;
  int start = 1;
  xs.clear();
  for (int r = 0; r < 5; r++) {
    if (hi[r] <= lo[r]) break;
    const double span = hi[r] - lo[r];
    const int n = std::max(1, (int)(span / dl[r] + 1.0 - 0.1));
    const double d0 = span / n;                       // Ranges_Add: delta = (t_end - t_start)/nstep
    const int steps = std::max(1, (int)(span / d0 + 1.0 - 0.1));
    GcRegion& A = R.R[R.count++];
    A.Low = lo[r];
    A.High = hi[r];
    A.delta = span / steps;
    A.start_index = start;
    A.IsLog = 0;
    for (int j = 0; j < steps; j++) xs.push_back(A.Low + A.delta * j);
    start += steps;
  }
  R.Lowest = 0;
  R.Highest = R.R[R.count - 1].High;
  xs.push_back(R.Highest);
  R.npoints = (int)xs.size();
}

int grow(double** p, size_t* cap, size_t need) {
  if (need <= *cap) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  if (cudaMalloc((void**)p, need * sizeof(double)) != cudaSuccess) return 1;
  *cap = need;
  return 0;
}

}  // namespace

extern "C" {

const char* galcamb_last_error_(void) { return g.err.c_str(); }
int galcamb_sizeof_model_(void) { return (int)sizeof(galcamb_model); }

int galcamb_init_(const int* device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_none.err = "galcamb_init_: no CUDA device (this library has no CPU fallback)";
    return 100;
  }
  const int dev = device ? *device : 0;
  if (dev < 0 || dev >= ndev) {
    g_none.err = "galcamb_init_: no such device";
    return 5;
  }
  {  // the calling thread works on this device's context from now on
    std::lock_guard<std::mutex> lk(g_ctx_mutex);
    Ctx*& c = g_ctxs[dev];
    if (!c) c = new Ctx();
    tl_ctx = c;
  }
  if (g.ready) return 0;
  g.device = dev;
  CK(cudaSetDevice(g.device));
  CK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  for (auto& e : g.ev) CK(cudaEventCreate(&e));
  CK(cudaEventCreateWithFlags(&g.ev_block, cudaEventBlockingSync | cudaEventDisableTiming));
  gc_upload_constants();
  gc_upload_constants_batch();
  gc_upload_constants_tr();
  gc_upload_constants_batch_tr();
  gc_upload_constants_batch_nu1();
  CK(cudaMalloc((void**)&g.d_counters, 24 * sizeof(unsigned long long)));
  CK(cudaMalloc((void**)&g.d_queue, sizeof(unsigned)));
  CK(cudaDeviceGetAttribute(&g.num_sms, cudaDevAttrMultiProcessorCount, g.device));
  CK(cudaMalloc((void**)&g.d_nu_tab, (size_t)GC_NRHOPN * 6 * sizeof(double)));
  CK(cudaDeviceSetLimit(cudaLimitStackSize, 8192));
  g.ready = true;
  return 0;
}

// the calling thread switches to the (already initialised) context of another device
int galcamb_select_device_(const int* device) {
  std::lock_guard<std::mutex> lk(g_ctx_mutex);
  auto it = g_ctxs.find(*device);
  if (it == g_ctxs.end() || !it->second->ready) {
    g_none.err = "galcamb_select_device_: device not initialised";
    return 5;
  }
  tl_ctx = it->second;
  return 0;
}

void galcamb_free_(void) {
  if (!g.ready) return;
  cudaSetDevice(g.device);
  cudaStreamSynchronize(g.stream);
  for (auto* v : {&g.slots, &g.spare}) {
    for (auto& s : *v) {
      if (s.d_tables) cudaFree(s.d_tables);
      if (s.d_out) cudaFree(s.d_out);
      if (s.staging) cudaFreeHost(s.staging);
    }
    v->clear();
  }
  for (void* p : {(void*)g.d_ls, (void*)g.d_xs, (void*)g.d_bess, (void*)g.d_tmpl, (void*)g.d_nu_tab, (void*)g.d_models,
                  (void*)g.d_items, (void*)g.d_status, (void*)g.d_queue, (void*)g.d_ws, (void*)g.d_counters, (void*)g.d_tmpl_pp, (void*)g.d_lens_ta,
                  (void*)g.d_lens_tb, (void*)g.d_lens_cin, (void*)g.d_lens_apod, (void*)g.d_lens_partial, (void*)g.d_lensed,
                  (void*)g.d_lens_status})
    if (p) cudaFree(p);
  for (auto& e : g.ev) cudaEventDestroy(e);
  if (g.ev_block) cudaEventDestroy(g.ev_block);
  g.ev_block = nullptr;
  cudaStreamDestroy(g.stream);
  g = Ctx();  // the context object stays registered for its device; a later galcamb_init_ sets it up again
}

int galcamb_sync_(void) {
  if (!g.ready) return fail(100, "not initialised");
  CK(cudaStreamSynchronize(g.stream));
  return 0;
}

int galcamb_set_bessels_(const int* ls, const int* nl, const int* kmaxfile, const double* AccuracyBoost) {
  if (!g.ready) return fail(100, "not initialised");
  CK(cudaSetDevice(g.device));
  const int n = *nl;
  // cache test of InitSpherBessels, camb/bessels.f90:40-41
  if (g.d_bess && n <= g.nl && *kmaxfile <= g.kmaxfile && g.bess_boost == *AccuracyBoost &&
      std::equal(ls, ls + n, g.ls.begin()) && n == g.nl)
    return 0;
  CK(cudaEventRecord(g.ev[8], g.stream));
  GcRegions R{};
  build_bess_ranges(*kmaxfile, *AccuracyBoost, R, g.xs);
  g.ls.assign(ls, ls + n);
  g.nx = (int)g.xs.size();
  g.kmaxfile = *kmaxfile;
  g.bess_boost = *AccuracyBoost;
  for (void* p : {(void*)g.d_ls, (void*)g.d_xs, (void*)g.d_bess})
    if (p) cudaFree(p);
  g.d_ls = nullptr;
  g.d_xs = nullptr;
  g.d_bess = nullptr;
  const int old_nl = g.nl;
  g.nl = 0;  // a failure below leaves the context without Bessel tables, not with dangling ones
  if (n != old_nl)
    for (auto& sl : g.slots) sl.set = false;  // their Delta / iCl buffers were sized for the previous l sampling
  CK(cudaMalloc((void**)&g.d_ls, n * sizeof(int)));
  CK(cudaMalloc((void**)&g.d_xs, g.nx * sizeof(double)));
  CK(cudaMalloc((void**)&g.d_bess, (size_t)g.nx * n * sizeof(double2)));
  double* scratch = nullptr;
  CK(cudaMalloc((void**)&scratch, (size_t)g.nx * n * sizeof(double)));
  CK(cudaMemcpyAsync(g.d_ls, ls, n * sizeof(int), cudaMemcpyHostToDevice, g.stream));
  CK(cudaMemcpyAsync(g.d_xs, g.xs.data(), g.nx * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CK(gc_set_bess_regions(&R));
  CK(gc_launch_bessels(g.d_bess, g.d_xs, g.d_ls, scratch, g.nx, n, g.stream));
  CK(cudaEventRecord(g.ev[9], g.stream));
  CK(cudaStreamSynchronize(g.stream));
  cudaFree(scratch);
  g.nl = n;
  float ms = 0;
  cudaEventElapsedTime(&ms, g.ev[8], g.ev[9]);
  g.ms[6] = ms;
  return 0;
}

// debugging / tests: copy the Bessel x grid and tables back (ajl, ajlpr as [nl][nx])
int galcamb_get_bessels_(int* nx, double* xs, double* ajl, double* ajlpr) {
  if (!g.d_bess) return fail(100, "no Bessel table");
  *nx = g.nx;
  if (xs) std::copy(g.xs.begin(), g.xs.end(), xs);
  if (ajl || ajlpr) {
    std::vector<double2> h((size_t)g.nx * g.nl);
    CK(cudaMemcpy(h.data(), g.d_bess, h.size() * sizeof(double2), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < h.size(); i++) {
      if (ajl) ajl[i] = h[i].x;
      if (ajlpr) ajlpr[i] = h[i].y;
    }
  }
  return 0;
}

// host copies of the templates: the workers of galcamb_multi_* bring their device's context up to date from here
struct SharedTemplates {
  std::mutex mu;
  std::vector<double> tmpl, pp;
  int lmax_tmpl = 0, lmax_pp = 0;
  unsigned long version = 0;
} g_shared_tmpl;

int galcamb_set_template_(const double* tmpl, const int* lmax_tmpl) {
  if (!g.ready) return fail(100, "not initialised");
  {
    std::lock_guard<std::mutex> lk(g_shared_tmpl.mu);
    const bool on = tmpl && *lmax_tmpl >= 2;
    g_shared_tmpl.tmpl.assign(on ? tmpl : nullptr, on ? tmpl + (size_t)3 * (*lmax_tmpl - 1) : nullptr);
    g_shared_tmpl.lmax_tmpl = on ? *lmax_tmpl : 0;
    g_shared_tmpl.version++;
    g.tmpl_version = g_shared_tmpl.version;
  }
  CK(cudaSetDevice(g.device));
  if (g.d_tmpl) cudaFree(g.d_tmpl);
  g.d_tmpl = nullptr;
  g.lmax_tmpl = 0;
  if (!tmpl || *lmax_tmpl < 2) return 0;
  const size_t n = (size_t)3 * (*lmax_tmpl - 1);
  CK(cudaMalloc((void**)&g.d_tmpl, n * sizeof(double)));
  CK(cudaMemcpy(g.d_tmpl, tmpl, n * sizeof(double), cudaMemcpyHostToDevice));
  g.lmax_tmpl = *lmax_tmpl;
  return 0;
}

int galcamb_set_lensing_template_(const double* pp, const int* lmax_tmpl) {
  if (!g.ready) return fail(100, "not initialised");
  {
    std::lock_guard<std::mutex> lk(g_shared_tmpl.mu);
    const bool on = pp && *lmax_tmpl >= 2;
    g_shared_tmpl.pp.assign(on ? pp : nullptr, on ? pp + (size_t)(*lmax_tmpl - 1) : nullptr);
    g_shared_tmpl.lmax_pp = on ? *lmax_tmpl : 0;
    g_shared_tmpl.version++;
    g.tmpl_version = g_shared_tmpl.version;
  }
  CK(cudaSetDevice(g.device));
  if (g.d_tmpl_pp) cudaFree(g.d_tmpl_pp);
  g.d_tmpl_pp = nullptr;
  g.lmax_tmpl_pp = 0;
  if (!pp || *lmax_tmpl < 2) return 0;
  const size_t n = (size_t)(*lmax_tmpl - 1);
  CK(cudaMalloc((void**)&g.d_tmpl_pp, n * sizeof(double)));
  CK(cudaMemcpy(g.d_tmpl_pp, pp, n * sizeof(double), cudaMemcpyHostToDevice));
  g.lmax_tmpl_pp = *lmax_tmpl;
  return 0;
}

int galcamb_batch_begin_(const int* nmodels) {
  if (!g.ready) return fail(100, "not initialised");
  if (*nmodels < 1) return fail(5, "nmodels < 1");
  // only nmodels slots take part; slots a smaller batch does not use are parked (with their buffers), not dropped
  while ((int)g.slots.size() > *nmodels) {
    g.spare.push_back(g.slots.back());
    g.slots.pop_back();
  }
  while ((int)g.slots.size() < *nmodels) {
    if (!g.spare.empty()) {
      g.slots.push_back(g.spare.back());
      g.spare.pop_back();
    } else {
      g.slots.emplace_back();
    }
  }
  for (auto& s : g.slots) s.set = false;
  g.items.clear();
  return 0;
}

static int pack_model(int islot, const galcamb_model* m) {
  const int* slot = &islot;
  if (!g.ready) return fail(100, "not initialised");
  if (*slot < 0 || *slot >= (int)g.slots.size()) return fail(5, "bad slot");
  if (g.nl == 0) return fail(5, "galcamb_set_bessels_ must be called first");
  if (m->Nu_mass_eigenstates > GC_MAX_NU || m->nqmax > GC_MAX_NQ) return fail(5, "too many neutrino eigenstates / q nodes");
  if (m->nthermo < 2 || !(m->tauminn > 0) || !(m->dlntau > 0)) return fail(5, "empty or degenerate thermo table in galcamb_model");
  if (m->use_galileon && m->ngal < 2) return fail(5, "use_galileon needs a tabulated background (ngal >= 2)");
  if (m->Nu_mass_eigenstates > 0 && !(m->dlnam > 0)) return fail(5, "massive neutrinos need their tables (dlnam > 0)");
  if (m->n_tau_regions > GC_MAX_REGIONS) return fail(5, "too many TimeSteps regions");
  if (m->nk < 2 || m->nq < 1 || m->nt < 2 || m->lmax < 2) return fail(5, "empty k, q or time grid in galcamb_model");
  if (m->WantTransfer && (m->transfer_num_redshifts < 1 || m->transfer_num_redshifts > GC_MAX_TF || m->nk_transfer < 0 ||
                          (m->nk_transfer > 0 && !m->k_transfer) || !(m->H0 > 0)))
    return fail(5, "WantTransfer needs 1..16 transfer times, H0 and the transfer wavenumbers");
  for (int i = 1; m->WantTransfer && i < m->transfer_num_redshifts; i++)
    if (!(m->tautf[i] > m->tautf[i - 1])) return fail(5, "Transfer redshifts not set or out of order");  // camb/cmbmain.f90:787-790
  Slot& s = g.slots[*slot];
  DevModel& D = s.host;
  D = DevModel{};
  D.grhom = m->grhom; D.grhog = m->grhog; D.grhor = m->grhor; D.grhob = m->grhob; D.grhoc = m->grhoc; D.grhov = m->grhov;
  D.grhornomass = m->grhornomass; D.grhok = m->grhok; D.adotrad = m->adotrad; D.tau0 = m->tau0; D.taurst = m->taurst;
  D.taurend = m->taurend; D.tau_maxvis = m->tau_maxvis; D.tight_tau = m->tight_tau;
  D.matter_verydom_tau = m->matter_verydom_tau; D.epsw = m->epsw; D.max_etak_scalar = m->max_etak_scalar;
  D.AccuracyBoost = m->AccuracyBoost; D.lAccuracyBoost = m->lAccuracyBoost;
  D.tol1 = 1.0e-4 / std::exp(m->AccuracyBoost - 1);  // camb/cmbmain.f90:1019
  D.use_galileon = m->use_galileon; D.DoLateRadTruncation = m->DoLateRadTruncation;
  D.HighAccuracyDefault = m->HighAccuracyDefault; D.AccuratePolarization = m->AccuratePolarization;
  D.AccurateReionization = m->AccurateReionization; D.WantLateTime = m->WantLateTime;
  D.SourceNum = m->WantLateTime ? 3 : 2;  // GetSourceMem, camb/cmbmain.f90:691-710
  D.NuMethod = m->MassiveNuMethod == 3 ? 1 : m->MassiveNuMethod;  // Nu_best -> Nu_trunc, equations_galileon.f90:820-821
  D.n_eig = m->Nu_mass_eigenstates; D.Num_Nu_massive = m->Num_Nu_massive; D.nqmax = m->nqmax;
  for (int i = 0; i < GC_MAX_NU; i++) {
    D.grhormass[i] = m->grhormass[i]; D.nu_masses[i] = m->nu_masses[i];
    D.nu_tau_nonrelativistic[i] = m->nu_tau_nonrelativistic[i]; D.nu_tau_massive[i] = m->nu_tau_massive[i];
    for (int qn = 0; qn < GC_MAX_NQ; qn++) D.nu_tau_notmassless[qn][i] = m->nu_tau_notmassless[qn][i];
  }
  for (int qn = 0; qn < GC_MAX_NQ; qn++) { D.nu_q[qn] = m->nu_q[qn]; D.nu_int_kernel[qn] = m->nu_int_kernel[qn]; }
  D.dlnam = m->dlnam;
  D.inv_dlnam = m->dlnam != 0 ? 1.0 / m->dlnam : 0.0;
  for (int qn = 0; qn < GC_MAX_NQ; qn++) D.inv_nu_q[qn] = m->nu_q[qn] != 0 ? 1.0 / m->nu_q[qn] : 0.0;
  D.nu_tab = g.d_nu_tab;
  D.nthermo = m->nthermo; D.tauminn = m->tauminn; D.dlntau = m->dlntau;
  D.inv_tauminn = 1.0 / m->tauminn; D.inv_dlntau = 1.0 / m->dlntau;
  D.nt = m->nt; D.nk = m->nk; D.nq = m->nq;
  D.c2 = m->c2; D.c3 = m->c3; D.c4 = m->c4; D.c5 = m->c5; D.cG = m->cG; D.h0 = m->h0; D.orad = m->orad;
  D.ScalarPowerAmp = m->ScalarPowerAmp; D.an = m->an; D.n_run = m->n_run; D.n_runrun = m->n_runrun; D.k_0_scalar = m->k_0_scalar;
  D.WantTransfer = m->WantTransfer ? 1 : 0;
  D.ntf = m->WantTransfer ? m->transfer_num_redshifts : 0;
  D.nk_tr = m->WantTransfer ? m->nk_transfer : 0;
  for (int i = 0; i < GC_MAX_TF; i++) D.tautf[i] = i < D.ntf ? m->tautf[i] : 0.0;
  D.h0_100 = m->H0 / 100.0;
  if (m->ALens < 0 || m->ALens_Fiducial < 0) return fail(5, "negative lensing amplitude");
  D.ALens = m->ALens > 0 ? m->ALens : 1.0;  // a zero-filled struct means the default
  D.ALens_fid = m->ALens_Fiducial;
  D.TimeSteps.count = m->n_tau_regions;
  D.TimeSteps.npoints = m->nt;
  D.TimeSteps.Lowest = m->tau[0];
  D.TimeSteps.Highest = m->tau[m->nt - 1];
  for (int i = 0; i < m->n_tau_regions; i++) {
    D.TimeSteps.R[i].Low = m->tau_regions[i].Low; D.TimeSteps.R[i].High = m->tau_regions[i].High;
    D.TimeSteps.R[i].delta = m->tau_regions[i].delta; D.TimeSteps.R[i].start_index = m->tau_regions[i].start_index;
    D.TimeSteps.R[i].IsLog = m->tau_regions[i].IsLog;
  }
  s.lmax = m->lmax;
  s.ntype = D.SourceNum > 2 ? 6 : 3;
  // ---- pack the tables (csrc/pack.h); a model that comes out of the library's own pre-stage was packed there already, on
  // the pre-stage threads and -- in the pipelined form -- behind the device work of the previous batch
  gcpack::Info I;
  gcpack::sizes(m, D.nk_tr, I);
  D.ngal = I.ngal;
  const size_t n_th = I.n_th, n_pm = I.n_pm, n_ts = I.n_ts, n_gal = I.n_gal, total = I.total;
  if (s.ensure_staging(total)) return fail(100, "cudaMallocHost staging");
  size_t pre_n = 0;
  gcpack::Info preI;
  if (const double* pre = gc_prestage_prepacked(m, &pre_n, &preI); pre && pre_n == total) {
    std::memcpy(s.staging, pre, total * sizeof(double));
    I = preI;
  } else {
    gcpack::tables(m, D.nk_tr, s.staging, I);
  }
  D.gal_a0 = I.gal_a0;
  D.gal_lnq_inv = I.gal_lnq_inv;
  s.m_nk_tr = D.nk_tr;
  s.total = total; s.n_th = n_th; s.n_pm = n_pm; s.n_ts = n_ts; s.n_gal = n_gal; s.m_nk = m->nk; s.m_nq = m->nq; s.m_lmax = m->lmax;
  return 0;
}

static int upload_model(int islot, const galcamb_model* m) {
  Slot& s = g.slots[islot];
  DevModel& D = s.host;
  if (m && m->Num_Nu_massive != 0 && !g.nu_tab_set) {
    if (!m->nu_r1) return fail(5, "massive neutrinos without tables");
    std::vector<double> t((size_t)GC_NRHOPN * 6);
    for (int i = 0; i < GC_NRHOPN; i++) {
      t[i * 6 + 0] = m->nu_r1[i]; t[i * 6 + 1] = m->nu_dr1[i]; t[i * 6 + 2] = m->nu_p1[i];
      t[i * 6 + 3] = m->nu_dp1[i]; t[i * 6 + 4] = m->nu_ddr1[i]; t[i * 6 + 5] = m->nu_ddp1[i];
    }
    CK(cudaMemcpy(g.d_nu_tab, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
    g.nu_tab_set = true;
  }
  const size_t total = s.total, n_th = s.n_th, n_pm = s.n_pm, n_ts = s.n_ts, n_gal = s.n_gal;
  if (grow(&s.d_tables, &s.tables_cap, total)) return fail(100, "cudaMalloc tables");
  CK(cudaMemcpyAsync(s.d_tables, s.staging, total * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  double* d = s.d_tables;
  D.th = d; d += n_th;
  D.dotmu_pmin = d; d += n_pm;
  D.ts = d; d += n_ts;
  D.gal = d; d += n_gal;
  D.k = d; d += s.m_nk;
  D.q = d; d += s.m_nq;
  D.dq = d; d += s.m_nq;
  D.k_tr = d; d += s.m_nk_tr;
  // ---- outputs
  const size_t n_src = (size_t)s.m_nk * D.SourceNum * m->nt, n_delta = (size_t)D.SourceNum * g.nl * s.m_nq;
  const size_t n_icl = (size_t)s.ntype * g.nl, n_cl = (size_t)s.ntype * (s.m_lmax - 1);
  const size_t n_tr = (size_t)D.ntf * (s.m_nk + s.m_nk_tr) * GC_TRANSFER_MAX;
  if (grow(&s.d_out, &s.out_cap, 2 * n_src + n_delta + n_icl + n_cl + n_tr)) return fail(100, "cudaMalloc outputs");
  d = s.d_out;
  D.Src = d; d += n_src;
  D.ddSrc = d; d += n_src;
  D.Delta = d; d += n_delta;
  D.iCl = d; d += n_icl;
  D.Cl = d; d += n_cl;
  D.Transfer = n_tr ? d : nullptr; d += n_tr;
  if (n_tr) CK(cudaMemsetAsync(D.Transfer, 0, n_tr * sizeof(double), g.stream));
  s.set = true;
  s.status = 0;
  return 0;
}

int galcamb_set_model_(const int* slot, const galcamb_model* m) {
  if (!g.ready) return fail(100, "not initialised");
  if (*slot < 0 || *slot >= (int)g.slots.size()) return fail(5, "bad slot");
  CK(cudaSetDevice(g.device));
  if (int rc = pack_model(*slot, m)) return rc;
  return upload_model(*slot, m);
}

static int upload_models() {
  const size_t n = g.slots.size();
  for (auto& s : g.slots)
    if (!s.set) return fail(5, "a slot of the batch has no model");
  if (n > g.models_cap) {
    if (g.d_models) cudaFree(g.d_models);
    CK(cudaMalloc((void**)&g.d_models, n * sizeof(DevModel)));
    g.models_cap = n;
  }
  std::vector<DevModel> hm(n);
  for (size_t i = 0; i < n; i++) hm[i] = g.slots[i].host;
  CK(cudaMemcpyAsync(g.d_models, hm.data(), n * sizeof(DevModel), cudaMemcpyHostToDevice, g.stream));
  CK(cudaStreamSynchronize(g.stream));  // hm is a local
  return 0;
}

// wavenumber kk of a slot's source grid, from its packed host copy (th | pmin | ts | gal | k | q | dq)
static double slot_k(const Slot& s, int kk) { return s.staging ? s.staging[s.n_th + s.n_pm + s.n_ts + s.n_gal + kk] : 0.0; }

int galcamb_evolve_sources_(void) {
  if (!g.ready) return fail(100, "not initialised");
  CK(cudaSetDevice(g.device));
  if (int rc = upload_models()) return rc;
  const int nm = (int)g.slots.size();
  // work items ordered by (k index ascending, model): with >= 32 models a warp holds the same k index of 32
  // cosmologies (lock-step control flow).  Low k first: those modes are integrated all the way to tau0
  // (~7600 trial steps) while k*tau > max_etak_scalar stops the high-k modes at ~800 (camb/cmbmain.f90:1034),
  // so the longest items start first and the short ones back-fill the SMs.
  int max_nk = 0, max_ntr = 0, NV = 0;
  for (auto& s : g.slots) {
    max_nk = std::max(max_nk, s.host.nk);
    max_ntr = std::max(max_ntr, s.host.nk_tr);
    const DevModel& D = s.host;
    // upper bound of nvar over k, from the two branches of GetNumEqns (camb/equations_galileon.f90:854-870)
    const double lAB = D.lAccuracyBoost;
    double mx = 0;
    for (int i = 0; i < D.n_eig; i++) mx = std::max(mx, D.nu_masses[i]);
    const int lmaxnu = D.Num_Nu_massive ? std::max(3, (int)std::lround((mx > 700 ? 15 : 10) * lAB)) : 0;
    const int nu_eqs = D.Num_Nu_massive ? (lmaxnu + 1) + D.n_eig * D.nqmax * (lmaxnu + 1) : 0;
    // q >= 0.05
    int gB = std::max((int)std::lround(((D.HighAccuracyDefault && D.AccuratePolarization) ? 11 : 8) * lAB), 3);
    int pB = D.AccuratePolarization ? gB : std::max((int)std::lround(4 * lAB), 3);
    int rB = std::max((int)std::lround(((mx > 700 && D.HighAccuracyDefault) ? 32 : 14) * lAB), 3);
    // q < 0.05
    int gA = std::max(3, (int)std::lround(8 * lAB)), pA = gA, rA = std::max(3, (int)std::lround(7 * lAB));
    if (D.AccurateReionization) { gA *= 4; pA *= 2; }
    const int nvA = 5 + (gA + 1) + (rA + 1) + pA - 1 + 2 + nu_eqs, nvB = 5 + (gB + 1) + (rB + 1) + pB - 1 + 2 + nu_eqs;
    const int nv = std::max(nvA, nvB);
    NV = std::max(NV, nv);
  }
  // Two forms of K1 (DESIGN.md "K1"):
  //  * eight lanes per item, integrator state in shared memory (csrc/evolve.cu): lowest time per trial step, but an SM
  //    holds only ~20 items -- the choice while the batch is small enough that the critical path of the slowest k mode
  //    sets the run time (single spectra, MCMC chain steps, k-sharded calls);
  //  * one lane per item, state in a lane-interleaved L2/HBM workspace (csrc/evolve_batch.cu): 256 items per SM and
  //    every lane busy in the scalar part of the right-hand side -- the choice for large parameter batches.
  // Measured crossover on config 2: ~64 models (13 000 items; 32 models 0.88 s vs 1.54 s, 64: 1.63 vs 1.64, 96: 2.45 vs 1.73).
  long total_items = 0;
  for (auto& s : g.slots) total_items += (s.host.nk + s.host.nk_tr + g.kshard_world - 1) / g.kshard_world;
  static long octet_items = -1;
  if (octet_items < 0) {
    const char* e = getenv("GALCAMB_OCTET_ITEMS");  // experiment knob: item count up to which the eight-lane form is used
    octet_items = e ? atol(e) : 13000;
  }
  const int NVP = NV;
  bool want_tr = false;
  for (auto& s : g.slots) want_tr = want_tr || s.host.WantTransfer != 0;
  bool want_nu1 = !want_tr && !getenv("GALCAMB_NO_NU1");  // (experiment knob: always the general kernels)
  for (auto& s : g.slots) want_nu1 = want_nu1 && s.host.n_eig <= 1;
  const int warps_per_sm = want_tr ? gc_evolve_warps_per_cta_tr(NVP) : gc_evolve_warps_per_cta(NVP);
  const bool batch_form = total_items > octet_items || warps_per_sm < 1;
  g.k1_form = batch_form ? 1 : 0;
  g.items.clear();
  int ipw = 4;
  static int sort_items = -1;
  if (sort_items < 0) {
    const char* e = getenv("GALCAMB_SORT_ITEMS");  // experiment knob: 0 = models in slot order within a k index
    sort_items = e ? atoi(e) : 1;
  }
  if (batch_form) {
    // every k index starts on a warp boundary (padding items have model = -1): a warp never mixes different k, whose
    // step sequences and switch times differ -- mixing them costs 2-4x.  Low k first: those modes are integrated all the
    // way to tau0 (~7600 trial steps) while k*tau > max_etak_scalar stops the high-k modes at ~800
    // (camb/cmbmain.f90:1034), so the longest items start first and the short ones back-fill the SMs.
    // Within one k index the lanes of a warp should take similar adaptive step sequences: the models are dealt to the
    // warps in order of their wavenumber at this index (the k grids of different cosmologies differ by a few per cent
    // through taurst and tau0), so neighbours in k -- whose oscillation phases and switch times are closest -- share a warp.
    std::vector<std::pair<double, int>> order;
    for (int kk = 0; kk < max_nk; kk++) {
      order.clear();
      if (kk % g.kshard_world != g.kshard_rank) continue;
      for (int m = 0; m < nm; m++)
        if (kk < g.slots[m].host.nk) order.emplace_back(sort_items ? slot_k(g.slots[m], kk) : 0.0, m);
      if (sort_items) std::stable_sort(order.begin(), order.end());
      for (auto& om : order) g.items.push_back(make_int2(om.second, kk));
      while (g.items.size() % 32) g.items.push_back(make_int2(-1, -1));
    }
    // the wavenumbers beyond the source grid (TransferOut, camb/cmbmain.f90:1164-1177): k index nk + t of their model
    for (int t = 0; t < max_ntr; t++) {
      if (t % g.kshard_world != g.kshard_rank) continue;
      for (int m = 0; m < nm; m++)
        if (t < g.slots[m].host.nk_tr) g.items.push_back(make_int2(m, g.slots[m].host.nk + t));
      while (g.items.size() % 32) g.items.push_back(make_int2(-1, -1));
    }
  } else {
    // Work queue: groups of GC_IPW = 4 items, one group per warp at a time, ordered by k index ascending (longest
    // first, the reference's SCHEDULE(DYNAMIC) over k, camb/cmbmain.f90:201).  A group holds the SAME k index of up to
    // four cosmologies (near-identical control flow across the warp).  With few items (a single model: ~200) a group
    // holds fewer items so that every item gets a warp of its own.
    const long slots_hw = (long)g.num_sms * warps_per_sm;
    ipw = (int)((total_items + slots_hw - 1) / slots_hw);
    ipw = std::max(1, std::min(ipw, 4));
    if (const char* e = getenv("GALCAMB_ITEMS_PER_WARP")) ipw = std::max(1, std::min(atoi(e), 4));  // experiment knob
    std::vector<std::pair<double, int>> order;
    for (int kk = 0; kk < max_nk; kk++) {
      int in_group = 0;
      order.clear();
      if (kk % g.kshard_world != g.kshard_rank) continue;
      for (int m = 0; m < nm; m++)
        if (kk < g.slots[m].host.nk) order.emplace_back(sort_items ? slot_k(g.slots[m], kk) : 0.0, m);
      if (sort_items) std::stable_sort(order.begin(), order.end());
      for (auto& om : order) {
        const int m = om.second;
        g.items.push_back(make_int2(m, kk));
        if (++in_group == ipw) {
          for (; in_group < 4; in_group++) g.items.push_back(make_int2(-1, -1));
          in_group = 0;
        }
      }
      if (in_group)
        for (; in_group < 4; in_group++) g.items.push_back(make_int2(-1, -1));
    }
    for (int t = 0; t < max_ntr; t++) {  // transfer-only wavenumbers (see the batch form above)
      if (t % g.kshard_world != g.kshard_rank) continue;
      int in_group = 0;
      for (int m = 0; m < nm; m++) {
        if (t >= g.slots[m].host.nk_tr) continue;
        g.items.push_back(make_int2(m, g.slots[m].host.nk + t));
        if (++in_group == ipw) {
          for (; in_group < 4; in_group++) g.items.push_back(make_int2(-1, -1));
          in_group = 0;
        }
      }
      if (in_group)
        for (; in_group < 4; in_group++) g.items.push_back(make_int2(-1, -1));
    }
  }
  const int nitems = (int)g.items.size();
  const int ngroups = nitems / 4;
  if ((size_t)nitems > g.items_cap) {
    if (g.d_items) cudaFree(g.d_items);
    if (g.d_status) cudaFree(g.d_status);
    g.d_items = nullptr;
    g.d_status = nullptr;
    g.items_cap = 0;
    CK(cudaMalloc((void**)&g.d_items, nitems * sizeof(int2)));
    CK(cudaMalloc((void**)&g.d_status, nitems * sizeof(int)));
    g.items_cap = nitems;
  }
  CK(cudaMemcpyAsync(g.d_items, g.items.data(), nitems * sizeof(int2), cudaMemcpyHostToDevice, g.stream));
  CK(cudaMemsetAsync(g.d_counters, 0, 24 * sizeof(unsigned long long), g.stream));
  if (batch_form) {
    // one slot per RESIDENT warp (persistent CTAs of up to 8 warps, csrc/evolve_batch.cu): never more than the warps of
    // the item list rounded up to whole CTAs
    const size_t nwarps = (nitems + 31) / 32 + 8;
    const size_t ws_need = nwarps * 12 * (size_t)NV * 32;  // GC_NCOL columns
    if (grow(&g.d_ws, &g.ws_cap, ws_need)) return fail(100, "cudaMalloc workspace");
  }
  CK(cudaEventRecord(g.ev[0], g.stream));
  if (batch_form) {
    CK((want_tr ? gc_launch_evolve_batch_tr : want_nu1 ? gc_launch_evolve_batch_nu1 : gc_launch_evolve_batch)(
        g.d_models, g.d_items, nitems, g.d_ws, NV, g.d_status, g.d_counters, g.stream));
  } else {
    CK((want_tr ? gc_launch_evolve_tr : gc_launch_evolve)(g.d_models, g.d_items, ngroups, NVP, g.num_sms, g.d_queue, g.d_status,
                                                            g.d_counters, g.stream));
  }
  CK(cudaEventRecord(g.ev[1], g.stream));
  std::vector<int> st(nitems);
  CK(cudaMemcpyAsync(st.data(), g.d_status, nitems * sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  unsigned long long c[24];
  CK(cudaMemcpyAsync(c, g.d_counters, 24 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, g.stream));
  // K1 runs for seconds on a large batch: wait on a blocking-sync event so that this thread's core stays free for the
  // host pre-stage of the next batch (one process per GPU: N spinning threads are N cores less for it)
  CK(cudaEventRecord(g.ev_block, g.stream));
  CK(cudaEventSynchronize(g.ev_block));
  float ms = 0;
  cudaEventElapsedTime(&ms, g.ev[0], g.ev[1]);
  g.ms[0] = ms;
  // c[0] = right-hand-side evaluations of the reference algorithm: k1..k8 of every trial step
  // (k1 not after a rejected one) + one per output (camb/equations_galileon.f90:1297).  K1 executes c[5] of them:
  // the y' of an output is the k1 of the next step (c[4] outputs served that way), a rejected step re-forms its k1.
  // (The lane-per-item form counts the executed evaluations in c[0]; the algorithmic count is c[0] + c[4].)
  g.counters[0] = batch_form ? (double)(c[0] + c[4]) : (double)c[0];
  g.counters[1] = (double)c[1];
  g.counters[2] = (double)c[2];
  g.evolve_stats[0] = batch_form ? (double)c[0] : (double)c[5];
  g.evolve_stats[1] = (double)c[4];
  g.evolve_stats[2] = batch_form ? 0.0 : (double)c[6];
  if (!batch_form && getenv("GALCAMB_PHASE_PROFILE")) {  // only meaningful for a -DGC_PROFILE_PHASES build: warp-cycles per phase
    const char* nm[9] = {"window", "phaseA", "phaseB", "combine", "scalar", "hier", "output", "final", "advance"};
    double tot = 0;
    for (int i = 0; i < 9; i++) tot += (double)c[8 + i];
    fprintf(stderr, "[galcamb] warp-cycles per trial step (%.0f trial steps, %d items per warp):", (double)c[6], ipw);
    for (int i = 0; i < 9; i++)
      fprintf(stderr, " %s %.0f (%.1f%%)", nm[i], (double)c[8 + i] / ((double)c[6] / ipw), 100.0 * c[8 + i] / tot);
    fprintf(stderr, "\n");
  }
  int rc = 0;
  for (int i = 0; i < nitems; i++)
    if (st[i] && g.items[i].x >= 0) {
      // dverk's own failure codes (101: bad re-entry, 102: hmin > hmax, 103: more equations than the kernel was sized
      // for) all surface as CAMB's error_evolution (camb/equations_galileon.f90:285-292); the raw code stays in the message
      g.slots[g.items[i].x].status = st[i] >= 100 ? 4 : st[i];
      rc = 4;  // error_evolution
      char buf[160];
      snprintf(buf, sizeof buf, "evolve_sources: status %d for model %d k index %d", st[i], g.items[i].x, g.items[i].y);
      g.err = buf;
    }
  return rc;
}

int galcamb_los_integrate_(void) {
  if (!g.ready) return fail(100, "not initialised");
  CK(cudaSetDevice(g.device));
  if (int rc = upload_models()) return rc;
  const int nm = (int)g.slots.size();
  int max_rows = 0, max_nq = 0, max_nt = 0, S = 2;
  for (auto& s : g.slots) {
    max_rows = std::max(max_rows, s.host.nt * s.host.SourceNum);
    max_nq = std::max(max_nq, s.host.nq);
    max_nt = std::max(max_nt, s.host.nt);
    S = std::max(S, s.host.SourceNum);
    if (s.host.nk > 768) return fail(5, "nk > 768 not supported by spline_k");
  }
  CK(cudaEventRecord(g.ev[2], g.stream));
  CK(gc_launch_spline_k(g.d_models, nm, max_rows, g.stream));
  CK(cudaEventRecord(g.ev[3], g.stream));
  CK(gc_launch_zero_delta(g.d_models, nm, g.nl, g.stream));
  CK(cudaMemsetAsync(g.d_counters + 3, 0, sizeof(unsigned long long), g.stream));
  {
    const int q_lo = g.q_hi >= 0 ? g.q_lo : 0, q_n = g.q_hi >= 0 ? std::max(0, std::min(g.q_hi, max_nq) - g.q_lo) : max_nq;
    if (q_n > 0)
      CK(gc_launch_los(g.d_models, nm, q_n, max_nt, S, g.d_bess, g.d_xs, g.d_ls, g.nx, g.nl, g.d_counters + 3, q_lo, g.stream));
  }
  CK(cudaEventRecord(g.ev[4], g.stream));
  unsigned long long c = 0;
  CK(cudaMemcpyAsync(&c, g.d_counters + 3, sizeof c, cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, g.ev[2], g.ev[3]);
  g.ms[1] = ms;
  cudaEventElapsedTime(&ms, g.ev[3], g.ev[4]);
  g.ms[2] = ms;
  g.counters[3] = (double)c;
  return 0;
}

int galcamb_scalar_cls_(void) {
  if (!g.ready) return fail(100, "not initialised");
  CK(cudaSetDevice(g.device));
  if (int rc = upload_models()) return rc;
  const int nm = (int)g.slots.size();
  const int lmax = g.slots[0].lmax, ntype = g.slots[0].ntype;
  for (auto& s : g.slots)
    if (s.lmax != lmax || s.ntype != ntype) return fail(5, "all models of a batch must share lmax and source count");
  if (g.nl > 512) return fail(5, "nl > 512 not supported");
  // InterpolateClArrTemplated (camb/modules.f90:1046-1089) is restated for a template that covers Max_l; its
  // direct-interpolation tail beyond lmax_extrap_highl = 8000 is not
  if (g.d_tmpl && (g.lmax_tmpl < lmax || lmax > 8000)) return fail(5, "high-l template shorter than Max_l (or Max_l > 8000)");
  CK(cudaEventRecord(g.ev[5], g.stream));
  CK(gc_launch_cls(g.d_models, nm, g.d_ls, g.nl, ntype, lmax, g.d_tmpl, g.lmax_tmpl, g.cls_what, g.stream));
  CK(cudaEventRecord(g.ev[6], g.stream));
  CK(cudaStreamSynchronize(g.stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, g.ev[5], g.ev[6]);
  g.ms[3] = ms;
  return 0;
}

// lens_Cls / CorrFuncFullSky, camb/lensing.f90:78-215 : geometry of the theta and l sampling on the host, the
// correlation functions and their transform on the device (csrc/lensing.cu)
int galcamb_lens_cls_(void) {
  if (!g.ready) return fail(100, "not initialised");
  CK(cudaSetDevice(g.device));
  const int nm = (int)g.slots.size();
  if (nm < 1) return fail(5, "no models");
  const int Max_l = g.slots[0].lmax;
  const double AB = g.slots[0].host.AccuracyBoost;
  const int HAD = g.slots[0].host.HighAccuracyDefault;
  for (auto& s : g.slots) {
    if (!s.set || s.ntype < 6) return fail(5, "lens_cls needs the lensing potential (DoLensing models, 3 sources)");
    if (s.lmax != Max_l || s.host.AccuracyBoost != AB || s.host.HighAccuracyDefault != HAD)
      return fail(5, "all models of a lensing batch must share lmax and the accuracy settings");
  }
  const double pi = 3.14159265358979323846264338328;
  int lmax_extrap = Max_l - 100 + 450;
  if (HAD) lmax_extrap += 300;
  lmax_extrap = std::min(8000, lmax_extrap);
  const int lmax = std::max(lmax_extrap, Max_l);
  if (lmax > Max_l && (!g.d_tmpl || !g.d_tmpl_pp || g.lmax_tmpl < lmax || g.lmax_tmpl_pp != g.lmax_tmpl))
    return fail(5, "lens_cls needs the high-L template (galcamb_set_template_ + galcamb_set_lensing_template_)");
  int ix = g.nl - 1;
  if (Max_l - 100 < g.ls[0]) return fail(5, "lens_cls: Max_l too small for the lensed_convolution_margin of 100");
  while (ix > 0 && g.ls[ix] > Max_l - 100) ix--;  // lensed_convolution_margin, :99-103
  const int lmax_lensed = g.ls[ix];
  int npoints = (int)(Max_l * 2 * AB);
  double dtheta = pi / npoints;
  if (Max_l > 3500) dtheta = dtheta / (double)1.3f;
  const int apw = (int)std::lround((double)0.003f / dtheta);
  npoints = (int)(pi / dtheta);
  const double range_fac = std::max(1.0, 32 / AB);
  npoints = (int)(npoints / range_fac);
  const int interp_fac = std::max(1, std::min((int)std::lround(10 / AB), (int)(range_fac * 2) - 1));
  const int nblk = (npoints - 1 + 127) / 128;
  const int geom[7] = {lmax, Max_l, lmax_lensed, npoints, interp_fac, 3 * apw, nblk};
  if (grow(g.d_lens_ta, g.lens_tab_cap, (size_t)lmax + 5) || grow(g.d_lens_tb, g.lens_tabb_cap, (size_t)lmax + 5))
    return fail(100, "out of device memory");
  if (g.lens_tab_lmax != lmax) {
    CK(gc_launch_lens_ltab(g.d_lens_ta, g.d_lens_tb, lmax, g.stream));
    g.lens_tab_lmax = lmax;
  }
  {  // apodisation window, :467-470 (default-real arithmetic)
    std::vector<float> w(3 * apw + 1, 1.0f);
    for (int d = 1; d <= 3 * apw; d++) w[d] = std::exp(-(float)(d * d) / (float)(2 * apw * apw));
    if (grow(g.d_lens_apod, g.lens_apod_cap, w.size())) return fail(100, "out of device memory");
    CK(cudaMemcpyAsync(g.d_lens_apod, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice, g.stream));
    CK(cudaStreamSynchronize(g.stream));
  }
  if (grow(g.d_lens_cin, g.lens_cin_cap, (size_t)nm * (lmax + 1)) ||
      grow(g.d_lens_partial, g.lens_partial_cap, (size_t)nm * nblk * 4 * (lmax_lensed + 1)) ||
      grow(g.d_lensed, g.lensed_cap, (size_t)nm * 4 * (lmax_lensed - 1)) || grow(g.d_lens_status, g.lens_status_cap, (size_t)nm))
    return fail(100, "out of device memory");
  if (int rc = upload_models()) return rc;
  CK(cudaEventRecord(g.ev[7], g.stream));
  CK(gc_launch_lensing(g.d_models, nm, geom, dtheta, g.d_lens_ta, g.d_lens_tb, g.d_tmpl, g.d_tmpl_pp, g.lmax_tmpl, g.d_lens_cin,
                       g.d_lens_apod, g.d_lens_partial, g.d_lens_status, g.d_lensed, g.stream));
  CK(cudaEventRecord(g.ev[8], g.stream));
  g.lens_status.assign(nm, 0);
  CK(cudaMemcpyAsync(g.lens_status.data(), g.d_lens_status, nm * sizeof(int), cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, g.ev[7], g.ev[8]);
  g.ms[4] = ms;
  g.lmax_lensed = lmax_lensed;
  g.lens_nm = nm;
  int rc = 0;
  for (int i = 0; i < nm; i++)
    if (g.lens_status[i] && !g.slots[i].status) {
      g.slots[i].status = 8;  // error_norm_lensing
      rc = 8;
    }
  if (rc) g.err = "You need to normalize realistically to use lensing (camb/lensing.f90:230-235)";
  return rc;
}

int galcamb_get_lensed_cls_(const int* slot, int* lmax_lensed, double* Cl_lensed) {
  if (*slot < 0 || *slot >= g.lens_nm || !g.d_lensed) return fail(5, "bad slot / lens_cls not run");
  *lmax_lensed = g.lmax_lensed;
  if (Cl_lensed) {
    const size_t per = (size_t)4 * (g.lmax_lensed - 1);
    CK(cudaMemcpyAsync(Cl_lensed, g.d_lensed + per * *slot, per * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    if (g.slots[*slot].status)
      for (size_t j = 0; j < per; j++) Cl_lensed[j] = std::nan("");
  }
  return 0;
}

// CAMB_TransfersToPowers, camb/camb.f90:87-102 (the fast-parameter path of source/Calculator_CAMB.f90:258): new
// primordial parameters for a slot whose Delta_p_l_k is still on the device, then ClTransferToCl (+ lens_Cls) only.
int galcamb_set_primordial_(const int* slot, const double* ScalarPowerAmp, const double* an, const double* n_run,
                            const double* n_runrun, const double* k_0_scalar) {
  if (*slot < 0 || *slot >= (int)g.slots.size() || !g.slots[*slot].set) return fail(5, "bad slot");
  if (!(*ScalarPowerAmp > 0) || !(*k_0_scalar > 0)) return fail(3, "bad primordial power spectrum");  // error_inital_power
  DevModel& D = g.slots[*slot].host;
  D.ScalarPowerAmp = *ScalarPowerAmp;
  D.an = *an;
  D.n_run = *n_run;
  D.n_runrun = *n_runrun;
  D.k_0_scalar = *k_0_scalar;
  return 0;
}

int galcamb_set_alens_(const int* slot, const double* ALens, const double* ALens_Fiducial) {
  if (*slot < 0 || *slot >= (int)g.slots.size() || !g.slots[*slot].set) return fail(5, "bad slot");
  if (!(*ALens > 0) || *ALens_Fiducial < 0) return fail(5, "bad lensing amplitude");
  g.slots[*slot].host.ALens = *ALens;
  g.slots[*slot].host.ALens_fid = *ALens_Fiducial;
  return 0;
}

int galcamb_transfers_to_powers_(const int* do_lensing) {
  if (int rc = galcamb_scalar_cls_()) return rc;
  if (*do_lensing) return galcamb_lens_cls_();
  return 0;
}

int galcamb_get_sources_(const int* slot, double* Src) {
  if (*slot < 0 || *slot >= (int)g.slots.size() || !g.slots[*slot].set) return fail(5, "bad slot");
  const DevModel& D = g.slots[*slot].host;
  CK(cudaMemcpyAsync(Src, D.Src, (size_t)D.nk * D.SourceNum * D.nt * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return 0;
}

int galcamb_put_sources_(const int* slot, const double* Src) {
  if (*slot < 0 || *slot >= (int)g.slots.size() || !g.slots[*slot].set) return fail(5, "bad slot");
  const DevModel& D = g.slots[*slot].host;
  CK(cudaMemcpyAsync(D.Src, Src, (size_t)D.nk * D.SourceNum * D.nt * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return 0;
}

int galcamb_get_transfers_(const int* slot, double* Delta) {
  if (*slot < 0 || *slot >= (int)g.slots.size() || !g.slots[*slot].set) return fail(5, "bad slot");
  const DevModel& D = g.slots[*slot].host;
  CK(cudaMemcpyAsync(Delta, D.Delta, (size_t)D.SourceNum * g.nl * D.nq * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return 0;
}

int galcamb_get_cls_(const int* slot, double* iCl, double* Cl) {
  if (*slot < 0 || *slot >= (int)g.slots.size() || !g.slots[*slot].set) return fail(5, "bad slot");
  const Slot& s = g.slots[*slot];
  if (iCl) CK(cudaMemcpyAsync(iCl, s.host.iCl, (size_t)s.ntype * g.nl * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  if (Cl) CK(cudaMemcpyAsync(Cl, s.host.Cl, (size_t)s.ntype * (s.lmax - 1) * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return 0;
}

int galcamb_put_cls_(const int* slot, const double* Cl) {
  if (*slot < 0 || *slot >= (int)g.slots.size() || !g.slots[*slot].set) return fail(5, "bad slot");
  const Slot& s = g.slots[*slot];
  CK(cudaMemcpyAsync(s.host.Cl, Cl, (size_t)s.ntype * (s.lmax - 1) * sizeof(double), cudaMemcpyHostToDevice, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return 0;
}

// MT%TransferData (FP64) + Transfer_Get_sigmas (camb/modules.f90:2304-2370, :2450-2472: total matter, R = 8 Mpc/h,
// trapezoid in ln k over MT%q_trans with the primordial power of the slot)
int galcamb_get_transfer_(const int* slot, int* num_q_trans, int* num_redshifts, double* q_trans, double* TransferData,
                          double* sigma8) {
  if (!g.ready) return fail(100, "not initialised");
  if (*slot < 0 || *slot >= (int)g.slots.size() || !g.slots[*slot].set) return fail(5, "bad slot");
  CK(cudaSetDevice(g.device));
  const Slot& s = g.slots[*slot];
  const DevModel& D = s.host;
  const int nqt = D.WantTransfer ? D.nk + D.nk_tr : 0, nz = D.ntf;
  if (num_q_trans) *num_q_trans = nqt;
  if (num_redshifts) *num_redshifts = nz;
  if (!nqt || !nz) return 0;
  if (q_trans) {  // the packed host copy holds both wavenumber lists (th | pmin | ts | gal | k | q | dq | k_tr)
    const double* kk = s.staging + s.n_th + s.n_pm + s.n_ts + s.n_gal;
    std::copy(kk, kk + D.nk, q_trans);
    std::copy(kk + D.nk + 2 * (size_t)D.nq, kk + D.nk + 2 * (size_t)D.nq + D.nk_tr, q_trans + D.nk);
  }
  if (!TransferData && !sigma8) return 0;
  const size_t n = (size_t)nz * nqt * GC_TRANSFER_MAX;
  std::vector<double> tmp;
  double* T = TransferData;
  if (!T) {
    tmp.resize(n);
    T = tmp.data();
  }
  CK(cudaMemcpyAsync(T, D.Transfer, n * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  if (sigma8) {
    const double radius = 8.0, H = D.h0_100;
    std::vector<double> dsig8o(nz, 0.0), sig8(nz, 0.0);
    double lnko = 0;
    for (int ik = 0; ik < nqt; ik++) {
      const double kh = T[(size_t)ik * GC_TRANSFER_MAX];
      if (kh == 0) continue;
      const double k = kh * H, x = kh * radius;
      const double win = 3 * (std::sin(x) - x * std::cos(x)) / (x * x * x);
      const double lnk = std::log(k);
      const double dlnk = ik == 0 ? 0.5 : lnk - lnko;
      // ScalarPower, camb/power_tilt.f90:149-172
      const double lnrat = std::log(k / D.k_0_scalar);
      const double powers = D.ScalarPowerAmp * std::exp(lnrat * (D.an - 1 + lnrat * (D.n_run / 2 + D.n_runrun / 6 * lnrat)));
      for (int iz = 0; iz < nz; iz++) {
        const double t = T[((size_t)iz * nqt + ik) * GC_TRANSFER_MAX + 6];  // Transfer_tot
        const double dsig8 = ((win * (k * k)) * (win * (k * k))) * powers * (t * t);
        sig8[iz] = sig8[iz] + (dsig8 + dsig8o[iz]) * dlnk / 2;
        dsig8o[iz] = dsig8;
      }
      lnko = lnk;
    }
    for (int iz = 0; iz < nz; iz++) sigma8[iz] = std::sqrt(sig8[iz]);
  }
  return 0;
}

int galcamb_get_status_(const int* slot, int* status) {
  if (*slot < 0 || *slot >= (int)g.slots.size()) return fail(5, "bad slot");
  *status = g.slots[*slot].status;
  return 0;
}

// models -> device slots: AoS packing (+ background spline) on the host threads, then one upload per slot
int galcamb_batch_upload_(const galcamb_model* models, const int* nmodels) {
  if (!g.ready) return fail(100, "not initialised");
  int rc = galcamb_batch_begin_(nmodels);
  if (rc) return rc;
  CK(cudaSetDevice(g.device));
  const int n = *nmodels;
  const auto t_pack0 = std::chrono::steady_clock::now();
  const int hostt = galcamb_get_host_threads_();
  const int nth = std::max(1, std::min<int>(n, hostt > 0 ? hostt : (int)std::thread::hardware_concurrency()));
  std::vector<int> rcs(nth, 0);
  std::vector<std::string> errs(nth);
  std::vector<std::thread> pool;
  std::atomic<int> next(0);
  Ctx* const parent = tl_ctx;
  for (int t = 0; t < nth; t++)
    pool.emplace_back([&, t]() {
      tl_ctx = parent;          // the workers pack into the slots of the caller's context
      cudaSetDevice(g.device);  // pinned-buffer growth inside pack_model
      for (;;) {
        const int i = next.fetch_add(1);
        if (i >= n) break;
        if (int r = pack_model(i, &models[i])) rcs[t] = r;
      }
    });
  for (auto& th : pool) th.join();
  for (int r : rcs)
    if (r) return r;
  const auto t_pack1 = std::chrono::steady_clock::now();
  for (int i = 0; i < n; i++)
    if ((rc = upload_model(i, &models[i]))) return rc;
  g.host_ms[0] = std::chrono::duration<double, std::milli>(t_pack1 - t_pack0).count();                            // packing
  g.host_ms[1] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_pack1).count();  // H2D enqueue
  return 0;
}

// the current pre-staged batch (csrc/prestage.cu) -> device slots; every point must be valid
int galcamb_prestaged_upload_(void) {
  const int n = galcamb_prestage_count_();
  if (n < 1) return fail(5, "no pre-staged batch");
  std::vector<galcamb_model> models(n);
  for (int i = 0; i < n; i++)
    if (galcamb_prestage_model_(&i, &models[i])) return fail(5, "a point of the pre-staged batch was rejected");
  int nl = 0, kmaxfile = 0, zero = 0;
  galcamb_prestage_lsamples_(&zero, &nl, nullptr, &kmaxfile);
  std::vector<int> ls(nl);
  galcamb_prestage_lsamples_(&zero, &nl, ls.data(), &kmaxfile);
  const double boost = models[0].AccuracyBoost;
  if (int rc = galcamb_set_bessels_(ls.data(), &nl, &kmaxfile, &boost)) return rc;
  return galcamb_batch_upload_(models.data(), &n);
}

int galcamb_batch_cls_(const galcamb_model* models, const int* nmodels, double* Cl_out) {
  if (!g.ready) return fail(100, "not initialised");
  const auto t_call0 = std::chrono::steady_clock::now();
  int rc = galcamb_batch_upload_(models, nmodels);
  if (rc) return rc;
  const auto t_call1 = std::chrono::steady_clock::now();
  // A parameter point whose integration fails (CAMB error_evolution) must not sink the batch: its slot keeps the
  // error id (galcamb_get_status_), its C_l come back as NaN, the other points are unaffected (CosmoMC rejects
  // such points one by one, source/Calculator_CAMB.f90:235-243).
  const int rc_ev = galcamb_evolve_sources_();
  if (rc_ev != 0 && rc_ev != 4) return rc_ev;
  if ((rc = galcamb_los_integrate_())) return rc;
  if ((rc = galcamb_scalar_cls_())) return rc;
  const auto t_call2 = std::chrono::steady_clock::now();
  const size_t per = (size_t)3 * (models[0].lmax - 1);
  for (int i = 0; i < *nmodels; i++)
    CK(cudaMemcpyAsync(Cl_out + per * i, g.slots[i].host.Cl, per * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  g.host_ms[2] = std::chrono::duration<double, std::milli>(t_call1 - t_call0).count();                            // upload, whole
  g.host_ms[3] = std::chrono::duration<double, std::milli>(t_call2 - t_call1).count();                            // stages 1-4
  g.host_ms[4] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call2).count();  // D2H
  for (int i = 0; i < *nmodels; i++)
    if (g.slots[i].status)
      for (size_t j = 0; j < per; j++) Cl_out[per * i + j] = std::nan("");
  return rc_ev;
}

// Parameters -> C_l: the host pre-stage (csrc/prestage.cu, one host thread per point) followed by stages 1-4 on the
// device.  Points the pre-stage rejects (unstable Galileon background, reionization not converged ...) are left out
// of the device batch and come back as NaN rows, like points whose integration fails.
int galcamb_prestaged_to_cls_(double* Cl_out) {
  if (!g.ready) return fail(100, "not initialised");
  const int n = galcamb_prestage_count_();
  if (n < 1) return fail(5, "no pre-staged batch");
  const auto t_col0 = std::chrono::steady_clock::now();
  std::vector<galcamb_model> models;
  std::vector<int> where;
  for (int i = 0; i < n; i++) {
    galcamb_model m;
    if (galcamb_prestage_model_(&i, &m) == 0) {
      models.push_back(m);
      where.push_back(i);
    }
  }
  if (models.empty()) return fail(5, "no valid point in the batch");
  const size_t per = (size_t)3 * (models[0].lmax - 1);
  for (size_t j = 0; j < per * (size_t)n; j++) Cl_out[j] = std::nan("");
  {  // Bessel tables for this l sampling / x range (cached across calls like camb/bessels.f90:40-41)
    int nl = 0, kmaxfile = 0;
    galcamb_prestage_lsamples_(&where[0], &nl, nullptr, &kmaxfile);
    std::vector<int> ls(nl);
    galcamb_prestage_lsamples_(&where[0], &nl, ls.data(), &kmaxfile);
    const double boost = models[0].AccuracyBoost;
    if (int rc = galcamb_set_bessels_(ls.data(), &nl, &kmaxfile, &boost)) return rc;  // returns at once when cached
  }
  const int nm = (int)models.size();
  g.host_ms[5] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_col0).count();
  if (nm == n) return galcamb_batch_cls_(models.data(), &nm, Cl_out);
  std::vector<double> out(per * nm);
  const int rc = galcamb_batch_cls_(models.data(), &nm, out.data());
  if (rc != 0 && rc != 4) return rc;
  for (int j = 0; j < nm; j++) std::copy(out.begin() + per * j, out.begin() + per * (j + 1), Cl_out + per * where[j]);
  return rc;
}

int galcamb_params_to_cls_(const galcamb_params* params, const int* n, const int* nthreads, double* Cl_out) {
  if (!g.ready) return fail(100, "not initialised");
  const int rc_pre = galcamb_prestage_(params, n, nthreads);
  const int rc = galcamb_prestaged_to_cls_(Cl_out);
  return rc ? rc : rc_pre;
}

// galcamb_batch_cls_ followed by the lensing post-step: what CAMB_GetResults hands CosmoMC when DoLensing = T
// (Cl_lensed, camb/camb.f90:167-171 -> camb/lensing.f90:78).  Cl_out may be NULL.
int galcamb_batch_lensed_cls_(const galcamb_model* models, const int* nmodels, double* Cl_out, int* lmax_lensed,
                              double* Cl_lensed_out) {
  if (!g.ready) return fail(100, "not initialised");
  const int lmax = models[0].lmax;
  std::vector<double> tmp;
  if (!Cl_out) {
    tmp.resize((size_t)3 * (lmax - 1) * *nmodels);
    Cl_out = tmp.data();
  }
  const int rc_ev = galcamb_batch_cls_(models, nmodels, Cl_out);
  if (rc_ev != 0 && rc_ev != 4) return rc_ev;
  const int rc_l = galcamb_lens_cls_();
  if (rc_l != 0 && rc_l != 8) return rc_l;
  *lmax_lensed = g.lmax_lensed;
  const size_t per = (size_t)4 * (g.lmax_lensed - 1);
  CK(cudaMemcpyAsync(Cl_lensed_out, g.d_lensed, per * *nmodels * sizeof(double), cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  for (int i = 0; i < *nmodels; i++)
    if (g.slots[i].status)
      for (size_t j = 0; j < per; j++) Cl_lensed_out[per * i + j] = std::nan("");
  return rc_ev ? rc_ev : rc_l;
}

int galcamb_get_timings_(double* ms) {
  for (int i = 0; i < 8; i++) ms[i] = g.ms[i];
  ms[7] = g.ms[0] + g.ms[1] + g.ms[2] + g.ms[3];  // the unlensed path; ms[4] = lensing
  return 0;
}

// wall clock of the host-side phases of the last galcamb_batch_cls_ call [ms]: 0 table packing on the host threads,
// 1 H2D enqueue, 2 upload as a whole, 3 device stages 1-4 (incl. their host set-up), 4 D2H of the C_l,
// 5 the collection of the pre-staged models in galcamb_prestaged_to_cls_
int galcamb_get_host_timings_(double* ms) {
  for (int i = 0; i < 8; i++) ms[i] = g.host_ms[i];
  return 0;
}

int galcamb_get_evolve_stats_(double* s) {
  s[0] = g.evolve_stats[0];  // right-hand sides executed
  s[1] = g.evolve_stats[1];  // outputs that took their y' from the next step's k1
  return 0;
}

int galcamb_get_counters_(double* c) {
  for (int i = 0; i < 4; i++) c[i] = g.counters[i];
  return 0;
}

// ================================================================================================ several GPUs, one process
// SURVEY.md section 8e behind the C ABI: a Fortran host (or any single process) reaches every GPU of the box through
// these two calls.  One worker thread per device drives that device's context.
//   * galcamb_multi_params_to_cls_ : parameter-batch sharding.  The points are pre-staged once on the host threads, the
//     valid ones are dealt to the devices in contiguous blocks, every device runs stages 1-4 on its block and writes its
//     C_l rows straight into the caller's array -- the path has no data exchange between devices at all.
//   * galcamb_multi_single_cls_    : ONE spectrum shared by the devices (the chain-step call).  Source wavenumbers are
//     dealt round-robin in cost order (kk mod G; low k = long integrations first, camb/cmbmain.f90:201), the Src rows
//     are merged with an NCCL all-reduce over NVLink (the one coupling of the k loop, camb/cmbmain.f90:1274: the spline
//     along k needs every k), every device integrates a slice of the ~2700 integration wavenumbers (:263-267) and forms
//     the partial sums of C_l over its slice, a second all-reduce of iCl (nl x ntype doubles) completes them.
// NCCL is loaded at run time (libnccl.so.2) the first time the second call is used with more than one device.
namespace {

struct NcclApi {
  void* h = nullptr;
  int (*CommInitAll)(void**, int, const int*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::vector<void*> comms;
  bool load() {
    if (h) return true;
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
    CommInitAll = (int (*)(void**, int, const int*))dlsym(h, "ncclCommInitAll");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
    CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    return CommInitAll && AllReduce && CommDestroy;
  }
} g_nccl;
const int kNcclFloat64 = 8, kNcclSum = 0;  // ncclDataType_t / ncclRedOp_t values of nccl.h

// The workers agree before every collective: a device that failed must not leave the others waiting inside NCCL.
struct Agreement {
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0, generation = 0, world;
  bool all_ok = true, result = true;
  explicit Agreement(int w) : world(w) {}
  bool arrive(bool ok) {  // true iff every worker arrived with ok
    std::unique_lock<std::mutex> lk(mu);
    all_ok = all_ok && ok;
    const int gen = generation;
    if (++arrived == world) {
      result = all_ok;
      arrived = 0;
      all_ok = true;
      generation++;
      cv.notify_all();
      return result;
    }
    cv.wait(lk, [&] { return generation != gen; });
    return result;
  }
};

// bring the calling worker's context up to date with the templates the host registered on any context
int sync_templates() {
  std::vector<double> t, pp;
  int lt = 0, lp = 0;
  unsigned long v;
  {
    std::lock_guard<std::mutex> lk(g_shared_tmpl.mu);
    v = g_shared_tmpl.version;
    if (g.tmpl_version == v) return 0;
    t = g_shared_tmpl.tmpl;
    pp = g_shared_tmpl.pp;
    lt = g_shared_tmpl.lmax_tmpl;
    lp = g_shared_tmpl.lmax_pp;
  }
  // (the two setters below take the same mutex and bump the version; re-stamp afterwards)
  if (int rc = galcamb_set_template_(t.empty() ? nullptr : t.data(), &lt)) return rc;
  if (int rc = galcamb_set_lensing_template_(pp.empty() ? nullptr : pp.data(), &lp)) return rc;
  return 0;
}

int set_bessels_for_point(int i, double boost) {
  int nl = 0, kmaxfile = 0;
  galcamb_prestage_lsamples_(&i, &nl, nullptr, &kmaxfile);
  std::vector<int> ls(nl);
  galcamb_prestage_lsamples_(&i, &nl, ls.data(), &kmaxfile);
  return galcamb_set_bessels_(ls.data(), &nl, &kmaxfile, &boost);
}

}  // namespace

int galcamb_device_count_(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int galcamb_multi_params_to_cls_(const int* ngpu, const galcamb_params* params, const int* n, const int* nthreads, double* Cl_out) {
  const int G = *ngpu;
  if (G < 1 || G > galcamb_device_count_()) {
    g_none.err = "galcamb_multi_params_to_cls_: ngpu out of range";
    return 5;
  }
  Ctx* const caller = tl_ctx;
  const int rc_pre = galcamb_prestage_(params, n, nthreads);
  std::vector<galcamb_model> models;
  std::vector<int> where;
  for (int i = 0; i < *n; i++) {
    galcamb_model m;
    if (galcamb_prestage_model_(&i, &m) == 0) {
      models.push_back(m);
      where.push_back(i);
    }
  }
  const size_t per = (size_t)3 * (params[0].Max_l - 1);
  for (size_t j = 0; j < per * (size_t)*n; j++) Cl_out[j] = std::nan("");
  if (models.empty()) return rc_pre ? rc_pre : 5;
  const int nm = (int)models.size();
  const int old_threads = galcamb_get_host_threads_();
  int per_dev_threads = std::max(1, (old_threads > 0 ? old_threads : (int)std::thread::hardware_concurrency()) / G);
  galcamb_set_host_threads_(&per_dev_threads);
  std::vector<int> rcs(G, 0);
  std::vector<std::string> errs(G);
  std::vector<std::thread> workers;
  for (int r = 0; r < G; r++)
    workers.emplace_back([&, r]() {
      const int lo = (int)((long)nm * r / G), hi = (int)((long)nm * (r + 1) / G);
      if (hi <= lo) return;
      int rc = galcamb_init_(&r);
      if (!rc) rc = sync_templates();
      if (!rc) rc = set_bessels_for_point(where[lo], models[lo].AccuracyBoost);
      const int cnt = hi - lo;
      std::vector<double> out(per * cnt);
      if (!rc) rc = galcamb_batch_cls_(models.data() + lo, &cnt, out.data());
      if (rc == 0 || rc == 4)
        for (int j = 0; j < cnt; j++) std::copy(out.begin() + per * j, out.begin() + per * (j + 1), Cl_out + per * where[lo + j]);
      rcs[r] = rc;
      if (rc) errs[r] = g.err;
    });
  for (auto& w : workers) w.join();
  galcamb_set_host_threads_(&old_threads);
  tl_ctx = caller;
  for (int r = 0; r < G; r++)
    if (rcs[r] && rcs[r] != 4) {
      (caller ? *caller : g_none).err = "device " + std::to_string(r) + ": " + errs[r];
      return rcs[r];
    }
  for (int r = 0; r < G; r++)
    if (rcs[r]) return rcs[r];
  return rc_pre;
}

int galcamb_multi_single_cls_(const int* ngpu, const galcamb_params* params, double* Cl_out, double* ms_out) {
  const int G = *ngpu;
  if (G < 1 || G > galcamb_device_count_()) {
    g_none.err = "galcamb_multi_single_cls_: ngpu out of range";
    return 5;
  }
  Ctx* const caller = tl_ctx;
  const auto t_begin = std::chrono::steady_clock::now();
  int one = 1, zero = 0;
  if (int rc = galcamb_prestage_(params, &one, &one)) return rc;
  galcamb_model m;
  if (int rc = galcamb_prestage_model_(&zero, &m)) return rc;
  const auto t_pre = std::chrono::steady_clock::now();
  if (G > 1) {
    if (!g_nccl.load()) {
      (caller ? *caller : g_none).err = "libnccl.so.2 not found";
      return 100;
    }
    if ((int)g_nccl.comms.size() != G) {
      for (void* c : g_nccl.comms) g_nccl.CommDestroy(c);
      g_nccl.comms.assign(G, nullptr);
      for (int r = 0; r < G; r++) {  // every device needs a context (and a primary CUDA context) before the communicators
        if (int rc = galcamb_init_(&r)) return rc;
      }
      std::vector<int> devs(G);
      for (int r = 0; r < G; r++) devs[r] = r;
      const int e = g_nccl.CommInitAll(g_nccl.comms.data(), G, devs.data());
      if (e != 0) {
        g_nccl.comms.clear();
        (caller ? *caller : g_none).err = std::string("ncclCommInitAll: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "?");
        return 100;
      }
    }
  }
  const size_t per = (size_t)3 * (m.lmax - 1);
  std::vector<int> rcs(G, 0);
  std::vector<std::string> errs(G);
  std::vector<double> k1_ms(G, 0.0), los_ms(G, 0.0);
  Agreement agree(G);
  std::vector<std::thread> workers;
  for (int r = 0; r < G; r++)
    workers.emplace_back([&, r]() {
      auto run = [&]() -> int {
        int rc0 = galcamb_init_(&r);
        if (!rc0) rc0 = sync_templates();
        if (!rc0) rc0 = set_bessels_for_point(0, m.AccuracyBoost);
        if (!rc0) rc0 = galcamb_batch_begin_(&one);
        if (!rc0) rc0 = galcamb_set_model_(&zero, &m);
        if (rc0) {  // still take part in the first agreement (everybody leaves there) so that no device waits for this one
          if (G > 1) agree.arrive(false);
          return rc0;
        }
        Ctx& c = g;
        const DevModel& D = c.slots[0].host;
        const size_t nsrc = (size_t)D.nk * D.SourceNum * D.nt;
        c.kshard_rank = r;
        c.kshard_world = G;
        c.q_lo = (int)((long)D.nq * r / G);
        c.q_hi = (int)((long)D.nq * (r + 1) / G);
        c.cls_what = 1;
        struct Restore {
          Ctx& c;
          ~Restore() {
            c.kshard_rank = 0;
            c.kshard_world = 1;
            c.q_lo = 0;
            c.q_hi = -1;
            c.cls_what = 3;
          }
        } restore{c};
        int rc = 0;
        if (G > 1 && cudaMemsetAsync(D.Src, 0, nsrc * sizeof(double), c.stream) != cudaSuccess) rc = fail(100, "cudaMemsetAsync(Src)");
        if (!rc) rc = galcamb_evolve_sources_();  // rows of the other devices' k stay zero and are summed in below
        k1_ms[r] = c.ms[0];
        if (G > 1) {
          if (!agree.arrive(rc == 0)) return rc ? rc : fail(100, "another device failed");
          const int e = g_nccl.AllReduce(D.Src, D.Src, nsrc, kNcclFloat64, kNcclSum, g_nccl.comms[r], c.stream);
          if (e != 0) rc = fail(100, "ncclAllReduce(Src) failed");
        }
        if (!rc) rc = galcamb_los_integrate_();
        los_ms[r] = c.ms[2];
        if (!rc) rc = galcamb_scalar_cls_();  // partial sums over this device's q slice
        if (G > 1) {
          if (!agree.arrive(rc == 0)) return rc ? rc : fail(100, "another device failed");
          const int e = g_nccl.AllReduce(D.iCl, D.iCl, (size_t)c.slots[0].ntype * c.nl, kNcclFloat64, kNcclSum, g_nccl.comms[r], c.stream);
          if (e != 0) rc = fail(100, "ncclAllReduce(iCl) failed");
        }
        if (rc) return rc;
        if (r == 0) {
          c.cls_what = 2;
          if (int rc2 = galcamb_scalar_cls_()) return rc2;
          CK(cudaMemcpyAsync(Cl_out, D.Cl, per * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        }
        CK(cudaStreamSynchronize(c.stream));
        return 0;
      };
      rcs[r] = run();
      if (rcs[r]) errs[r] = g.err;
    });
  for (auto& w : workers) w.join();
  tl_ctx = caller;
  const auto t_end = std::chrono::steady_clock::now();
  if (ms_out) {  // [0] whole call, [1] host pre-stage, [2] slowest device's K1, [3] slowest device's LOS slice
    ms_out[0] = std::chrono::duration<double, std::milli>(t_end - t_begin).count();
    ms_out[1] = std::chrono::duration<double, std::milli>(t_pre - t_begin).count();
    ms_out[2] = *std::max_element(k1_ms.begin(), k1_ms.end());
    ms_out[3] = *std::max_element(los_ms.begin(), los_ms.end());
  }
  for (int r = 0; r < G; r++)
    if (rcs[r]) {
      (caller ? *caller : g_none).err = "device " + std::to_string(r) + ": " + errs[r];
      return rcs[r];
    }
  return 0;
}

}  // extern "C"
