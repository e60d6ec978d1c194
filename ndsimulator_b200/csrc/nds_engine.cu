// nds_engine.cu -- host side of the C ABI (include/ndsb200.h): configuration checks, device memory,
// the launch loop that replaces NDRun.run (ndrun.py:214-251) with fused K-step launches, the
// metadynamics deposition schedule (bias/gaus.py:130-146) and the data movers.
//
// There is no CPU path in this file: every entry point that computes launches a kernel.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <vector>

#include <dlfcn.h>

#include <cub/device/device_scan.cuh>

#include "nds_variants.h"
#include "nds_mc_kernel.cuh"
#include "nds_grid_kernels.cuh"
#include "nds_map50.cuh"

namespace nds { constexpr int MTD_TILE_HOST = 2048; }  // >= mtd_tile_records<R>() of nds_mtd_kernel.cuh

using namespace nds;

// ---------------------------------------------------------------------------------- errors
static thread_local std::string g_last_error;

struct nds_engine {
  std::string err;
  virtual ~nds_engine() {}
  virtual int set_state(const double* x, const double* v) = 0;
  virtual int set_walker_bias(const double* k, const double* r0) = 0;
  virtual int set_noise(const double* noise, int64_t first, int64_t nsteps) = 0;
  virtual int set_path(int64_t n_ref, int path_dim, const double* ref, const double* ind, double sigma) = 0;
  virtual int begin() = 0;
  virtual int run(int64_t nsteps) = 0;
  virtual int sync() = 0;
  virtual int get_state(double* x, double* v, double* f, double* fixf, double* colv, double* thermo) = 0;
  virtual int set_state_f32(const float* x, const float* v) = 0;
  virtual int get_state_f32(float* x, float* v, float* f, float* fixf, float* colv, float* thermo) = 0;
  virtual int eval_at(const double* x, int64_t n, double* V, double* F, double* colv, double* J,
                      double* Vb, double* Fb) = 0;
  virtual int get_hills(int64_t walker, int64_t* n, double* centers, double* heights, int64_t cap) = 0;
  virtual int set_hills(int64_t walker, int64_t n, const double* centers, const double* heights) = 0;
  virtual int64_t hill_table_size(int64_t n) const = 0;
  virtual int get_hill_table_raw(double* out, int64_t cap) = 0;
  virtual int set_hill_table_raw(int64_t n, const double* in) = 0;
  virtual int get_hill_sigma2(int64_t walker, double* sigma2, int64_t cap) = 0;
  virtual int set_hill_sigma2(int64_t walker, int64_t n, const double* sigma2) = 0;
  virtual int get_vr(double* vr) = 0;
  virtual int get_commit(int32_t* basin, int64_t* exit_step) = 0;
  virtual int get_alive(uint8_t* alive) = 0;
  virtual int set_commit(const int32_t* basin, const int64_t* exit_step, const uint8_t* alive) = 0;
  virtual int get_mc(int64_t* accepted) = 0;
  virtual int mc_sync_walker(int64_t walker) = 0;
  virtual int get_wl_histogram(int64_t walker, double* hE) = 0;
  virtual int set_wl_histogram(int64_t walker, const double* hE) = 0;
  virtual int get_clock(int64_t* step, double* time, int64_t* dep, int64_t* resc) const = 0;
  virtual int set_clock(int64_t step, double time, int64_t dep, int64_t resc) = 0;
  virtual int set_vr(const double* vr) = 0;
  virtual int64_t count_nonfinite() = 0;
  virtual int64_t count_alive() = 0;
  virtual int drain_frames(double* out, int64_t max_frames, int64_t* n_frames) = 0;
  virtual int fes_grid(int64_t walker, double x0, double dx, int64_t nx, double y0, double dy,
                       int64_t ny, double* grid) = 0;
  virtual int histogram(double x0, double dx, int64_t nx, double y0, double dy, int64_t ny,
                        int64_t* counts) = 0;
  virtual int device_pointers(void** x, void** v, void** hills, int64_t* stride) = 0;
  virtual int hills_ipc_handle(void* handle) = 0;
  virtual int set_peer_tables(void* const* tables, int nranks, int rank, bool opened_by_ipc) = 0;
  virtual double last_run_ms() = 0;
  virtual int set_comm(const void* unique_id, void* comm, int rank, int nranks) = 0;
  virtual int set_peer_barrier(int on_device) = 0;
  virtual int get_mtd_grid(double* out, int64_t cap) = 0;
  virtual int set_mtd_grid(const double* in, int64_t n) = 0;
  virtual int64_t mtd_grid_outside() = 0;
  virtual int get_mc_state(int64_t* accepted, double* time, int64_t* dep, double* wl, int64_t* max_dep) = 0;
  virtual int set_mc_state(const int64_t* accepted, const double* time, const int64_t* dep, const double* wl) = 0;
  virtual int set_thermo(const double* thermo) = 0;
  virtual int get_umb_vr(double* vr, int64_t cap, int64_t* n) = 0;
  virtual int set_umb_vr(const double* vr, int64_t n) = 0;
  virtual int get_cv_history(double* ring, int64_t cap_elems, int64_t* n_samples, int64_t* window) = 0;
  virtual int set_cv_history(const double* ring, int64_t n_samples) = 0;
  virtual int submit(const void* host_in, int64_t nsteps, void* host_out, int chunks) = 0;
  virtual int exchange_info(double* exchange_ms, int64_t* n_exchanges) = 0;
  virtual int get_histogram(int64_t* counts, int clear) = 0;
  virtual int frames_f32() const = 0;
  virtual int table_checksum(uint64_t* sum, int64_t first, int64_t n) = 0;

  nds_config cfg;
  int cd = 0, tcd = 0;
  int64_t step = 0;
  double time = 0.0;
  int64_t n_hills = 0;
  int64_t launches = 0;
  int64_t frames_written = 0;  // frames in the ring since the last drain
  int frame_width = 0;
  cudaStream_t stream = nullptr;
  const char* kernel_name = "";
  nds_exchange_fn exchange = nullptr;
  void* exchange_user = nullptr;
};

static int fail(nds_engine* e, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  if (e) e->err = buf;
  return 1;
}

#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess)                                                                \
      return fail(this, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, \
                  __LINE__, #call);                                                       \
  } while (0)


// ------------------------------------------------------------------------------- NCCL (dlopen)
// The library does not link against NCCL: the few entry points it needs are resolved at run time from the
// libnccl.so.2 already loaded in the process (torch's bundled copy) or found by the dynamic loader
// (NDS_NCCL_LIB overrides the name).  Types are restated from nccl.h (stable ABI: ncclUniqueId is 128 bytes,
// ncclChar = 0, ncclFloat32 = 7, ncclFloat64 = 8, ncclSum = 0).
struct NcclApi {
  typedef struct ncclComm* comm_t;
  struct unique_id { char internal[128]; };
  int (*GetUniqueId)(unique_id*) = nullptr;
  int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, comm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  void* handle = nullptr;
  std::string why;
  bool ok = false;
};

static NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  const char* names[] = {getenv("NDS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm) continue;
    api.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);  // the copy that is already mapped (torch's), if any
    if (!api.handle) api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) { api.why = "libnccl.so.2 not found (set NDS_NCCL_LIB)"; return api; }
  auto sym = [&](const char* n) { void* p = dlsym(api.handle, n); if (!p) api.why = std::string("missing symbol ") + n; return p; };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.AllReduce && api.GetErrorString;
  return api;
}

// ------------------------------------------------------------------------------ validation
static const int kPotNdim[NDS_POT_COUNT] = {2, 5, 5, 1, 2, 2, 5, 2, 2, 1, 2, 0};  // 0: any supported ndim
static const bool kPotNeedsTcv[NDS_POT_COUNT] = {false, true, true, false, false, true, true,
                                                 false, false, false, false};

extern "C" int nds_colvar_dim(int cv, int ndim) {
  switch (cv) {
    case NDS_CV_ORIGINAL: return (ndim == 1 || ndim == 2 || ndim == 5) ? ndim : -1;
    case NDS_CV_EXTRACTX: case NDS_CV_EXTRACTY: case NDS_CV_XMY: case NDS_CV_XPY:
    case NDS_CV_PROJ2DTO1D: return ndim == 2 ? 1 : -1;
    case NDS_CV_ROTATE2D: return ndim == 2 ? 2 : -1;
    case NDS_CV_PROJ5DTO1DX0: case NDS_CV_PROJ5DTO1DX2: case NDS_CV_PROJ5DTO1D:
      return ndim == 5 ? 1 : -1;
    case NDS_CV_PROJ5DTO2D: case NDS_CV_PROJ5DTO2DV2: return ndim == 5 ? 2 : -1;
    case NDS_CV_PATH: return (ndim == 1 || ndim == 2 || ndim == 5) ? 1 : -1;
    case NDS_CV_PATH5DTO2D: return ndim == 5 ? 1 : -1;
    default: return -1;
  }
}

// Gaussian.tao_d = np.min([dep_freq * 50, 100]) (gaus.py:94): window of the colvar history; 0 if not an integer
static int nds_tao_d(const nds_config* c) {
  const double t = c->mtd_dep_freq * 50 < 100 ? c->mtd_dep_freq * 50 : 100;
  return (t >= 1 && t == std::floor(t)) ? (int)t : 0;
}

static int validate(const nds_config* c) {
  if (!c) return fail(nullptr, "nds_create: null config");
  if (c->abi_version != NDS_ABI_VERSION)
    return fail(nullptr, "nds_create: abi_version %d != %d", c->abi_version, NDS_ABI_VERSION);
  if (c->precision != NDS_F32 && c->precision != NDS_F64) return fail(nullptr, "bad precision");
  if (c->ndim != 1 && c->ndim != 2 && c->ndim != 5)
    return fail(nullptr, "ndim=%d: the built-in potentials have ndim 1, 2 or 5", c->ndim);
  if (c->potential < 0 || c->potential >= NDS_POT_COUNT) return fail(nullptr, "unknown potential id %d", c->potential);
  if (kPotNdim[c->potential] != 0 && kPotNdim[c->potential] != c->ndim)  // potentials/potential.py:12 (assert pointer.ndim == self.ndim)
    return fail(nullptr, "potential %d has ndim=%d, run has ndim=%d", c->potential, kPotNdim[c->potential], c->ndim);
  if (nds_colvar_dim(c->colvar, c->ndim) < 0) return fail(nullptr, "colvar %d is not defined for ndim=%d", c->colvar, c->ndim);
  if (nds_colvar_dim(c->true_colvar, c->ndim) < 0) return fail(nullptr, "true_colvar %d is not defined for ndim=%d", c->true_colvar, c->ndim);
  if (c->true_colvar == NDS_CV_PATH || c->true_colvar == NDS_CV_PATH5DTO2D)
    return fail(nullptr, "path colvars are supported as colvar_name, not as true_colvar_name");
  if (kPotNeedsTcv[c->potential] && nds_colvar_dim(c->true_colvar, c->ndim) != 2)
    return fail(nullptr, "potential %d is evaluated in a 2-D true_colvar space", c->potential);
  if (c->integrator < 0 || c->integrator > NDS_INT_MINIMIZE)  // md.py:59
    return fail(nullptr, "unknown integrator %d", c->integrator);
  if (c->integrator == NDS_INT_RESCALE && !(c->rescale_freq > 0))  // md.py:220 (TypeError on None)
    return fail(nullptr, "integrate=rescale needs rescale_freq > 0");
  if (c->method < NDS_METHOD_MD || c->method > NDS_METHOD_WL) return fail(nullptr, "unknown method %d", c->method);
  if (c->method == NDS_METHOD_MC || c->method == NDS_METHOD_WL) {
    if (c->mtd_enabled && c->mtd_shared)  // ndrun.py:234-237: time (hence the deposition schedule) is per walker
      return fail(nullptr, "metadynamics under MC needs mtd_shared = 0: NDRun.time advances on accepted moves only, "
                           "so every walker deposits on its own clock");
    for (int d = 0; d < c->ndim; d++)
      if (c->mc_periodic && !(c->mc_global_bounds[d] != 0)) return fail(nullptr, "mc_global_bounds[%d] must be non-zero", d);
  }
  if (c->n_walkers <= 0) return fail(nullptr, "n_walkers must be positive");
  if (c->n_walkers_global < c->walker_offset + c->n_walkers) return fail(nullptr, "n_walkers_global too small");
  if (!(c->mass > 0) || !(c->dt > 0)) return fail(nullptr, "mass and dt must be positive");
  if (c->n_umb < 0 || c->n_umb > NDS_MAX_UMB) return fail(nullptr, "n_umb out of range");
  if (c->umb_per_walker && c->n_umb < 1) return fail(nullptr, "umb_per_walker needs n_umb >= 1");
  if (c->mtd_enabled) {
    if (!(c->mtd_biasf > 1.0))  // gaus.py:102-103,187 divide by kB*(biasf-1)*T
      return fail(nullptr, "mtd_biasf must be > 1");
    if (!(c->mtd_dep_freq > 0)) return fail(nullptr, "mtd_dep_freq must be positive");
    if (c->mtd_capacity <= 0) return fail(nullptr, "mtd_capacity must be positive");
    for (int k = 0; k < nds_colvar_dim(c->colvar, c->ndim); k++)
      if (!(c->mtd_sigma[k] > 0)) return fail(nullptr, "mtd_sigma[%d] must be positive", k);
    if (c->mtd_adapsig_geo && nds_colvar_dim(c->colvar, c->ndim) != 1)
      return fail(nullptr, "mtd_adapsig_geo needs a 1-D colvar: for colvardim > 1 the reference takes the "
                           "element-wise reciprocal of J.J^T (gaus.py:268-275), which is inf / NaN");
    if (c->mtd_grid) {
      const int gcd_ = nds_colvar_dim(c->colvar, c->ndim);
      if (!c->mtd_shared) return fail(nullptr, "mtd_grid needs the shared hill table (mtd_shared = 1)");
      if (c->mtd_adapsig_geo || c->mtd_adapsig_hist) return fail(nullptr, "mtd_grid needs a fixed sigma (no adapsig)");
      if (gcd_ > 2) return fail(nullptr, "mtd_grid: colvardim %d > 2", gcd_);
      if (c->method != NDS_METHOD_MD && c->method != NDS_METHOD_COMMITTOR) return fail(nullptr, "mtd_grid needs method md / committor");
      if (c->n_walkers_global % c->n_walkers != 0 || c->walker_offset % c->n_walkers != 0)
        return fail(nullptr, "mtd_grid needs the same number of walkers on every rank");
      if (c->n_walkers_global / c->n_walkers > nds::NDS_MAX_RANKS) return fail(nullptr, "mtd_grid: more than %d ranks", nds::NDS_MAX_RANKS);
      for (int k = 0; k < gcd_; k++) {
        if (c->mtd_grid_n[k] < 2) return fail(nullptr, "mtd_grid_n[%d] must be >= 2", k);
        if (!(c->mtd_grid_hi[k] > c->mtd_grid_lo[k])) return fail(nullptr, "mtd_grid_hi[%d] must exceed mtd_grid_lo[%d]", k, k);
      }
      const long long pts = (long long)c->mtd_grid_n[0] * (gcd_ > 1 ? c->mtd_grid_n[1] : 1);
      if (pts > (1 << 24)) return fail(nullptr, "mtd_grid: %lld grid points > 2^24", pts);
    }
  }
  if (c->mtd_enabled && c->mtd_adapsig_hist) {  // gaus.py:166-183, 193-212: variance from the Stat history of the colvar
    if (c->mtd_adapsig_geo) return fail(nullptr, "mtd_adapsig_hist and mtd_adapsig_geo exclude each other (gaus.py:163,191)");
    if (nds_colvar_dim(c->colvar, c->ndim) != 1)
      return fail(nullptr, "mtd_adapsig needs a 1-D colvar: for colvardim > 1 the reference sums exp(-dc_i dc_j / (2 S_ij)) over "
                           "the ELEMENTS of the covariance matrix (gaus.py:268-277), which diverges for negative covariances");
    if (c->method != NDS_METHOD_MD) return fail(nullptr, "mtd_adapsig (history-based) needs method md");
    if (nds_tao_d(c) < 1) return fail(nullptr, "mtd_adapsig: tao_d = min(50 * mtd_dep_freq, 100) must be a positive integer "
                                                "(gaus.py:94; the reference slices stat.colv with it)");
  }
  if (c->method == NDS_METHOD_COMMITTOR) {
    if (c->n_basins < 1 || c->n_basins > NDS_MAX_BASINS) return fail(nullptr, "n_basins out of range");
    if (c->temperature == 0 && c->integrator != NDS_INT_MINIMIZE)  // engines/__init__.py:22-23
      return fail(nullptr, "committor at temperature 0 runs the Minimize engine: set integrator = NDS_INT_MINIMIZE");
  }
  if (c->steps_per_launch <= 0) return fail(nullptr, "steps_per_launch must be positive");
  if (c->dump_freq > 0 && c->dump_walkers > 0 && c->dump_capacity <= 0)
    return fail(nullptr, "dump_capacity must be positive when frames are recorded");
  if (c->hist_n[0] > 0) {
    const int hcd = nds_colvar_dim(c->colvar, c->ndim);
    if (hcd > 2) return fail(nullptr, "hist_n: colvardim %d > 2", hcd);
    if (c->dump_freq <= 0) return fail(nullptr, "hist_n needs dump_freq > 0 (the histogram is sampled at dump cadence)");
    if (c->method != NDS_METHOD_MD && c->method != NDS_METHOD_COMMITTOR) return fail(nullptr, "hist_n needs method md / committor");
    const long long nb = (long long)c->hist_n[0] * (hcd > 1 ? c->hist_n[1] : 1);
    if (nb <= 0 || nb > 12288) return fail(nullptr, "hist_n: %lld bins do not fit the shared-memory histogram (max 12288)", nb);
    for (int k = 0; k < hcd; k++)
      if (!(c->hist_hi[k] > c->hist_lo[k])) return fail(nullptr, "hist_hi[%d] must exceed hist_lo[%d]", k, k);
  }
  return 0;
}

// ------------------------------------------------------------------------- constant block
static const double kB_ = 8.6173303e-5;  // constant.py:2
static double au_mass_() {               // constant.py:12
  const double eV = 1.60217733e-19, angstrom = 1.0e-10, fs = 1e-15;
  return 1.67e-27 / (eV / angstrom * fs / angstrom * fs);
}

template <typename R>
static void fill_consts(const nds_config& c, Consts<R>& K) {
  memset(&K, 0, sizeof K);
  const double kcal = 0.04337209302, escale = 0.2, lscale = 16.0;  // constant.py:10-15
  K.ndim = c.ndim;
  K.cd = nds_colvar_dim(c.colvar, c.ndim);
  K.tcd = nds_colvar_dim(c.true_colvar, c.ndim);
  K.pot = c.potential; K.cv = c.colvar; K.tcv = c.true_colvar; K.integ = c.integrator;
  K.method = c.method; K.noise = c.noise_mode; K.n_umb = c.n_umb; K.n_basins = c.n_basins;
  K.bias_mask = (c.n_umb > 0 ? BIAS_UMB : 0) | (c.mtd_enabled ? BIAS_MTD : 0) |
                (c.pe_enabled ? BIAS_PE : 0) | ((c.n_umb > 0 && c.umb_per_walker) ? BIAS_UMB_WALKER : 0);
  K.mtd_shared = c.mtd_shared;
  K.mtd_adaptive = (c.mtd_enabled && (c.mtd_adapsig_geo || c.mtd_adapsig_hist)) ? 1 : 0;
  if (c.mtd_enabled && c.mtd_adapsig_hist) K.bias_mask |= BIAS_CVHIST;
  K.mtd_dep_freq64 = c.mtd_dep_freq; K.mtd_capacity = c.mtd_capacity;
  // 4-word records (c0, c1 | 0, w, log2 w) for every shared table with colvardim <= 2 (mtd_block streams them
  // through TMA tiles); adaptive records are (c, w, 1/s, s); everything else is (c_0..c_{cd-1}, w)
  const bool rec4 = !K.mtd_adaptive && (K.cd == 2 || (K.cd == 1 && c.mtd_shared));
  K.hill_stride = (rec4 || K.mtd_adaptive) ? 4 : K.cd + 1;
  K.hill_hidx = rec4 ? 2 : K.cd;
  const double m = au_mass_() * c.mass;        // data.py:27-28
  const double kBT = kB_ * c.temperature;      // data.py:26
  // Minimize multiplies the configured time step by 10 (engines/minimize.py:18)
  const double dt = (c.integrator == NDS_INT_MINIMIZE) ? c.dt * 10 : c.dt, gamma = c.gamma;
  K.m = (R)m; K.inv_m = (R)(1.0 / m); K.dt = (R)dt; K.kB = (R)kB_; K.kBT = (R)kBT;
  K.dt64 = dt; K.rescale_freq = c.rescale_freq;
  K.hdtm = (R)(dt / (2.0 * m));
  K.vel_scale = (R)std::sqrt(kB_ * c.initial_temperature / m);  // modify.py:45
  if (c.integrator == NDS_INT_LANGEVIN) {      // md.py:83-86
    const double c1 = std::exp(-gamma * dt / 2.0);
    K.c1 = (R)c1; K.c2 = (R)std::sqrt((1 - c1 * c1) * kBT / m);
    K.bm_c2 = bm_scale(std::sqrt((1 - c1 * c1) * kBT / m));
  } else if (c.integrator == NDS_INT_LANGEVIN2) {  // md.py:87-97
    const double frdt = (1.0 - std::exp(-gamma * dt / 2.0)) * 8.0;
    const double sigma = std::sqrt(2 * kBT * frdt / m);
    const double c5 = std::sqrt(dt) * sigma / (2 * std::sqrt((double)c.ndim));
    K.c1 = (R)(dt / 2.0 - dt * frdt / 8.0);
    K.c2 = (R)(frdt / 2 - frdt * frdt / 8.0);
    K.c3 = (R)(std::sqrt(dt) * sigma / 2.0 * (1.0 - frdt / 4.0));
    K.c5 = (R)c5;
    K.c4 = (R)(frdt / 2.0 * c5);
    K.l2_a = (R)(1.0 - (frdt / 2 - frdt * frdt / 8.0));
    K.l2_c1m = (R)((dt / 2.0 - dt * frdt / 8.0) / m);
    K.l2_c5dt = (R)(c5 * dt);
  } else if (c.integrator == NDS_INT_RESCALE) {  // md.py:76-79
    K.v_target = (R)std::sqrt(kBT * c.ndim / m);
  }
  double A[4] = {0, 0, 0, 0}, a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}, cc[4] = {0, 0, 0, 0};
  double gx[4] = {0, 0, 0, 0}, gy[4] = {0, 0, 0, 0};
  const int pot = c.potential;
  if (pot == NDS_POT_MUELLER2D || pot == NDS_POT_MUELLER5D || pot == NDS_POT_MUELLER5DSHL) {
    const double A0[4] = {-200, -100, -170, 15}, a0[4] = {-1, -1, -6.5, 0.7};  // mueller.py:8-11
    const double b0[4] = {0, 0, 11, 0.6}, c0[4] = {-10, -10, -6.5, 0.7};
    for (int i = 0; i < 4; i++) {
      A[i] = A0[i] * kcal * escale;
      if (pot == NDS_POT_MUELLER5DSHL) A[i] = A[i] / 5;  // mueller.py:157
      a[i] = a0[i] / (lscale * lscale); b[i] = b0[i] / (lscale * lscale); cc[i] = c0[i] / (lscale * lscale);
      gx[i] = c.pot_params[i]; gy[i] = c.pot_params[4 + i];  // mueller.py:15-16
    }
  } else if (pot == NDS_POT_THREEHOLE2D || pot == NDS_POT_THREEHOLE5D) {  // three_hole.py:10-22
    const double es = 2.0, ls = 4.0;
    const double A0[4] = {3, -3, -5, -5}, x0r[4] = {0, 0, 1, -1}, y0r[4] = {1 / 3.0, 5 / 3.0, 0, 0};
    for (int i = 0; i < 4; i++) {
      A[i] = A0[i] / es; a[i] = -1 / (ls * ls); b[i] = 0 / (ls * ls); cc[i] = -1 / (ls * ls);
      gx[i] = x0r[i] * ls + 7.5; gy[i] = y0r[i] * ls + 7.5;
    }
    K.th_D = (R)(0.2 / std::pow(ls, 4) / es); K.th_E = (R)(0.2 / std::pow(ls, 4) / es);
    K.th_x1 = (R)(0 * ls + 7.5); K.th_y1 = (R)(1 / 3.0 * ls + 7.5); K.th_ref = (R)2;
  } else if (pot == NDS_POT_GAUSSIAN2D) {  // misc.py:9-13
    A[0] = -200.0 * kcal * escale; a[0] = -1.0 / (lscale * lscale); cc[0] = -10.0 / (lscale * lscale);
  }
  const double kfast = 0.5 * 1.4426950408889634;  // log2(e) / 2
  for (int i = 0; i < 4; i++) {
    K.gA[i] = (R)A[i]; K.ga[i] = (R)a[i]; K.gb[i] = (R)b[i]; K.gc[i] = (R)cc[i];
    K.gx[i] = (R)gx[i]; K.gy[i] = (R)gy[i];
    K.fa[i] = (R)(2 * a[i] * kfast); K.fb[i] = (R)(b[i] * kfast); K.fc[i] = (R)(2 * cc[i] * kfast);
    K.fA[i] = (R)(A[i] / kfast);
    K.flA[i] = (R)(A[i] != 0 ? std::log2(std::fabs(A[i] / kfast)) : 0.0);
  }
  for (int d = 0; d < 5; d++) { K.hq_k[d] = (R)c.pot_params[d]; K.hq_q[d] = (R)c.pot_params[5 + d]; K.hq_c[d] = (R)c.pot_params[10 + d]; }
  K.rot = (R)(1.0 / std::sqrt(2.0));  // two.py:53
  if (pot == NDS_POT_DOUBLEWELL1D) {  // double_well.py:9-14
    K.dw_K1 = (R)c.pot_params[0]; K.dw_K2 = (R)c.pot_params[1];
    K.dw_x1[0] = (R)c.pot_params[2]; K.dw_x2[0] = (R)c.pot_params[3];
    K.dw_invsig2 = (R)(1.0 / c.pot_params[4] / c.pot_params[4]);
  } else if (pot == NDS_POT_DOUBLEWELL) {  // double_well.py:37-42
    K.dw_K1 = (R)c.pot_params[0]; K.dw_K2 = (R)c.pot_params[1];
    K.dw_x1[0] = (R)c.pot_params[2]; K.dw_x1[1] = (R)c.pot_params[3];
    K.dw_x2[0] = (R)c.pot_params[4]; K.dw_x2[1] = (R)c.pot_params[5];
    K.dw_invsig2 = (R)(1.0 / c.pot_params[6] / c.pot_params[6]);
  }
  for (int u = 0; u < NDS_MAX_UMB; u++)
    for (int k = 0; k < NDS_MAX_CV; k++) { K.umb_k[u][k] = (R)c.umb_k[u][k]; K.umb_r0[u][k] = (R)c.umb_r0[u][k]; }
  if (c.mtd_enabled) {
    K.mtd_w = (R)c.mtd_w;
    for (int k = 0; k < K.cd; k++) {
      K.mtd_invsig2[k] = (R)(1.0 / (c.mtd_sigma[k] * c.mtd_sigma[k]));  // gaus.py:90-91
      K.mtd_sig2[k] = (R)(c.mtd_sigma[k] * c.mtd_sigma[k]);
      K.mtd_pre[k] = (R)(std::sqrt(kfast) / c.mtd_sigma[k]);
    }
    for (int k = K.cd; k < NDS_MAX_CV; k++) { K.mtd_invsig2[k] = (R)0; K.mtd_pre[k] = (R)1; }  // unused components
    K.mtd_kBdT = (R)(kB_ * ((c.mtd_biasf - 1) * c.temperature));  // gaus.py:102-103
    if (c.mtd_grid) {
      K.bias_mask |= BIAS_MTD_GRID;
      K.grid_nx = c.mtd_grid_n[0];
      K.grid_ny = K.cd > 1 ? c.mtd_grid_n[1] : 1;
      for (int k = 0; k < 2; k++) {
        const int n = k == 0 ? K.grid_nx : K.grid_ny;
        const double h = n > 1 ? (c.mtd_grid_hi[k] - c.mtd_grid_lo[k]) / (n - 1) : 1.0;
        K.grid_lo[k] = (R)c.mtd_grid_lo[k]; K.grid_h[k] = (R)h; K.grid_invh[k] = (R)(1.0 / h);
      }
    }
  }
  K.pe_thred = (R)c.pe_thred;
  for (int bb = 0; bb < NDS_MAX_BASINS; bb++)
    for (int k = 0; k < NDS_MAX_CV; k++) { K.crit[bb][k][0] = (R)c.criteria[bb][k][0]; K.crit[bb][k][1] = (R)c.criteria[bb][k][1]; }
  K.diffusion_time = c.diffusion_time;
  K.min_gtol = (R)c.min_gtol; K.emin = (R)c.emin; K.min_maxiter = c.min_maxiter;
  K.zero_temperature = (c.temperature == 0) ? 1 : 0;
  const bool dumping = c.dump_freq > 0 && c.dump_walkers > 0;
  const bool hist = c.dump_freq > 0 && c.hist_n[0] > 0;
  if (hist) {
    K.hist_nx = c.hist_n[0];
    K.hist_ny = K.cd > 1 ? c.hist_n[1] : 1;
    for (int k = 0; k < 2; k++) {
      const int n = k == 0 ? K.hist_nx : K.hist_ny;
      const bool used = k < K.cd;
      K.hist_lo[k] = used ? (float)c.hist_lo[k] : 0.f;
      K.hist_inv[k] = used ? (float)(n / (c.hist_hi[k] - c.hist_lo[k])) : 1.f;
    }
  }
  K.dump_freq = (dumping || hist) ? c.dump_freq : 0;
  K.stat_freq = c.stat_freq > 0 ? c.stat_freq : 1;
  K.dump_walkers = dumping ? (c.dump_walkers < c.n_walkers ? c.dump_walkers : c.n_walkers) : 0;
  K.key0 = (unsigned)(c.seed & 0xffffffffull);
  K.key1 = (unsigned)(c.seed >> 32);
  K.rk_step = philox_round_keys(K.key0, K.key1);
  K.rk_init = philox_round_keys(K.key0, K.key1 ^ 1u);
  K.rk_mc = philox_round_keys(K.key0, K.key1 ^ 2u);
  for (int d = 0; d < NDS_MAX_DIM; d++) { K.mc_bounds[d] = (R)c.mc_bounds[d]; K.mc_gb[d] = (R)c.mc_global_bounds[d]; }
  K.mc_all = c.mc_propose_all; K.mc_periodic = c.mc_periodic;
}

// -------------------------------------------------------------------------- variant lookup
// dump: bit 0 = frames, bit 1 = in-kernel histogram (the DUMP field of Spec)
static const Variant* pick_variant(const nds_config& c, int bias_mask, int dump, int prec_override = -1) {
  typedef const Variant* (*table_fn)(int*);
  static const table_fn tables[] = {variants_hot, variants_grid,
                                    variants_generic_f32_nd1, variants_generic_f32_nd2,
                                    variants_generic_f32_nd5, variants_generic_f64_nd1,
                                    variants_generic_f64_nd2, variants_generic_f64_nd5};
  const Variant* best = nullptr;
  int best_score = -1;
  const int want[11] = {prec_override >= 0 ? prec_override : c.precision, c.ndim, c.potential, c.colvar,
                        c.true_colvar, c.integrator, bias_mask, c.method, c.noise_mode, dump,
                        c.n_basins};
  for (table_fn t : tables) {
    int n = 0;
    const Variant* v = t(&n);
    for (int i = 0; i < n; i++) {
      const int have[11] = {v[i].key.prec, v[i].key.nd, v[i].key.pot, v[i].key.cv, v[i].key.tcv,
                            v[i].key.integ, v[i].key.bias, v[i].key.method, v[i].key.noise, v[i].key.dump,
                            v[i].key.nbasin};
      bool ok = true;
      int score = 0;
      for (int k = 0; k < 11 && ok; k++) {
        if (have[k] == DYN) continue;
        if (have[k] != want[k]) ok = false; else score++;
      }
      if (ok && score > best_score) { best = &v[i]; best_score = score; }
    }
  }
  return best;
}

// shared-hill metadynamics: pick the mtd_block variant with the best SM fill for M walkers
static const MtdVariant* pick_mtd_variant(const nds_config& c, int bias_mask, int n_sm) {
  typedef const MtdVariant* (*table_fn)(int*);
  static const table_fn tables[] = {mtd_variants_f32_nd1, mtd_variants_f64_nd1, mtd_variants_f32_nd2,
                                    mtd_variants_f32_nd5, mtd_variants_f64_nd2, mtd_variants_f64_nd5};
  const int want[10] = {c.precision, c.ndim, c.potential, c.colvar, c.true_colvar, c.integrator,
                        bias_mask, c.method, c.noise_mode, 0};
  const MtdVariant* best = nullptr;
  long long best_cost = 0;
  int best_score = -1;
  for (table_fn t : tables) {
    int n = 0;
    const MtdVariant* v = t(&n);
    for (int i = 0; i < n; i++) {
      const int have[10] = {v[i].key.prec, v[i].key.nd, v[i].key.pot, v[i].key.cv, v[i].key.tcv,
                            v[i].key.integ, v[i].key.bias, v[i].key.method, v[i].key.noise, v[i].key.dump};
      bool ok = true;
      int score = 0;
      for (int k = 0; k < 10 && ok; k++) {
        if (have[k] == DYN) continue;
        if (have[k] != want[k]) ok = false; else score++;
      }
      if (!ok) continue;
      // time ~ waves x walkers-per-warp (every warp streams the whole table once per walker it owns)
      const long long ctas = (c.n_walkers + (long long)v[i].W * v[i].NW - 1) / ((long long)v[i].W * v[i].NW);
      // CTAs with a warp count that is not a multiple of 4 leave schedulers idle at every tile barrier
      const long long slot = (v[i].NW % 2 == 0) ? 100 : 115;
      long long cost = ((ctas + n_sm - 1) / n_sm) * v[i].W * slot;
      if (const char* force = getenv("NDS_MTD_VARIANT")) {  // experiments: e.g. "W8,NW7"
        char tag[32];
        snprintf(tag, sizeof tag, "W%d,NW%d", v[i].W, v[i].NW);
        if (strcmp(force, tag) == 0) cost = 0;
      }
      if (!best || cost < best_cost || (cost == best_cost && score > best_score)) {
        best = &v[i]; best_cost = cost; best_score = score;
      }
    }
  }
  return best;
}

static __global__ void philox_kat_kernel(u32x4 c, uint32_t k0, uint32_t k1, uint32_t* out) {
  const u32x4 r = philox4x32_10(c, k0, k1);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// Tail census of the step-noise stream (tests): every thread draws `calls` Philox calls of one
// walker through normals6<R> -- the function the step kernels use -- and counts |z| above each threshold.
constexpr int NDS_MAX_THR = 8;
struct TailArgs { double thr[NDS_MAX_THR]; int n_thr; };
template <typename R>
static __global__ void noise_tail_kernel(PhiloxKeys rk, long long n_walkers, long long calls, TailArgs T,
                                         unsigned long long* counts, double* stats) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long cnt[NDS_MAX_THR];
  R thr[NDS_MAX_THR];
#pragma unroll
  for (int t = 0; t < NDS_MAX_THR; t++) { cnt[t] = 0ull; thr[t] = t < T.n_thr ? (R)T.thr[t] : (R)1e30; }
  double s1 = 0.0, s2 = 0.0, s4 = 0.0, mx = 0.0;
  if (w < n_walkers) {
    for (long long c = 0; c < calls; c++) {
      R z[NDS_NPC];
      normals6<R>(rk, (unsigned long long)c, (unsigned long long)w, z);
      R p1 = (R)0, p2 = (R)0, p4 = (R)0;
#pragma unroll
      for (int j = 0; j < NDS_NPC; j++) {
        const R a = z[j] < (R)0 ? -z[j] : z[j];
        p1 += z[j]; p2 += a * a; p4 += (a * a) * (a * a);
        mx = (double)a > mx ? (double)a : mx;
#pragma unroll
        for (int t = 0; t < NDS_MAX_THR; t++) cnt[t] += a > thr[t] ? 1ull : 0ull;
      }
      s1 += (double)p1; s2 += (double)p2; s4 += (double)p4;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_down_sync(0xffffffffu, s1, o); s2 += __shfl_down_sync(0xffffffffu, s2, o);
    s4 += __shfl_down_sync(0xffffffffu, s4, o);
    const double m2 = __shfl_down_sync(0xffffffffu, mx, o); mx = m2 > mx ? m2 : mx;
#pragma unroll
    for (int t = 0; t < NDS_MAX_THR; t++) cnt[t] += __shfl_down_sync(0xffffffffu, cnt[t], o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&stats[0], s1); atomicAdd(&stats[1], s2); atomicAdd(&stats[2], s4);
    atomicMax(reinterpret_cast<unsigned long long*>(&stats[3]), (unsigned long long)__double_as_longlong(mx));
    for (int t = 0; t < T.n_thr; t++) atomicAdd(&counts[t], cnt[t]);
  }
}

static __global__ void compact_lists_kernel(long long n, long long n_alive, const unsigned char* alive,
                                     const int* scan, int* holes, int* movers, int* n_holes) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int front_alive = n_alive < n ? scan[n_alive] : (int)n_alive;
  if (i == 0) *n_holes = (int)n_alive - front_alive;
  if (i >= n) return;
  if (i < n_alive) { if (!alive[i]) holes[i - scan[i]] = (int)i; }
  else if (alive[i]) movers[scan[i] - front_alive] = (int)i;
}

static __global__ void iota_kernel(long long n, long long* ids) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ids[i] = i;
}
static __global__ void alive_flags_kernel(long long n, const unsigned char* alive, int* flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = alive[i] ? 1 : 0;
}


template <typename R>
static __global__ void count_nonfinite_kernel(long long M, int nd, const R* x, const R* v, unsigned long long* out) {
  unsigned long long local = 0;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < M; w += (long long)gridDim.x * blockDim.x) {
    bool bad = false;
    for (int d = 0; d < nd; d++) bad |= !isfinite((double)x[d * M + w]) || !isfinite((double)v[d * M + w]);
    local += bad ? 1ull : 0ull;
  }
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

static __global__ void count_alive_kernel(long long M, const unsigned char* alive, unsigned long long* out) {
  unsigned long long local = 0;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < M; w += (long long)gridDim.x * blockDim.x)
    local += alive[w] ? 1ull : 0ull;
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

// out[row * M + w] = in[row] for every walker w (install one hill list into all per-walker tables)
template <typename R>
static __global__ void broadcast_rows_kernel(long long total, long long M, const R* __restrict__ in, R* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) out[i] = in[i / M];
}

// position-weighted 64-bit checksum of a word array: sum_i (i + 1) * 0x9E3779B97F4A7C15 * (word_i + 1)  (mod 2^64)
static __global__ void checksum_kernel(const uint32_t* __restrict__ words, long long n, unsigned long long* out) {
  unsigned long long local = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    local += (unsigned long long)(i + 1) * 0x9E3779B97F4A7C15ull * ((unsigned long long)words[i] + 1ull);
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, local);
}

// ----------------------------------------------------------------------------------- engine
template <typename R>
struct Engine : nds_engine {
  Consts<R> K;
  int64_t M = 0;
  int nd = 0;
  BlockLaunchFn<R> launch_block = nullptr;
  // packed-FP32 variant of the step kernel (fp32 engines only): two walkers per thread
  Consts<f2> K2;
  BlockLaunchFn<f2> launch_packed = nullptr;
  const char* packed_name = nullptr;
  // device state
  R *d_x = nullptr, *d_v = nullptr, *d_f = nullptr, *d_fixf = nullptr, *d_colv = nullptr;
  R *d_thermo = nullptr, *d_vr = nullptr, *d_umb_k = nullptr, *d_umb_r0 = nullptr, *d_hills = nullptr;
  unsigned char* d_alive = nullptr;
  int* d_basin = nullptr;
  long long* d_exit = nullptr;
  R* d_frames = nullptr;               // frame ring in the engine's real type
  unsigned long long* d_hist = nullptr;  // in-kernel colvar histogram [ny][nx]
  bool frames_on = false, hist_on = false;
  // subset dumps: walkers [0, split_at) run the frame-recording variant on a second stream, the rest the plain one
  BlockLaunchFn<R> launch_block_rest = nullptr;
  BlockLaunchFn<f2> launch_packed_rest = nullptr;
  Consts<R> K_rest;
  Consts<f2> K2_rest;
  int64_t split_at = 0;
  int dyn_smem = 0, hist_smem_off = 0;
  R *d_path_ref = nullptr, *d_path_ind = nullptr;
  long long* d_mc_accepted = nullptr;
  double* d_mc_time = nullptr;   // MC: per-walker NDRun.time (accepted moves * dt)
  long long* d_mc_dep = nullptr; // MC + metadynamics: per-walker deposition count
  R* d_wl = nullptr;
  double* d_noise = nullptr;
  int64_t noise_first = 0, noise_steps = 0;
  double* d_stage = nullptr;  // fp64 staging for host <-> device conversion
  int64_t stage_elems = 0;
  unsigned long long* d_counter = nullptr;
  bool have_v = false, begun = false, need_refresh_vr = true;
  // committor lane compaction
  long long* d_ids = nullptr;
  int *d_flags = nullptr, *d_scan = nullptr, *d_holes = nullptr, *d_movers = nullptr, *d_nholes = nullptr;
  void* d_cub = nullptr;
  size_t cub_bytes = 0;
  int64_t n_active = 0;
  bool compactable = false, permuted = false;
  // peer-memory exchange: device array of every rank's hill table (own included)
  R** d_peer_tables = nullptr;
  std::vector<void*> ipc_opened;
  int n_peers = 0, peer_rank = 0;
  // one device slab holds x | v | thermo | f | fixf | colv | vr: (x, v) and (x, v, thermo) are contiguous, so the
  // asynchronous host path moves each direction with ONE copy (nds_submit)
  R* d_slab = nullptr;
  // exchange window behind the hill table (same allocation, hence the same IPC handle): barrier flags + grid slots
  int nranks_cfg = 1;
  int64_t xwin_off = 0;                    // bytes from d_hills to the window
  unsigned long long* d_flags_own = nullptr;  // [NDS_MAX_RANKS] arrival epochs written by the peers, then the error word
  unsigned long long** d_peer_flags = nullptr;
  unsigned long long barrier_epoch = 0;
  bool device_barrier = false;
  // grid-accumulated bias
  bool grid_mode = false, grid_in_smem = false, variant_wide = false;
  int64_t grid_elems = 0;                  // 4 * nx * ny
  R* d_grid = nullptr;
  R* d_ginc = nullptr;                     // increment of one stride (single slot: all-reduce path)
  unsigned int* d_ticket = nullptr;        // arrival counter of grid_exchange_kernel
  unsigned long long* d_xtrace = nullptr;  // NDS_EXCHANGE_TRACE: phase times of grid_exchange_kernel
  int sm_count = 0;
  bool cluster_ready = false;              // grid_accumulate_kernel opted in to clusters of GRID_SLICES > 8 CTAs
  R* d_vr_umb = nullptr;                   // Harmonic.VR per restraint [n_umb][M] (umbrella + metadynamics with frames)
  R* d_cvhist = nullptr;                   // Stat.colv ring [tao_d][M] (mtd_adapsig_hist)
  int64_t cv_n = 0;                        // samples taken so far (len(stat.colv))
  int tao_d = 0;
  R** d_slot_dst = nullptr;                // [nranks] slot of THIS rank on every peer, parity 0 then parity 1
  R** d_slot_src = nullptr;                // [nranks] slots of every rank in THIS rank's window, parity 0 then 1
  R** d_ginc_list = nullptr;               // {d_ginc}
  unsigned long long* d_grid_outside = nullptr;
  // native NCCL communicator (nds_set_comm)
  NcclApi::comm_t comm = nullptr;
  bool comm_owned = false;
  int comm_rank = 0, comm_nranks = 1;
  cudaEvent_t xev0 = nullptr, xev1 = nullptr;
  double exchange_ms_total = 0.0;
  int64_t n_exchanges = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> xev_pending;
  // asynchronous host path (nds_submit): copy streams + events
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_alt = nullptr;
  std::vector<cudaEvent_t> pipe_events;
  // scratch of nds_eval_at, grown on demand
  R* d_eval = nullptr;
  size_t eval_elems = 0;
  std::vector<long long> ids_host;  // slot -> walker of the current permutation (downloaded once per change)
  bool ids_host_valid = false;
  int64_t dep_count = 0, rescale_count = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;

  ~Engine() override {
    cudaSetDevice(cfg.device);
    if (d_xtrace) {  // NDS_EXCHANGE_TRACE=1
      unsigned long long h[8] = {0};
      cudaStreamSynchronize(stream);
      if (cudaMemcpy(h, d_xtrace, sizeof h, cudaMemcpyDeviceToHost) == cudaSuccess && h[5] > 0)
        fprintf(stderr, "[nds exchange trace] rank %d: %llu fused exchanges; mean ns from kernel entry: own push fenced %.0f, "
                        "signal sent %.0f, all peers seen %.0f, slots added %.0f\n", peer_rank, h[5],
                (double)h[0] / h[5], (double)h[1] / h[5], (double)h[2] / h[5], (double)h[3] / h[5]);
      cudaFree(d_xtrace);
    }
    if (comm && comm_owned && nccl_api().ok) nccl_api().CommDestroy(comm);
    for (void* p : {(void*)d_slab, (void*)d_grid, (void*)d_ginc, (void*)d_ticket, (void*)d_cvhist, (void*)d_vr_umb, (void*)d_slot_dst, (void*)d_slot_src, (void*)d_ginc_list,
                    (void*)d_grid_outside, (void*)d_peer_flags, (void*)d_eval,
                    (void*)d_umb_k, (void*)d_umb_r0, (void*)d_hills, (void*)d_alive,
                    (void*)d_basin, (void*)d_exit, (void*)d_frames, (void*)d_hist, (void*)d_noise, (void*)d_stage,
                    (void*)d_path_ref, (void*)d_path_ind, (void*)d_mc_accepted, (void*)d_mc_time, (void*)d_mc_dep, (void*)d_wl, (void*)d_ids, (void*)d_flags, (void*)d_scan,
                    (void*)d_holes, (void*)d_movers, (void*)d_nholes, d_cub,
                    (void*)d_counter})
      if (p) cudaFree(p);
    for (void* p : ipc_opened) cudaIpcCloseMemHandle(p);
    if (d_peer_tables) cudaFree(d_peer_tables);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    for (auto& pr : xev_pending) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (cudaEvent_t ev : pipe_events) cudaEventDestroy(ev);
    for (cudaStream_t st : {s_h2d, s_d2h, s_alt}) if (st) cudaStreamDestroy(st);
    if (stream) cudaStreamDestroy(stream);
  }

  double step_dt() const { return cfg.integrator == NDS_INT_MINIMIZE ? cfg.dt * 10 : cfg.dt; }

  // records one deposition appends to this rank's table, and the records the table can hold.  mtd_grid keeps only
  // the LOCAL hills (for fix.dat listings / checkpoints): the bias the walkers feel lives in the grid.
  int64_t table_add() const { return !cfg.mtd_shared ? 1 : (grid_mode ? M : cfg.n_walkers_global); }
  int64_t table_cap() const {
    return (grid_mode && cfg.mtd_shared) ? (cfg.mtd_capacity + nranks_cfg - 1) / nranks_cfg : cfg.mtd_capacity;
  }
  int64_t hill_elems() const {
    if (!cfg.mtd_shared) return cfg.mtd_capacity * (int64_t)K.hill_stride * M;
    if (grid_mode) return (table_cap() + 1) * (int64_t)K.hill_stride;
    // shared table: padded with zero-height records to whole TMA tiles (mtd_block reads full tiles)
    const int64_t cap = (cfg.mtd_capacity + MTD_TILE_HOST - 1) / MTD_TILE_HOST * MTD_TILE_HOST + MTD_TILE_HOST;
    return cap * (int64_t)K.hill_stride;
  }

  int clear_hills() {
    const int64_t n = hill_elems();
    if (cfg.mtd_shared) {
      const int64_t nrec = n / K.hill_stride;
      fill_hills_kernel<R><<<(unsigned)((nrec + 255) / 256), 256, 0, stream>>>(nrec, K.hill_stride, K.hill_hidx, d_hills);
      CU(cudaGetLastError());
    } else {
      CU(cudaMemsetAsync(d_hills, 0, sizeof(R) * (size_t)n, stream));
    }
    return 0;
  }

  int init(const nds_config& c) {
    cfg = c;
    M = c.n_walkers;
    nd = c.ndim;
    fill_consts<R>(c, K);
    cd = K.cd; tcd = K.tcd;
    CU(cudaSetDevice(c.device));
    CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&ev0));
    CU(cudaEventCreate(&ev1));
    frames_on = K.dump_freq > 0 && K.dump_walkers > 0;
    hist_on = K.hist_nx > 0;
    const bool dumping = frames_on || hist_on;
    const int dump_code = (frames_on ? 1 : 0) | (hist_on ? 2 : 0);
    grid_mode = c.mtd_enabled && c.mtd_grid;
    nranks_cfg = (int)((c.n_walkers_global + c.n_walkers - 1) / c.n_walkers);
    const Variant* var = pick_variant(c, K.bias_mask, dump_code);
    if (!var) return fail(this, "no compiled md_block variant matches this configuration");
    launch_block = (BlockLaunchFn<R>)var->launch;
    kernel_name = var->name;
    variant_wide = var->wide;
    // packed kernels see pairs of walkers: the frame prefix must then be a whole number of pairs
    const bool can_pack = std::is_same<R, float>::value && M % 2 == 0 && K.dump_walkers % 2 == 0 && !getenv("NDS_NO_PACKED");
    if (can_pack) {
      const Variant* pv = pick_variant(c, K.bias_mask, dump_code, PREC_F32_PACKED);
      if (pv) {
        fill_consts<f2>(c, K2);
        K2.dump_walkers = K.dump_walkers / 2;
        launch_packed = (BlockLaunchFn<f2>)pv->launch;
        packed_name = pv->name;
        kernel_name = pv->name;
      }
    }
    // subset dumps ("full state of walkers 0..4095"): only the CTAs that own dumping walkers run the
    // frame-recording variant (on a second stream); all other walkers keep the kernel without frames
    if (frames_on && K.dump_walkers < M && c.method == NDS_METHOD_MD && !c.mtd_enabled && !getenv("NDS_NO_SPLIT")) {
      const int64_t unit = 256;  // whole CTAs of either kernel (128 threads, two walkers per thread when packed)
      const int64_t pad = (K.dump_walkers + unit - 1) / unit * unit;
      const Variant* rv = pick_variant(c, K.bias_mask, dump_code & 2);
      if (pad < M && rv) {
        split_at = pad;
        launch_block_rest = (BlockLaunchFn<R>)rv->launch;
        K_rest = K;
        K_rest.dump_walkers = 0;
        if (!hist_on) K_rest.dump_freq = 0;
        if (launch_packed) {
          const Variant* rp = pick_variant(c, K.bias_mask, dump_code & 2, PREC_F32_PACKED);
          if (rp) { launch_packed_rest = (BlockLaunchFn<f2>)rp->launch; K2_rest = K2; K2_rest.dump_walkers = 0; if (!hist_on) K2_rest.dump_freq = 0; }
          else { launch_packed = nullptr; kernel_name = var->name; }  // both halves must agree on the walker layout
        }
      }
    }
    if (c.mtd_enabled && c.mtd_shared && !grid_mode && (cd == 2 || cd == 1) && !K.mtd_adaptive && c.method == NDS_METHOD_MD && !dumping &&
        !getenv("NDS_NO_MTD_BLOCK")) {
      cudaDeviceProp prop;
      CU(cudaGetDeviceProperties(&prop, c.device));
      const MtdVariant* mv = pick_mtd_variant(c, K.bias_mask, prop.multiProcessorCount);
      if (mv) { launch_block = (BlockLaunchFn<R>)mv->launch; kernel_name = mv->name; }
    }
    const size_t sM = (size_t)M;
    CU(cudaMalloc(&d_slab, sizeof(R) * (size_t)(4 * nd + cd + 6) * sM));
    d_x = d_slab; d_v = d_x + nd * sM; d_thermo = d_v + nd * sM; d_f = d_thermo + 5 * sM;
    d_fixf = d_f + nd * sM; d_colv = d_fixf + nd * sM; d_vr = d_colv + cd * sM;
    CU(cudaMemsetAsync(d_slab, 0, sizeof(R) * (size_t)(4 * nd + cd + 6) * sM, stream));
    if (K.bias_mask & BIAS_UMB_WALKER) {
      CU(cudaMalloc(&d_umb_k, sizeof(R) * cd * sM));
      CU(cudaMalloc(&d_umb_r0, sizeof(R) * cd * sM));
      std::vector<double> k((size_t)cd * sM), r0((size_t)cd * sM);
      for (int q = 0; q < cd; q++)
        for (size_t w = 0; w < sM; w++) { k[q * sM + w] = c.umb_k[0][q]; r0[q * sM + w] = c.umb_r0[0][q]; }
      if (upload(k.data(), d_umb_k, (int64_t)cd * M)) return 1;
      if (upload(r0.data(), d_umb_r0, (int64_t)cd * M)) return 1;
    }
    if (c.mtd_enabled) {
      // hill table, then (same allocation => same IPC handle) the exchange window: barrier flags + grid slots
      grid_elems = grid_mode ? 4 * (int64_t)K.grid_nx * K.grid_ny : 0;
      xwin_off = ((int64_t)sizeof(R) * hill_elems() + 255) / 256 * 256;
      const int64_t flag_bytes = 256;  // NDS_MAX_RANKS arrival words + the error word
      const int64_t slot_bytes = (int64_t)sizeof(R) * grid_elems;
      const int64_t total = xwin_off + flag_bytes + (c.mtd_shared ? 2 * nranks_cfg * slot_bytes : 0);
      CU(cudaMalloc(&d_hills, (size_t)total));
      CU(cudaMemsetAsync(d_hills, 0, (size_t)total, stream));
      d_flags_own = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(d_hills) + xwin_off);
      if (clear_hills()) return 1;
      if (grid_mode) {
        CU(cudaMalloc(&d_grid, sizeof(R) * (size_t)grid_elems));
        CU(cudaMalloc(&d_ginc, sizeof(R) * (size_t)grid_elems));
        CU(cudaMemsetAsync(d_grid, 0, sizeof(R) * (size_t)grid_elems, stream));
        CU(cudaMemsetAsync(d_ginc, 0, sizeof(R) * (size_t)grid_elems, stream));
        CU(cudaMalloc(&d_ginc_list, sizeof(R*)));
        CU(cudaMemcpyAsync(d_ginc_list, &d_ginc, sizeof(R*), cudaMemcpyHostToDevice, stream));
        CU(cudaMalloc(&d_grid_outside, sizeof(unsigned long long)));
        CU(cudaMemsetAsync(d_grid_outside, 0, sizeof(unsigned long long), stream));
        // the whole grid rides in shared memory when the wide variant was picked and it fits beside the static part
        int dev_smem = 0;
        CU(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c.device));
        CU(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, c.device));
        grid_in_smem = variant_wide && (int64_t)sizeof(R) * grid_elems <= (int64_t)dev_smem - 1024 && !getenv("NDS_GRID_NO_SMEM");
      }
      if (c.mtd_adapsig_hist) {
        tao_d = nds_tao_d(&c);
        CU(cudaMalloc(&d_cvhist, sizeof(R) * (size_t)tao_d * sM));
        CU(cudaMemsetAsync(d_cvhist, 0, sizeof(R) * (size_t)tao_d * sM, stream));
      }
    }
    if (c.method == NDS_METHOD_COMMITTOR) {
      CU(cudaMalloc(&d_alive, sM));
      CU(cudaMalloc(&d_basin, sizeof(int) * sM));
      CU(cudaMalloc(&d_exit, sizeof(long long) * sM));
      CU(cudaMemsetAsync(d_alive, 1, sM, stream));
      CU(cudaMemsetAsync(d_basin, 0xff, sizeof(int) * sM, stream));
      CU(cudaMemsetAsync(d_exit, 0, sizeof(long long) * sM, stream));
    }
    n_active = M;
    compactable = c.method == NDS_METHOD_COMMITTOR && c.integrator != NDS_INT_MINIMIZE && !dumping && !c.mtd_enabled &&
                  c.noise_mode == NDS_NOISE_PHILOX && !getenv("NDS_NO_COMPACTION");
    if (compactable) {
      CU(cudaMalloc(&d_ids, sizeof(long long) * sM));
      CU(cudaMalloc(&d_flags, sizeof(int) * sM));
      CU(cudaMalloc(&d_scan, sizeof(int) * (sM + 1)));
      CU(cudaMalloc(&d_holes, sizeof(int) * sM));
      CU(cudaMalloc(&d_movers, sizeof(int) * sM));
      CU(cudaMalloc(&d_nholes, sizeof(int)));
      cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, d_flags, d_scan, (int)M, stream);
      CU(cudaMalloc(&d_cub, cub_bytes));
      iota_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(M, d_ids);
    }
    if (c.method == NDS_METHOD_MC || c.method == NDS_METHOD_WL) {
      CU(cudaMalloc(&d_mc_accepted, sizeof(long long) * sM));
      CU(cudaMemsetAsync(d_mc_accepted, 0, sizeof(long long) * sM, stream));
      if (c.mtd_enabled) {
        CU(cudaMalloc(&d_mc_time, sizeof(double) * sM));
        CU(cudaMalloc(&d_mc_dep, sizeof(long long) * (sM + 1)));  // [M]: overflow flag (first walker + 1)
        CU(cudaMemsetAsync(d_mc_time, 0, sizeof(double) * sM, stream));
        CU(cudaMemsetAsync(d_mc_dep, 0, sizeof(long long) * (sM + 1), stream));
      }
      kernel_name = sizeof(R) == 4 ? "mc_block<f32>" : "mc_block<f64>";
      if (c.method == NDS_METHOD_WL) {
        CU(cudaMalloc(&d_wl, sizeof(R) * NDS_WL_GRID * sM));
        CU(cudaMemsetAsync(d_wl, 0, sizeof(R) * NDS_WL_GRID * sM, stream));
        kernel_name = sizeof(R) == 4 ? "mc_block<f32,wang-landau>" : "mc_block<f64,wang-landau>";
      }
    }
    CU(cudaMalloc(&d_counter, sizeof(unsigned long long)));
    frame_width = 2 + 3 * nd + cd + 6;
    if (frames_on && c.mtd_enabled && c.n_umb > 0 && c.method == NDS_METHOD_MD) {
      // Gaussian.update recomputes EVERY fix after a deposition (gaus.py:138-146): Harmonic.VR follows, one frame field each
      CU(cudaMalloc(&d_vr_umb, sizeof(R) * (size_t)c.n_umb * (size_t)M));
      CU(cudaMemsetAsync(d_vr_umb, 0, sizeof(R) * (size_t)c.n_umb * (size_t)M, stream));
      frame_width += c.n_umb;
    }
    if (frames_on) {
      const size_t n = (size_t)c.dump_capacity * frame_width * (size_t)K.dump_walkers;
      CU(cudaMalloc(&d_frames, sizeof(R) * n));
      CU(cudaMemsetAsync(d_frames, 0xff, sizeof(R) * n, stream));  // NaN pattern
    }
    hist_smem_off = grid_in_smem ? (int)(sizeof(R) * grid_elems) : 0;
    dyn_smem = hist_smem_off;
    if (hist_on) {
      const size_t nb = (size_t)K.hist_nx * K.hist_ny;
      CU(cudaMalloc(&d_hist, sizeof(unsigned long long) * nb));
      CU(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nb, stream));
      dyn_smem += (int)(sizeof(unsigned int) * nb);
    }
    CU(cudaStreamSynchronize(stream));
    return 0;
  }

  // ---- host <-> device movers (host side is always fp64)
  int ensure_stage(int64_t n) {
    if (n <= stage_elems) return 0;
    if (d_stage) CU(cudaFree(d_stage));
    CU(cudaMalloc(&d_stage, sizeof(double) * (size_t)n));
    stage_elems = n;
    return 0;
  }
  int upload(const double* h, R* d, int64_t n) {
    if (sizeof(R) == sizeof(double)) {
      CU(cudaMemcpyAsync(d, h, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, stream));
    } else {
      if (ensure_stage(n)) return 1;
      CU(cudaMemcpyAsync(d_stage, h, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, stream));
      convert_kernel<R><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, d_stage, d);
      CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(stream));  // the host buffer may be reused by the caller
    return 0;
  }
  int download(const R* d, double* h, int64_t n) {
    if (sizeof(R) == sizeof(double)) {
      CU(cudaMemcpyAsync(h, d, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, stream));
    } else {
      if (ensure_stage(n)) return 1;
      convert_back_kernel<R><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, d, d_stage);
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(h, d_stage, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, stream));
    }
    CU(cudaStreamSynchronize(stream));
    return 0;
  }

  int set_state(const double* x, const double* v) override {
    CU(cudaSetDevice(cfg.device));
    if (!x) return fail(this, "nds_set_state: x is null");
    if (permuted) {  // new state for every walker: slots go back to walker order, all walkers live again
      iota_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(M, d_ids);
      permuted = false;
      n_active = M;
      CU(cudaMemsetAsync(d_alive, 1, (size_t)M, stream));
    }
    if (upload(x, d_x, (int64_t)nd * M)) return 1;
    if (v) { if (upload(v, d_v, (int64_t)nd * M)) return 1; }
    // positions + velocities given: a running simulation continues from the new state (the next
    // launch re-evaluates forces at entry); positions only: velocities are re-drawn by nds_begin()
    if (!v) { have_v = false; begun = false; }
    else have_v = true;
    return 0;
  }

  // fp32 host buffers: direct copies for fp32 engines, through the staging buffer otherwise
  int upload_f32(const float* h, R* d, int64_t n) {
    if (sizeof(R) == sizeof(float)) {
      CU(cudaMemcpyAsync(d, h, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, stream));
    } else {
      if (ensure_stage(n)) return 1;
      float* st = reinterpret_cast<float*>(d_stage);
      CU(cudaMemcpyAsync(st, h, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, stream));
      cast_kernel<float, R><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, st, d);
      CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(stream));
    return 0;
  }
  int download_f32(const R* d, float* h, int64_t n) {
    if (sizeof(R) == sizeof(float)) {
      CU(cudaMemcpyAsync(h, d, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, stream));
    } else {
      if (ensure_stage(n)) return 1;
      float* st = reinterpret_cast<float*>(d_stage);
      cast_kernel<R, float><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, d, st);
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(h, st, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, stream));
    }
    CU(cudaStreamSynchronize(stream));
    return 0;
  }

  int set_state_f32(const float* x, const float* v) override {
    CU(cudaSetDevice(cfg.device));
    if (!x) return fail(this, "nds_set_state_f32: x is null");
    if (permuted) {
      iota_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(M, d_ids);
      permuted = false;
      n_active = M;
      CU(cudaMemsetAsync(d_alive, 1, (size_t)M, stream));
    }
    if (upload_f32(x, d_x, (int64_t)nd * M)) return 1;
    if (v) { if (upload_f32(v, d_v, (int64_t)nd * M)) return 1; }
    if (!v) { have_v = false; begun = false; }
    else have_v = true;
    return 0;
  }

  int get_state_f32(float* x, float* v, float* f, float* fixf, float* colv, float* thermo) override {
    CU(cudaSetDevice(cfg.device));
    if (x && (download_f32(d_x, x, (int64_t)nd * M) || unpermute_host(x, nd))) return 1;
    if (v && (download_f32(d_v, v, (int64_t)nd * M) || unpermute_host(v, nd))) return 1;
    if (f && (download_f32(d_f, f, (int64_t)nd * M) || unpermute_host(f, nd))) return 1;
    if (fixf && (download_f32(d_fixf, fixf, (int64_t)nd * M) || unpermute_host(fixf, nd))) return 1;
    if (colv && (download_f32(d_colv, colv, (int64_t)cd * M) || unpermute_host(colv, cd))) return 1;
    if (thermo && (download_f32(d_thermo, thermo, (int64_t)5 * M) || unpermute_host(thermo, 5))) return 1;
    return 0;
  }

  int set_walker_bias(const double* k, const double* r0) override {
    CU(cudaSetDevice(cfg.device));
    if (!(K.bias_mask & BIAS_UMB_WALKER)) return fail(this, "nds_set_walker_bias: config has umb_per_walker = 0");
    if (upload(k, d_umb_k, (int64_t)cd * M)) return 1;
    return upload(r0, d_umb_r0, (int64_t)cd * M);
  }

  int set_noise(const double* noise, int64_t first, int64_t nsteps) override {
    CU(cudaSetDevice(cfg.device));
    if (cfg.noise_mode != NDS_NOISE_EXTERNAL) return fail(this, "nds_set_noise: noise_mode is not NDS_NOISE_EXTERNAL");
    const int nn = (cfg.method == NDS_METHOD_MC || cfg.method == NDS_METHOD_WL) ? (cfg.mc_propose_all ? nd + 1 : 3) : normals_per_step(cfg.integrator, nd);
    if (d_noise) { CU(cudaFree(d_noise)); d_noise = nullptr; }
    const size_t n = (size_t)nsteps * nn * (size_t)M;
    if (n) {
      CU(cudaMalloc(&d_noise, sizeof(double) * n));
      CU(cudaMemcpyAsync(d_noise, noise, sizeof(double) * n, cudaMemcpyHostToDevice, stream));
      CU(cudaStreamSynchronize(stream));
    }
    noise_first = first; noise_steps = nsteps;
    return 0;
  }

  int set_path(int64_t n_ref, int path_dim, const double* ref, const double* ind, double sigma) override {
    CU(cudaSetDevice(cfg.device));  // Path.initialize, colvars/path.py:51-71
    if (cfg.colvar != NDS_CV_PATH && cfg.colvar != NDS_CV_PATH5DTO2D) return fail(this, "nds_set_path: colvar is not a path variable");
    if (n_ref < 1 || n_ref > NDS_MAX_PATH_REF) return fail(this, "nds_set_path: n_ref must be in [1, %d]", NDS_MAX_PATH_REF);
    const int want = cfg.colvar == NDS_CV_PATH ? nd : 2;
    if (path_dim != want) return fail(this, "nds_set_path: path_dim %d, expected %d", path_dim, want);
    if (!(sigma > 0)) return fail(this, "nds_set_path: sigma must be positive");
    if (d_path_ref) { CU(cudaFree(d_path_ref)); CU(cudaFree(d_path_ind)); }
    CU(cudaMalloc(&d_path_ref, sizeof(R) * (size_t)(n_ref * path_dim)));
    CU(cudaMalloc(&d_path_ind, sizeof(R) * (size_t)n_ref));
    if (upload(ref, d_path_ref, n_ref * path_dim)) return 1;
    if (upload(ind, d_path_ind, n_ref)) return 1;
    K.path_ref = d_path_ref; K.path_ind = d_path_ind; K.n_path = (int)n_ref; K.path_dim = path_dim;
    K.path_sig2_invh = (R)(1 / (sigma * sigma) / 2.0);
    K.path_inv_sig2 = (R)(1 / (sigma * sigma));
    K.path_N = (R)(double)(int)ind[n_ref - 1];
    begun = false;
    return 0;
  }

  BlockArgs<R> make_args(int64_t nsteps) {
    BlockArgs<R> A;
    memset(&A, 0, sizeof A);
    A.M = M; A.walker_offset = cfg.walker_offset;
    A.step0 = step; A.nsteps = nsteps; A.time0 = time; A.rescale_count = rescale_count;
    A.x = d_x; A.v = d_v; A.f = d_f; A.fixf = d_fixf; A.colv = d_colv; A.thermo = d_thermo; A.vr = d_vr;
    A.umb_k = d_umb_k; A.umb_r0 = d_umb_r0;
    A.alive = d_alive; A.basin = d_basin; A.exit_step = d_exit;
    A.n_active = n_active; A.ids = permuted ? d_ids : nullptr;
    A.hills = d_hills; A.n_hills = n_hills;
    A.grid = d_grid; A.grid_smem = grid_in_smem ? 1 : 0; A.grid_outside = d_grid_outside;
    A.noise = d_noise; A.noise_first = noise_first;
    A.frames = d_frames; A.frame_base = frames_written; A.frame_cap = cfg.dump_capacity;
    A.frame_width = frame_width;
    A.hist = d_hist; A.hist_smem_off = hist_smem_off; A.dyn_smem = dyn_smem;
    A.cvhist = d_cvhist; A.cvhist_n = cv_n; A.cvhist_cap = tao_d;
    A.vr_umb = d_vr_umb;
    return A;
  }

  int begin() override {  // NDRun.begin (ndrun.py:182-206) + MD.begin (md.py:101-110)
    CU(cudaSetDevice(cfg.device));
    if ((cfg.colvar == NDS_CV_PATH || cfg.colvar == NDS_CV_PATH5DTO2D) && !d_path_ref)
      return fail(this, "path colvar: call nds_set_path() before nds_begin()");
    // Gaussian._dep_count and _history survive a second NDRun.begin() in the reference (only step / time restart,
    // ndrun.py:199-200): neither the deposition count nor the hills are reset here; MC keeps its per-walker
    // clocks on the device and restarts both below.
    step = 0; time = 0.0; rescale_count = 0; frames_written = 0;
    if (d_mc_accepted) CU(cudaMemsetAsync(d_mc_accepted, 0, sizeof(long long) * (size_t)M, stream));
    if (d_mc_time) CU(cudaMemsetAsync(d_mc_time, 0, sizeof(double) * (size_t)M, stream));
    if (d_mc_dep) { CU(cudaMemsetAsync(d_mc_dep, 0, sizeof(long long) * (size_t)(M + 1), stream)); n_hills = 0; dep_count = 0; }
    if (d_wl) CU(cudaMemsetAsync(d_wl, 0, sizeof(R) * NDS_WL_GRID * (size_t)M, stream));
    if (cfg.method == NDS_METHOD_COMMITTOR) {
      if (permuted && unpermute_device()) return 1;
      CU(cudaMemsetAsync(d_alive, 1, (size_t)M, stream));
      CU(cudaMemsetAsync(d_basin, 0xff, sizeof(int) * (size_t)M, stream));
      CU(cudaMemsetAsync(d_exit, 0, sizeof(long long) * (size_t)M, stream));
      n_active = M;
    }
    BlockArgs<R> A = make_args(0);
    A.refresh_vr = 1;
    A.need_init_v = have_v ? 0 : 1;
    A.begin_mode = 1;
    cudaError_t err = launch_block(K, A, stream);
    if (err != cudaSuccess) return fail(this, "md_block launch failed: %s", cudaGetErrorString(err));
    launches++;
    have_v = true; begun = true; need_refresh_vr = false;
    if (d_cvhist) cv_n += 1;  // stat.append of NDRun.begin (ndrun.py:201); Stat survives a second begin()
    if (frames_on) frames_written = 1;
    if (d_hist) CU(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * (size_t)(K.hist_nx * K.hist_ny), stream));
    return 0;
  }

  // ---- exchange plumbing shared by the exact and the grid path
  // events around every exchange (deposit kernel excluded): nds_exchange_info() reports their sum
  void exchange_mark(bool begin_) {
    if (begin_) {
      cudaEvent_t a = nullptr, b = nullptr;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
      cudaEventRecord(a, stream);
      xev_pending.push_back({a, b});
    } else if (!xev_pending.empty()) {
      cudaEventRecord(xev_pending.back().second, stream);
      if (xev_pending.size() > 1024) exchange_collect();
    }
  }
  void exchange_collect() {
    for (auto& pr : xev_pending) {
      float ms = 0.f;
      if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
        exchange_ms_total += ms;
        n_exchanges += 1;
      }
      cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
    }
    xev_pending.clear();
  }
  int exchange_info(double* ms, int64_t* n) override {
    CU(cudaSetDevice(cfg.device));
    exchange_collect();
    if (ms) *ms = exchange_ms_total;
    if (n) *n = n_exchanges;
    return 0;
  }

  // all ranks have finished their peer stores: device-side flag barrier (one GPU per rank) or the host callback
  int peer_barrier() {
    if (device_barrier) {
      barrier_epoch += 1;
      peer_barrier_kernel<<<1, 32, 0, stream>>>(d_peer_flags, peer_rank, n_peers, barrier_epoch,
                                                10000000000ull, d_flags_own + NDS_MAX_RANKS);
      CU(cudaGetLastError());
      launches++;
      return 0;
    }
    CU(cudaStreamSynchronize(stream));
    if (!exchange) return fail(this, "peer tables need a barrier: nds_set_peer_barrier(e, 1) or a callback (nds_set_exchange)");
    const int rc = exchange(exchange_user, nullptr, 0, 0);  // bytes == 0: barrier only
    if (rc) return fail(this, "barrier callback failed (%d)", rc);
    return 0;
  }

  int nccl_check(int rc, const char* what) {
    if (rc == 0) return 0;
    return fail(this, "%s failed: %s", what, nccl_api().GetErrorString(rc));
  }

  int deposit() {  // Gaussian.update -> _deposite (gaus.py:130-213), all live walkers together
    const int64_t add = table_add();
    if (n_hills + add > table_cap())
      return fail(this, "hill table full: %lld + %lld > mtd_capacity %lld", (long long)n_hills,
                  (long long)add, (long long)table_cap());
    const bool multi = cfg.mtd_shared && cfg.n_walkers_global > M;
    const int first = dep_count == 0 ? 1 : 0;
    const unsigned g = (unsigned)((M + 127) / 128);
    if (grid_mode) {
      // the new records of this rank (local table), then their footprint on the grid
      deposit_kernel<R><<<g, 128, 0, stream>>>(K, M, 0, d_colv, d_vr, d_alive, d_hills, n_hills, first);
      CU(cudaGetLastError());
      const R* rec = d_hills + n_hills * K.hill_stride;
      const unsigned ga = (unsigned)((grid_elems + 255) / 256);
      launches += 1;
      if (!multi) {
        if (accumulate(rec, M, nullptr, 0, true)) return 1;
      } else if (d_slot_dst) {
        // the increment is pushed into this rank's slot on EVERY peer over NVLink; after the barrier each rank adds
        // the slots in rank order (bit-identical grids everywhere)
        const int par = (int)(dep_count & 1);
        // (the exchange clock starts at the kernel that writes to the peers, not at the footprint computation)
        if (accumulate(rec, M, d_ginc_list, 1, false)) return 1;
        exchange_mark(true);
        const long long n16 = grid_elems * (long long)sizeof(R) / 16;  // 4 reals per point: a multiple of 16 bytes
        const unsigned gx = (unsigned)((n16 + 255) / 256);
        if (device_barrier && gx <= (unsigned)sm_count && !getenv("NDS_NO_FUSED_EXCHANGE")) {
          // push + barrier + apply in one launch (all CTAs co-resident: at most one per SM)
          if (!d_ticket) {
            CU(cudaMalloc(&d_ticket, sizeof(unsigned int)));
            CU(cudaMemsetAsync(d_ticket, 0, sizeof(unsigned int), stream));
            if (getenv("NDS_EXCHANGE_TRACE")) {  // phase times of the fused exchange, printed by the destructor
              CU(cudaMalloc(&d_xtrace, sizeof(unsigned long long) * 8));
              CU(cudaMemsetAsync(d_xtrace, 0, sizeof(unsigned long long) * 8, stream));
            }
          }
          barrier_epoch += 1;
          grid_exchange_kernel<R><<<gx, 256, 0, stream>>>(n16, d_ginc, d_slot_dst + par * n_peers, d_slot_src + par * n_peers, n_peers,
                                                       d_grid, d_peer_flags, peer_rank, barrier_epoch, 10000000000ull,
                                                       d_flags_own + NDS_MAX_RANKS, d_ticket, d_xtrace);
          CU(cudaGetLastError());
          exchange_mark(false);
          launches++;
        } else {
          grid_push_kernel<R><<<gx, 256, 0, stream>>>(n16, d_ginc, d_slot_dst + par * n_peers, n_peers);
          CU(cudaGetLastError());
          launches++;
          if (peer_barrier()) return 1;
          grid_apply_kernel<R><<<ga, 256, 0, stream>>>(grid_elems, d_grid, d_slot_src + par * n_peers, n_peers);
          CU(cudaGetLastError());
          exchange_mark(false);
          launches++;
        }
      } else if (comm) {
        if (accumulate(rec, M, d_ginc_list, 1, false)) return 1;
        exchange_mark(true);
        if (nccl_check(nccl_api().AllReduce(d_ginc, d_ginc, (size_t)grid_elems, sizeof(R) == 4 ? 7 : 8, 0, comm, stream), "ncclAllReduce")) return 1;
        exchange_mark(false);
        grid_apply_kernel<R><<<ga, 256, 0, stream>>>(grid_elems, d_grid, d_ginc_list, 1);
        CU(cudaGetLastError());
        launches++;
      } else if (exchange) {
        if (accumulate(rec, M, d_ginc_list, 1, false)) return 1;
        exchange_mark(true);
        CU(cudaStreamSynchronize(stream));
        // rank_offset_bytes == -1: all-reduce (sum) of bytes / sizeof(real) reals of the engine's precision, in place
        const int rc = exchange(exchange_user, (void*)d_ginc, grid_elems * (int64_t)sizeof(R), -1);
        if (rc) return fail(this, "grid exchange callback failed (%d)", rc);
        exchange_mark(false);
        grid_apply_kernel<R><<<ga, 256, 0, stream>>>(grid_elems, d_grid, d_ginc_list, 1);
        CU(cudaGetLastError());
        launches++;
      } else {
        return fail(this, "shared hills over several ranks need nds_set_comm(), nds_open_peer_tables() or nds_set_exchange()");
      }
      n_hills += add;
      dep_count += 1;
      need_refresh_vr = true;
      return 0;
    }
    if (multi && d_peer_tables && !K.mtd_adaptive) {
      // fused deposition + exchange: records go straight into every rank's table over NVLink
      exchange_mark(true);
      deposit_p2p_kernel<R><<<g, 128, 0, stream>>>(
          K, M, cfg.walker_offset, d_colv, d_vr, d_alive, d_peer_tables, n_peers, n_hills, first);
      CU(cudaGetLastError());
      launches++;
      if (peer_barrier()) return 1;
      exchange_mark(false);
      n_hills += add;
      dep_count += 1;
      need_refresh_vr = true;
      return 0;
    }
    deposit_kernel<R><<<g, 128, 0, stream>>>(
        K, M, cfg.walker_offset, d_colv, d_vr, d_alive, d_hills, n_hills, first);
    CU(cudaGetLastError());
    launches++;
    if (d_cvhist) {  // per-hill variance from the walker's colvar history (gaus.py:166-183, 193-212)
      hill_variance_hist_kernel<R><<<g, 128, 0, stream>>>(K, M, cfg.walker_offset, d_cvhist, tao_d, cv_n, tao_d, d_hills, n_hills, first);
      CU(cudaGetLastError());
      launches++;
    } else if (K.mtd_adaptive) {  // per-hill variance J.J^T (* sigma^2) at the deposition point
      if (nd == 1) hill_variance_kernel<R, 1><<<g, 128, 0, stream>>>(K, M, cfg.walker_offset, d_x, d_hills, n_hills, first);
      else if (nd == 2) hill_variance_kernel<R, 2><<<g, 128, 0, stream>>>(K, M, cfg.walker_offset, d_x, d_hills, n_hills, first);
      else hill_variance_kernel<R, 5><<<g, 128, 0, stream>>>(K, M, cfg.walker_offset, d_x, d_hills, n_hills, first);
      CU(cudaGetLastError());
      launches++;
    }
    if (multi) {
      // multi-GPU: complete the segment [n_hills, n_hills + n_walkers_global) with every rank's records
      // (SURVEY 8e): ONE in-place allgather, on the engine's stream (native NCCL) or through the callback
      const int64_t rec = (int64_t)K.hill_stride * (int64_t)sizeof(R);
      R* seg = d_hills + n_hills * K.hill_stride;
      exchange_mark(true);
      if (comm) {
        if (cfg.walker_offset != (int64_t)comm_rank * M)
          return fail(this, "nds_set_comm: walker_offset %lld != rank %d * n_walkers", (long long)cfg.walker_offset, comm_rank);
        if (nccl_check(nccl_api().AllGather((const char*)seg + cfg.walker_offset * rec, seg, (size_t)(M * rec), 0, comm, stream), "ncclAllGather")) return 1;
      } else if (exchange) {
        CU(cudaStreamSynchronize(stream));
        const int rc = exchange(exchange_user, (void*)seg, M * rec, cfg.walker_offset * rec);
        if (rc) return fail(this, "hill exchange callback failed (%d)", rc);
      } else {
        return fail(this, "shared hills over several ranks need nds_set_comm(), nds_open_peer_tables() or nds_set_exchange()");
      }
      exchange_mark(false);
    }
    n_hills += add;
    dep_count += 1;
    need_refresh_vr = true;
    return 0;
  }

  // pack the live walkers into slots [0, n_alive) (see compact_swap_kernel)
  int compact() {
    if (n_active <= 0) return 0;
    CU(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream));
    count_alive_kernel<<<296, 256, 0, stream>>>(n_active, d_alive, d_counter);
    unsigned long long h_alive = 0;
    CU(cudaMemcpyAsync(&h_alive, d_counter, sizeof h_alive, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    launches++;
    const int64_t n_alive = (int64_t)h_alive;
    if (n_alive == n_active || n_active - n_alive < n_active / 16) return 0;  // not worth a pass yet
    if (n_alive > 0) {
      const unsigned g = (unsigned)((n_active + 255) / 256);
      alive_flags_kernel<<<g, 256, 0, stream>>>(n_active, d_alive, d_flags);
      size_t bytes = cub_bytes;
      CU(cub::DeviceScan::ExclusiveSum(d_cub, bytes, d_flags, d_scan, (int)n_active, stream));
      compact_lists_kernel<<<g, 256, 0, stream>>>(n_active, n_alive, d_alive, d_scan, d_holes, d_movers, d_nholes);
      const int64_t hb = n_alive < n_active - n_alive ? n_alive : n_active - n_alive;
      if (hb > 0)
        compact_swap_kernel<R><<<(unsigned)((hb + 127) / 128), 128, 0, stream>>>(
            M, nd, cd, d_holes, d_movers, d_nholes, d_x, d_v, d_f, d_fixf, d_colv, d_thermo, d_vr, d_umb_k,
            d_umb_r0, d_alive, d_basin, d_exit, d_ids);
      CU(cudaGetLastError());
      launches += 4;
      permuted = true;
      ids_host_valid = false;
    }
    n_active = n_alive;
    return 0;
  }

  // host-side inverse permutation of a downloaded [rows][M] array: out[ids[slot]] = in[slot]
  template <typename T>
  int unpermute_host(T* buf, int rows) {
    if (!permuted) return 0;
    if (!ids_host_valid) {  // one download per permutation, shared by all arrays of a nds_get_state call
      ids_host.resize((size_t)M);
      CU(cudaMemcpyAsync(ids_host.data(), d_ids, sizeof(long long) * (size_t)M, cudaMemcpyDeviceToHost, stream));
      CU(cudaStreamSynchronize(stream));
      ids_host_valid = true;
    }
    const std::vector<long long>& ids = ids_host;
    std::vector<T> tmp((size_t)M);
    for (int r = 0; r < rows; r++) {
      T* row = buf + (size_t)r * M;
      for (int64_t s = 0; s < M; s++) tmp[(size_t)ids[(size_t)s]] = row[s];
      memcpy(row, tmp.data(), sizeof(T) * (size_t)M);
    }
    return 0;
  }

  // restore slot == walker on the device (before a new begin)
  int unpermute_device() {
    std::vector<double> buf;
    auto fix = [&](R* d, int rows) -> int {
      if (!d) return 0;
      buf.resize((size_t)rows * M);
      if (download(d, buf.data(), (int64_t)rows * M)) return 1;
      if (unpermute_host<double>(buf.data(), rows)) return 1;
      return upload(buf.data(), d, (int64_t)rows * M);
    };
    if (fix(d_x, nd) || fix(d_v, nd) || fix(d_umb_k, cd) || fix(d_umb_r0, cd)) return 1;
    iota_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(M, d_ids);
    permuted = false;
    return 0;
  }

  // Metropolis MC: NDRun.run with engine = MC (ndrun.py:230-241, engines/mc.py:42-101)
  int run_mc(int64_t nsteps) {
    const int nu = cfg.mc_propose_all ? nd + 1 : 3;
    if (cfg.noise_mode == NDS_NOISE_EXTERNAL && (!d_noise || step < noise_first || step + nsteps > noise_first + noise_steps))
      return fail(this, "external uniform table does not cover moves [%lld, %lld)", (long long)step, (long long)(step + nsteps));
    (void)nu;
    CU(cudaEventRecord(ev0, stream));
    int64_t remaining = nsteps;
    while (remaining > 0) {
      const int64_t k = remaining < cfg.steps_per_launch ? remaining : cfg.steps_per_launch;
      BlockArgs<R> A = make_args(k);
      // metadynamics: every walker deposits inside the kernel on its own clock (accepted moves * dt)
      cudaError_t err = launch_mc_block<R>(nd, K, A, d_mc_accepted, d_wl, d_mc_time, d_mc_dep, d_hills, stream);
      if (err != cudaSuccess) return fail(this, "mc_block launch failed: %s", cudaGetErrorString(err));
      launches++;
      step += k;
      remaining -= k;
    }
    CU(cudaEventRecord(ev1, stream));
    timed = true;
    if (cfg.mtd_enabled) {  // a walker that had a hill to deposit and no room: same failure as the MD path's deposit()
      long long over = 0;
      CU(cudaMemcpyAsync(&over, d_mc_dep + M, sizeof over, cudaMemcpyDeviceToHost, stream));
      CU(cudaStreamSynchronize(stream));
      if (over) return fail(this, "hill table full: walker %lld needed more than mtd_capacity %lld hills (MC deposits on the "
                                  "accepted-move clock); raise mtd_capacity", over - 1, (long long)K.mtd_capacity);
    }
    return 0;
  }

  // MC: NDRun.time and Gaussian._dep_count are per walker; the engine-level values mirror one walker
  int mc_sync_walker(int64_t walker) override {
    if (!(cfg.method == NDS_METHOD_MC || cfg.method == NDS_METHOD_WL)) return 0;
    if (walker < 0 || walker >= M) walker = 0;
    CU(cudaStreamSynchronize(stream));
    long long acc = 0;
    CU(cudaMemcpy(&acc, d_mc_accepted + walker, sizeof acc, cudaMemcpyDeviceToHost));
    time = 0.0;
    for (long long i = 0; i < acc; i++) time += cfg.dt;  // repeated addition like ndrun.py:236
    if (cfg.mtd_enabled) {
      long long dep = 0;
      CU(cudaMemcpy(&dep, d_mc_dep + walker, sizeof dep, cudaMemcpyDeviceToHost));
      n_hills = dep; dep_count = dep;
    }
    return 0;
  }

  int get_wl_histogram(int64_t walker, double* hE) override {
    CU(cudaSetDevice(cfg.device));
    if (cfg.method != NDS_METHOD_WL) return fail(this, "nds_get_wl_histogram: method is not wl-mc");
    if (walker < 0 || walker >= M) return fail(this, "nds_get_wl_histogram: bad walker");
    CU(cudaStreamSynchronize(stream));
    std::vector<R> h(NDS_WL_GRID);
    CU(cudaMemcpy2D(h.data(), sizeof(R), d_wl + walker, sizeof(R) * (size_t)M, sizeof(R), NDS_WL_GRID, cudaMemcpyDeviceToHost));
    for (int i = 0; i < NDS_WL_GRID; i++) hE[i] = (double)h[(size_t)i];
    return 0;
  }

  int set_wl_histogram(int64_t walker, const double* hE) override {
    CU(cudaSetDevice(cfg.device));
    if (cfg.method != NDS_METHOD_WL) return fail(this, "nds_set_wl_histogram: method is not wl-mc");
    if (walker >= M) return fail(this, "nds_set_wl_histogram: bad walker");
    CU(cudaStreamSynchronize(stream));
    std::vector<R> h((size_t)NDS_WL_GRID);
    for (int i = 0; i < NDS_WL_GRID; i++) h[(size_t)i] = (R)hE[i];
    if (walker >= 0) {  // one column of the [NDS_WL_GRID][M] storage
      CU(cudaMemcpy2D(d_wl + walker, sizeof(R) * (size_t)M, h.data(), sizeof(R), sizeof(R), NDS_WL_GRID, cudaMemcpyHostToDevice));
      return 0;
    }
    R* d_tmp = nullptr;
    CU(cudaMalloc(&d_tmp, sizeof(R) * NDS_WL_GRID));
    cudaError_t err = cudaMemcpy(d_tmp, h.data(), sizeof(R) * NDS_WL_GRID, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) {
      const long long tot = (long long)NDS_WL_GRID * M;
      broadcast_rows_kernel<R><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(tot, M, d_tmp, d_wl);
      err = cudaGetLastError();
      launches++;
    }
    cudaStreamSynchronize(stream);
    cudaFree(d_tmp);
    if (err != cudaSuccess) return fail(this, "nds_set_wl_histogram: %s", cudaGetErrorString(err));
    return 0;
  }

  int get_clock(int64_t* s, double* t, int64_t* dep, int64_t* resc) const override {
    if (s) *s = step;
    if (t) *t = time;
    if (dep) *dep = dep_count;
    if (resc) *resc = rescale_count;
    return 0;
  }
  int set_clock(int64_t s, double t, int64_t dep, int64_t resc) override {
    if (!begun) return fail(this, "nds_set_clock before nds_begin");
    step = s; time = t; dep_count = dep; rescale_count = resc;
    need_refresh_vr = false;  // VR comes from the checkpoint (nds_set_vr), not from this launch entry
    return 0;
  }
  int set_vr(const double* vr) override {
    CU(cudaSetDevice(cfg.device));
    return upload(vr, d_vr, M);
  }
  // Harmonic.VR of every restraint, [n_umb][M] (kept only when umbrella + metadynamics record frames): checkpoints
  int get_umb_vr(double* vr, int64_t cap, int64_t* n) override {
    CU(cudaSetDevice(cfg.device));
    const int64_t want = d_vr_umb ? (int64_t)cfg.n_umb * M : 0;
    if (n) *n = want;
    if (!vr || want == 0) return 0;
    if (cap < want) return fail(this, "nds_get_umb_vr: buffer of %lld < %lld elements", (long long)cap, (long long)want);
    return download(d_vr_umb, vr, want);
  }
  int set_umb_vr(const double* vr, int64_t n) override {
    CU(cudaSetDevice(cfg.device));
    if (!d_vr_umb) return fail(this, "nds_set_umb_vr: this engine keeps no umbrella VR (needs umb + mtd biases and frames)");
    if (n != (int64_t)cfg.n_umb * M) return fail(this, "nds_set_umb_vr: expected %lld elements", (long long)cfg.n_umb * M);
    return upload(vr, d_vr_umb, n);
  }
  int64_t count_nonfinite() override {
    cudaSetDevice(cfg.device);
    cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream);
    count_nonfinite_kernel<R><<<296, 256, 0, stream>>>(M, nd, d_x, d_v, d_counter);
    launches++;
    unsigned long long h = 0;
    cudaMemcpyAsync(&h, d_counter, sizeof h, cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    return (int64_t)h;
  }

  int get_mc(int64_t* accepted) override {
    CU(cudaSetDevice(cfg.device));
    if (cfg.method != NDS_METHOD_MC && cfg.method != NDS_METHOD_WL) return fail(this, "nds_get_mc: method is not mc / wl-mc");
    CU(cudaStreamSynchronize(stream));
    CU(cudaMemcpy(accepted, d_mc_accepted, sizeof(long long) * (size_t)M, cudaMemcpyDeviceToHost));
    return 0;
  }

  // Per-walker Monte Carlo bookkeeping for checkpoints: accepted moves, and -- with a metadynamics bias -- every walker's
  // own clock and deposition count; Wang-Landau histograms of all walkers as [NDS_WL_GRID][M].  Null pointers are
  // skipped.  get also makes the raw hill-table getters cover the longest table (n_hills = max deposition count).
  int get_mc_state(int64_t* accepted, double* tm, int64_t* dep, double* wl, int64_t* max_dep) override {
    CU(cudaSetDevice(cfg.device));
    if (cfg.method != NDS_METHOD_MC && cfg.method != NDS_METHOD_WL) return fail(this, "nds_get_mc_state: method is not mc / wl-mc");
    CU(cudaStreamSynchronize(stream));
    if (accepted) CU(cudaMemcpy(accepted, d_mc_accepted, sizeof(long long) * (size_t)M, cudaMemcpyDeviceToHost));
    if (max_dep) *max_dep = 0;
    if (cfg.mtd_enabled) {
      std::vector<long long> d((size_t)M);
      CU(cudaMemcpy(d.data(), d_mc_dep, sizeof(long long) * (size_t)M, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (long long v : d) mx = v > mx ? v : mx;
      if (dep) for (int64_t i = 0; i < M; i++) dep[i] = d[(size_t)i];
      if (tm) CU(cudaMemcpy(tm, d_mc_time, sizeof(double) * (size_t)M, cudaMemcpyDeviceToHost));
      if (max_dep) *max_dep = mx;
      n_hills = mx;
    } else if (tm || dep) {
      return fail(this, "nds_get_mc_state: clocks and deposition counts exist only with a metadynamics bias");
    }
    if (wl) {
      if (!d_wl) return fail(this, "nds_get_mc_state: method is not wl-mc");
      if (download(d_wl, wl, (int64_t)NDS_WL_GRID * M)) return 1;
    }
    return 0;
  }
  int set_mc_state(const int64_t* accepted, const double* tm, const int64_t* dep, const double* wl) override {
    CU(cudaSetDevice(cfg.device));
    if (cfg.method != NDS_METHOD_MC && cfg.method != NDS_METHOD_WL) return fail(this, "nds_set_mc_state: method is not mc / wl-mc");
    CU(cudaStreamSynchronize(stream));
    if (accepted) CU(cudaMemcpy(d_mc_accepted, accepted, sizeof(long long) * (size_t)M, cudaMemcpyHostToDevice));
    if ((tm || dep) && !cfg.mtd_enabled) return fail(this, "nds_set_mc_state: clocks and deposition counts exist only with a metadynamics bias");
    if (tm) CU(cudaMemcpy(d_mc_time, tm, sizeof(double) * (size_t)M, cudaMemcpyHostToDevice));
    if (dep) {
      std::vector<long long> d((size_t)M);
      for (int64_t i = 0; i < M; i++) {
        if (dep[i] < 0 || dep[i] > cfg.mtd_capacity) return fail(this, "nds_set_mc_state: deposition count of walker %lld out of range", (long long)i);
        d[(size_t)i] = dep[i];
      }
      CU(cudaMemcpy(d_mc_dep, d.data(), sizeof(long long) * (size_t)M, cudaMemcpyHostToDevice));
    }
    if (wl) {
      if (!d_wl) return fail(this, "nds_set_mc_state: method is not wl-mc");
      if (upload(wl, d_wl, (int64_t)NDS_WL_GRID * M)) return 1;
    }
    return 0;
  }

  // thermo[5][M] = (T, pe, ke, biase, totale): under MC these follow the last EVENT (mc.py:93-95 zeroes ke on the first
  // accepted move, a deposition sets totale, gaus.py:146), so a resumed run has to get them back
  int set_thermo(const double* thermo) override {
    CU(cudaSetDevice(cfg.device));
    return upload(thermo, d_thermo, (int64_t)5 * M);
  }

  // one fused launch of k steps over slots [slot0, slot0 + nslots) (nslots == 0: every launched slot) on stream st;
  // rest = true: the variant without frames (walkers beyond the dump prefix of a split launch)
  cudaError_t launch_steps(int64_t k, int64_t slot0, int64_t nslots, cudaStream_t st, bool rest = false) {
    BlockArgs<R> A = make_args(k);
    A.refresh_vr = need_refresh_vr ? 1 : 0;
    if (nslots > 0) { A.slot0 = slot0; A.n_active = slot0 + nslots; }
    if (launch_packed) {
      // same buffers, read as pairs: SoA float[d][M] == f2[d][M/2]
      BlockArgs<f2> P;
      memset(&P, 0, sizeof P);
      P.M = M / 2; P.walker_offset = A.walker_offset; P.step0 = A.step0; P.nsteps = A.nsteps;
      P.time0 = A.time0; P.rescale_count = A.rescale_count;
      P.x = (f2*)A.x; P.v = (f2*)A.v; P.f = (f2*)A.f; P.fixf = (f2*)A.fixf; P.colv = (f2*)A.colv;
      P.thermo = (f2*)A.thermo; P.vr = (f2*)A.vr;
      P.umb_k = (const f2*)A.umb_k; P.umb_r0 = (const f2*)A.umb_r0;
      // committor: a thread owns slots 2t and 2t+1 of the (compacted) slot order
      P.alive = A.alive; P.basin = A.basin; P.exit_step = A.exit_step; P.ids = A.ids;
      P.n_active = A.n_active > 0 ? (A.n_active + 1) / 2 : 0;
      P.slot0 = A.slot0 / 2;
      P.frames = (f2*)A.frames; P.frame_base = A.frame_base; P.frame_cap = A.frame_cap; P.frame_width = A.frame_width;
      P.hist = A.hist; P.hist_smem_off = A.hist_smem_off; P.dyn_smem = A.dyn_smem;
      P.begin_mode = A.begin_mode; P.need_init_v = A.need_init_v;
      return rest ? launch_packed_rest(K2_rest, P, st) : launch_packed(K2, P, st);
    }
    return rest ? launch_block_rest(K_rest, A, st) : launch_block(K, A, st);
  }

  // k steps for every walker: one launch, or -- subset dumps -- the dump prefix on a second stream beside the rest
  cudaError_t launch_all(int64_t k) {
    if (split_at <= 0 || split_at >= M) return launch_steps(k, 0, 0, stream);
    if (!s_alt) {
      cudaError_t e = cudaStreamCreateWithFlags(&s_alt, cudaStreamNonBlocking);
      if (e != cudaSuccess) return e;
    }
    while (pipe_events.size() < 2) {
      cudaEvent_t ev;
      cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
      pipe_events.push_back(ev);
    }
    cudaEventRecord(pipe_events[0], stream);
    cudaStreamWaitEvent(s_alt, pipe_events[0], 0);
    cudaError_t e = launch_steps(k, 0, split_at, s_alt);
    if (e != cudaSuccess) return e;
    e = launch_steps(k, split_at, M - split_at, stream, true);
    if (e != cudaSuccess) return e;
    launches++;
    cudaEventRecord(pipe_events[1], s_alt);
    return cudaStreamWaitEvent(stream, pipe_events[1], 0);
  }

  // asynchronous host path: ONE pinned block per direction.  host_in = [2 ndim][M] (x rows, then v rows), host_out =
  // [2 ndim + 5][M] (x, v, thermo), both in the engine's precision; either may be NULL.  chunks > 1 pipelines the
  // pass over walker chunks (H2D of chunk c+1 and D2H of chunk c-1 overlap the kernel of chunk c on their own
  // streams): legal because the walkers of these ensembles are independent.  Returns after enqueueing;
  // nds_wait() (= nds_sync) completes the pass.
  int submit(const void* host_in, int64_t nsteps, void* host_out, int chunks) override {
    CU(cudaSetDevice(cfg.device));
    if (!begun) return fail(this, "nds_submit before nds_begin");
    if (permuted || compactable) return fail(this, "nds_submit: committor ensembles with lane compaction use nds_set_state / nds_run / nds_get_state");
    const size_t row = sizeof(R) * (size_t)M;
    const bool can_chunk = chunks > 1 && !cfg.mtd_enabled && cfg.method == NDS_METHOD_MD && (K.dump_freq == 0 || split_at > 0) &&
                           nsteps <= cfg.steps_per_launch && cfg.integrator != NDS_INT_RESCALE &&
                           (cfg.noise_mode == NDS_NOISE_PHILOX || normals_per_step(cfg.integrator, nd) == 0);
    if (!can_chunk) {
      if (host_in) { CU(cudaMemcpyAsync(d_x, host_in, row * 2 * nd, cudaMemcpyHostToDevice, stream)); have_v = true; }
      if (run(nsteps)) return 1;
      if (host_out) CU(cudaMemcpyAsync(host_out, d_x, row * (2 * nd + 5), cudaMemcpyDeviceToHost, stream));
      return 0;
    }
    if (!s_h2d) {
      CU(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
      CU(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
      CU(cudaStreamCreateWithFlags(&s_alt, cudaStreamNonBlocking));
    }
    // chunk = whole CTAs of the step kernel (128 threads, two walkers per thread when packed)
    const int64_t unit = 128 * (launch_packed ? 2 : 1);
    int64_t per = ((M + chunks - 1) / chunks + unit - 1) / unit * unit;
    const int nch = (int)((M + per - 1) / per);
    while ((int)pipe_events.size() < 2 * nch + 2) {
      cudaEvent_t ev;
      CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      pipe_events.push_back(ev);
    }
    int64_t events = 0;
    if (frames_on) {
      for (int64_t s = step + 1; s <= step + nsteps; s++) events += frame_event(s, K.dump_freq, K.stat_freq) ? 1 : 0;
      if (frames_written + events > cfg.dump_capacity)
        return fail(this, "frame ring full (%lld frames pending, capacity %lld): call nds_drain_frames",
                    (long long)frames_written, (long long)cfg.dump_capacity);
    }
    cudaEvent_t ev_start = pipe_events[(size_t)(2 * nch)], ev_end = pipe_events[(size_t)(2 * nch + 1)];
    CU(cudaEventRecord(ev0, stream));
    CU(cudaEventRecord(ev_start, stream));  // everything queued so far (previous pass) precedes the new copies
    CU(cudaStreamWaitEvent(s_h2d, ev_start, 0));
    CU(cudaStreamWaitEvent(s_alt, ev_start, 0));
    CU(cudaStreamWaitEvent(s_d2h, ev_start, 0));
    if (host_in) have_v = true;
    for (int c = 0; c < nch; c++) {
      const int64_t w0 = c * per, wn = (w0 + per <= M ? per : M - w0);
      cudaStream_t cs = (c & 1) ? s_alt : stream;
      cudaEvent_t e_in = pipe_events[(size_t)(2 * c)], e_k = pipe_events[(size_t)(2 * c + 1)];
      if (host_in) {
        CU(cudaMemcpy2DAsync(d_x + w0, row, (const R*)host_in + w0, row, sizeof(R) * (size_t)wn, (size_t)(2 * nd),
                             cudaMemcpyHostToDevice, s_h2d));
        CU(cudaEventRecord(e_in, s_h2d));
        CU(cudaStreamWaitEvent(cs, e_in, 0));
      }
      // subset dumps: the part of the chunk inside the dump prefix runs the frame-recording variant
      const int64_t npre = split_at > w0 ? (split_at - w0 < wn ? split_at - w0 : wn) : 0;
      cudaError_t err = cudaSuccess;
      if (split_at == 0) { err = launch_steps(nsteps, w0, wn, cs); launches++; }
      else {
        if (npre > 0) { err = launch_steps(nsteps, w0, npre, cs); launches++; }
        if (err == cudaSuccess && wn > npre) { err = launch_steps(nsteps, w0 + npre, wn - npre, cs, true); launches++; }
      }
      if (err != cudaSuccess) return fail(this, "md_block launch failed: %s", cudaGetErrorString(err));
      if (host_out) {
        CU(cudaEventRecord(e_k, cs));
        CU(cudaStreamWaitEvent(s_d2h, e_k, 0));
        CU(cudaMemcpy2DAsync((R*)host_out + w0, row, d_x + w0, row, sizeof(R) * (size_t)wn, (size_t)(2 * nd + 5),
                             cudaMemcpyDeviceToHost, s_d2h));
      }
    }
    // join: the engine stream completes when every chunk has
    CU(cudaEventRecord(ev_end, s_alt));
    CU(cudaStreamWaitEvent(stream, ev_end, 0));
    CU(cudaEventRecord(ev_end, s_d2h));
    CU(cudaStreamWaitEvent(stream, ev_end, 0));
    CU(cudaEventRecord(ev1, stream));
    timed = true;
    need_refresh_vr = false;
    for (int64_t j = 0; j < nsteps; j++) time += step_dt();  // ndrun.py:240
    step += nsteps;
    frames_written += events;
    return 0;
  }

  int run(int64_t nsteps) override {  // the NDRun.run loop (ndrun.py:214-251)
    CU(cudaSetDevice(cfg.device));
    if (!begun) return fail(this, "nds_run before nds_begin");
    if (cfg.method == NDS_METHOD_MC || cfg.method == NDS_METHOD_WL) return run_mc(nsteps);
    if (cfg.noise_mode == NDS_NOISE_EXTERNAL && normals_per_step(cfg.integrator, nd) > 0) {
      if (!d_noise || step < noise_first || step + nsteps > noise_first + noise_steps)
        return fail(this, "external noise table does not cover steps [%lld, %lld)", (long long)step,
                    (long long)(step + nsteps));
    }
    CU(cudaEventRecord(ev0, stream));
    int64_t remaining = nsteps;
    while (remaining > 0) {
      int64_t k = remaining < cfg.steps_per_launch ? remaining : cfg.steps_per_launch;
      if (cfg.mtd_enabled) {
        // fix.update(step, time) before the step (ndrun.py:227-228, gaus.py:134-136)
        if ((int64_t)std::floor(time / cfg.mtd_dep_freq) + 1 > dep_count) {
          if (deposit()) return 1;
        }
        // the launch ends right before the next step that would deposit
        double t = time;
        int64_t j = 0;
        while (j < k) {
          t += step_dt(); j++;
          if ((int64_t)std::floor(t / cfg.mtd_dep_freq) + 1 > dep_count) break;
        }
        k = j;
      }
      int64_t events = 0;
      if (frames_on) {
        for (int64_t s = step + 1; s <= step + k; s++) events += frame_event(s, K.dump_freq, K.stat_freq) ? 1 : 0;
        if (frames_written + events > cfg.dump_capacity)
          return fail(this, "frame ring full (%lld frames pending, capacity %lld): call nds_drain_frames",
                      (long long)frames_written, (long long)cfg.dump_capacity);
      }
      cudaError_t err = launch_all(k);
      if (err != cudaSuccess) return fail(this, "md_block launch failed: %s", cudaGetErrorString(err));
      launches++;
      need_refresh_vr = false;
      for (int64_t j = 0; j < k; j++) {  // replay the walker-independent schedules on the host
        if (cfg.integrator == NDS_INT_RESCALE) {
          const int64_t count = (int64_t)std::floor(time / cfg.rescale_freq) + 1;  // md.py:220-222
          if (count > rescale_count) rescale_count = count;
        }
        time += step_dt();  // ndrun.py:240 (engine.current_dt)
      }
      if (d_cvhist) cv_n += (step + k) / K.stat_freq - step / K.stat_freq;  // samples the launch appended
      step += k;
      frames_written += events;
      remaining -= k;
      if (compactable && compact()) return 1;
    }
    CU(cudaEventRecord(ev1, stream));
    timed = true;
    return 0;
  }

  int sync() override {
    CU(cudaSetDevice(cfg.device));
    CU(cudaStreamSynchronize(stream));
    if (device_barrier && barrier_epoch > 0) {
      unsigned long long bad = 0;
      CU(cudaMemcpy(&bad, d_flags_own + NDS_MAX_RANKS, sizeof bad, cudaMemcpyDeviceToHost));
      if (bad) return fail(this, "peer barrier %llu timed out: a rank did not arrive within 10 s", bad);
    }
    return 0;
  }

  double last_run_ms() override {
    if (!timed) return 0.0;
    cudaSetDevice(cfg.device);
    if (cudaEventSynchronize(ev1) != cudaSuccess) return -1.0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev0, ev1) != cudaSuccess) return -1.0;
    return (double)ms;
  }

  int get_state(double* x, double* v, double* f, double* fixf, double* colv, double* thermo) override {
    CU(cudaSetDevice(cfg.device));
    if (x && (download(d_x, x, (int64_t)nd * M) || unpermute_host(x, nd))) return 1;
    if (v && (download(d_v, v, (int64_t)nd * M) || unpermute_host(v, nd))) return 1;
    if (f && (download(d_f, f, (int64_t)nd * M) || unpermute_host(f, nd))) return 1;
    if (fixf && (download(d_fixf, fixf, (int64_t)nd * M) || unpermute_host(fixf, nd))) return 1;
    if (colv && (download(d_colv, colv, (int64_t)cd * M) || unpermute_host(colv, cd))) return 1;
    if (thermo && (download(d_thermo, thermo, (int64_t)5 * M) || unpermute_host(thermo, 5))) return 1;
    return 0;
  }

  int eval_at(const double* x, int64_t n, double* V, double* F, double* colv, double* J, double* Vb,
              double* Fb) override {
    CU(cudaSetDevice(cfg.device));
    if (n <= 0) return 0;
    if (mc_sync_walker(0)) return 1;  // MC: walker 0's table
    if ((cfg.colvar == NDS_CV_PATH || cfg.colvar == NDS_CV_PATH5DTO2D) && !d_path_ref)
      return fail(this, "path colvar: call nds_set_path() before nds_eval_at()");
    const size_t sn = (size_t)n;
    const size_t total = sn * (size_t)(nd + 1 + nd + cd + cd * nd + 1 + nd);
    if (total > eval_elems) {  // scratch grows on demand and is kept
      if (d_eval) CU(cudaFree(d_eval));
      d_eval = nullptr; eval_elems = 0;
      CU(cudaMalloc(&d_eval, sizeof(R) * total));
      eval_elems = total;
    }
    R* buf = d_eval;
    R* dx = buf; R* dV = dx + nd * sn; R* dF = dV + sn; R* dc = dF + nd * sn; R* dJ = dc + cd * sn;
    R* dVb = dJ + (size_t)cd * nd * sn; R* dFb = dVb + sn;
    int rc = upload(x, dx, (int64_t)nd * n);
    if (!rc) {
      const unsigned grid = (unsigned)((n + NDS_BLOCK - 1) / NDS_BLOCK);
      const long long es = cfg.mtd_shared ? 1 : M;
      if (nd == 1) eval_at_kernel<R, 1><<<grid, NDS_BLOCK, 0, stream>>>(K, n, dx, d_hills, n_hills, es, dV, dF, dc, dJ, dVb, dFb, d_grid);
      else if (nd == 2) eval_at_kernel<R, 2><<<grid, NDS_BLOCK, 0, stream>>>(K, n, dx, d_hills, n_hills, es, dV, dF, dc, dJ, dVb, dFb, d_grid);
      else eval_at_kernel<R, 5><<<grid, NDS_BLOCK, 0, stream>>>(K, n, dx, d_hills, n_hills, es, dV, dF, dc, dJ, dVb, dFb, d_grid);
      cudaError_t err = cudaGetLastError();
      if (err != cudaSuccess) rc = fail(this, "eval_at launch failed: %s", cudaGetErrorString(err));
      launches++;
    }
    if (!rc && V) rc = download(dV, V, n);
    if (!rc && F) rc = download(dF, F, (int64_t)nd * n);
    if (!rc && colv) rc = download(dc, colv, (int64_t)cd * n);
    if (!rc && J) rc = download(dJ, J, (int64_t)cd * nd * n);
    if (!rc && Vb) rc = download(dVb, Vb, n);
    if (!rc && Fb) rc = download(dFb, Fb, (int64_t)nd * n);
    cudaStreamSynchronize(stream);
    return rc;
  }

  int64_t hill_table_size(int64_t n) const override {
    return n * (int64_t)K.hill_stride * (cfg.mtd_shared ? 1 : M);
  }
  int get_hill_table_raw(double* out, int64_t cap) override {
    CU(cudaSetDevice(cfg.device));
    if (!cfg.mtd_enabled) return fail(this, "nds_get_hill_table_raw: metadynamics is not enabled");
    const int64_t n = hill_table_size(n_hills);
    if (cap < n) return fail(this, "nds_get_hill_table_raw: capacity %lld < %lld", (long long)cap, (long long)n);
    return n ? download(d_hills, out, n) : 0;
  }
  int set_hill_table_raw(int64_t n, const double* in) override {
    CU(cudaSetDevice(cfg.device));
    if (!cfg.mtd_enabled) return fail(this, "nds_set_hill_table_raw: metadynamics is not enabled");
    if (n < 0 || n > cfg.mtd_capacity) return fail(this, "nds_set_hill_table_raw: %lld records > capacity", (long long)n);
    if (clear_hills()) return 1;
    if (n && upload(in, d_hills, hill_table_size(n))) return 1;
    n_hills = n;
    need_refresh_vr = false;
    return 0;
  }

  // first n records of one table as doubles, record-major [n][hill_stride].  Per-walker tables live
  // field-interleaved over the walkers (field j of record i of walker w at (i * stride + j) * M + w): a strided
  // 2-D copy moves exactly the n * stride values of that walker.
  int download_table(int64_t walker, int64_t n, std::vector<double>& out) {
    const int rs = K.hill_stride;
    out.assign((size_t)(n * rs), 0.0);
    if (n <= 0) return 0;
    if (cfg.mtd_shared) return download(d_hills, out.data(), n * rs);
    if (walker < 0 || walker >= M) return fail(this, "hill table: bad walker");
    std::vector<R> tmp((size_t)(n * rs));
    CU(cudaMemcpy2DAsync(tmp.data(), sizeof(R), d_hills + walker, sizeof(R) * (size_t)M, sizeof(R), (size_t)(n * rs),
                         cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    for (size_t i = 0; i < tmp.size(); i++) out[i] = (double)tmp[i];
    return 0;
  }
  // install n records (record-major doubles) into one walker's table, or into every walker's (walker < 0)
  int upload_table(int64_t walker, int64_t n, const std::vector<double>& in) {
    const int rs = K.hill_stride;
    if (n <= 0) return 0;
    if (cfg.mtd_shared) return upload(in.data(), d_hills, n * rs);
    std::vector<R> tmp((size_t)(n * rs));
    for (size_t i = 0; i < tmp.size(); i++) tmp[i] = (R)in[i];
    if (walker >= 0) {
      CU(cudaMemcpy2DAsync(d_hills + walker, sizeof(R) * (size_t)M, tmp.data(), sizeof(R), sizeof(R), (size_t)(n * rs),
                           cudaMemcpyHostToDevice, stream));
      CU(cudaStreamSynchronize(stream));
      return 0;
    }
    // every walker: upload the n * stride values once, broadcast on the device
    R* d_tmp = nullptr;
    CU(cudaMalloc(&d_tmp, sizeof(R) * tmp.size()));
    cudaError_t err = cudaMemcpyAsync(d_tmp, tmp.data(), sizeof(R) * tmp.size(), cudaMemcpyHostToDevice, stream);
    if (err == cudaSuccess) {
      const long long tot = (long long)tmp.size() * M;
      broadcast_rows_kernel<R><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(tot, M, d_tmp, d_hills);
      err = cudaGetLastError();
      launches++;
    }
    cudaStreamSynchronize(stream);
    cudaFree(d_tmp);
    if (err != cudaSuccess) return fail(this, "hill table upload: %s", cudaGetErrorString(err));
    return 0;
  }

  int get_hill_sigma2(int64_t walker, double* sigma2, int64_t cap) override {
    CU(cudaSetDevice(cfg.device));
    if (mc_sync_walker(walker)) return 1;
    if (!cfg.mtd_enabled) return fail(this, "nds_get_hill_sigma2: metadynamics is not enabled");
    if (n_hills == 0 || !sigma2) return 0;
    if (cap < n_hills) return fail(this, "nds_get_hill_sigma2: capacity %lld < %lld hills", (long long)cap, (long long)n_hills);
    if (!cfg.mtd_shared && (walker < 0 || walker >= M)) return fail(this, "nds_get_hill_sigma2: bad walker");
    if (!K.mtd_adaptive) {
      for (int64_t i = 0; i < n_hills; i++) sigma2[i] = cfg.mtd_sigma[0] * cfg.mtd_sigma[0];
      return 0;
    }
    const int rs = K.hill_stride;
    std::vector<double> h;
    if (download_table(walker, n_hills, h)) return 1;
    for (int64_t i = 0; i < n_hills; i++) sigma2[i] = h[(size_t)(i * rs + 3)];
    return 0;
  }

  // follows set_hills() when an adaptive table is restored
  int set_hill_sigma2(int64_t walker, int64_t n, const double* sigma2) override {
    CU(cudaSetDevice(cfg.device));
    if (!K.mtd_adaptive) return fail(this, "nds_set_hill_sigma2: the table is not adaptive (mtd_adapsig_geo = 0)");
    if (n != n_hills) return fail(this, "nds_set_hill_sigma2: %lld values for %lld hills", (long long)n, (long long)n_hills);
    if (n == 0) return 0;
    const int rs = K.hill_stride;
    // fields 2, 3 of every record: 1/s and s; one walker's table, or (walker < 0) the same values for all of them
    std::vector<double> h;
    if (download_table(walker < 0 ? 0 : walker, n, h)) return 1;
    for (int64_t i = 0; i < n; i++) { h[(size_t)(i * rs + 2)] = 1.0 / sigma2[i]; h[(size_t)(i * rs + 3)] = sigma2[i]; }
    return upload_table(walker, n, h);
  }

  int get_hills(int64_t walker, int64_t* n, double* centers, double* heights, int64_t cap) override {
    CU(cudaSetDevice(cfg.device));
    if (mc_sync_walker(walker)) return 1;
    if (n) *n = n_hills;
    if (!cfg.mtd_enabled || n_hills == 0 || (!centers && !heights)) return 0;
    if (cap < n_hills) return fail(this, "nds_get_hills: capacity %lld < %lld hills", (long long)cap, (long long)n_hills);
    if (!cfg.mtd_shared && (walker < 0 || walker >= M)) return fail(this, "nds_get_hills: bad walker");
    const int rs = K.hill_stride;
    std::vector<double> h;
    if (download_table(walker, n_hills, h)) return 1;
    for (int64_t i = 0; i < n_hills; i++) {
      for (int k = 0; k < cd; k++)
        if (centers) centers[i * cd + k] = h[(size_t)(i * rs + k)];
      if (heights) heights[i] = h[(size_t)(i * rs + K.hill_hidx)];
    }
    return 0;
  }

  int set_hills(int64_t walker, int64_t n, const double* centers, const double* heights) override {
    CU(cudaSetDevice(cfg.device));
    if (!cfg.mtd_enabled) return fail(this, "nds_set_hills: metadynamics is not enabled");
    if (n > cfg.mtd_capacity) return fail(this, "nds_set_hills: %lld hills > capacity", (long long)n);
    const int rs = K.hill_stride;
    if (cfg.mtd_shared) {
      std::vector<double> h((size_t)(n * rs), 0.0);
      for (int64_t i = 0; i < n; i++) {
        for (int k = 0; k < cd; k++) h[(size_t)(i * rs + k)] = centers[i * cd + k];
        h[(size_t)(i * rs + K.hill_hidx)] = heights[i];
        for (int k = K.hill_hidx + 1; k < rs; k++) h[(size_t)(i * rs + k)] = std::log2(heights[i]);
        if (K.mtd_adaptive) {  // default variance sigma^2 until nds_set_hill_sigma2 supplies the history
          h[(size_t)(i * rs + 2)] = 1.0 / (cfg.mtd_sigma[0] * cfg.mtd_sigma[0]);
          h[(size_t)(i * rs + 3)] = cfg.mtd_sigma[0] * cfg.mtd_sigma[0];
        }
      }
      if (n > table_cap()) return fail(this, "nds_set_hills: %lld hills > capacity %lld of this rank's table", (long long)n, (long long)table_cap());
      if (clear_hills()) return 1;  // padding records stay (0, 0, 0, -inf)
      if (n && upload(h.data(), d_hills, n * rs)) return 1;
      if (grid_mode && rebuild_grid(n)) return 1;
    } else {
      // per-walker tables: the given hills are installed for EVERY walker (walker < 0) or one walker
      if (walker >= M) return fail(this, "nds_set_hills: bad walker");
      std::vector<double> h((size_t)(n * rs), 0.0);
      for (int64_t i = 0; i < n; i++) {
        for (int k = 0; k < cd; k++) h[(size_t)(i * rs + k)] = centers[i * cd + k];
        h[(size_t)(i * rs + K.hill_hidx)] = heights[i];
        if (K.mtd_adaptive) {
          h[(size_t)(i * rs + 2)] = 1.0 / (cfg.mtd_sigma[0] * cfg.mtd_sigma[0]);
          h[(size_t)(i * rs + 3)] = cfg.mtd_sigma[0] * cfg.mtd_sigma[0];
        }
      }
      if (n && upload_table(walker, n, h)) return 1;
    }
    n_hills = n;
    need_refresh_vr = false;
    return 0;
  }

  int get_vr(double* vr) override {
    CU(cudaSetDevice(cfg.device));
    return download(d_vr, vr, M) || unpermute_host(vr, 1);
  }

  int get_commit(int32_t* basin, int64_t* exit_step) override {
    CU(cudaSetDevice(cfg.device));
    if (cfg.method != NDS_METHOD_COMMITTOR) return fail(this, "nds_get_commit: method is not committor");
    CU(cudaStreamSynchronize(stream));
    if (basin) {
      CU(cudaMemcpy(basin, d_basin, sizeof(int) * (size_t)M, cudaMemcpyDeviceToHost));
      if (unpermute_host<int32_t>(basin, 1)) return 1;
    }
    if (exit_step) {
      std::vector<unsigned char> alive((size_t)M);
      std::vector<long long> ex((size_t)M);
      CU(cudaMemcpy(ex.data(), d_exit, sizeof(long long) * (size_t)M, cudaMemcpyDeviceToHost));
      CU(cudaMemcpy(alive.data(), d_alive, (size_t)M, cudaMemcpyDeviceToHost));
      for (int64_t w = 0; w < M; w++)
        if (alive[(size_t)w]) ex[(size_t)w] = step;  // uncommitted: ran all steps (ndrun.py:220)
      if (unpermute_host<long long>(ex.data(), 1)) return 1;
      for (int64_t w = 0; w < M; w++) exit_step[w] = ex[(size_t)w];
    }
    return 0;
  }

  int get_alive(uint8_t* alive) override {
    CU(cudaSetDevice(cfg.device));
    if (cfg.method != NDS_METHOD_COMMITTOR) return fail(this, "nds_get_alive: method is not committor");
    CU(cudaStreamSynchronize(stream));
    CU(cudaMemcpy(alive, d_alive, (size_t)M, cudaMemcpyDeviceToHost));
    return unpermute_host<uint8_t>(alive, 1);
  }

  int set_commit(const int32_t* basin, const int64_t* exit_step, const uint8_t* alive) override {
    CU(cudaSetDevice(cfg.device));
    if (cfg.method != NDS_METHOD_COMMITTOR) return fail(this, "nds_set_commit: method is not committor");
    if (permuted && unpermute_device()) return 1;  // slots back to walker order
    n_active = M;
    std::vector<long long> ex((size_t)M);
    for (int64_t w = 0; w < M; w++) ex[(size_t)w] = alive[w] ? 0 : exit_step[w];
    CU(cudaStreamSynchronize(stream));
    CU(cudaMemcpy(d_basin, basin, sizeof(int) * (size_t)M, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_exit, ex.data(), sizeof(long long) * (size_t)M, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_alive, alive, (size_t)M, cudaMemcpyHostToDevice));
    return 0;
  }

  int64_t count_alive() override {
    if (cfg.method != NDS_METHOD_COMMITTOR) return M;
    // after compaction the live walkers are exactly the launched slots that have not stopped since
    cudaSetDevice(cfg.device);
    cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream);
    count_alive_kernel<<<296, 256, 0, stream>>>(M, d_alive, d_counter);
    launches++;
    unsigned long long h = 0;
    cudaMemcpyAsync(&h, d_counter, sizeof h, cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    return (int64_t)h;
  }

  int drain_frames(double* out, int64_t max_frames, int64_t* n_frames) override {
    CU(cudaSetDevice(cfg.device));
    if (!out) {  // discard: rewind the ring without copying (frames stay on the device)
      if (n_frames) *n_frames = frames_written;
      frames_written = 0;
      return 0;
    }
    const int64_t n = frames_written < max_frames ? frames_written : max_frames;
    if (n_frames) *n_frames = n;
    if (n <= 0) return 0;
    if (n != frames_written) return fail(this, "nds_drain_frames: buffer holds %lld frames, %lld pending", (long long)max_frames, (long long)frames_written);
    const size_t cnt = (size_t)n * frame_width * (size_t)K.dump_walkers;
    if (sizeof(R) == sizeof(double)) {
      CU(cudaMemcpyAsync(out, d_frames, sizeof(double) * cnt, cudaMemcpyDeviceToHost, stream));
    } else {  // float ring: widen on the host (the API delivers doubles)
      std::vector<R> tmp(cnt);
      CU(cudaMemcpyAsync(tmp.data(), d_frames, sizeof(R) * cnt, cudaMemcpyDeviceToHost, stream));
      CU(cudaStreamSynchronize(stream));
      for (size_t i = 0; i < cnt; i++) out[i] = (double)tmp[i];
    }
    CU(cudaMemsetAsync(d_frames, 0xff, sizeof(R) * cnt, stream));
    CU(cudaStreamSynchronize(stream));
    frames_written = 0;
    return 0;
  }

  int fes_grid(int64_t walker, double x0, double dx, int64_t nx, double y0, double dy, int64_t ny,
               double* grid) override {
    CU(cudaSetDevice(cfg.device));
    if (!cfg.mtd_enabled) return fail(this, "nds_fes_grid: metadynamics is not enabled");
    if (mc_sync_walker(walker)) return 1;
    if (cd > 2) return fail(this, "nds_fes_grid: colvardim %d > 2", cd);
    if (cd == 1) ny = 1;
    const int64_t np = nx * ny;
    double* d_out = nullptr;
    CU(cudaMalloc(&d_out, sizeof(double) * (size_t)np));
    const R* base = cfg.mtd_shared ? d_hills : d_hills + walker;
    if (grid_mode)  // the accumulated grid IS the sum of all ranks' hills: sample its interpolant
      grid_sample_kernel<R><<<(unsigned)((np + 127) / 128), 128, 0, stream>>>(K, d_grid, x0, dx, nx, y0, dy, ny, d_out);
    else
      fes_grid_kernel<R><<<(unsigned)((np + 127) / 128), 128, 0, stream>>>(
          K, base, n_hills, cfg.mtd_shared ? 1 : M, x0, dx, nx, y0, dy, ny, d_out);
    launches++;
    cudaError_t err = cudaMemcpyAsync(grid, d_out, sizeof(double) * (size_t)np, cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    cudaFree(d_out);
    if (err != cudaSuccess) return fail(this, "nds_fes_grid: %s", cudaGetErrorString(err));
    return 0;
  }

  int histogram(double x0, double dx, int64_t nx, double y0, double dy, int64_t ny, int64_t* counts) override {
    CU(cudaSetDevice(cfg.device));
    if (cd > 2) return fail(this, "nds_colvar_histogram: colvardim %d > 2", cd);
    if (cd == 1) ny = 1;
    const int64_t nb = nx * ny;
    if (nb <= 0 || nb * (int64_t)sizeof(unsigned) > 48 * 1024)
      return fail(this, "nds_colvar_histogram: %lld bins do not fit the shared-memory histogram (max 12288)", (long long)nb);
    unsigned long long* d_counts = nullptr;
    CU(cudaMalloc(&d_counts, sizeof(unsigned long long) * (size_t)nb));
    CU(cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * (size_t)nb, stream));
    histogram_kernel<R><<<296, 256, sizeof(unsigned) * (size_t)nb, stream>>>(
        M, cd, d_colv, d_alive, x0, 1.0 / dx, (int)nx, y0, 1.0 / dy, (int)ny, d_counts);
    launches++;
    cudaError_t err = cudaMemcpyAsync(counts, d_counts, sizeof(unsigned long long) * (size_t)nb, cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    cudaFree(d_counts);
    if (err != cudaSuccess) return fail(this, "nds_colvar_histogram: %s", cudaGetErrorString(err));
    return 0;
  }

  int frames_f32() const override { return sizeof(R) == 4 ? 1 : 0; }
  int get_histogram(int64_t* counts, int clear) override {
    CU(cudaSetDevice(cfg.device));
    if (!d_hist) return fail(this, "nds_get_histogram: hist_n is 0 (no in-kernel histogram configured)");
    const size_t nb = (size_t)K.hist_nx * K.hist_ny;
    CU(cudaMemcpyAsync(counts, d_hist, sizeof(unsigned long long) * nb, cudaMemcpyDeviceToHost, stream));
    if (clear) CU(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * nb, stream));
    CU(cudaStreamSynchronize(stream));
    return 0;
  }

  int hills_ipc_handle(void* handle) override {
    CU(cudaSetDevice(cfg.device));
    if (!d_hills) return fail(this, "nds_hills_ipc_handle: metadynamics is not enabled");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, d_hills));
    memcpy(handle, &h, sizeof h);
    return 0;
  }

  int set_peer_tables(void* const* tables, int nranks, int rank, bool opened_by_ipc) override {
    CU(cudaSetDevice(cfg.device));
    if (!d_hills || !cfg.mtd_shared) return fail(this, "peer tables need a shared hill table");
    if (K.mtd_adaptive) return fail(this, "peer tables do not carry per-hill variances (mtd_adapsig_geo): use the allgather exchange");
    if (nranks < 1 || rank < 0 || rank >= nranks || nranks > NDS_MAX_RANKS) return fail(this, "bad rank / nranks");
    if (grid_mode && nranks != nranks_cfg) return fail(this, "peer tables: %d ranks, config says %d", nranks, nranks_cfg);
    std::vector<R*> t((size_t)nranks);
    for (int r = 0; r < nranks; r++) t[(size_t)r] = (r == rank) ? d_hills : (R*)tables[r];
    if (d_peer_tables) CU(cudaFree(d_peer_tables));
    CU(cudaMalloc(&d_peer_tables, sizeof(R*) * (size_t)nranks));
    CU(cudaMemcpy(d_peer_tables, t.data(), sizeof(R*) * (size_t)nranks, cudaMemcpyHostToDevice));
    n_peers = nranks; peer_rank = rank;
    // the exchange window sits at the same offset behind every rank's table
    std::vector<unsigned long long*> fl((size_t)nranks);
    for (int r = 0; r < nranks; r++) fl[(size_t)r] = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(t[(size_t)r]) + xwin_off);
    if (d_peer_flags) CU(cudaFree(d_peer_flags));
    CU(cudaMalloc(&d_peer_flags, sizeof(void*) * (size_t)nranks));
    CU(cudaMemcpy(d_peer_flags, fl.data(), sizeof(void*) * (size_t)nranks, cudaMemcpyHostToDevice));
    if (grid_mode) {
      auto slot = [&](int owner, int par, int writer) {
        char* base = reinterpret_cast<char*>(t[(size_t)owner]) + xwin_off + 256;
        return reinterpret_cast<R*>(base + (int64_t)sizeof(R) * grid_elems * (par * nranks + writer));
      };
      std::vector<R*> dst((size_t)(2 * nranks)), src((size_t)(2 * nranks));
      for (int par = 0; par < 2; par++)
        for (int r = 0; r < nranks; r++) {
          dst[(size_t)(par * nranks + r)] = slot(r, par, rank);   // my slot on peer r
          src[(size_t)(par * nranks + r)] = slot(rank, par, r);   // rank r's slot in my window
        }
      if (d_slot_dst) { CU(cudaFree(d_slot_dst)); CU(cudaFree(d_slot_src)); }
      CU(cudaMalloc(&d_slot_dst, sizeof(R*) * dst.size()));
      CU(cudaMalloc(&d_slot_src, sizeof(R*) * src.size()));
      CU(cudaMemcpy(d_slot_dst, dst.data(), sizeof(R*) * dst.size(), cudaMemcpyHostToDevice));
      CU(cudaMemcpy(d_slot_src, src.data(), sizeof(R*) * src.size(), cudaMemcpyHostToDevice));
    }
    if (opened_by_ipc)
      for (int r = 0; r < nranks; r++)
        if (r != rank) ipc_opened.push_back(tables[r]);
    return 0;
  }

  int set_peer_barrier(int on_device) override {
    if (on_device && !d_peer_flags) return fail(this, "nds_set_peer_barrier: open the peer tables first");
    device_barrier = on_device != 0;
    return 0;
  }

  // native NCCL communicator for the hill / grid exchange: a ready ncclComm_t of the caller, or the 128-byte
  // unique id from nds_nccl_unique_id() (broadcast by the caller's launcher), from which the engine builds its own
  int set_comm(const void* unique_id, void* existing, int rank, int nranks) override {
    CU(cudaSetDevice(cfg.device));
    NcclApi& api = nccl_api();
    if (!api.ok) return fail(this, "nds_set_comm: NCCL is not available (%s)", api.why.c_str());
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(this, "nds_set_comm: bad rank / nranks");
    if (comm && comm_owned) api.CommDestroy(comm);
    comm = nullptr; comm_owned = false;
    if (existing) {
      comm = (NcclApi::comm_t)existing;
    } else {
      if (!unique_id) return fail(this, "nds_set_comm: neither a communicator nor a unique id");
      NcclApi::unique_id id;
      memcpy(&id, unique_id, sizeof id);
      if (nccl_check(api.CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank")) return 1;
      comm_owned = true;
    }
    comm_rank = rank; comm_nranks = nranks;
    return 0;
  }

  int get_mtd_grid(double* out, int64_t cap) override {
    CU(cudaSetDevice(cfg.device));
    if (!grid_mode) return fail(this, "nds_get_mtd_grid: mtd_grid is off");
    if (cap < grid_elems) return fail(this, "nds_get_mtd_grid: capacity %lld < %lld", (long long)cap, (long long)grid_elems);
    return download(d_grid, out, grid_elems);
  }
  int set_mtd_grid(const double* in, int64_t n) override {
    CU(cudaSetDevice(cfg.device));
    if (!grid_mode) return fail(this, "nds_set_mtd_grid: mtd_grid is off");
    if (n != grid_elems) return fail(this, "nds_set_mtd_grid: %lld values, the grid has %lld", (long long)n, (long long)grid_elems);
    need_refresh_vr = false;
    return upload(in, d_grid, grid_elems);
  }
  // Stat.colv window of every walker (mtd_adapsig_hist): ring[(j % window) * M + w] = sample j; checkpoint / resume
  int get_cv_history(double* ring, int64_t cap_elems, int64_t* n_samples, int64_t* window) override {
    CU(cudaSetDevice(cfg.device));
    if (n_samples) *n_samples = cv_n;
    if (window) *window = tao_d;
    if (!ring) return 0;
    if (!d_cvhist) return fail(this, "nds_get_cv_history: the config has no history-based mtd_adapsig");
    if (cap_elems < (int64_t)tao_d * M) return fail(this, "nds_get_cv_history: buffer of %lld < %lld elements", (long long)cap_elems, (long long)tao_d * M);
    return download(d_cvhist, ring, (int64_t)tao_d * M);
  }
  int set_cv_history(const double* ring, int64_t n_samples) override {
    CU(cudaSetDevice(cfg.device));
    if (!d_cvhist) return fail(this, "nds_set_cv_history: the config has no history-based mtd_adapsig");
    if (n_samples < 0) return fail(this, "nds_set_cv_history: negative sample count");
    if (upload(ring, d_cvhist, (int64_t)tao_d * M)) return 1;
    cv_n = n_samples;
    return 0;
  }

  int64_t mtd_grid_outside() override {
    if (!grid_mode) return 0;
    cudaSetDevice(cfg.device);
    unsigned long long h = 0;
    cudaMemcpyAsync(&h, d_grid_outside, sizeof h, cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    return (int64_t)h;
  }

  // footprint of n records on the grid: partial sums per record slice, then the slices in order -> grid += (add),
  // or -> the destination list (increment slots on the peers / the all-reduce buffer)
  int accumulate(const R* rec, int64_t n, R* const* dst, int ndst, bool add) {
    const int tiles = ((K.grid_nx + GRID_TX - 1) / GRID_TX) * ((K.grid_ny + GRID_TY - 1) / GRID_TY);
    if (GRID_SLICES > 8 && !cluster_ready) {  // clusters above the portable size of 8 CTAs need the opt-in
      CU(cudaFuncSetAttribute(grid_accumulate_kernel<R>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      cluster_ready = true;
    }
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(tiles * GRID_SLICES));
    lc.blockDim = dim3(GRID_NT);
    lc.dynamicSmemBytes = 0;
    lc.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = GRID_SLICES; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    CU(cudaLaunchKernelEx(&lc, grid_accumulate_kernel<R>, K, rec, (long long)n, dst, ndst, d_grid, add ? 1 : 0));
    launches += 1;
    return 0;
  }

  // grid <- footprint of the first n records of the (local) hill table
  int rebuild_grid(int64_t n) {
    CU(cudaMemsetAsync(d_grid, 0, sizeof(R) * (size_t)grid_elems, stream));
    if (n > 0 && accumulate(d_hills, n, nullptr, 0, true)) return 1;
    return 0;
  }

  // order-independent checksum of hill records [first, first + n) of this rank's table (multi-rank exchange checks)
  int table_checksum(uint64_t* sum, int64_t first, int64_t n) override {
    CU(cudaSetDevice(cfg.device));
    if (!d_hills || !cfg.mtd_shared) return fail(this, "nds_table_checksum: no shared hill table");
    if (first < 0 || n < 0 || first + n > n_hills) return fail(this, "nds_table_checksum: records [%lld, %lld) of %lld", (long long)first, (long long)(first + n), (long long)n_hills);
    CU(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), stream));
    const long long words = n * K.hill_stride * (long long)(sizeof(R) / 4);
    if (words > 0) {
      checksum_kernel<<<296, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(d_hills + first * K.hill_stride), words, d_counter);
      CU(cudaGetLastError());
      launches++;
    }
    unsigned long long h = 0;
    CU(cudaMemcpyAsync(&h, d_counter, sizeof h, cudaMemcpyDeviceToHost, stream));
    CU(cudaStreamSynchronize(stream));
    *sum = h;
    return 0;
  }

  int device_pointers(void** x, void** v, void** hills, int64_t* stride) override {
    if (x) *x = d_x;
    if (v) *v = d_v;
    if (hills) *hills = d_hills;
    if (stride) *stride = K.hill_stride;
    return 0;
  }
};

// ------------------------------------------------------------------------------------ C ABI
extern "C" {

const char* nds_last_error(const nds_engine* e) { return e ? e->err.c_str() : g_last_error.c_str(); }

void nds_config_defaults(nds_config* c) {
  memset(c, 0, sizeof *c);
  c->abi_version = NDS_ABI_VERSION;
  c->precision = NDS_F64;
  c->ndim = 2;
  c->method = NDS_METHOD_MD;
  c->integrator = NDS_INT_LANGEVIN;      // md.py:52
  c->potential = NDS_POT_GAUSSIAN;       // potentials/_build.py:10
  c->colvar = NDS_CV_ORIGINAL;           // colvars/_build.py:10
  c->true_colvar = NDS_CV_ORIGINAL;
  c->noise_mode = NDS_NOISE_PHILOX;
  c->steps_per_launch = 1000;
  c->n_walkers = 1; c->n_walkers_global = 1;
  c->temperature = 300.0;                // ndrun.py:80
  c->initial_temperature = 300.0;
  c->mass = 5;                           // ndrun.py:81
  c->dt = 0.5; c->gamma = 0.005;         // md.py:50-51
  c->rescale_freq = -1.0;
  const double xc[4] = {48, 32, 24, 16}, yc[4] = {8, 16, 32, 24};  // mueller.py:15-16
  for (int i = 0; i < 4; i++) { c->pot_params[i] = xc[i]; c->pot_params[4 + i] = yc[i]; }
  c->mtd_dep_freq = 10; c->mtd_biasf = 1.0;  // gaus.py:52-53
  c->mtd_shared = 1;
  c->stat_freq = 1;
}

int nds_create(const nds_config* cfg, nds_engine** out) {
  if (!out) return fail(nullptr, "nds_create: out is null");
  *out = nullptr;
  if (validate(cfg)) return 1;
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0)
    return fail(nullptr, "nds_create: no CUDA device (%s); this engine has no CPU fallback",
                err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, "nds_create: device %d out of range", cfg->device);
  int rc;
  nds_engine* e;
  if (cfg->precision == NDS_F32) { auto* p = new Engine<float>(); rc = p->init(*cfg); e = p; }
  else { auto* p = new Engine<double>(); rc = p->init(*cfg); e = p; }
  if (rc) { g_last_error = e->err; delete e; return rc; }
  *out = e;
  return 0;
}

void nds_destroy(nds_engine* e) { delete e; }
int nds_set_state(nds_engine* e, const double* x, const double* v) { return e->set_state(x, v); }
int nds_set_walker_bias(nds_engine* e, const double* k, const double* r0) { return e->set_walker_bias(k, r0); }
int nds_set_noise(nds_engine* e, const double* n, int64_t first, int64_t ns) { return e->set_noise(n, first, ns); }
int nds_set_path(nds_engine* e, int64_t n_ref, int32_t path_dim, const double* ref, const double* ind, double sigma) {
  return e->set_path(n_ref, path_dim, ref, ind, sigma);
}
int nds_begin(nds_engine* e) { return e->begin(); }
int nds_run(nds_engine* e, int64_t n) { return e->run(n); }
int nds_sync(nds_engine* e) { return e->sync(); }
int nds_get_state(nds_engine* e, double* x, double* v, double* f, double* fixf, double* colv, double* thermo) {
  return e->get_state(x, v, f, fixf, colv, thermo);
}
int nds_set_state_f32(nds_engine* e, const float* x, const float* v) { return e->set_state_f32(x, v); }
int nds_get_state_f32(nds_engine* e, float* x, float* v, float* f, float* fixf, float* colv, float* thermo) {
  return e->get_state_f32(x, v, f, fixf, colv, thermo);
}
int64_t nds_step(const nds_engine* e) { return e->step; }
double nds_time(const nds_engine* e) {
  const_cast<nds_engine*>(e)->mc_sync_walker(0);  // MC: NDRun.time of walker 0 = accepted moves * dt
  return e->time;
}
int nds_eval_at(nds_engine* e, const double* x, int64_t n, double* V, double* F, double* colv, double* J,
                double* Vb, double* Fb) {
  return e->eval_at(x, n, V, F, colv, J, Vb, Fb);
}
int nds_get_hills(nds_engine* e, int64_t walker, int64_t* n, double* c, double* h, int64_t cap) {
  return e->get_hills(walker, n, c, h, cap);
}
int nds_set_hills(nds_engine* e, int64_t walker, int64_t n, const double* c, const double* h) {
  return e->set_hills(walker, n, c, h);
}
int nds_get_hill_sigma2(nds_engine* e, int64_t walker, double* sigma2, int64_t cap) {
  return e->get_hill_sigma2(walker, sigma2, cap);
}
int nds_set_hill_sigma2(nds_engine* e, int64_t walker, int64_t n, const double* sigma2) {
  return e->set_hill_sigma2(walker, n, sigma2);
}
int64_t nds_hill_table_size(const nds_engine* e, int64_t n) { return e->hill_table_size(n); }
int nds_get_hill_table_raw(nds_engine* e, double* out, int64_t cap) { return e->get_hill_table_raw(out, cap); }
int nds_set_hill_table_raw(nds_engine* e, int64_t n, const double* in) { return e->set_hill_table_raw(n, in); }
int nds_get_vr(nds_engine* e, double* vr) { return e->get_vr(vr); }
int nds_set_exchange(nds_engine* e, nds_exchange_fn cb, void* user) {
  e->exchange = cb; e->exchange_user = user;
  return 0;
}
int nds_get_commit(nds_engine* e, int32_t* basin, int64_t* exit_step) { return e->get_commit(basin, exit_step); }
int nds_get_alive(nds_engine* e, uint8_t* alive) { return e->get_alive(alive); }
int nds_set_commit(nds_engine* e, const int32_t* basin, const int64_t* exit_step, const uint8_t* alive) {
  return e->set_commit(basin, exit_step, alive);
}
int64_t nds_count_alive(nds_engine* e) { return e->count_alive(); }
int nds_get_mc(nds_engine* e, int64_t* accepted) { return e->get_mc(accepted); }
int nds_get_wl_histogram(nds_engine* e, int64_t walker, double* hE) { return e->get_wl_histogram(walker, hE); }
int nds_set_wl_histogram(nds_engine* e, int64_t walker, const double* hE) { return e->set_wl_histogram(walker, hE); }
int nds_get_clock(const nds_engine* e, int64_t* step, double* time, int64_t* dep, int64_t* resc) {
  return e->get_clock(step, time, dep, resc);
}
int nds_set_clock(nds_engine* e, int64_t step, double time, int64_t dep, int64_t resc) {
  return e->set_clock(step, time, dep, resc);
}
int nds_set_vr(nds_engine* e, const double* vr) { return e->set_vr(vr); }
int64_t nds_count_nonfinite(nds_engine* e) { return e->count_nonfinite(); }
int nds_frame_width(const nds_engine* e) { return e->frame_width; }
int64_t nds_frames_pending(const nds_engine* e) { return e->frames_written; }
int nds_drain_frames(nds_engine* e, double* out, int64_t max_frames, int64_t* n) { return e->drain_frames(out, max_frames, n); }
int nds_fes_grid(nds_engine* e, int64_t walker, double x0, double dx, int64_t nx, double y0, double dy,
                 int64_t ny, double* grid) {
  return e->fes_grid(walker, x0, dx, nx, y0, dy, ny, grid);
}
int nds_colvar_histogram(nds_engine* e, double x0, double dx, int64_t nx, double y0, double dy, int64_t ny,
                         int64_t* counts) {
  return e->histogram(x0, dx, nx, y0, dy, ny, counts);
}
int nds_device_pointers(nds_engine* e, void** x, void** v, void** hills, int64_t* stride) {
  return e->device_pointers(x, v, hills, stride);
}
int nds_hills_ipc_handle(nds_engine* e, void* handle) { return e->hills_ipc_handle(handle); }
int nds_set_peer_tables(nds_engine* e, void* const* tables, int nranks, int rank) {
  return e->set_peer_tables(tables, nranks, rank, false);
}
int nds_open_peer_tables(nds_engine* e, const void* handles, int nranks, int rank) {
  if (cudaSetDevice(e->cfg.device) != cudaSuccess) return fail(e, "nds_open_peer_tables: cudaSetDevice failed");
  std::vector<void*> ptrs((size_t)nranks, nullptr);
  for (int r = 0; r < nranks; r++) {
    if (r == rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * sizeof h, sizeof h);
    cudaError_t err = cudaIpcOpenMemHandle(&ptrs[(size_t)r], h, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) return fail(e, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(err));
  }
  return e->set_peer_tables(ptrs.data(), nranks, rank, true);
}
int nds_set_peer_barrier(nds_engine* e, int on_device) { return e->set_peer_barrier(on_device); }
int nds_nccl_unique_id(void* id128) {
  NcclApi& api = nccl_api();
  if (!api.ok) return fail(nullptr, "nds_nccl_unique_id: NCCL is not available (%s)", api.why.c_str());
  NcclApi::unique_id id;
  const int rc = api.GetUniqueId(&id);
  if (rc) return fail(nullptr, "ncclGetUniqueId failed: %s", api.GetErrorString(rc));
  memcpy(id128, &id, sizeof id);
  return 0;
}
int nds_set_comm(nds_engine* e, const void* unique_id, void* nccl_comm, int rank, int nranks) {
  return e->set_comm(unique_id, nccl_comm, rank, nranks);
}
int nds_exchange_info(nds_engine* e, double* exchange_ms, int64_t* n_exchanges) { return e->exchange_info(exchange_ms, n_exchanges); }
int nds_table_checksum(nds_engine* e, int64_t first, int64_t n, uint64_t* sum) { return e->table_checksum(sum, first, n); }
int nds_get_mtd_grid(nds_engine* e, double* out, int64_t cap) { return e->get_mtd_grid(out, cap); }
int nds_set_mtd_grid(nds_engine* e, const double* in, int64_t n) { return e->set_mtd_grid(in, n); }
int64_t nds_mtd_grid_outside(nds_engine* e) { return e->mtd_grid_outside(); }
int nds_get_mc_state(nds_engine* e, int64_t* accepted, double* time, int64_t* dep, double* wl, int64_t* max_dep) {
  return e->get_mc_state(accepted, time, dep, wl, max_dep);
}
int nds_set_mc_state(nds_engine* e, const int64_t* accepted, const double* time, const int64_t* dep, const double* wl) {
  return e->set_mc_state(accepted, time, dep, wl);
}
int nds_set_thermo(nds_engine* e, const double* thermo) { return e->set_thermo(thermo); }
int nds_get_umb_vr(nds_engine* e, double* vr, int64_t cap, int64_t* n) { return e->get_umb_vr(vr, cap, n); }
int nds_set_umb_vr(nds_engine* e, const double* vr, int64_t n) { return e->set_umb_vr(vr, n); }
int nds_get_cv_history(nds_engine* e, double* ring, int64_t cap_elems, int64_t* n_samples, int64_t* window) {
  return e->get_cv_history(ring, cap_elems, n_samples, window);
}
int nds_set_cv_history(nds_engine* e, const double* ring, int64_t n_samples) { return e->set_cv_history(ring, n_samples); }
int nds_get_histogram(nds_engine* e, int64_t* counts, int clear) { return e->get_histogram(counts, clear); }
int nds_frames_f32(const nds_engine* e) { return e->frames_f32(); }
int nds_submit(nds_engine* e, const void* host_xv, int64_t nsteps, void* host_xvt, int32_t chunks) {
  return e->submit(host_xv, nsteps, host_xvt, chunks);
}
int nds_wait(nds_engine* e) { return e->sync(); }
void* nds_stream(nds_engine* e) { return (void*)e->stream; }
int64_t nds_n_hills(const nds_engine* e, int64_t walker) {
  const_cast<nds_engine*>(e)->mc_sync_walker(walker);  // MC: every walker has its own count
  return e->n_hills;
}
int64_t nds_launch_count(const nds_engine* e) { return e->launches; }
double nds_last_run_ms(nds_engine* e) { return e->last_run_ms(); }
const char* nds_kernel_name(const nds_engine* e) { return e->kernel_name; }

int nds_philox_kat(int device, const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, "nds_philox_kat: no CUDA device");
  uint32_t* d = nullptr;
  if (cudaMalloc(&d, 16) != cudaSuccess) return fail(nullptr, "nds_philox_kat: cudaMalloc failed");
  philox_kat_kernel<<<1, 1>>>(u32x4{ctr[0], ctr[1], ctr[2], ctr[3]}, key[0], key[1], d);
  cudaError_t err = cudaMemcpy(out, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return err == cudaSuccess ? 0 : fail(nullptr, "nds_philox_kat: %s", cudaGetErrorString(err));
}

int nds_noise_tail_counts(int device, int precision, uint64_t seed, int64_t n_walkers, int64_t calls_per_walker,
                          const double* thresholds, int n_thresholds, uint64_t* counts, double* stats4) {
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, "nds_noise_tail_counts: no CUDA device");
  if (n_thresholds < 0 || n_thresholds > NDS_MAX_THR) return fail(nullptr, "nds_noise_tail_counts: at most %d thresholds", NDS_MAX_THR);
  if (n_walkers <= 0 || calls_per_walker <= 0) return fail(nullptr, "nds_noise_tail_counts: empty request");
  TailArgs T;
  for (int t = 0; t < NDS_MAX_THR; t++) T.thr[t] = t < n_thresholds ? thresholds[t] : 1e30;
  T.n_thr = n_thresholds;
  unsigned long long* d_counts = nullptr;
  double* d_stats = nullptr;
  if (cudaMalloc(&d_counts, sizeof(unsigned long long) * NDS_MAX_THR) != cudaSuccess || cudaMalloc(&d_stats, sizeof(double) * 4) != cudaSuccess)
    return fail(nullptr, "nds_noise_tail_counts: cudaMalloc failed");
  cudaMemset(d_counts, 0, sizeof(unsigned long long) * NDS_MAX_THR);
  cudaMemset(d_stats, 0, sizeof(double) * 4);
  // the per-step stream of a run with this seed (domain 0), see nds_philox.cuh
  const PhiloxKeys rk = philox_round_keys((uint32_t)seed, (uint32_t)(seed >> 32));
  const unsigned grid = (unsigned)((n_walkers + 127) / 128);
  if (precision == NDS_F64) noise_tail_kernel<double><<<grid, 128>>>(rk, n_walkers, calls_per_walker, T, d_counts, d_stats);
  else noise_tail_kernel<float><<<grid, 128>>>(rk, n_walkers, calls_per_walker, T, d_counts, d_stats);
  unsigned long long h[NDS_MAX_THR];
  cudaError_t err = cudaMemcpy(h, d_counts, sizeof h, cudaMemcpyDeviceToHost);
  if (err == cudaSuccess && stats4) err = cudaMemcpy(stats4, d_stats, sizeof(double) * 4, cudaMemcpyDeviceToHost);
  cudaFree(d_counts); cudaFree(d_stats);
  if (err != cudaSuccess) return fail(nullptr, "nds_noise_tail_counts: %s", cudaGetErrorString(err));
  for (int t = 0; t < n_thresholds; t++) counts[t] = h[t];
  return 0;
}

int nds_map50to2d(int device, const double* x, int64_t n, double* c, double* jac) {
  if (n < 0) return fail(nullptr, "nds_map50to2d: negative count");
  if (n == 0) return 0;
  if (!x) return fail(nullptr, "nds_map50to2d: x is NULL");
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, "nds_map50to2d: no CUDA device");
  double *dx = nullptr, *dc = nullptr, *dj = nullptr;
  const size_t row = sizeof(double) * (size_t)n;
  cudaError_t err = cudaMalloc(&dx, row * MAP50_ND);
  if (err == cudaSuccess && c) err = cudaMalloc(&dc, row * 2);
  if (err == cudaSuccess && jac) err = cudaMalloc(&dj, row * 2 * MAP50_ND);
  if (err == cudaSuccess) err = cudaMemcpy(dx, x, row * MAP50_ND, cudaMemcpyHostToDevice);
  if (err == cudaSuccess) {
    static const Map50Table T = map50_table();
    map50to2d_kernel<<<(unsigned)((n + 127) / 128), 128>>>(T, dx, (long long)n, dc, dj);
    err = cudaGetLastError();
  }
  if (err == cudaSuccess && c) err = cudaMemcpy(c, dc, row * 2, cudaMemcpyDeviceToHost);
  if (err == cudaSuccess && jac) err = cudaMemcpy(jac, dj, row * 2 * MAP50_ND, cudaMemcpyDeviceToHost);
  cudaFree(dx); cudaFree(dc); cudaFree(dj);
  return err == cudaSuccess ? 0 : fail(nullptr, "nds_map50to2d: %s", cudaGetErrorString(err));
}

}  // extern "C"
