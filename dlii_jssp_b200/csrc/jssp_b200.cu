// jssp_b200.cu - C ABI (include/jssp_b200.h) over the sm_100a kernels in jssp_kernels.cuh.
//
// Build: nvcc -std=c++17 -O3 --fmad=false -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC
// There is no CPU fallback anywhere in this file: without a CUDA device jssp_create returns JSSP_ERR_NO_DEVICE.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <new>
#include <vector>
#include <algorithm>
#include <limits>
#include <dlfcn.h>

#include "jssp_kernels.cuh"

using namespace jssp;

// tools/ A/B overrides (JSSP_B200_TUNE_* environment variables), read ONCE when a handle is created - never on the per-cycle path
struct Tuning {
  bool no_pdl = false;        // plain stream order instead of programmatic dependent launches
  bool full_walk = false;     // batched replay: walk every candidate to the end instead of stopping at its first veto
  int seed_cluster = 0;       // CTAs per seed-enumeration cluster (0 = by batch size)
  int seed_threads = 0;       // threads per seed-enumeration CTA (0 = by batch size)
  int chunks = 0;             // lane groups per seed of the chunked tiled kernel (0 = 2)
  int lockstep = 0;           // batched replay: 1 = force the lock-step kernels instead of the persistent scenario kernel
  static Tuning from_env() {
    Tuning t;
    t.no_pdl = getenv("JSSP_B200_TUNE_NO_PDL") != nullptr;
    t.full_walk = getenv("JSSP_B200_TUNE_FULL_WALK") != nullptr;
    if (const char* e = getenv("JSSP_B200_TUNE_SEED_CLUSTER")) t.seed_cluster = atoi(e);
    if (const char* e = getenv("JSSP_B200_TUNE_SEED_THREADS")) t.seed_threads = atoi(e);
    if (const char* e = getenv("JSSP_B200_TUNE_CHUNKS")) t.chunks = atoi(e);
    if (const char* e = getenv("JSSP_B200_TUNE_LOCKSTEP")) t.lockstep = atoi(e);
    return t;
  }
};

struct jssp_handle {
  jssp_config cfg;
  Tuning tune;
  int device = 0;
  int P = 0;   // samples per trajectory = n_steps + 1
  // sampling grids (device) + lengths
  double *d_time = nullptr, *d_j1 = nullptr, *d_t1 = nullptr, *d_t2 = nullptr, *d_j3 = nullptr, *d_t3 = nullptr, *d_t4 = nullptr,
         *d_jneg = nullptr, *d_jpos = nullptr, *d_ts = nullptr;
  SeedGrids grids{};
  uint32_t main_total = 0, stop_total = 0;
  // seeds
  double* d_seeds = nullptr;
  int* d_counts = nullptr;
  SeedWork seed_work{};
  size_t seed_smem = 0;
  int* h_counts = nullptr;   // pinned
  bool seeds_ready = false;
  // reference line
  double *d_knot = nullptr, *d_coef = nullptr;
  int K = 0, kpad = 0;
  double ox = 0, oy = 0, s_end = 0;
  // obstacles
  unsigned char *d_raw = nullptr, *h_raw = nullptr;   // one device block + pinned staging block: [state | size | prob | valid]
  size_t raw_bytes = 0;
  cudaEvent_t raw_done = nullptr;                     // the last upload out of h_raw has completed
  // host-buffer obstacle uploads run on a private stream, next to whatever the caller's stream is doing (e.g. the H2D copy of
  // the cycle's seeds): tables_ready = upload + pack finished, tables_free = the last evaluation has finished reading them
  cudaStream_t up_stream = nullptr;
  cudaEvent_t tables_ready = nullptr, tables_free = nullptr;
  bool wait_tables = false, tables_in_use = false;
  const double *src_state = nullptr, *src_size = nullptr, *src_prob = nullptr;   // what the pack kernel reads
  const uint8_t* src_valid = nullptr;
  float4 *d_broad = nullptr, *d_narrow = nullptr;
  double *d_obs64 = nullptr, *d_ext64 = nullptr;
  float2* d_sint = nullptr;       // reachable arc-length interval per obstacle state
  float* d_probf = nullptr;
  double* d_knot_xy = nullptr;
  double* d_lane_scratch = nullptr;   // jssp_predict_modes: per-lane arc-length / heading tables
  size_t lane_scratch_cap = 0;        // doubles
  double cull_reach = 0, cull_ds = 0;
  bool cull_ok = false;
  int O = 0, M = 0, Tt = 0, n_dict = 0;
  bool obs_set = false, obs_dirty = false;
  // selection
  double* d_blk_cost = nullptr;
  int* d_blk_idx = nullptr;
  int max_blocks = 0;
  unsigned int* d_ticket = nullptr;
  BestRecord* d_best = nullptr;
  BestRecord* h_best = nullptr;   // pinned
  double* d_best_traj = nullptr;
  double* h_best_traj = nullptr;  // pinned
  unsigned char *d_result = nullptr, *h_result = nullptr;   // contiguous [best | counts | trajectory]
  size_t result_bytes = 0;
  EvalParams last{};
  bool last_valid = false;
  int last_flags = 0;
  int64_t launches = 0;
  int max_smem_optin = 0;
  int sm_count = 148;
  double* d_fop_tab = nullptr;              // FOP grids: horizons | speeds, then n_samples (int) in d_fop_int
  int* d_fop_int = nullptr;
  int fop_cap_t = 0, fop_cap_v = 0;
  unsigned long long* phase_ts = nullptr;   // profiling builds only
  // candidate split over ranks (jssp_set_split / jssp_allreduce_argmin / jssp_peer_*)
  long long split_seed_lo = 0, split_ns_global = 0;
  SplitRecord* d_split_rec = nullptr;       // this rank's record
  SplitRecord* d_split_table = nullptr;     // [SPLIT_MAX_WORLD] all-gather target (NCCL transport)
  GlobalBest *h_gbest = nullptr, *d_gbest = nullptr;   // mapped pinned host memory
  PeerSlot* d_mailbox = nullptr;            // [2][world] own mailbox (peer transport)
  PeerView* d_peers = nullptr;              // device copy of the peer table
  void* peer_ptr[SPLIT_MAX_WORLD] = {nullptr};   // mappings opened with cudaIpcOpenMemHandle
  int peer_world = 0, peer_rank = 0;
  bool peers_connected = false;
  unsigned long long split_epoch = 0;
  char cuda_err[256] = {0};
};

namespace {

int fail_cuda(jssp_handle* h, cudaError_t e, const char* what) {
  if (h) snprintf(h->cuda_err, sizeof(h->cuda_err), "%s: %s", what, cudaGetErrorString(e));
  return JSSP_ERR_CUDA;
}
#define CK(call)                                                    \
  do {                                                              \
    cudaError_t e__ = (call);                                       \
    if (e__ != cudaSuccess) return fail_cuda(h, e__, #call);        \
  } while (0)

// numpy.linspace(start, stop, num) (endpoint=True), same operations as numpy/_core/function_base.py
std::vector<double> np_linspace(double start, double stop, int num) {
  std::vector<double> y((size_t)(num > 0 ? num : 0));
  if (num <= 0) return y;
  const int div = num - 1;
  const double delta = stop - start;
  if (div > 0) {
    const double step = delta / div;
    for (int i = 0; i < num; ++i) y[i] = (step == 0.0) ? ((double)i / div) * delta + start : (double)i * step + start;
    y[num - 1] = stop;
  } else {
    y[0] = start;
  }
  return y;
}
// numpy.arange(start, stop, step) for floats: len = ceil((stop - start) / step), value_i = start + i * step
std::vector<double> np_arange(double start, double stop, double step) {
  const long n = (long)std::ceil((stop - start) / step);
  std::vector<double> y((size_t)(n > 0 ? n : 0));
  for (long i = 0; i < n; ++i) y[i] = start + (double)i * step;
  return y;
}

struct HostGrids {
  std::vector<double> time, j1, t1, t2, j3, t3, t4, js, ts;
};
HostGrids make_grids(double total, double jerk, int ns, double delta_t = 0.1) {
  HostGrids g;
  g.time = np_linspace(0.0, total, (int)(total / delta_t) + 1);      // polyvt_sampling.py:44
  g.j1 = np_linspace(-jerk, jerk, ns);                                // :45, :181
  g.t1 = np_arange(0.0, total + 1e-9, 0.5);                           // :50
  g.t2 = np_arange(0.0, total + 1e-9, 0.8);                           // :51 (literal 5.0 prepended)
  g.t2.insert(g.t2.begin(), 5.0);
  g.t3 = np_arange(0.0, total + 1e-9, 1.0);                           // :52
  g.t4 = np_arange(0.0, total + 1e-9, 1.5);                           // :53
  g.j3 = np_linspace(-jerk * 0.6, jerk * 0.6, ns / 4);                // :55-56, :207  int(num_samples / 4)
  g.js = np_linspace(-jerk, jerk, ns * 2);                            // :274
  g.ts = np_arange(0.0, total + 1e-9, 0.1);                           // :278
  return g;
}

template <class T>
cudaError_t upload(T** dst, const std::vector<T>& v) {
  cudaError_t e = cudaMalloc((void**)dst, (v.size() ? v.size() : 1) * sizeof(T));
  if (e != cudaSuccess) return e;
  if (!v.empty()) e = cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return e;
}

// sampling grids -> device arrays (dst order: time, j1, t1, t2, j3, t3, t4, jneg, jpos, ts) and the kernel-side view
cudaError_t upload_grids(const jssp_config* cfg, double** dst[10], SeedGrids* sgp, int* n_time) {
  HostGrids g = make_grids(cfg->plan_horizon, cfg->highest_jerk, cfg->num_jerk, cfg->delta_t);
  std::vector<double> jneg, jpos;
  for (double j : g.js) { if (!(j > 0)) jneg.push_back(j); if (!(j < 0)) jpos.push_back(j); }   // polyvt_sampling.py:275,287
  std::vector<double> time_padded = g.time;           // bulk copies move an even number of doubles
  time_padded.resize((g.time.size() + 2) & ~(size_t)1, 0.0);
  const std::vector<double>* v[10] = {&time_padded, &g.j1, &g.t1, &g.t2, &g.j3, &g.t3, &g.t4, &jneg, &jpos, &g.ts};
  for (int k = 0; k < 10; ++k) {
    cudaError_t e = upload(dst[k], *v[k]);
    if (e != cudaSuccess) return e;
  }
  SeedGrids& sg = *sgp;
  sg.j1 = *dst[1]; sg.t1 = *dst[2]; sg.t2 = *dst[3]; sg.j3 = *dst[4]; sg.t3 = *dst[5]; sg.t4 = *dst[6];
  sg.n_j1 = (int)g.j1.size(); sg.n_t1 = (int)g.t1.size(); sg.n_t2 = (int)g.t2.size(); sg.n_j3 = (int)g.j3.size();
  sg.n_t3 = (int)g.t3.size(); sg.n_t4 = (int)g.t4.size();
  sg.jneg = *dst[7]; sg.jpos = *dst[8]; sg.ts = *dst[9];
  sg.n_jneg = (int)jneg.size(); sg.n_jpos = (int)jpos.size(); sg.n_ts = (int)g.ts.size();
  sg.jpos_uniform = jpos.size() >= 2 ? 1 : 0;
  sg.jpos0 = jpos.empty() ? 0.f : (float)jpos[0];
  sg.jpos_inv_step = 0.f;
  if (sg.jpos_uniform) {
    const double step = jpos[1] - jpos[0];
    for (size_t k = 1; k + 1 < jpos.size(); ++k)
      if (!(step > 0) || std::fabs((jpos[k + 1] - jpos[k]) - step) > 1e-9 * std::fabs(step)) sg.jpos_uniform = 0;
    if (sg.jpos_uniform) sg.jpos_inv_step = (float)(1.0 / step);
  }
  sg.total_time = cfg->plan_horizon;
  sg.acc_max = cfg->lon_accel_max - 1e-9; sg.vel_max = cfg->speed_max - 1e-9; sg.v_max = cfg->speed_max;
  sg.jmin5 = -cfg->highest_jerk * 0.3; sg.jmax5 = cfg->highest_jerk * 0.3;
  sg.lim2 = cfg->plan_horizon - 0.8 * 3; sg.lim3 = cfg->plan_horizon - 1.0 * 2; sg.lim4 = cfg->plan_horizon - 1.5;
  *n_time = (int)g.time.size();
  return cudaSuccess;
}

// planner constants of a plan cycle (everything of EvalParams that does not depend on the scenario or the cycle)
void fill_constants(EvalParams& p, const jssp_config& c, int P) {
  p.P = P; p.check_res = c.check_res;
  p.kmax = c.max_curvature; p.ea = c.ego_length / 2.0; p.eb = c.ego_width / 2.0; p.p_min = c.p_min; p.w_risk = c.w_risk;
  p.vmax = c.speed_max; p.inv_vmax = 1.0 / c.speed_max; p.max_distance = c.max_distance; p.half_w = c.max_road_width / 2.0;
  p.thr_lon_a = c.thr_lon_accel; p.thr_lon_j = c.thr_lon_jerk; p.thr_lat_a = c.thr_lat_accel; p.thr_lat_j = c.thr_lat_jerk;
  memcpy(p.cw, c.cost_weights, sizeof(p.cw));
}

// lateral targets: linspace(-(max_road_width - w)/2, +(max_road_width - w)/2, num_width) (jerk_space_sampling_planner.py:100-104)
std::vector<double> default_targets(const jssp_config& c, int W) {
  if (W <= 1) return std::vector<double>(1, 0.0);
  const double sw = c.max_road_width - c.ego_width;
  return np_linspace(-sw / 2, sw / 2, W);
}

bool config_ok(const jssp_config* cfg) {
  // int(T / dt) (time grid, polyvt_sampling.py:44) and round(T / dt) (lateral grid, jerk_space_sampling_planner.py:78) must agree
  if (cfg->delta_t > 0 && (long)(cfg->plan_horizon / cfg->delta_t) != std::lround(cfg->plan_horizon / cfg->delta_t)) return false;
  return !(cfg->plan_horizon <= 0 || cfg->delta_t <= 0 || cfg->num_jerk < 4 || cfg->num_width < 1 || cfg->num_width > JSSP_MAX_WIDTHS ||
           cfg->check_res < 1 || cfg->max_seeds < 1 || cfg->max_knots < 2 || cfg->max_obstacle_modes < 0 || cfg->max_total_steps < 1);
}

int pack_if_dirty(jssp_handle* h, cudaStream_t st) {
  if (!h->obs_dirty) return JSSP_OK;
  const int OM = h->O * h->M;
  const int OMpad = (OM + 31) & ~31;
  const int total = OMpad * h->Tt;
  if (total > 0) {
    const double r_ego = std::sqrt(h->cfg.ego_length * h->cfg.ego_length + h->cfg.ego_width * h->cfg.ego_width) / 2.0;
    pack_obstacles_kernel<<<(total * PACK_LANES + 255) / 256, 256, 0, st>>>(h->src_state, h->src_valid, h->src_size, h->src_prob, OM, OMpad, h->M, h->Tt,
                                                               h->ox, h->oy, h->cfg.obstacle_inflation, r_ego, h->d_broad, h->d_narrow,
                                                               h->d_obs64, h->d_ext64, h->d_probf, h->d_sint, h->d_knot, h->d_knot_xy, h->K,
                                                               h->cull_reach, h->cull_ds);
    h->launches++;
    CK(cudaGetLastError());
  }
  h->obs_dirty = false;
  return JSSP_OK;
}

template <typename T>
static int32_t measure_fma(int device, double* tflops) {
  jssp_handle* h = nullptr;
  if (cudaSetDevice(device) != cudaSuccess) return JSSP_ERR_NO_DEVICE;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int blocks = sms * 8, threads = 256, iters = 8192;
  T* out = nullptr;
  CK(cudaMalloc(&out, (size_t)blocks * threads * sizeof(T)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(e0, 0));
    fma_peak_kernel<T><<<blocks, threads>>>(out, iters, (T)1.0000001, (T)1e-7);
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 8.0 * (double)iters * blocks * threads;
    if (rep > 0 && ms > 0.f) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  *tflops = best;
  return JSSP_OK;
}


// opt a kernel in to the largest dynamic shared memory the device allows next to the kernel's own static shared memory;
// *budget is lowered to the smallest such limit seen (the host sizes launches against it)
template <class F>
cudaError_t allow_max_dynamic_smem(F* func, int optin, int* budget) {
  cudaFuncAttributes a;
  cudaError_t e = cudaFuncGetAttributes(&a, func);
  if (e != cudaSuccess) return e;
  const int dyn = optin - (int)a.sharedSizeBytes - 256;
  if (dyn < *budget) *budget = dyn;
  return cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
}

// launch of the seed enumeration: n_scenarios clusters of `csize` CTAs
template <bool kBatch>
cudaError_t launch_seed_enum(const SeedGrids& g, SeedState ss, const EvalParams* slots, const SeedWork& wk, int n_scenarios, int csize,
                             int threads, size_t smem, cudaStream_t st, bool pdl = false) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_scenarios * csize), 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, seed_enum_kernel<kBatch>, g, ss, slots, wk);
}

// launch of a lock-step kernel of the batched replay: chained to its predecessor by programmatic dependent launch (the
// kernels call pdl_wait() before they touch anything the predecessor writes), so that block scheduling and the launch
// latency overlap the predecessor's tail
template <class... KArgs, class... Args>
cudaError_t launch_chained(void (*kern)(KArgs...), dim3 grid, int threads, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

// shared-memory list capacities of the seed enumeration kernel for the handle's sampling grids
SeedWork make_seed_work(const SeedGrids& g, int* counts, double* seeds, int cap_seeds) {
  SeedWork w{};
  w.counts = counts; w.seeds = seeds; w.cap_seeds = cap_seeds;
  w.cap_a_main = g.n_j1 * g.n_t1;                                     // worst case: every (j1, t1) survives
  w.cap_a_stop = g.n_jneg * g.n_ts;
  w.cap_b = std::min(g.n_j1 * g.n_t1 * g.n_t2, 512);                  // (j1, t1, t2) prefixes: ~150 for the default grid
  int keys = 64;
  while (keys < cap_seeds && keys < 4096) keys <<= 1;                  // the sort works on a power of two; the default grid yields < 1000 seeds
  w.cap_keys = keys;
  return w;
}

}  // namespace

extern "C" {

#ifndef JSSP_BUILD_ID_STR
#define JSSP_BUILD_ID_STR "0000000000000000"
#endif
// hash of the sources and flags this library was compiled from (dlii_jssp_b200/build.py); the loader refuses a mismatch
static const char k_build_id[] = "JSSP_BUILD_ID=" JSSP_BUILD_ID_STR;
const char* jssp_build_id(void) { return k_build_id + 14; }
int32_t jssp_abi_version(void) { return JSSP_ABI_VERSION; }
int32_t jssp_struct_size(int32_t which) {
  const size_t sizes[] = {sizeof(jssp_config), sizeof(jssp_cycle_in), sizeof(jssp_cycle_out), sizeof(jssp_fop_in), sizeof(jssp_imm_params),
                          sizeof(jssp_cycle_record), sizeof(jssp_scenario_result), sizeof(jssp_scenario_in)};
  return which >= 0 && which < (int)(sizeof(sizes) / sizeof(sizes[0])) ? (int32_t)sizes[which] : -1;
}

const char* jssp_strerror(int32_t s) {
  switch (s) {
    case JSSP_OK: return "ok";
    case JSSP_ERR_INVALID_ARG: return "invalid argument";
    case JSSP_ERR_CUDA: return "CUDA error (see jssp_last_cuda_error)";
    case JSSP_ERR_CAPACITY: return "capacity of the handle exceeded (seeds / knots / obstacles / steps)";
    case JSSP_ERR_NOT_READY: return "reference line, obstacles or seeds have not been set";
    case JSSP_ERR_SEED_OVERFLOW: return "seed enumeration produced more seeds than max_seeds";
    case JSSP_ERR_NO_DEVICE: return "no CUDA device: this library has no CPU fallback";
    default: return "unknown status";
  }
}

const char* jssp_last_cuda_error(const jssp_handle* h) { return h ? h->cuda_err : ""; }

void jssp_default_config(jssp_config* c) {
  if (!c) return;
  memset(c, 0, sizeof(*c));
  c->delta_t = 0.1; c->plan_horizon = 5.0; c->max_road_width = 3.5;
  c->num_width = 3; c->num_jerk = 25; c->highest_jerk = 10.0;
  c->speed_max = 18.0; c->lon_accel_max = 8.0;
  c->ego_length = 4.924; c->ego_width = 1.872;   // ReplayInfo default ego (controller.py:78)
  c->max_curvature = 0.2; c->obstacle_inflation = 0.1; c->check_res = 2;
  const double w[8] = {0.0, 3.0, 5.4, 0.2, 0.1, 0.8, 0.4, 4.0};
  memcpy(c->cost_weights, w, sizeof(w));
  c->max_distance = 140.0;
  c->thr_lon_accel = 3.0; c->thr_lon_jerk = 6.0; c->thr_lat_accel = 0.5; c->thr_lat_jerk = 1.0;
  c->p_min = 0.0; c->w_risk = 0.0;
  c->max_seeds = 4096; c->max_knots = 1024; c->max_obstacle_modes = 256; c->max_total_steps = 512;
}

int32_t jssp_make_grids(double total_time, double jerk, int32_t num_samples, int32_t cap, double* time_array, double* j1,
                        double* t1, double* t2, double* j3, double* t3, double* t4, double* js, double* ts, int32_t n_out[9]) {
  if (!n_out || total_time <= 0 || num_samples < 1) return JSSP_ERR_INVALID_ARG;
  HostGrids g = make_grids(total_time, jerk, num_samples);
  const std::vector<double>* v[9] = {&g.time, &g.j1, &g.t1, &g.t2, &g.j3, &g.t3, &g.t4, &g.js, &g.ts};
  double* o[9] = {time_array, j1, t1, t2, j3, t3, t4, js, ts};
  for (int k = 0; k < 9; ++k) {
    n_out[k] = (int32_t)v[k]->size();
    if (!o[k]) continue;
    if ((int)v[k]->size() > cap) return JSSP_ERR_CAPACITY;
    memcpy(o[k], v[k]->data(), v[k]->size() * sizeof(double));
  }
  return JSSP_OK;
}

void jssp_quintic_coeffs(double xs, double vxs, double axs, double xe, double vxe, double axe, double T, double out[6]) {
  quintic_coeffs(xs, vxs, axs, xe, vxe, axe, T, out);
}

int32_t jssp_create(const jssp_config* cfg, int32_t device, jssp_handle** out) {
  if (!cfg || !out) return JSSP_ERR_INVALID_ARG;
  *out = nullptr;
  if (!config_ok(cfg)) return JSSP_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return JSSP_ERR_NO_DEVICE;
  }
  jssp_handle* h = new (std::nothrow) jssp_handle();
  if (!h) return JSSP_ERR_INVALID_ARG;
  h->cfg = *cfg;
  h->tune = Tuning::from_env();
  h->device = device;
  h->P = (int)(cfg->plan_horizon / cfg->delta_t) + 1;   // len(time_array), polyvt_sampling.py:44
  auto bail = [&](int code) { jssp_destroy(h); return code; };
#define CKC(call)                                                                     \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) { fprintf(stderr, "jssp_create: %s: %s\n", #call, cudaGetErrorString(e__)); return bail(JSSP_ERR_CUDA); } \
  } while (0)
  CKC(cudaSetDevice(device));
  int optin = 0;
  CKC(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  h->max_smem_optin = optin;   // lowered below to the tightest per-kernel limit (device limit - the kernel's static shared memory)
  CKC(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));
  {
    double** slots[10] = {&h->d_time, &h->d_j1, &h->d_t1, &h->d_t2, &h->d_j3, &h->d_t3, &h->d_t4, &h->d_jneg, &h->d_jpos, &h->d_ts};
    int n_time = 0;
    CKC(upload_grids(cfg, slots, &h->grids, &n_time));
    if (n_time != h->P) return bail(JSSP_ERR_INVALID_ARG);
  }
  h->main_total = (uint32_t)((size_t)h->grids.n_j1 * h->grids.n_t1 * h->grids.n_t2 * h->grids.n_j3 * h->grids.n_t3 * h->grids.n_t4);
  h->stop_total = (uint32_t)((size_t)h->grids.n_jneg * h->grids.n_ts * h->grids.n_jpos * h->grids.n_ts);

  CKC(cudaMalloc(&h->d_seeds, (size_t)cfg->max_seeds * 10 * sizeof(double)));
  h->kpad = (cfg->max_knots + 1) & ~1;   // capacity; the per-scenario stride is set in jssp_set_refline
  CKC(cudaMalloc(&h->d_knot, ((size_t)cfg->max_knots + 2) * sizeof(double)));   // the kernels bulk-copy kpad = K rounded up to even
  CKC(cudaMalloc(&h->d_coef, (size_t)8 * h->kpad * sizeof(double)));
  const size_t om = (size_t)(cfg->max_obstacle_modes > 0 ? cfg->max_obstacle_modes : 1), tt = (size_t)cfg->max_total_steps;
  h->raw_bytes = om * tt * 3 * sizeof(double) + om * 2 * sizeof(double) + om * sizeof(double) + om * tt;
  CKC(cudaMalloc(&h->d_raw, h->raw_bytes));
  CKC(cudaMallocHost(&h->h_raw, h->raw_bytes));
  CKC(cudaEventCreateWithFlags(&h->raw_done, cudaEventDisableTiming));
  CKC(cudaStreamCreateWithFlags(&h->up_stream, cudaStreamNonBlocking));
  CKC(cudaEventCreateWithFlags(&h->tables_ready, cudaEventDisableTiming));
  CKC(cudaEventCreateWithFlags(&h->tables_free, cudaEventDisableTiming));
  CKC(cudaMalloc(&h->d_broad, ((om + 31) & ~(size_t)31) * tt * sizeof(float4)));
  CKC(cudaMalloc(&h->d_narrow, om * tt * sizeof(float4)));
  CKC(cudaMalloc(&h->d_obs64, om * tt * 4 * sizeof(double)));
  CKC(cudaMalloc(&h->d_ext64, om * 2 * sizeof(double)));
  CKC(cudaMalloc(&h->d_sint, ((om + 31) & ~(size_t)31) * tt * sizeof(float2)));
  CKC(cudaMalloc(&h->d_probf, om * sizeof(float)));
  CKC(cudaMalloc(&h->d_knot_xy, (size_t)cfg->max_knots * 2 * sizeof(double)));
  h->max_blocks = (int)(((size_t)cfg->max_seeds * JSSP_MAX_WIDTHS + SMALL_WARPS - 1) / SMALL_WARPS);   // warp-per-candidate grid (the largest)
  CKC(cudaMalloc(&h->d_blk_cost, (size_t)h->max_blocks * sizeof(double)));
  CKC(cudaMalloc(&h->d_blk_idx, (size_t)h->max_blocks * sizeof(int)));
  CKC(cudaMalloc(&h->d_ticket, sizeof(unsigned int)));
  CKC(cudaMemset(h->d_ticket, 0, sizeof(unsigned int)));
  // one contiguous result block so that a cycle's results come back with a single copy:
  // [BestRecord, padded to 48 B][seed counts 16 B][selected trajectory P x 16 float64]
  static_assert(sizeof(BestRecord) <= 48, "result block layout");
  // The result block [BestRecord, padded to 64 B][selected trajectory P x 16 float64] lives in MAPPED pinned host memory: the
  // last block of the evaluation kernel writes the selection and the exported rows straight across the bus, so a cycle's
  // results need no copy call at all - the host only waits for the stream.  The seed counters stay in device memory (every
  // block reads them) with a 16-byte pinned mirror for the calls that report them.
  static_assert(sizeof(BestRecord) <= 64, "result block layout");
  h->result_bytes = 64 + (size_t)h->P * JSSP_TRAJ_COLS * sizeof(double);
  CKC(cudaHostAlloc(&h->h_result, h->result_bytes, cudaHostAllocMapped));
  memset(h->h_result, 0, h->result_bytes);
  CKC(cudaHostGetDevicePointer((void**)&h->d_result, h->h_result, 0));
  h->d_best = reinterpret_cast<BestRecord*>(h->d_result);
  h->d_best_traj = reinterpret_cast<double*>(h->d_result + 64);
  h->h_best = reinterpret_cast<BestRecord*>(h->h_result);
  h->h_best_traj = reinterpret_cast<double*>(h->h_result + 64);
  CKC(cudaMalloc(&h->d_counts, 4 * sizeof(int)));
  CKC(cudaMemset(h->d_counts, 0, 4 * sizeof(int)));
  CKC(cudaMallocHost(&h->h_counts, 4 * sizeof(int)));
  CKC(cudaMalloc(&h->d_split_rec, sizeof(SplitRecord)));
  CKC(cudaMemset(h->d_split_rec, 0xff, sizeof(SplitRecord)));
  CKC(cudaMalloc(&h->d_split_table, SPLIT_MAX_WORLD * sizeof(SplitRecord)));
  CKC(cudaHostAlloc(&h->h_gbest, sizeof(GlobalBest), cudaHostAllocMapped));
  memset(h->h_gbest, 0, sizeof(GlobalBest));
  CKC(cudaHostGetDevicePointer((void**)&h->d_gbest, h->h_gbest, 0));
  h->seed_work = make_seed_work(h->grids, h->d_counts, h->d_seeds, cfg->max_seeds);
  h->seed_smem = seed_smem_bytes(h->seed_work);
  if (h->seed_smem > (size_t)h->max_smem_optin) return bail(JSSP_ERR_CAPACITY);
  CKC(allow_max_dynamic_smem(seed_enum_kernel<false>, optin, &h->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_kernel<0>, optin, &h->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_kernel<1>, optin, &h->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_kernel<2>, optin, &h->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_small_kernel, optin, &h->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_group_kernel<0>, optin, &h->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_group_kernel<1>, optin, &h->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_group_kernel<2>, optin, &h->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_tiled_kernel<0>, optin, &h->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_tiled_kernel<1>, optin, &h->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_tiled_kernel<2>, optin, &h->max_smem_optin));
  CKC(allow_max_dynamic_smem(dump_states_kernel, optin, &h->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_fop_kernel, optin, &h->max_smem_optin));
  CKC(allow_max_dynamic_smem(fop_select_export_kernel, optin, &h->max_smem_optin));
  CKC(cudaDeviceSynchronize());
#undef CKC
  *out = h;
  return JSSP_OK;
}

int32_t jssp_destroy(jssp_handle* h) {
  if (!h) return JSSP_OK;
  cudaSetDevice(h->device);
  void* dev[] = {h->d_time, h->d_j1, h->d_t1, h->d_t2, h->d_j3, h->d_t3, h->d_t4, h->d_jneg, h->d_jpos, h->d_ts, h->d_seeds,
                 h->d_knot, h->d_coef, h->d_raw, h->d_broad, h->d_narrow, h->d_obs64, h->d_ext64, h->d_sint, h->d_probf, h->d_knot_xy, h->d_lane_scratch, h->d_blk_cost, h->d_blk_idx, h->d_ticket, h->d_counts};
  for (void* p : dev) if (p) cudaFree(p);
  if (h->h_result) cudaFreeHost(h->h_result);
  if (h->h_counts) cudaFreeHost(h->h_counts);
  if (h->d_fop_tab) cudaFree(h->d_fop_tab);
  if (h->d_fop_int) cudaFree(h->d_fop_int);
  if (h->h_raw) cudaFreeHost(h->h_raw);
  if (h->raw_done) cudaEventDestroy(h->raw_done);
  if (h->tables_ready) cudaEventDestroy(h->tables_ready);
  if (h->tables_free) cudaEventDestroy(h->tables_free);
  if (h->up_stream) cudaStreamDestroy(h->up_stream);
  jssp_peer_disconnect(h);
  if (h->d_split_rec) cudaFree(h->d_split_rec);
  if (h->d_split_table) cudaFree(h->d_split_table);
  if (h->h_gbest) cudaFreeHost(h->h_gbest);
  delete h;
  return JSSP_OK;
}

int32_t jssp_set_refline(jssp_handle* h, const double* knot_s, const double* coef, int32_t K, void* stream) {
  if (!h || !knot_s || !coef || K < 2) return JSSP_ERR_INVALID_ARG;
  if (K > h->cfg.max_knots) return JSSP_ERR_CAPACITY;
  for (int i = 1; i < K; ++i) if (!(knot_s[i] > knot_s[i - 1])) return JSSP_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  if (h->wait_tables) CK(cudaEventSynchronize(h->tables_ready));   // a pending obstacle packing on the private stream reads the knots
  if (h->tables_in_use) CK(cudaEventSynchronize(h->tables_free));  // ... and so does the last evaluation
  // structure-of-arrays [8][kpad] so that lanes on neighbouring segments hit different shared-memory banks
  h->kpad = (K + 1) & ~1;
  std::vector<double> soa((size_t)8 * h->kpad, 0.0);
  for (int i = 0; i < K - 1; ++i)
    for (int c = 0; c < 8; ++c) soa[(size_t)c * h->kpad + i] = coef[(size_t)i * 8 + c];
  CK(cudaMemcpyAsync(h->d_knot, knot_s, (size_t)K * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(h->d_coef, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  // block-level culling tables: knot positions, the longest knot spacing and a rigorous bound on |r'(s)| (triangle
  // inequality on the per-segment quadratics), so that "within reach of a knot" covers every point of the curve
  std::vector<double> kxy((size_t)K * 2);
  double hmax = 0.0, gmax = 0.0;
  for (int i = 0; i < K - 1; ++i) {
    const double* c = coef + (size_t)i * 8;
    const double hh = knot_s[i + 1] - knot_s[i];
    kxy[2 * i] = c[0]; kxy[2 * i + 1] = c[4];
    hmax = std::max(hmax, hh);
    const double bx = std::fabs(c[1]) + 2.0 * std::fabs(c[2]) * hh + 3.0 * std::fabs(c[3]) * hh * hh;
    const double by = std::fabs(c[5]) + 2.0 * std::fabs(c[6]) * hh + 3.0 * std::fabs(c[7]) * hh * hh;
    gmax = std::max(gmax, std::hypot(bx, by));
    if (i == K - 2) {
      kxy[2 * (K - 1)] = ((c[0] + c[1] * hh) + c[2] * (hh * hh)) + c[3] * ((hh * hh) * hh);
      kxy[2 * (K - 1) + 1] = ((c[4] + c[5] * hh) + c[6] * (hh * hh)) + c[7] * ((hh * hh) * hh);
    }
  }
  CK(cudaMemcpyAsync(h->d_knot_xy, kxy.data(), kxy.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));   // soa, kxy are temporaries
  h->K = K;
  h->s_end = knot_s[K - 1];
  h->cull_ok = std::isfinite(gmax) && gmax < 8.0;
  h->cull_reach = h->cfg.max_road_width + gmax * hmax;   // lateral cap + longest knot-to-knot arc (+ radii, added per obstacle)
  h->cull_ds = hmax;
  h->ox = coef[0]; h->oy = coef[4];   // first knot = local origin of the fp32 collision frame
  if (h->obs_set) h->obs_dirty = true;
  return JSSP_OK;
}

static int32_t set_obstacles_common(jssp_handle* h, int32_t O, int32_t M, int32_t Tt, int32_t n_dict) {
  if (O < 0 || M < 1 || Tt < 1) return JSSP_ERR_INVALID_ARG;
  if (O * M > h->cfg.max_obstacle_modes || Tt > h->cfg.max_total_steps) return JSSP_ERR_CAPACITY;
  h->O = O; h->M = M; h->Tt = Tt; h->n_dict = n_dict;
  h->obs_set = true; h->obs_dirty = true;
  return JSSP_OK;
}

int32_t jssp_set_obstacles(jssp_handle* h, const double* state, const uint8_t* valid, const double* size, const double* prob,
                           int32_t O, int32_t M, int32_t Tt, int32_t n_dict, void* stream) {
  if (!h) return JSSP_ERR_INVALID_ARG;
  if (O > 0 && (!state || !valid || !size || !prob)) return JSSP_ERR_INVALID_ARG;
  int rc = set_obstacles_common(h, O, M, Tt, n_dict);
  if (rc != JSSP_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  // one staged upload: the four host arrays are gathered in pinned memory and cross the bus as a single copy
  const size_t om = (size_t)O * M;
  const size_t b_state = om * Tt * 3 * sizeof(double), b_size = (size_t)O * 2 * sizeof(double), b_prob = om * sizeof(double), b_valid = om * Tt;
  unsigned char* d = h->d_raw;
  if (om > 0) {
    CK(cudaEventSynchronize(h->raw_done));             // the previous upload has left the staging block
    unsigned char* s = h->h_raw;
    memcpy(s, state, b_state); memcpy(s + b_state, size, b_size); memcpy(s + b_state + b_size, prob, b_prob);
    memcpy(s + b_state + b_size + b_prob, valid, b_valid);
    // upload + pack on the private stream, after the last evaluation has finished with the tables; the caller's stream is not
    // touched here, so e.g. the H2D copy of this cycle's seeds proceeds in parallel.  jssp_evaluate waits for tables_ready.
    (void)st;
    if (h->tables_in_use) CK(cudaStreamWaitEvent(h->up_stream, h->tables_free, 0));
    CK(cudaMemcpyAsync(d, s, b_state + b_size + b_prob + b_valid, cudaMemcpyHostToDevice, h->up_stream));
    CK(cudaEventRecord(h->raw_done, h->up_stream));
  }
  h->src_state = reinterpret_cast<const double*>(d); h->src_size = reinterpret_cast<const double*>(d + b_state);
  h->src_prob = reinterpret_cast<const double*>(d + b_state + b_size); h->src_valid = d + b_state + b_size + b_prob;
  rc = pack_if_dirty(h, h->up_stream);
  if (rc != JSSP_OK) return rc;
  CK(cudaEventRecord(h->tables_ready, h->up_stream));
  h->wait_tables = true;
  return JSSP_OK;
}

// make `st` wait for a pending private-stream upload of the obstacle tables
static int32_t join_tables(jssp_handle* h, cudaStream_t st) {
  if (h->wait_tables) {
    CK(cudaStreamWaitEvent(st, h->tables_ready, 0));
    h->wait_tables = false;
  }
  return JSSP_OK;
}
// the evaluation just enqueued on `st` reads the tables: the next private-stream upload must wait for it
static int32_t mark_tables_in_use(jssp_handle* h, cudaStream_t st) {
  CK(cudaEventRecord(h->tables_free, st));
  h->tables_in_use = true;
  return JSSP_OK;
}

int32_t jssp_set_obstacles_dev(jssp_handle* h, const double* state_dev, const uint8_t* valid_dev, const double* size_dev,
                               const double* prob_dev, int32_t O, int32_t M, int32_t Tt, int32_t n_dict, void* stream) {
  if (!h) return JSSP_ERR_INVALID_ARG;
  if (O > 0 && (!state_dev || !valid_dev || !size_dev || !prob_dev)) return JSSP_ERR_INVALID_ARG;
  int rc = set_obstacles_common(h, O, M, Tt, n_dict);
  if (rc != JSSP_OK) return rc;
  CK(cudaSetDevice(h->device));
  rc = join_tables(h, (cudaStream_t)stream);
  if (rc != JSSP_OK) return rc;
  h->src_state = state_dev; h->src_valid = valid_dev; h->src_size = size_dev; h->src_prob = prob_dev;
  return pack_if_dirty(h, (cudaStream_t)stream);
}

int32_t jssp_enumerate_seeds(jssp_handle* h, double s0, double v0, double a0, void* stream, int32_t* n_main, int32_t* n_stop) {
  if (!h) return JSSP_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  SeedState ss{s0, v0, a0};
  CK(launch_seed_enum<false>(h->grids, ss, nullptr, h->seed_work, 1, SEED_MAX_CLUSTER, SEED_THREADS, h->seed_smem, st));
  h->launches += 1;
  CK(cudaGetLastError());
  h->seeds_ready = true;
  if (n_main || n_stop) {
    CK(cudaMemcpyAsync(h->h_counts, h->d_counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (n_main) *n_main = h->h_counts[0];
    if (n_stop) *n_stop = h->h_counts[1];
    if (h->h_counts[3]) return JSSP_ERR_SEED_OVERFLOW;
  }
  return JSSP_OK;
}

int32_t jssp_get_seeds(jssp_handle* h, double* seeds_host, int32_t cap, int32_t* n, void* stream) {
  if (!h || !n) return JSSP_ERR_INVALID_ARG;
  if (!h->seeds_ready) return JSSP_ERR_NOT_READY;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->h_counts, h->d_counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (h->h_counts[3]) return JSSP_ERR_SEED_OVERFLOW;
  *n = h->h_counts[2];
  if (seeds_host) {
    if (*n > cap) return JSSP_ERR_CAPACITY;
    CK(cudaMemcpyAsync(seeds_host, h->d_seeds, (size_t)(*n) * 10 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  return JSSP_OK;
}

int32_t jssp_fetch_best(jssp_handle* h, void* stream, jssp_cycle_out* out) {
  if (!h || !out) return JSSP_ERR_INVALID_ARG;
  if (!h->last_valid) return JSSP_ERR_NOT_READY;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  const bool want_traj = out->best_traj_host && !(h->last_flags & JSSP_FLAG_NO_EXPORT);
  CK(cudaStreamSynchronize(st));      // the kernel wrote the result block into mapped host memory: nothing to copy
  out->n_candidates = h->h_best->n_candidates;
  out->n_seeds = h->h_best->n_seeds;
  out->best_idx = (int32_t)h->h_best->idx;
  out->best_cost = h->h_best->cost;
  out->best_n_pts = h->h_best->n_pts;
  if (want_traj && out->best_idx >= 0) memcpy(out->best_traj_host, h->h_best_traj, (size_t)h->P * JSSP_TRAJ_COLS * sizeof(double));
  // enumerated seeds: surface an overflow detected on the device
  return (h->last.n_seeds_ptr && h->h_best->pad) ? JSSP_ERR_SEED_OVERFLOW : JSSP_OK;   // pad = the enumeration's overflow flag
}

int32_t jssp_evaluate(jssp_handle* h, const jssp_cycle_in* in, void* stream, jssp_cycle_out* out) {
  if (!h || !in || !out) return JSSP_ERR_INVALID_ARG;
  if (h->K < 2 || !h->obs_set) return JSSP_ERR_NOT_READY;
  const bool ext = in->seeds_dev != nullptr;
  const bool enumerate = !ext && (in->flags & JSSP_FLAG_ENUMERATE);
  if (!ext && !enumerate && !h->seeds_ready) return JSSP_ERR_NOT_READY;
  if (ext && (in->n_seeds < 0 || in->n_seeds > h->cfg.max_seeds)) return JSSP_ERR_CAPACITY;
  const int W = in->n_widths > 0 ? in->n_widths : h->cfg.num_width;
  if (W < 1 || W > JSSP_MAX_WIDTHS) return JSSP_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  int rc = join_tables(h, st);
  if (rc == JSSP_OK) rc = pack_if_dirty(h, st);
  if (rc != JSSP_OK) return rc;
  if (enumerate) {             // last, so that the evaluation kernel directly follows the seed enumeration in the stream
    const int rc0 = jssp_enumerate_seeds(h, in->s0, in->v0, in->a0, stream, nullptr, nullptr);
    if (rc0 != JSSP_OK) return rc0;
  }

  EvalParams p{};
  const jssp_config& c = h->cfg;
  p.s0 = in->s0; p.v0 = in->v0; p.a0 = in->a0;
  // lateral targets: linspace(-(max_road_width - w)/2, +(max_road_width - w)/2, num_width) (jerk_space_sampling_planner.py:100-104)
  std::vector<double> tg;
  if (in->n_widths > 0) tg.assign(in->targets, in->targets + W);
  else tg = default_targets(c, W);
  for (int w = 0; w < W; ++w) quintic_coeffs(in->d0, in->d_d0, in->d_dd0, tg[w], 0.0, 0.0, c.plan_horizon, p.quintic[w]);
  fill_constants(p, c, h->P);
  p.W = W; p.now = in->time_step_now; p.tcap = in->final_time_step - in->time_step_now;
  p.n_dict = h->n_dict;
  p.time = h->d_time; p.knot = h->d_knot; p.coef = h->d_coef; p.K = h->K; p.kpad = h->kpad; p.ox = h->ox; p.oy = h->oy;
  p.seeds = ext ? in->seeds_dev : h->d_seeds;
  p.n_seeds_ptr = ext ? nullptr : h->d_counts;
  p.n_seeds_fixed = ext ? in->n_seeds : 0;
  p.seed_cap = c.max_seeds;
  p.broad = h->d_broad; p.narrow = h->d_narrow; p.obs64 = h->d_obs64; p.ext64 = h->d_ext64;
  p.OM = h->O * h->M; p.OMpad = (p.OM + 31) & ~31; p.Tt = h->Tt;
  int rows = 0;   // collision rows of the window: i = 0, check_res, ... < min(P, tcap)
  if (p.tcap > 0 && p.now >= 0) rows = (std::min(p.P, p.tcap) + c.check_res - 1) / c.check_res;
  if (p.now < 0) return JSSP_ERR_INVALID_ARG;
  p.win_rows = p.OM > 0 ? rows : 0;
  p.sint = h->d_sint; p.probf = h->d_probf;
  p.list_cap = std::min(p.win_rows * p.OMpad, 2048);
  // block-level culling is valid while every lateral offset of this cycle stays inside the cap the intervals were built for
  double dlat = 0.0;
  for (int w = 0; w < W; ++w)
    for (int i = 0; i < h->P; ++i) {
      double d, d1, d2, d3;
      quintic_eval(p.quintic[w], c.plan_horizon * i / (h->P - 1), d, d1, d2, d3);
      dlat = std::max(dlat, std::fabs(d));
    }
  int mode = (h->cull_ok && p.win_rows > 0 && dlat * 1.001 + 1e-6 <= c.max_road_width && !(in->flags & JSSP_FLAG_NO_CULL)) ? 2 : 1;
  SmemLayout L = eval_smem_layout(p.win_rows, p.OMpad, mode, p.K, p.kpad, p.P, p.W, p.list_cap);
  if (mode == 2 && L.total > (size_t)h->max_smem_optin) { mode = 1; L = eval_smem_layout(p.win_rows, p.OMpad, 1, p.K, p.kpad, p.P, p.W); }
  if (mode == 1 && (p.win_rows == 0 || L.total > (size_t)h->max_smem_optin)) {
    mode = 0;
    L = eval_smem_layout(p.win_rows, p.OMpad, 0, p.K, p.kpad, p.P, p.W);
    if (L.total > (size_t)h->max_smem_optin) return JSSP_ERR_CAPACITY;
  }
  p.smem_obs = mode;
  p.feasible = out->feasible_dev; p.collision = out->collision_dev; p.cost = out->cost_dev;
  p.select_only = (!p.feasible && !p.collision && !p.cost) ? 1 : 0;    // nobody reads per-candidate results: walks may stop at the first veto
  p.blk_cost = h->d_blk_cost; p.blk_idx = h->d_blk_idx; p.ticket = h->d_ticket; p.best = h->d_best;
  p.best_traj = (in->flags & JSSP_FLAG_NO_EXPORT) ? nullptr : h->d_best_traj;
  if (h->split_ns_global > 0) {           // candidate split: this launch covers seeds [seed_lo, seed_lo + n_seeds) of ns_global
    if (!ext || h->split_seed_lo + in->n_seeds > h->split_ns_global) return JSSP_ERR_INVALID_ARG;
    p.split_seed_lo = h->split_seed_lo; p.split_ns_global = h->split_ns_global; p.split_rec = h->d_split_rec;
    if (h->peers_connected) { p.peers = h->d_peers; p.split_epoch = ++h->split_epoch; }
  }
#ifdef JSSP_PROFILE_PHASES
  p.phase_ts = h->phase_ts;
#endif

  const long long n_max = (long long)W * (ext ? in->n_seeds : c.max_seeds);
  // small cycles (the default sampling grid: ~10^3 candidates) run one warp per candidate: latency, not throughput, matters
  const SmemLayout Ls = eval_smem_layout(0, p.OMpad, 0, p.K, p.kpad, p.P, p.W);
  const size_t small_bytes = Ls.total + small_scratch_bytes(p.P);
  bool small = n_max <= 16384 && small_bytes <= (size_t)h->max_smem_optin;
  if (in->flags & JSSP_FLAG_FORCE_THREAD) small = false;
  if ((in->flags & JSSP_FLAG_FORCE_WARP) && small_bytes <= (size_t)h->max_smem_optin) small = true;
  if (in->flags & (JSSP_FLAG_FORCE_GROUP | JSSP_FLAG_FORCE_TILED)) small = false;
  const bool group = !small && !(in->flags & JSSP_FLAG_FORCE_THREAD);
  // large cycles: the tiled two-phase kernel unless the shuffle form is asked for (or its scratch does not fit)
  // tile length of the tiled kernel: the longest (<= ~16 samples) with which two blocks still share an SM, else the longest that fits
  int tile_rounds = 0;
  size_t tiled_total = 0;
  {
    cudaFuncAttributes fa{};
    cudaFuncGetAttributes(&fa, evaluate_tiled_kernel<2>);
    const size_t per_sm = 228u * 1024u, two = per_sm / 2 - fa.sharedSizeBytes - 1024;
    const size_t lbase = (L.total + 15) & ~(size_t)15;
    for (int r = tile_rounds_wanted(W); r >= 1; --r) {
      tile_rounds = r; tiled_total = lbase + tile_scratch_bytes(W, EVAL_THREADS, r);
      if (tiled_total <= two) break;
    }
    if (tiled_total > two)                                  // not even the shortest tile leaves room for two blocks: one block, long tiles
      for (int r = tile_rounds_wanted(W); r >= 1; --r) {
        tile_rounds = r; tiled_total = lbase + tile_scratch_bytes(W, EVAL_THREADS, r);
        if (tiled_total <= (size_t)h->max_smem_optin) break;
      }
  }
  p.tile_rounds = tile_rounds;
  const bool tiled = group && !(in->flags & JSSP_FLAG_FORCE_GROUP) && tiled_total <= (size_t)h->max_smem_optin;
  const long long seeds_max = ext ? in->n_seeds : c.max_seeds;
  const int spb = group_seeds_per_block(W);
  // warp-per-candidate grid: grid-stride kernel, sized from the last cycle's candidate count (x2) when the seeds are
  // enumerated on the device, never above what the capacity needs nor above a few resident waves
  long long small_need = (n_max + SMALL_WARPS - 1) / SMALL_WARPS;
  if (small && !ext) {
    const long long hint = h->last_valid && h->h_best->n_candidates > 0 ? ((long long)h->h_best->n_candidates * 2 + SMALL_WARPS - 1) / SMALL_WARPS : 256;
    small_need = std::min(small_need, std::max(hint, (long long)h->sm_count));
  }
  small_need = std::min(small_need, (long long)h->sm_count * 8);
  int blocks = small ? (int)small_need
                     : (group ? (int)((seeds_max + spb - 1) / spb) : (int)((n_max + EVAL_THREADS - 1) / EVAL_THREADS));
  if (blocks < 1) blocks = 1;
  // tiled kernel, single wave (at most two blocks per SM): fewer warps per block so that EVERY SM holds the same number of warps -
  // 256 blocks of 8 warps leave 108 SMs with 16 warps and 40 with 8, and the kernel ends with the slow ones; 293 blocks of 7 warps
  // give every SM 14 (BASELINE configs[2]: 16,384 seeds x 4 lateral targets)
  int tiled_threads = EVAL_THREADS;
  if (tiled && blocks <= 2 * h->sm_count && blocks > h->sm_count / 2 && !h->tune.lockstep) {
    const int gpw_t = 32 / W;
    const long long per_block = (seeds_max + 2 * h->sm_count - 1) / (2 * h->sm_count);
    const int warps = (int)std::min<long long>(EVAL_THREADS / 32, std::max<long long>(2, (per_block + gpw_t - 1) / gpw_t));
    tiled_threads = warps * 32;
    blocks = (int)((seeds_max + (long long)warps * gpw_t - 1) / ((long long)warps * gpw_t));
  }
  if (blocks > h->max_blocks) return JSSP_ERR_CAPACITY;
  if (small) {
    p.smem_obs = 0; p.win_rows = 0; p.list_cap = 0;
    // one-call plan cycle: the evaluation is chained to the seed enumeration just enqueued (programmatic dependent launch)
    const bool chained = !ext && (in->flags & JSSP_FLAG_ENUMERATE) && !h->tune.no_pdl;
    CK(launch_chained(evaluate_small_kernel, dim3(blocks), SMALL_WARPS * 32, small_bytes, st, chained, p));
  }
  else if (tiled && mode == 2) evaluate_tiled_kernel<2><<<blocks, tiled_threads, tiled_total, st>>>(p);
  else if (tiled && mode == 1) evaluate_tiled_kernel<1><<<blocks, tiled_threads, tiled_total, st>>>(p);
  else if (tiled) evaluate_tiled_kernel<0><<<blocks, tiled_threads, tiled_total, st>>>(p);
  else if (group && mode == 2) evaluate_group_kernel<2><<<blocks, EVAL_THREADS, L.total, st>>>(p);
  else if (group && mode == 1) evaluate_group_kernel<1><<<blocks, EVAL_THREADS, L.total, st>>>(p);
  else if (group) evaluate_group_kernel<0><<<blocks, EVAL_THREADS, L.total, st>>>(p);
  else if (mode == 2) evaluate_kernel<2><<<blocks, EVAL_THREADS, L.total, st>>>(p);
  else if (mode == 1) evaluate_kernel<1><<<blocks, EVAL_THREADS, L.total, st>>>(p);
  else evaluate_kernel<0><<<blocks, EVAL_THREADS, L.total, st>>>(p);
  h->launches++;
  CK(cudaGetLastError());
  rc = mark_tables_in_use(h, st);
  if (rc != JSSP_OK) return rc;
  EvalParams q = p;
  q.smem_obs = 0; q.win_rows = 0; q.list_cap = 0;
  const SmemLayout L2 = eval_smem_layout(0, q.OMpad, 0, q.K, q.kpad, q.P, q.W);
  if (out->states_dev) {
    const size_t dump_bytes = ((L2.total + 15) & ~(size_t)15) + dump_scratch_bytes(q.P);
    if (dump_bytes > (size_t)h->max_smem_optin) return JSSP_ERR_CAPACITY;
    const long long want = (seeds_max + DUMP_WARPS - 1) / DUMP_WARPS;
    const int dump_blocks = (int)std::max<long long>(1, std::min<long long>(want, (long long)h->sm_count * 4));
    dump_states_kernel<<<dump_blocks, DUMP_WARPS * 32, dump_bytes, st>>>(q, out->states_dev);
    h->launches++;
    CK(cudaGetLastError());
  }
  h->last = p; h->last_valid = true; h->last_flags = in->flags;
  if (in->flags & JSSP_FLAG_NO_SYNC) return JSSP_OK;
  return jssp_fetch_best(h, stream, out);
}

int32_t jssp_best_record_dev(jssp_handle* h, int64_t index_offset, void* stream, uint64_t* record_dev) {
  if (!h || !record_dev) return JSSP_ERR_INVALID_ARG;
  if (!h->last_valid) return JSSP_ERR_NOT_READY;
  CK(cudaSetDevice(h->device));
  best_record_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(h->d_best, (long long)index_offset, (unsigned long long*)record_dev);
  h->launches++;
  CK(cudaGetLastError());
  return JSSP_OK;
}


// -- candidate split over ranks ----------------------------------------------------------------------------------------
namespace {
// NCCL is reached through dlopen of the library the process already uses (torch bundles libnccl.so.2): no link-time dependency.
// Minimal declarations after nccl.h (ncclResult_t = int, ncclSuccess = 0, ncclUint64 = 5, ncclUniqueId = 128 opaque bytes).
struct NcclApi {
  struct Id128 { char b[128]; };      // ncclUniqueId, passed by value
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok() const { return lib && GetUniqueId && CommInitRank && CommDestroy && AllGather; }
};
NcclApi& nccl_api() {
  static NcclApi api;
  if (api.lib) return api;
  const char* env = getenv("JSSP_NCCL_LIB");
  void* lib = env ? dlopen(env, RTLD_NOW | RTLD_GLOBAL) : nullptr;
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // already in the process (torch.distributed)
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW);
  if (!lib) return api;
  api.lib = lib;
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(lib, "ncclAllGather"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  return api;
}
}  // namespace

int32_t jssp_nccl_unique_id(uint8_t id_out[128]) {
  if (!id_out) return JSSP_ERR_INVALID_ARG;
  NcclApi& n = nccl_api();
  if (!n.ok()) return JSSP_ERR_NOT_READY;
  return n.GetUniqueId(id_out) == 0 ? JSSP_OK : JSSP_ERR_CUDA;
}

int32_t jssp_nccl_comm_create(const uint8_t id[128], int32_t world, int32_t rank, int32_t device, void** comm_out) {
  if (!id || !comm_out || world < 1 || world > SPLIT_MAX_WORLD || rank < 0 || rank >= world) return JSSP_ERR_INVALID_ARG;
  NcclApi& n = nccl_api();
  if (!n.ok()) return JSSP_ERR_NOT_READY;
  if (cudaSetDevice(device) != cudaSuccess) return JSSP_ERR_NO_DEVICE;
  NcclApi::Id128 uid;
  memcpy(uid.b, id, 128);
  void* comm = nullptr;
  const int rc = n.CommInitRank(&comm, world, uid, rank);
  if (rc != 0) { fprintf(stderr, "jssp_nccl_comm_create: %s\n", n.GetErrorString ? n.GetErrorString(rc) : "nccl error"); return JSSP_ERR_CUDA; }
  *comm_out = comm;
  return JSSP_OK;
}

int32_t jssp_nccl_comm_destroy(void* comm) {
  if (!comm) return JSSP_OK;
  NcclApi& n = nccl_api();
  if (!n.ok()) return JSSP_ERR_NOT_READY;
  return n.CommDestroy(comm) == 0 ? JSSP_OK : JSSP_ERR_CUDA;
}

int32_t jssp_set_split(jssp_handle* h, int64_t seed_lo, int64_t n_seeds_global) {
  if (!h || seed_lo < 0 || n_seeds_global < 0 || seed_lo > n_seeds_global) return JSSP_ERR_INVALID_ARG;
  h->split_seed_lo = seed_lo; h->split_ns_global = n_seeds_global;
  return JSSP_OK;
}

int32_t jssp_allreduce_argmin(jssp_handle* h, void* nccl_comm, int32_t world, void* stream) {
  if (!h || world < 1 || world > SPLIT_MAX_WORLD || (world > 1 && !nccl_comm)) return JSSP_ERR_INVALID_ARG;
  if (!h->last_valid || h->split_ns_global <= 0) return JSSP_ERR_NOT_READY;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  const SplitRecord* table = h->d_split_rec;
  if (world > 1) {
    NcclApi& n = nccl_api();
    if (!n.ok()) return JSSP_ERR_NOT_READY;
    const int rc = n.AllGather(h->d_split_rec, h->d_split_table, 2, /*ncclUint64*/ 5, nccl_comm, st);
    if (rc != 0) { snprintf(h->cuda_err, sizeof(h->cuda_err), "ncclAllGather: %s", n.GetErrorString ? n.GetErrorString(rc) : "nccl error"); return JSSP_ERR_CUDA; }
    table = h->d_split_table;
  }
  split_reduce_kernel<<<1, 32, 0, st>>>(table, world, h->d_gbest);
  h->launches++;
  CK(cudaGetLastError());
  return JSSP_OK;
}

int32_t jssp_fetch_global_best(jssp_handle* h, void* stream, int64_t* best_idx, double* best_cost, int32_t* winner_rank) {
  if (!h) return JSSP_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize((cudaStream_t)stream));      // GlobalBest lives in mapped pinned memory: nothing to copy
  if (best_idx) *best_idx = h->h_gbest->idx;
  if (best_cost) *best_cost = h->h_gbest->cost;
  if (winner_rank) *winner_rank = h->h_gbest->rank;
  if (h->h_gbest->timed_out) { snprintf(h->cuda_err, sizeof(h->cuda_err), "peer exchange: a rank's record did not arrive (epoch %llu)", h->h_gbest->epoch); return JSSP_ERR_NOT_READY; }
  return JSSP_OK;
}

int32_t jssp_peer_export(jssp_handle* h, int32_t world, uint8_t handle_out[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!h || !handle_out || world < 1 || world > SPLIT_MAX_WORLD) return JSSP_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  jssp_peer_disconnect(h);
  CK(cudaMalloc(&h->d_mailbox, (size_t)2 * world * sizeof(PeerSlot)));
  CK(cudaMemset(h->d_mailbox, 0, (size_t)2 * world * sizeof(PeerSlot)));
  CK(cudaDeviceSynchronize());
  h->peer_world = world;
  cudaIpcMemHandle_t ipc;
  CK(cudaIpcGetMemHandle(&ipc, h->d_mailbox));
  memcpy(handle_out, &ipc, 64);
  return JSSP_OK;
}

int32_t jssp_peer_connect(jssp_handle* h, const uint8_t* handles, int32_t world, int32_t rank) {
  if (!h || !handles || world != h->peer_world || rank < 0 || rank >= world || !h->d_mailbox) return JSSP_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  PeerView pv{};
  pv.world = world; pv.rank = rank; pv.out = h->d_gbest;
  int khz = 0;
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device));
  pv.spin_budget = (long long)khz * 1000ll * 5ll;                      // ~5 s of SM clocks, then the exchange reports a time-out
  for (int r = 0; r < world; ++r) {
    if (r == rank) { pv.box[r] = h->d_mailbox; continue; }
    cudaIpcMemHandle_t ipc;
    memcpy(&ipc, handles + (size_t)r * 64, 64);
    void* ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->peer_ptr[r] = ptr;
    pv.box[r] = reinterpret_cast<PeerSlot*>(ptr);
  }
  if (!h->d_peers) CK(cudaMalloc(&h->d_peers, sizeof(PeerView)));
  CK(cudaMemcpy(h->d_peers, &pv, sizeof(pv), cudaMemcpyHostToDevice));
  h->peer_rank = rank; h->peers_connected = true; h->split_epoch = 0;
  return JSSP_OK;
}

int32_t jssp_peer_disconnect(jssp_handle* h) {
  if (!h) return JSSP_OK;
  cudaSetDevice(h->device);
  if (h->peers_connected || h->d_mailbox) cudaDeviceSynchronize();
  for (int r = 0; r < SPLIT_MAX_WORLD; ++r) if (h->peer_ptr[r]) { cudaIpcCloseMemHandle(h->peer_ptr[r]); h->peer_ptr[r] = nullptr; }
  if (h->d_mailbox) { cudaFree(h->d_mailbox); h->d_mailbox = nullptr; }
  if (h->d_peers) { cudaFree(h->d_peers); h->d_peers = nullptr; }
  h->peers_connected = false; h->peer_world = 0;
  return JSSP_OK;
}

int64_t jssp_launch_count(const jssp_handle* h) { return h ? h->launches : 0; }

#ifdef JSSP_PROFILE_PHASES
/* profiling builds only (tools/phase_probe.py): device buffer [blocks][8] receiving per-block globaltimer stamps */
int32_t jssp_debug_phase_buffer(jssp_handle* h, unsigned long long* dev) { if (!h) return JSSP_ERR_INVALID_ARG; h->phase_ts = dev; return JSSP_OK; }
int32_t jssp_debug_seed_stamps(unsigned long long* host16) { return cudaMemcpyFromSymbol(host16, g_seed_ts, sizeof(g_seed_ts)) == cudaSuccess ? JSSP_OK : JSSP_ERR_CUDA; }
#endif

int32_t jssp_measure_fma_peak(int32_t device, int32_t bits, double* tflops) {
  if (!tflops || (bits != 32 && bits != 64)) return JSSP_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return JSSP_ERR_NO_DEVICE; }
  return bits == 32 ? measure_fma<float>(device, tflops) : measure_fma<double>(device, tflops);
}

int32_t jssp_fop_evaluate(jssp_handle* h, const jssp_fop_in* in, void* stream, jssp_cycle_out* out) {
  if (!h || !in || !out) return JSSP_ERR_INVALID_ARG;
  if (h->K < 2 || !h->obs_set) return JSSP_ERR_NOT_READY;
  const jssp_config& c = h->cfg;
  if (in->num_width < 1 || in->num_width > JSSP_MAX_WIDTHS || in->num_speed < 1 || in->num_t < 1 || !(in->min_t > 0) || !(in->max_t >= in->min_t) ||
      in->time_step_now < 0)
    return JSSP_ERR_INVALID_ARG;
  if (in->max_t > c.plan_horizon + 1e-9) return JSSP_ERR_CAPACITY;      // the handle's sample capacity follows plan_horizon
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  int rc = join_tables(h, st);
  if (rc == JSSP_OK) rc = pack_if_dirty(h, st);
  if (rc != JSSP_OK) return rc;
  const int W = in->num_width, NT = in->num_t, NV = in->num_speed;
  const long long N = (long long)W * NT * NV;
  const int blocks = (int)((N + EVAL_THREADS - 1) / EVAL_THREADS);
  if (blocks > h->max_blocks) return JSSP_ERR_CAPACITY;
  // grids: np.linspace(min_t, max_t, num_t), np.linspace(lowest, highest, num_speed), len(np.arange(0, T, dt)) = ceil(T / dt)
  std::vector<double> tab = np_linspace(in->min_t, in->max_t, NT);
  std::vector<int> ns((size_t)NT);
  for (int i = 0; i < NT; ++i) {
    ns[i] = (int)std::ceil((tab[i] - 0.0) / c.delta_t);
    if (ns[i] < 1 || ns[i] > h->P) return JSSP_ERR_CAPACITY;
  }
  const std::vector<double> sp = np_linspace(in->lowest_speed, in->highest_speed, NV);
  tab.insert(tab.end(), sp.begin(), sp.end());
  if (NT > h->fop_cap_t || NV > h->fop_cap_v) {                          // first FOP call on this handle (or a larger grid)
    CK(cudaStreamSynchronize(st));
    if (h->d_fop_tab) cudaFree(h->d_fop_tab);
    if (h->d_fop_int) cudaFree(h->d_fop_int);
    h->fop_cap_t = std::max(NT, h->fop_cap_t); h->fop_cap_v = std::max(NV, h->fop_cap_v);
    CK(cudaMalloc(&h->d_fop_tab, (size_t)(h->fop_cap_t + h->fop_cap_v) * sizeof(double)));
    CK(cudaMalloc(&h->d_fop_int, (size_t)h->fop_cap_t * sizeof(int)));
  }
  CK(cudaMemcpyAsync(h->d_fop_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(h->d_fop_int, ns.data(), ns.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));                                          // tab, ns are temporaries

  FopParams fp{};
  EvalParams& p = fp.base;
  fill_constants(p, c, h->P);
  p.s0 = in->s0; p.v0 = in->v0; p.a0 = in->a0;
  p.W = W; p.now = in->time_step_now; p.tcap = in->final_time_step - in->time_step_now; p.n_dict = h->n_dict;
  p.time = h->d_time; p.knot = h->d_knot; p.coef = h->d_coef; p.K = h->K; p.kpad = h->kpad; p.ox = h->ox; p.oy = h->oy;
  p.seeds = nullptr; p.n_seeds_ptr = nullptr; p.n_seeds_fixed = NT * NV; p.seed_cap = c.max_seeds;
  p.broad = h->d_broad; p.narrow = h->d_narrow; p.obs64 = h->d_obs64; p.ext64 = h->d_ext64;
  p.OM = h->O * h->M; p.OMpad = (p.OM + 31) & ~31; p.Tt = h->Tt;
  p.win_rows = 0; p.smem_obs = 0; p.sint = h->d_sint; p.probf = h->d_probf; p.list_cap = 0;
  p.feasible = out->feasible_dev; p.collision = out->collision_dev; p.cost = out->cost_dev;
  p.select_only = (!p.feasible && !p.collision && !p.cost) ? 1 : 0;    // nobody reads per-candidate results: walks may stop at the first veto
  p.blk_cost = h->d_blk_cost; p.blk_idx = h->d_blk_idx; p.ticket = h->d_ticket; p.best = h->d_best;
  p.best_traj = (in->flags & JSSP_FLAG_NO_EXPORT) ? nullptr : h->d_best_traj;
  fp.num_t = NT; fp.num_speed = NV;
  fp.horizon = h->d_fop_tab; fp.speed = h->d_fop_tab + NT; fp.n_samples = h->d_fop_int;
  fp.d0 = in->d0; fp.d_d0 = in->d_d0; fp.d_dd0 = in->d_dd0;
  const std::vector<double> tg = default_targets(c, W);                   // :99-101, same expression as the JSSP lateral grid
  for (int w = 0; w < W; ++w) fp.targets[w] = tg[w];
  fp.dt = c.delta_t; fp.max_accel = in->max_accel; fp.max_jerk = in->max_jerk; fp.t_min = in->min_t; fp.t_max = in->max_t;
  const FopSmem L = fop_smem_layout(p.K, p.kpad, p.P, W, NT);
  if (L.total > (size_t)h->max_smem_optin) return JSSP_ERR_CAPACITY;
  evaluate_fop_kernel<<<blocks, EVAL_THREADS, L.total, st>>>(fp);
  fop_select_export_kernel<<<1, EVAL_THREADS, L.total, st>>>(fp, blocks);
  h->launches += 2;
  CK(cudaGetLastError());
  rc = mark_tables_in_use(h, st);
  if (rc != JSSP_OK) return rc;
  h->last = p; h->last_valid = true; h->last_flags = in->flags;
  if (in->flags & JSSP_FLAG_NO_SYNC) return JSSP_OK;
  return jssp_fetch_best(h, stream, out);
}

int32_t jssp_project_states(jssp_handle* h, const double* state_dev, int32_t n, double step, double* frenet_dev, void* stream) {
  if (!h || n < 0 || !(step > 0) || (n > 0 && (!state_dev || !frenet_dev))) return JSSP_ERR_INVALID_ARG;
  if (h->K < 2) return JSSP_ERR_NOT_READY;
  if (n == 0) return JSSP_OK;
  CK(cudaSetDevice(h->device));
  // samples of np.arange(0, s_end, step): ceil(s_end / step)
  const int n_samples = (int)std::ceil(h->s_end / step);
  if (n_samples < 1) return JSSP_ERR_INVALID_ARG;
  project_states_kernel<<<n, PROJECT_THREADS, 0, (cudaStream_t)stream>>>(h->d_knot, h->d_coef, h->K, h->kpad, step, n_samples, state_dev, frenet_dev);
  h->launches++;
  CK(cudaGetLastError());
  return JSSP_OK;
}

int32_t jssp_imm_update(jssp_handle* h, const jssp_imm_params* prm, double* mean_dev, double* cov_dev, double* mu_dev,
                        const double* z_dev, int32_t O, int32_t M, void* stream) {
  if (!h || !prm || !mean_dev || !cov_dev || !mu_dev || !z_dev || O < 0 || M < 1 || M > IMM_MAX_MODES) return JSSP_ERR_INVALID_ARG;
  if (O == 0) return JSSP_OK;
  CK(cudaSetDevice(h->device));
  imm_update_kernel<<<(O + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*prm, mean_dev, cov_dev, mu_dev, z_dev, O, M);
  h->launches++;
  CK(cudaGetLastError());
  return JSSP_OK;
}

int32_t jssp_predict_modes(jssp_handle* h, const double* lanes_dev, int32_t L, int32_t P, const int32_t* lane_of_dev,
                           const double* s_start_dev, const double* speed_dev, double dt, int32_t O, int32_t M, int32_t Tt,
                           double* state_dev, uint8_t* valid_dev, void* stream) {
  if (!h || !lanes_dev || !lane_of_dev || !s_start_dev || !speed_dev || !state_dev || !valid_dev || L < 1 || P < 2 || O < 0 ||
      M < 1 || Tt < 1)
    return JSSP_ERR_INVALID_ARG;
  if (O == 0) return JSSP_OK;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  // lane tables: cum [L][P], len [L][P-1], yaw [L][P-1] in a handle-owned scratch buffer that only ever grows
  const size_t n_cum = (size_t)L * P, n_seg = (size_t)L * (P - 1);
  if (h->lane_scratch_cap < n_cum + 2 * n_seg) {
    if (h->d_lane_scratch) { CK(cudaFree(h->d_lane_scratch)); h->d_lane_scratch = nullptr; h->lane_scratch_cap = 0; }
    CK(cudaMalloc(&h->d_lane_scratch, (n_cum + 2 * n_seg) * sizeof(double)));
    h->lane_scratch_cap = n_cum + 2 * n_seg;
  }
  double *cum = h->d_lane_scratch, *len = cum + n_cum, *yaw = len + n_seg;
  lane_tables_kernel<<<L, 128, 0, st>>>(lanes_dev, P, cum, len, yaw);
  const int pt = Tt >= 192 ? 256 : (Tt >= 96 ? 128 : 64);
  if ((Tt + pt - 1) / pt > 65535) return JSSP_ERR_CAPACITY;
  predict_modes_kernel<<<dim3((unsigned)(O * M), (unsigned)((Tt + pt - 1) / pt)), pt, 0, st>>>(lanes_dev, L, P, lane_of_dev, s_start_dev, speed_dev, dt, O, M, Tt,
                                                                        cum, len, yaw, state_dev, valid_dev);
  h->launches += 2;
  CK(cudaGetLastError());
  CK(cudaGetLastError());
  return JSSP_OK;
}

}  // extern "C"

// =====================================================================================================================
// Batched, device-resident replay (include/jssp_b200.h, "batched replay")
// =====================================================================================================================
struct jssp_batch {
  jssp_config cfg;
  Tuning tune;
  int device = 0, P = 0, S = 0, C = 0;
  int W = 0;
  double* d_grid[10] = {nullptr};
  SeedGrids grids{};
  // strides (elements per scenario)
  size_t n_knot = 0, n_coef = 0, n_om = 0, n_ompad = 0, n_tt = 0;
  int blocks_thread = 0, blocks_warp = 0, max_blocks = 0;
  // everything jssp_batch_set_scenario uploads for one scenario (reference line, raw obstacle tables, ground truth, clock
  // tables) lives in ONE arena slice per scenario, packed in the order of the upload blob: one H2D copy per scenario
  unsigned char* d_arena = nullptr;
  size_t arena_cap = 0;             // bytes per scenario
  // per-scenario strided device arrays (written by the pack kernel / the replay)
  float4 *d_broad = nullptr, *d_narrow = nullptr;
  double *d_obs64 = nullptr, *d_ext64 = nullptr;
  float* d_probf = nullptr;
  double* d_seeds = nullptr;
  int* d_counts = nullptr;
  SeedWork seed_work{};
  size_t seed_smem = 0;
  double* d_blk_cost = nullptr;
  int* d_blk_idx = nullptr;
  unsigned int* d_ticket = nullptr;
  BestRecord* d_best = nullptr;
  double* d_best_traj = nullptr;
  EvalParams *d_slots = nullptr, *d_slots0 = nullptr;    // live records and their pristine copies (jssp_batch_reset)
  BatchSlot *d_bslots = nullptr, *d_bslots0 = nullptr;
  jssp_cycle_record* d_records = nullptr;
  std::vector<unsigned char> pristine_stale;   // slots whose pristine copy (d_slots0 ...) has not been taken yet
  // FOP baseline in the batch (jssp_batch_set_fop): per-scenario FopParams records instead of EvalParams
  bool fop = false;
  jssp_fop_in fop_grid{};
  FopParams *d_fslots = nullptr, *d_fslots0 = nullptr;
  double* d_fop_tab = nullptr;
  int* d_fop_int = nullptr;
  int fop_blocks = 0;
  size_t fop_smem = 0;
  // pinned staging ring of the uploads: host arrays are copied here, the H2D copies out of it are truly asynchronous
  unsigned char* h_stage = nullptr;
  size_t stage_cap = 0, stage_off = 0;
  // launch sizing: maxima over the scenarios set so far
  int max_K = 2, max_kpad = 2, max_OMpad = 0;
  std::vector<unsigned char> is_set;
#ifdef JSSP_PROFILE_PHASES
  unsigned long long* phase_ts = nullptr;
#endif
  std::vector<double*> d_poly;      // kinematics mode: the scenario's re-discretised reference line [n][6] (allocated on first use)
  std::vector<int> poly_cap;
  int max_smem_optin = 0;
  int sm_count = 148;
  int64_t launches = 0;
  char cuda_err[256] = {0};
};

namespace {
int fail_cuda_b(jssp_batch* b, cudaError_t e, const char* what) {
  if (b) snprintf(b->cuda_err, sizeof(b->cuda_err), "%s: %s", what, cudaGetErrorString(e));
  return JSSP_ERR_CUDA;
}
#define CKB(call)                                                   \
  do {                                                              \
    cudaError_t e__ = (call);                                       \
    if (e__ != cudaSuccess) return fail_cuda_b(b, e__, #call);      \
  } while (0)

// host -> device through the pinned staging ring: on return `src` may be released, the copy runs asynchronously on `st`
cudaError_t upload_staged(jssp_batch* b, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return cudaSuccess;
  const size_t need = (bytes + 15) & ~(size_t)15;
  if (need > b->stage_cap) {                                 // larger than the ring: plain (synchronous) copy
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
    return e != cudaSuccess ? e : cudaStreamSynchronize(st);
  }
  if (b->stage_off + need > b->stage_cap) {                  // wrap: everything staged so far must have left the ring
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    b->stage_off = 0;
  }
  unsigned char* h = b->h_stage + b->stage_off;
  b->stage_off += need;
  memcpy(h, src, bytes);
  return cudaMemcpyAsync(dst, h, bytes, cudaMemcpyHostToDevice, st);
}
}  // namespace

extern "C" {

const char* jssp_batch_last_cuda_error(const jssp_batch* b) { return b ? b->cuda_err : ""; }
int64_t jssp_batch_launch_count(const jssp_batch* b) { return b ? b->launches : 0; }

int32_t jssp_batch_destroy(jssp_batch* b) {
  if (!b) return JSSP_OK;
  cudaSetDevice(b->device);
  void* dev[] = {b->d_arena, b->d_broad, b->d_narrow, b->d_obs64, b->d_ext64, b->d_probf, b->d_seeds,
                 b->d_counts, b->d_blk_cost, b->d_blk_idx, b->d_ticket, b->d_best, b->d_best_traj, b->d_slots, b->d_bslots,
                 b->d_slots0, b->d_bslots0, b->d_records};
  for (void* q : dev) if (q) cudaFree(q);
  for (double* q : b->d_grid) if (q) cudaFree(q);
  if (b->h_stage) cudaFreeHost(b->h_stage);
  void* fopdev[] = {b->d_fslots, b->d_fslots0, b->d_fop_tab, b->d_fop_int};
  for (void* q : fopdev) if (q) cudaFree(q);
  for (double* q : b->d_poly) if (q) cudaFree(q);
  delete b;
  return JSSP_OK;
}

int32_t jssp_batch_create(const jssp_config* cfg, int32_t device, int32_t max_scenarios, int32_t max_cycles, jssp_batch** out) {
  if (!cfg || !out || max_scenarios < 1 || max_scenarios > 65535 || max_cycles < 1) return JSSP_ERR_INVALID_ARG;
  *out = nullptr;
  if (!config_ok(cfg)) return JSSP_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return JSSP_ERR_NO_DEVICE;
  }
  jssp_batch* b = new (std::nothrow) jssp_batch();
  if (!b) return JSSP_ERR_INVALID_ARG;
  b->cfg = *cfg; b->tune = Tuning::from_env(); b->device = device; b->S = max_scenarios; b->C = max_cycles; b->W = cfg->num_width;
  b->P = (int)(cfg->plan_horizon / cfg->delta_t) + 1;
  auto bail = [&](int code) { jssp_batch_destroy(b); return code; };
#define CKC(call)                                                                     \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) { fprintf(stderr, "jssp_batch_create: %s: %s\n", #call, cudaGetErrorString(e__)); return bail(JSSP_ERR_CUDA); } \
  } while (0)
  CKC(cudaSetDevice(device));
  int optin = 0;
  CKC(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  b->max_smem_optin = optin;   // lowered below to the tightest per-kernel limit
  {
    double** slots[10];
    for (int k = 0; k < 10; ++k) slots[k] = &b->d_grid[k];
    int n_time = 0;
    CKC(upload_grids(cfg, slots, &b->grids, &n_time));
    if (n_time != b->P) return bail(JSSP_ERR_INVALID_ARG);
  }
  const size_t S = (size_t)b->S;
  b->n_knot = ((size_t)cfg->max_knots + 2) & ~(size_t)1;   // even stride: every scenario's knots stay 16-byte aligned
  b->n_coef = (size_t)8 * ((cfg->max_knots + 1) & ~1);
  b->n_om = (size_t)(cfg->max_obstacle_modes > 0 ? cfg->max_obstacle_modes : 1);
  b->n_ompad = (b->n_om + 31) & ~(size_t)31;
  b->n_tt = (size_t)cfg->max_total_steps;
  b->blocks_thread = (int)(((size_t)cfg->max_seeds * b->W + BATCH_EVAL_THREADS - 1) / BATCH_EVAL_THREADS);
  b->blocks_warp = (int)(((size_t)cfg->max_seeds * b->W + SMALL_WARPS - 1) / SMALL_WARPS);
  b->max_blocks = std::max(b->blocks_thread, b->blocks_warp);
  {
    auto a16 = [](size_t n) { return (n + 15) & ~(size_t)15; };
    const size_t om_tt = b->n_om * b->n_tt;
    b->arena_cap = a16(b->n_knot * 8) + a16(b->n_coef * 8) + 2 * (a16(om_tt * 24) + a16(b->n_om * 16) + a16(om_tt)) + a16(b->n_om * 8) +
                   a16((size_t)(b->C + 1) * 8) + 2 * a16((size_t)(b->C + 1) * 4);
    CKC(cudaMalloc(&b->d_arena, S * b->arena_cap));
  }
  CKC(cudaMalloc(&b->d_broad, S * b->n_ompad * b->n_tt * sizeof(float4)));
  CKC(cudaMalloc(&b->d_narrow, S * b->n_om * b->n_tt * sizeof(float4)));
  CKC(cudaMalloc(&b->d_obs64, S * b->n_om * b->n_tt * 4 * sizeof(double)));
  CKC(cudaMalloc(&b->d_ext64, S * b->n_om * 2 * sizeof(double)));
  CKC(cudaMalloc(&b->d_probf, S * b->n_om * sizeof(float)));
  CKC(cudaMalloc(&b->d_seeds, S * (size_t)cfg->max_seeds * 10 * sizeof(double)));
  CKC(cudaMalloc(&b->d_counts, S * 4 * sizeof(int)));
  CKC(cudaMemset(b->d_counts, 0, S * 4 * sizeof(int)));
  b->seed_work = make_seed_work(b->grids, b->d_counts, b->d_seeds, cfg->max_seeds);
  b->seed_smem = seed_smem_bytes(b->seed_work);
  if (b->seed_smem > (size_t)b->max_smem_optin) return bail(JSSP_ERR_CAPACITY);
  CKC(allow_max_dynamic_smem(seed_enum_kernel<true>, optin, &b->max_smem_optin));
  CKC(cudaMalloc(&b->d_blk_cost, S * b->max_blocks * sizeof(double)));
  CKC(cudaMalloc(&b->d_blk_idx, S * b->max_blocks * sizeof(int)));
  CKC(cudaMalloc(&b->d_ticket, S * sizeof(unsigned int)));
  CKC(cudaMemset(b->d_ticket, 0, S * sizeof(unsigned int)));
  CKC(cudaMalloc(&b->d_best, S * sizeof(BestRecord)));
  CKC(cudaMemset(b->d_best, 0, S * sizeof(BestRecord)));
  CKC(cudaMalloc(&b->d_best_traj, S * (size_t)b->P * JSSP_TRAJ_COLS * sizeof(double)));
  CKC(cudaMalloc(&b->d_slots, S * sizeof(EvalParams)));
  CKC(cudaMemset(b->d_slots, 0, S * sizeof(EvalParams)));
  CKC(cudaMalloc(&b->d_bslots, S * sizeof(BatchSlot)));
  CKC(cudaMemset(b->d_bslots, 0, S * sizeof(BatchSlot)));
  CKC(cudaMalloc(&b->d_slots0, S * sizeof(EvalParams)));
  CKC(cudaMalloc(&b->d_bslots0, S * sizeof(BatchSlot)));
  CKC(cudaMalloc(&b->d_records, S * (size_t)b->C * sizeof(jssp_cycle_record)));
  b->stage_cap = (size_t)32 << 20;
  CKC(cudaMallocHost(&b->h_stage, b->stage_cap));
  b->is_set.assign(S, 0);
  b->pristine_stale.assign(S, 0);
  b->d_poly.assign(S, nullptr);
  b->poly_cap.assign(S, 0);
  CKC(allow_max_dynamic_smem(evaluate_batch_kernel<0>, optin, &b->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_batch_kernel<1>, optin, &b->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_small_batch_kernel, optin, &b->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_group_batch_kernel<0>, optin, &b->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_group_batch_kernel<1>, optin, &b->max_smem_optin));
  CKC(allow_max_dynamic_smem(replay_persistent_kernel<512>, optin, &b->max_smem_optin));
  CKC(allow_max_dynamic_smem(replay_persistent_kernel<1024>, optin, &b->max_smem_optin));
  CKC(allow_max_dynamic_smem(replay_persistent_kernel<512, true>, optin, &b->max_smem_optin));
  CKC(cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, device));
  CKC(allow_max_dynamic_smem(evaluate_tiled_batch_kernel<0>, optin, &b->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_tiled_batch_kernel<1>, optin, &b->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_tiled_chunked_batch_kernel<0>, optin, &b->max_smem_optin));
  CKC(allow_max_dynamic_smem(evaluate_tiled_chunked_batch_kernel<1>, optin, &b->max_smem_optin));
  CKC(cudaDeviceSynchronize());
#undef CKC
  *out = b;
  return JSSP_OK;
}

int32_t jssp_batch_set_fop(jssp_batch* b, const jssp_fop_in* grid) {
  if (!b || !grid) return JSSP_ERR_INVALID_ARG;
  const jssp_config& c = b->cfg;
  if (grid->num_width != c.num_width || grid->num_speed < 1 || grid->num_t < 1 || !(grid->min_t > 0) || !(grid->max_t >= grid->min_t))
    return JSSP_ERR_INVALID_ARG;
  if (grid->max_t > c.plan_horizon + 1e-9) return JSSP_ERR_CAPACITY;
  CKB(cudaSetDevice(b->device));
  const int NT = grid->num_t, NV = grid->num_speed;
  const long long N = (long long)b->W * NT * NV;
  b->fop_blocks = (int)((N + EVAL_THREADS - 1) / EVAL_THREADS);
  if (b->fop_blocks > b->max_blocks) return JSSP_ERR_CAPACITY;      // per-scenario block records follow max_seeds * num_width
  std::vector<double> tab = np_linspace(grid->min_t, grid->max_t, NT);
  std::vector<int> ns((size_t)NT);
  for (int i = 0; i < NT; ++i) {
    ns[i] = (int)std::ceil((tab[i] - 0.0) / c.delta_t);
    if (ns[i] < 1 || ns[i] > b->P) return JSSP_ERR_CAPACITY;
  }
  const std::vector<double> sp = np_linspace(grid->lowest_speed, grid->highest_speed, NV);
  tab.insert(tab.end(), sp.begin(), sp.end());
  void* old[] = {b->d_fslots, b->d_fslots0, b->d_fop_tab, b->d_fop_int};
  for (void* q : old) if (q) cudaFree(q);
  b->d_fslots = b->d_fslots0 = nullptr; b->d_fop_tab = nullptr; b->d_fop_int = nullptr;
  CKB(cudaMalloc(&b->d_fslots, (size_t)b->S * sizeof(FopParams)));
  CKB(cudaMalloc(&b->d_fslots0, (size_t)b->S * sizeof(FopParams)));
  CKB(cudaMalloc(&b->d_fop_tab, tab.size() * sizeof(double)));
  CKB(cudaMalloc(&b->d_fop_int, ns.size() * sizeof(int)));
  CKB(cudaMemcpy(b->d_fop_tab, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
  CKB(cudaMemcpy(b->d_fop_int, ns.data(), ns.size() * sizeof(int), cudaMemcpyHostToDevice));
  int optin = 0;
  CKB(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, b->device));
  CKB(allow_max_dynamic_smem(evaluate_fop_batch_kernel, optin, &b->max_smem_optin));
  CKB(allow_max_dynamic_smem(fop_select_export_batch_kernel, optin, &b->max_smem_optin));
  b->fop_smem = fop_smem_layout(c.max_knots, (c.max_knots + 1) & ~1, b->P, b->W, NT).total;   // sized for the knot capacity
  if (b->fop_smem > (size_t)b->max_smem_optin) return JSSP_ERR_CAPACITY;
  b->fop_grid = *grid;
  b->fop = true;
  std::fill(b->is_set.begin(), b->is_set.end(), 0);      // scenarios must be (re)loaded in FOP mode
  return JSSP_OK;
}

int32_t jssp_batch_set_scenario(jssp_batch* b, int32_t slot, const jssp_scenario_in* in, void* stream) {
  if (!b || !in || slot < 0 || slot >= b->S) return JSSP_ERR_INVALID_ARG;
  if (!in->knot_s || !in->coef || in->K < 2 || in->O < 0 || in->M < 1 || in->Tt < 1 || !in->cycle_time || !in->cycle_now || !in->cycle_end_step ||
      in->n_cycles < 1)
    return JSSP_ERR_INVALID_ARG;
  if (in->O > 0 && (!in->state || !in->valid || !in->size || !in->prob)) return JSSP_ERR_INVALID_ARG;
  if (in->gt_state && (!in->gt_valid || !in->gt_size || in->gt_O < 0)) return JSSP_ERR_INVALID_ARG;
  if (!in->gt_state && in->M != 1 && in->O > 0) return JSSP_ERR_INVALID_ARG;   // the ground truth must be given for multi-modal predictions
  if (in->ego_update_mode != 0 && in->ego_update_mode != 1) return JSSP_ERR_INVALID_ARG;
  if (in->ego_update_mode == 1 && (!in->polyline || in->n_polyline < 1 || !(in->pp_wheel_base > 0.0))) return JSSP_ERR_INVALID_ARG;
  const int OM = in->O * in->M;
  const int gtO = in->gt_state ? in->gt_O : in->O;
  if (in->K > b->cfg.max_knots || OM > b->cfg.max_obstacle_modes || gtO > b->cfg.max_obstacle_modes || in->Tt > b->cfg.max_total_steps ||
      in->n_cycles > b->C)
    return JSSP_ERR_CAPACITY;
  for (int i = 1; i < in->K; ++i) if (!(in->knot_s[i] > in->knot_s[i - 1])) return JSSP_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CKB(cudaSetDevice(b->device));
  const size_t s = (size_t)slot;
  const double ego_l = in->ego_length > 0 ? in->ego_length : b->cfg.ego_length, ego_w = in->ego_width > 0 ? in->ego_width : b->cfg.ego_width;
  const int K = in->K, kpad = (K + 1) & ~1, Tt = in->Tt, OMpad = (OM + 31) & ~31;
  // one upload blob per scenario: [knots | coefficients (structure of arrays) | obstacle states | sizes | probabilities |
  // ground-truth states | sizes | clock tables | validity masks], every part 16-byte aligned, assembled in the pinned ring
  const size_t nc = (size_t)in->n_cycles + 1;
  auto a16 = [](size_t n) { return (n + 15) & ~(size_t)15; };
  const bool own_gt = in->gt_state != nullptr && gtO > 0;
  size_t off = 0;
  const size_t o_knot = off; off += a16((size_t)kpad * 8);
  const size_t o_coef = off; off += a16((size_t)8 * kpad * 8);
  const size_t o_state = off; off += a16((size_t)OM * Tt * 24);
  const size_t o_size = off; off += a16((size_t)in->O * 16);
  const size_t o_prob = off; off += a16((size_t)OM * 8);
  const size_t o_gstate = off; off += own_gt ? a16((size_t)gtO * Tt * 24) : 0;
  const size_t o_gsize = off; off += own_gt ? a16((size_t)gtO * 16) : 0;
  const size_t o_ctime = off; off += a16(nc * 8);
  const size_t o_cnow = off; off += a16(nc * 4);
  const size_t o_cend = off; off += a16(nc * 4);
  const size_t o_valid = off; off += a16((size_t)OM * Tt);
  const size_t o_gvalid = off; off += own_gt ? a16((size_t)gtO * Tt) : 0;
  const size_t blob = off;
  if (blob > b->arena_cap) return JSSP_ERR_CAPACITY;
  std::vector<unsigned char> big;                            // a scenario larger than the ring: pageable copy, synchronised below
  unsigned char* hb;
  if (blob > b->stage_cap) { big.resize(blob); hb = big.data(); }
  else {
    if (b->stage_off + blob > b->stage_cap) {                // wrap: everything staged so far must have left the ring
      CKB(cudaStreamSynchronize(st));
      b->stage_off = 0;
    }
    hb = b->h_stage + b->stage_off;
    b->stage_off += blob;
  }
  unsigned char* arena = b->d_arena + s * b->arena_cap;
  {
    double* hk = reinterpret_cast<double*>(hb + o_knot);
    memcpy(hk, in->knot_s, (size_t)K * 8);
    if (kpad > K) hk[K] = 0.0;
    double* hc = reinterpret_cast<double*>(hb + o_coef);
    for (int c8 = 0; c8 < 8; ++c8) {
      for (int i = 0; i < K - 1; ++i) hc[(size_t)c8 * kpad + i] = in->coef[(size_t)i * 8 + c8];
      for (int i = K - 1; i < kpad; ++i) hc[(size_t)c8 * kpad + i] = 0.0;
    }
    if (OM > 0) {
      memcpy(hb + o_state, in->state, (size_t)OM * Tt * 24);
      memcpy(hb + o_size, in->size, (size_t)in->O * 16);
      memcpy(hb + o_prob, in->prob, (size_t)OM * 8);
      memcpy(hb + o_valid, in->valid, (size_t)OM * Tt);
    }
    if (own_gt) {
      memcpy(hb + o_gstate, in->gt_state, (size_t)gtO * Tt * 24);
      memcpy(hb + o_gsize, in->gt_size, (size_t)gtO * 16);
      memcpy(hb + o_gvalid, in->gt_valid, (size_t)gtO * Tt);
    }
    memcpy(hb + o_ctime, in->cycle_time, nc * 8);
    memcpy(hb + o_cnow, in->cycle_now, nc * 4);
    memcpy(hb + o_cend, in->cycle_end_step, nc * 4);
  }
  CKB(cudaMemcpyAsync(arena, hb, blob, cudaMemcpyHostToDevice, st));
  if (!big.empty()) CKB(cudaStreamSynchronize(st));
  double* knot = reinterpret_cast<double*>(arena + o_knot);
  double* coef = reinterpret_cast<double*>(arena + o_coef);
  const double ox = in->coef[0], oy = in->coef[4];
  double* raw_state = reinterpret_cast<double*>(arena + o_state);
  uint8_t* raw_valid = arena + o_valid;
  double* raw_size = reinterpret_cast<double*>(arena + o_size);
  double* raw_prob = reinterpret_cast<double*>(arena + o_prob);
  float4* broad = b->d_broad + s * b->n_ompad * b->n_tt;
  float4* narrow = b->d_narrow + s * b->n_om * b->n_tt;
  double* obs64 = b->d_obs64 + s * b->n_om * b->n_tt * 4;
  double* ext64 = b->d_ext64 + s * b->n_om * 2;
  float* probf = b->d_probf + s * b->n_om;
  if (OM > 0) {
    const double r_ego = std::sqrt(ego_l * ego_l + ego_w * ego_w) / 2.0;
    const int total = OMpad * Tt;
    pack_obstacles_kernel<<<(total * PACK_LANES + 255) / 256, 256, 0, st>>>(raw_state, raw_valid, raw_size, raw_prob, OM, OMpad, in->M, Tt, ox, oy,
                                                               b->cfg.obstacle_inflation, r_ego, broad, narrow, obs64, ext64, probf, nullptr,
                                                               nullptr, nullptr, 0, 0.0, 0.0);
    b->launches++;
    CKB(cudaGetLastError());
  }
  const double *gt_state = raw_state, *gt_size = raw_size;
  const uint8_t* gt_valid = raw_valid;
  if (in->gt_state) {
    gt_state = reinterpret_cast<const double*>(arena + o_gstate); gt_valid = arena + o_gvalid;
    gt_size = reinterpret_cast<const double*>(arena + o_gsize);
  }
  double* ctime = reinterpret_cast<double*>(arena + o_ctime);
  int* cnow = reinterpret_cast<int*>(arena + o_cnow);
  int* cend = reinterpret_cast<int*>(arena + o_cend);

  jssp_config c = b->cfg;
  c.ego_length = ego_l; c.ego_width = ego_w;
  EvalParams p{};
  fill_constants(p, c, b->P);
  p.s0 = in->frenet0[0]; p.v0 = in->frenet0[1]; p.a0 = in->frenet0[2];
  const std::vector<double> tg = default_targets(c, b->W);
  for (int w = 0; w < b->W; ++w) quintic_coeffs(in->frenet0[3], in->frenet0[4], in->frenet0[5], tg[w], 0.0, 0.0, c.plan_horizon, p.quintic[w]);
  p.W = b->W; p.now = in->cycle_now[0]; p.tcap = in->final_time_step - p.now; p.n_dict = in->n_dict;
  p.time = b->d_grid[0]; p.knot = knot; p.coef = coef; p.K = K; p.kpad = kpad; p.ox = ox; p.oy = oy;
  p.seeds = b->d_seeds + s * (size_t)c.max_seeds * 10;
  p.n_seeds_ptr = b->d_counts + s * 4; p.n_seeds_fixed = 0; p.seed_cap = c.max_seeds;
  p.broad = broad; p.narrow = narrow; p.obs64 = obs64; p.ext64 = ext64;
  p.OM = OM; p.OMpad = OMpad; p.Tt = Tt;
  int rows = 0;
  if (p.tcap > 0) rows = (std::min(p.P, p.tcap) + c.check_res - 1) / c.check_res;
  p.win_rows = OM > 0 ? rows : 0;
  p.smem_obs = 1; p.sint = nullptr; p.probf = probf; p.list_cap = 0;
  p.feasible = nullptr; p.collision = nullptr; p.cost = nullptr;
  p.select_only = b->tune.full_walk ? 0 : 1;    // the replay only needs the selected trajectory (tools/: A/B override)
  p.blk_cost = b->d_blk_cost + s * b->max_blocks; p.blk_idx = b->d_blk_idx + s * b->max_blocks;
  p.ticket = b->d_ticket + s; p.best = b->d_best + s;
  p.best_traj = b->d_best_traj + s * (size_t)b->P * JSSP_TRAJ_COLS;
  p.batch = b->d_bslots + s; p.self = b->d_slots + s;
#ifdef JSSP_PROFILE_PHASES
  p.phase_ts = b->phase_ts;
#endif
  BatchSlot q{};
  q.live = 1; q.cycle = 0; q.n_cycles_cap = in->n_cycles; q.final_step = in->final_time_step;
  q.gt_O = gtO; q.gt_Tt = Tt;
  q.end = -1.0; q.max_t = in->max_t; q.horizon = c.plan_horizon;
  q.goal_x = in->goal_x; q.goal_y = in->goal_y; q.goal_cos = std::cos(in->goal_yaw); q.goal_sin = std::sin(in->goal_yaw);
  q.ego_l = c.ego_length; q.ego_w = c.ego_width;
  q.ego_x = in->ego0[0]; q.ego_y = in->ego0[1]; q.ego_yaw = in->ego0[2]; q.ego_v = in->ego0[3];
  q.dt = c.delta_t;
  const double inf = std::numeric_limits<double>::infinity();
  q.bx_min = in->has_bounds ? in->bounds[0] : -inf; q.bx_max = in->has_bounds ? in->bounds[1] : inf;
  q.by_min = in->has_bounds ? in->bounds[2] : -inf; q.by_max = in->has_bounds ? in->bounds[3] : inf;
  for (int w = 0; w < b->W; ++w) q.targets[w] = tg[w];
  q.cycle_time = ctime; q.cycle_now = cnow; q.cycle_end_step = cend;
  q.gt_state = gt_state; q.gt_valid = gt_valid; q.gt_size = gt_size;
  q.records = b->d_records + s * (size_t)b->C;
  if (in->ego_update_mode == 1) {
    if (b->poly_cap[s] < in->n_polyline) {
      if (b->d_poly[s]) { CKB(cudaFree(b->d_poly[s])); b->d_poly[s] = nullptr; b->poly_cap[s] = 0; }
      CKB(cudaMalloc(&b->d_poly[s], (size_t)in->n_polyline * 6 * sizeof(double)));
      b->poly_cap[s] = in->n_polyline;
    }
    CKB(upload_staged(b, b->d_poly[s], in->polyline, (size_t)in->n_polyline * 6 * sizeof(double), st));
    q.kin = 1; q.n_poly = in->n_polyline; q.poly = b->d_poly[s];
    q.pp_kv = in->pp_kv; q.pp_ld0 = in->pp_ld0; q.pp_wb = in->pp_wheel_base;
  }
  if (b->fop) {
    const jssp_fop_in& g = b->fop_grid;
    q.is_fop = 1;
    p.n_seeds_ptr = nullptr; p.n_seeds_fixed = g.num_t * g.num_speed; p.seeds = nullptr;
    p.smem_obs = 0; p.win_rows = 0;
    p.self = &b->d_fslots[s].base;
    FopParams fp{};
    fp.base = p;
    fp.num_t = g.num_t; fp.num_speed = g.num_speed;
    fp.horizon = b->d_fop_tab; fp.speed = b->d_fop_tab + g.num_t; fp.n_samples = b->d_fop_int;
    fp.d0 = in->frenet0[3]; fp.d_d0 = in->frenet0[4]; fp.d_dd0 = in->frenet0[5];
    for (int w = 0; w < b->W; ++w) fp.targets[w] = tg[w];
    fp.dt = c.delta_t; fp.max_accel = g.max_accel; fp.max_jerk = g.max_jerk; fp.t_min = g.min_t; fp.t_max = g.max_t;
    CKB(upload_staged(b, b->d_fslots + s, &fp, sizeof(fp), st));
  }
  CKB(upload_staged(b, b->d_slots + s, &p, sizeof(p), st));
  CKB(upload_staged(b, b->d_bslots + s, &q, sizeof(q), st));
  b->pristine_stale[s] = 1;     // the pristine copies (jssp_batch_reset) and the ticket are taken care of before the next run / reset
  // everything went through the staging ring: soa, p, q and the caller's arrays may be released, the copies are in flight
  b->max_K = std::max(b->max_K, K); b->max_kpad = std::max(b->max_kpad, kpad); b->max_OMpad = std::max(b->max_OMpad, OMpad);
  b->is_set[s] = 1;
  return JSSP_OK;
}

// Freshly loaded slots: copy their records to the pristine arrays and clear their tickets, one device-to-device copy per
// contiguous run of slots (a whole batch loaded in order = one copy of each array).
static int32_t snapshot_fresh_slots(jssp_batch* b, cudaStream_t st) {
  int s = 0;
  while (s < b->S) {
    if (!b->pristine_stale[s]) { ++s; continue; }
    int e = s;
    while (e < b->S && b->pristine_stale[e]) b->pristine_stale[e++] = 0;
    const size_t n = (size_t)(e - s);
    CKB(cudaMemcpyAsync(b->d_slots0 + s, b->d_slots + s, n * sizeof(EvalParams), cudaMemcpyDeviceToDevice, st));
    CKB(cudaMemcpyAsync(b->d_bslots0 + s, b->d_bslots + s, n * sizeof(BatchSlot), cudaMemcpyDeviceToDevice, st));
    if (b->fop) CKB(cudaMemcpyAsync(b->d_fslots0 + s, b->d_fslots + s, n * sizeof(FopParams), cudaMemcpyDeviceToDevice, st));
    CKB(cudaMemsetAsync(b->d_ticket + s, 0, n * sizeof(unsigned int), st));
    s = e;
  }
  return JSSP_OK;
}

int32_t jssp_batch_reset(jssp_batch* b, int32_t n_scenarios, void* stream) {
  if (!b || n_scenarios < 1 || n_scenarios > b->S) return JSSP_ERR_INVALID_ARG;
  for (int s = 0; s < n_scenarios; ++s) if (!b->is_set[s]) return JSSP_ERR_NOT_READY;
  cudaStream_t st = (cudaStream_t)stream;
  CKB(cudaSetDevice(b->device));
  const size_t n = (size_t)n_scenarios;
  { const int32_t rc = snapshot_fresh_slots(b, st); if (rc != JSSP_OK) return rc; }
  CKB(cudaMemcpyAsync(b->d_slots, b->d_slots0, n * sizeof(EvalParams), cudaMemcpyDeviceToDevice, st));
  CKB(cudaMemcpyAsync(b->d_bslots, b->d_bslots0, n * sizeof(BatchSlot), cudaMemcpyDeviceToDevice, st));
  if (b->fop) CKB(cudaMemcpyAsync(b->d_fslots, b->d_fslots0, n * sizeof(FopParams), cudaMemcpyDeviceToDevice, st));
  CKB(cudaMemsetAsync(b->d_ticket, 0, n * sizeof(unsigned int), st));
  return JSSP_OK;
}

int32_t jssp_batch_run(jssp_batch* b, int32_t n_scenarios, int32_t n_cycles, int32_t flags, void* stream) {
  if (!b || n_scenarios < 1 || n_scenarios > b->S || n_cycles < 0) return JSSP_ERR_INVALID_ARG;
  for (int s = 0; s < n_scenarios; ++s) if (!b->is_set[s]) return JSSP_ERR_NOT_READY;
  cudaStream_t st = (cudaStream_t)stream;
  CKB(cudaSetDevice(b->device));
  { const int32_t rc = snapshot_fresh_slots(b, st); if (rc != JSSP_OK) return rc; }
  const bool pdl = !b->tune.no_pdl;     // tools/ only: plain stream order instead of chained launches
  if (b->fop) {                       // FOP baseline: no seed enumeration, evaluation + (selection, export, hand-off) per step
    for (int c = 0; c < n_cycles; ++c) {
      CKB(launch_chained(evaluate_fop_batch_kernel, dim3(b->fop_blocks, n_scenarios), EVAL_THREADS, b->fop_smem, st, pdl, (const FopParams*)b->d_fslots));
      CKB(launch_chained(fop_select_export_batch_kernel, dim3(n_scenarios), EVAL_THREADS, b->fop_smem, st, pdl, (const FopParams*)b->d_fslots, b->fop_blocks));
      b->launches += 2;
    }
    CKB(cudaGetLastError());
    return JSSP_OK;
  }
  const int max_rows = (b->P + b->cfg.check_res - 1) / b->cfg.check_res;
  // JSSP_FLAG_PERSISTENT: the persistent replay kernel - one CTA per scenario, every cycle of the call in ONE launch.  Few scenarios
  // (at most one per SM) get 512 threads each, more get 256 so that two CTAs share an SM.
  // JSSP_FLAG_BEST_FIRST on a batch whose every cycle qualifies (select-only slots, w_risk >= 0): the kernel variant without the
  // exhaustive passes - 512 threads per CTA (16 warps checking candidates: the sweep ends with its heaviest scenario) up to three
  // scenarios per SM, 256-thread CTAs, two per SM, above that or when the caller runs several sub-batches at once (JSSP_FLAG_SMALL_CTAS)
  if ((flags & JSSP_FLAG_BEST_FIRST) && !(flags & JSSP_FLAG_LOCKSTEP) && !b->tune.full_walk && b->cfg.w_risk >= 0.0) {
    int threads = (n_scenarios <= 3 * b->sm_count && !(flags & JSSP_FLAG_SMALL_CTAS)) ? REPLAY_MAX_THREADS : 256;
    if (b->tune.seed_threads == 256 || b->tune.seed_threads == 512) threads = b->tune.seed_threads;      // tools/ only
    const int rows = b->max_OMpad > 0 ? max_rows : 0;
    const int bf_keys = b->cfg.max_seeds * b->W;
    ReplaySmem RL = replay_smem_layout(b->max_kpad, b->P, b->W, rows, b->max_OMpad, b->seed_smem, threads, -1, bf_keys);
    if (RL.total > (size_t)b->max_smem_optin && threads > 256) {
      threads = 256;
      RL = replay_smem_layout(b->max_kpad, b->P, b->W, rows, b->max_OMpad, b->seed_smem, threads, -1, bf_keys);
    }
    if (RL.total <= (size_t)b->max_smem_optin) {
      if (n_cycles > 0) {
        replay_persistent_kernel<512, true><<<dim3((unsigned)n_scenarios, 1, 1), threads, RL.total, st>>>(b->grids, b->d_slots, b->seed_work, n_cycles, rows, b->max_OMpad, b->max_kpad, -1, bf_keys);
        b->launches += 1;
        CKB(cudaGetLastError());
      }
      return JSSP_OK;
    }
  }
  if ((flags & (JSSP_FLAG_PERSISTENT | JSSP_FLAG_BEST_FIRST)) && !(flags & JSSP_FLAG_LOCKSTEP)) {
    int threads = n_scenarios <= b->sm_count ? REPLAY_MAX_THREADS : 256;
    if (b->tune.seed_threads == 256 || b->tune.seed_threads == 512 || b->tune.seed_threads == 1024) threads = b->tune.seed_threads;      // tools/ only
    const int rows = b->max_OMpad > 0 ? max_rows : 0;
    // tile length: the default (~16 samples), shortened until the warps' tiles fit next to the tables
    int tile_rounds = 0;
    // JSSP_FLAG_BEST_FIRST: best-first selection inside the persistent kernel (jssp_bestfirst.cuh); its keys and per-warp scratch share
    // the region of the tiles, which then only serve cycles the best-first rule does not cover and may be as short as needed
    const int bf_keys = (flags & JSSP_FLAG_BEST_FIRST) ? b->cfg.max_seeds * b->W : 0;
    auto layout = [&](int thr, int tr) { return replay_smem_layout(b->max_kpad, b->P, b->W, rows, b->max_OMpad, b->seed_smem, thr, tr, bf_keys); };
    ReplaySmem RL = layout(threads, tile_rounds);
    // persistent kernel: as long a tile as fits (<= ~16 samples); with 256-thread CTAs two of them share an SM
    const size_t budget = threads <= 256 ? std::min((size_t)b->max_smem_optin, (size_t)(228 * 1024 / 2 - 10 * 1024)) : (size_t)b->max_smem_optin;
    for (int r = bf_keys > 0 ? 1 : tile_rounds_wanted(b->W); r >= 1; --r) {
      tile_rounds = r;
      RL = layout(threads, tile_rounds);
      if (RL.total <= budget) break;
    }
    for (int r = tile_rounds - 1; RL.total > (size_t)b->max_smem_optin && r >= 1; --r) {
      tile_rounds = r;
      RL = layout(threads, tile_rounds);
    }
    if (RL.total > (size_t)b->max_smem_optin && threads > 256) {
      threads = 256; tile_rounds = bf_keys > 0 ? 1 : 0;
      RL = layout(threads, tile_rounds);
    }
    if (RL.total <= (size_t)b->max_smem_optin) {
      if (n_cycles > 0) {
        if (threads > 512) replay_persistent_kernel<1024><<<dim3((unsigned)n_scenarios, 1, 1), threads, RL.total, st>>>(b->grids, b->d_slots, b->seed_work, n_cycles, rows, b->max_OMpad, b->max_kpad, tile_rounds, bf_keys);
        else replay_persistent_kernel<512><<<dim3((unsigned)n_scenarios, 1, 1), threads, RL.total, st>>>(b->grids, b->d_slots, b->seed_work, n_cycles, rows, b->max_OMpad, b->max_kpad, tile_rounds, bf_keys);
        b->launches += 1;
        CKB(cudaGetLastError());
      }
      return JSSP_OK;
    }
  }
  // the evaluation kernel: one warp per candidate while the whole batch is small (latency), one thread per candidate otherwise
  const SmemLayout Ls = eval_smem_layout(0, b->max_OMpad, 0, b->max_K, b->max_kpad, b->P, b->W);
  const size_t small_bytes = Ls.total + small_scratch_bytes(b->P);
  bool small = (long long)n_scenarios * b->cfg.max_seeds * b->W <= 65536 && small_bytes <= (size_t)b->max_smem_optin;
  if (flags & (JSSP_FLAG_FORCE_THREAD | JSSP_FLAG_FORCE_GROUP | JSSP_FLAG_FORCE_TILED)) small = false;
  if ((flags & JSSP_FLAG_FORCE_WARP) && small_bytes <= (size_t)b->max_smem_optin) small = true;
  // thread per candidate for the default 3 lateral targets (a scenario's ~150 seeds fill the group kernel's 40-seed blocks and
  // 30-of-32 lanes poorly: measured 49.0 vs 53.9 ms on the 800-scenario sweep), the group kernel from 4 targets on
  const bool group = !small && !(flags & JSSP_FLAG_FORCE_THREAD) && ((flags & JSSP_FLAG_FORCE_GROUP) || b->W >= 4);
  // 128-thread blocks in the batch: a scenario has only a few hundred candidates, so smaller blocks leave fewer idle
  // lanes in its last block and spread its work over more SMs (measured: 800-scenario sweep 50.3 -> 47.7 ms)
  const int bthreads = BATCH_EVAL_THREADS;
  const int spb = group_seeds_per_block(b->W, bthreads);
  const int blocks_group = (b->cfg.max_seeds + spb - 1) / spb;
  const int blocks_thread = (int)(((size_t)b->cfg.max_seeds * b->W + bthreads - 1) / bthreads);
  int mode = b->max_OMpad > 0 ? 1 : 0;
  SmemLayout L = eval_smem_layout(max_rows, b->max_OMpad, mode, b->max_K, b->max_kpad, b->P, b->W);
  if (mode == 1 && L.total > (size_t)b->max_smem_optin) { mode = 0; L = eval_smem_layout(0, b->max_OMpad, 0, b->max_K, b->max_kpad, b->P, b->W); }
  if (!small && L.total > (size_t)b->max_smem_optin) return JSSP_ERR_CAPACITY;
  const size_t tiled_total = ((L.total + 15) & ~(size_t)15) + tile_scratch_bytes(b->W, bthreads);     // default tile length (slot records carry tile_rounds = 0)
  // the tiled walk when a rank holds few scenarios (latency of a plan cycle matters: measured 11.8 vs 13.1 ms for 100 scenarios x <= 100
  // cycles), the thread-per-candidate walk for large batches (throughput: 42.2 vs 48.4 ms for 800 scenarios)
  const bool tiled_auto = !(flags & (JSSP_FLAG_FORCE_THREAD | JSSP_FLAG_FORCE_GROUP | JSSP_FLAG_FORCE_WARP)) && n_scenarios <= 256;
  const bool tiled = !small && ((flags & JSSP_FLAG_FORCE_TILED) || tiled_auto) && tiled_total <= (size_t)b->max_smem_optin;
  // the chunked tiled walk: two lane groups per seed split the horizon (half the serial chain, twice the lanes) - while the lanes of the
  // whole batch still fit the SMs in one wave
  int chunks = b->tune.chunks > 0 ? b->tune.chunks : 2;
  chunks = std::max(1, std::min(chunks, 32 / b->W));
  const size_t chunked_total = ((L.total + 15) & ~(size_t)15) + tile_scratch_bytes(b->W, bthreads, 0, chunks, b->P);
  const bool chunk_fits = chunks > 1 && mode != 2 && chunked_total <= (size_t)b->max_smem_optin;
  const int spb_chunked = (bthreads / 32) * ((32 / b->W) / chunks);
  const int blocks_chunked = (b->cfg.max_seeds + spb_chunked - 1) / spb_chunked;
  const bool chunked = !small && chunk_fits && blocks_chunked <= b->max_blocks && (flags & JSSP_FLAG_CHUNKED);
  // few scenarios: clusters of up to 8 CTAs x 1024 threads answer with the lowest latency; many: one 512-thread CTA each
  int seed_cluster = 1;
  while (seed_cluster < SEED_MAX_CLUSTER && n_scenarios * seed_cluster * 2 <= 148) seed_cluster *= 2;
  int seed_threads = n_scenarios >= 296 ? 512 : SEED_THREADS;
  if (b->tune.seed_cluster > 0) seed_cluster = std::max(1, std::min(SEED_MAX_CLUSTER, b->tune.seed_cluster));    // tools/ only
  if (b->tune.seed_threads > 0) seed_threads = b->tune.seed_threads == 512 ? 512 : SEED_THREADS;
  for (int c = 0; c < n_cycles; ++c) {
    CKB(launch_seed_enum<true>(b->grids, SeedState{0, 0, 0}, b->d_slots, b->seed_work, n_scenarios, seed_cluster, seed_threads, b->seed_smem, st, pdl));
    const EvalParams* slots = b->d_slots;
    if (small) CKB(launch_chained(evaluate_small_batch_kernel, dim3(n_scenarios, std::min(b->blocks_warp, 128)), SMALL_WARPS * 32, small_bytes, st, pdl, slots));
    else if (chunked && mode == 1) CKB(launch_chained(evaluate_tiled_chunked_batch_kernel<1>, dim3(n_scenarios, blocks_chunked), bthreads, chunked_total, st, pdl, slots, chunks));
    else if (chunked) CKB(launch_chained(evaluate_tiled_chunked_batch_kernel<0>, dim3(n_scenarios, blocks_chunked), bthreads, chunked_total, st, pdl, slots, chunks));
    else if (tiled && mode == 1) CKB(launch_chained(evaluate_tiled_batch_kernel<1>, dim3(n_scenarios, blocks_group), bthreads, tiled_total, st, pdl, slots));
    else if (tiled) CKB(launch_chained(evaluate_tiled_batch_kernel<0>, dim3(n_scenarios, blocks_group), bthreads, tiled_total, st, pdl, slots));
    else if (group && mode == 1) CKB(launch_chained(evaluate_group_batch_kernel<1>, dim3(n_scenarios, blocks_group), bthreads, L.total, st, pdl, slots));
    else if (group) CKB(launch_chained(evaluate_group_batch_kernel<0>, dim3(n_scenarios, blocks_group), bthreads, L.total, st, pdl, slots));
    else if (mode == 1) CKB(launch_chained(evaluate_batch_kernel<1>, dim3(n_scenarios, blocks_thread), bthreads, L.total, st, pdl, slots));
    else CKB(launch_chained(evaluate_batch_kernel<0>, dim3(n_scenarios, blocks_thread), bthreads, L.total, st, pdl, slots));
    b->launches += 2;
  }
  CKB(cudaGetLastError());
  return JSSP_OK;
}

int32_t jssp_batch_fetch(jssp_batch* b, int32_t n_scenarios, jssp_scenario_result* results_host, jssp_cycle_record* records_host,
                         void* stream) {
  if (!b || !results_host || n_scenarios < 1 || n_scenarios > b->S) return JSSP_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CKB(cudaSetDevice(b->device));
  std::vector<BatchSlot> q((size_t)n_scenarios);
  CKB(cudaMemcpyAsync(q.data(), b->d_bslots, q.size() * sizeof(BatchSlot), cudaMemcpyDeviceToHost, st));
  if (records_host)
    CKB(cudaMemcpyAsync(records_host, b->d_records, (size_t)n_scenarios * b->C * sizeof(jssp_cycle_record), cudaMemcpyDeviceToHost, st));
  CKB(cudaStreamSynchronize(st));
  bool overflow = false;
  for (int s = 0; s < n_scenarios; ++s) {
    results_host[s].end = q[s].end; results_host[s].cycles = q[s].cycle; results_host[s].seed_overflow = q[s].seed_overflow;
    overflow |= q[s].seed_overflow != 0;
  }
  return overflow ? JSSP_ERR_SEED_OVERFLOW : JSSP_OK;
}

int32_t jssp_batch_best_traj(jssp_batch* b, int32_t slot, double* out, void* stream) {
  if (!b || !out || slot < 0 || slot >= b->S) return JSSP_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CKB(cudaSetDevice(b->device));
  CKB(cudaMemcpyAsync(out, b->d_best_traj + (size_t)slot * b->P * JSSP_TRAJ_COLS, (size_t)b->P * JSSP_TRAJ_COLS * sizeof(double),
                      cudaMemcpyDefault, st));
  CKB(cudaStreamSynchronize(st));
  return JSSP_OK;
}

#ifdef JSSP_PROFILE_PHASES
/* persistent replay kernel: per-stage nanoseconds of scenario 0 (see jssp_replay.cuh) */
int32_t jssp_debug_replay_phases(unsigned long long* host8) { return cudaMemcpyFromSymbol(host8, g_replay_ns, sizeof(g_replay_ns)) == cudaSuccess ? JSSP_OK : JSSP_ERR_CUDA; }
int32_t jssp_debug_bestfirst(unsigned long long* host8) { return cudaMemcpyFromSymbol(host8, g_bestfirst_ns, sizeof(g_bestfirst_ns)) == cudaSuccess ? JSSP_OK : JSSP_ERR_CUDA; }
/* same for the batched replay (tools/batch_phase_probe.py): [scenarios * blocks][8]; call before jssp_batch_set_scenario */
int32_t jssp_batch_debug_phase_buffer(jssp_batch* b, unsigned long long* dev) { if (!b) return JSSP_ERR_INVALID_ARG; b->phase_ts = dev; return JSSP_OK; }
#endif

}  // extern "C"
