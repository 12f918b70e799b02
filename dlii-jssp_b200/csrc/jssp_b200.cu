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

#include "jssp_kernels.cuh"

using namespace jssp;

struct jssp_handle {
  jssp_config cfg;
  int device = 0;
  int P = 0;   // samples per trajectory = n_steps + 1
  // sampling grids (device) + lengths
  double *d_time = nullptr, *d_j1 = nullptr, *d_t1 = nullptr, *d_t2 = nullptr, *d_j3 = nullptr, *d_t3 = nullptr, *d_t4 = nullptr,
         *d_jneg = nullptr, *d_jpos = nullptr, *d_ts = nullptr;
  SeedGrids grids{};
  uint32_t main_total = 0, stop_total = 0;
  // seeds
  double* d_seeds = nullptr;
  uint32_t *d_list_main = nullptr, *d_list_stop = nullptr;
  int* d_counts = nullptr;
  int* h_counts = nullptr;   // pinned
  bool seeds_ready = false;
  // reference line
  double *d_knot = nullptr, *d_coef = nullptr;
  int K = 0, kpad = 0;
  double ox = 0, oy = 0;
  // obstacles
  double *d_raw_state = nullptr, *d_raw_size = nullptr, *d_raw_prob = nullptr;
  uint8_t* d_raw_valid = nullptr;
  const double *src_state = nullptr, *src_size = nullptr, *src_prob = nullptr;   // what the pack kernel reads
  const uint8_t* src_valid = nullptr;
  float4 *d_broad = nullptr, *d_narrow = nullptr;
  double *d_obs64 = nullptr, *d_ext64 = nullptr;
  float2* d_sint = nullptr;       // reachable arc-length interval per obstacle state
  float* d_probf = nullptr;
  double* d_knot_xy = nullptr;
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

int pack_if_dirty(jssp_handle* h, cudaStream_t st) {
  if (!h->obs_dirty) return JSSP_OK;
  const int OM = h->O * h->M;
  const int OMpad = (OM + 31) & ~31;
  const int total = OMpad * h->Tt;
  if (total > 0) {
    const double r_ego = std::sqrt(h->cfg.ego_length * h->cfg.ego_length + h->cfg.ego_width * h->cfg.ego_width) / 2.0;
    pack_obstacles_kernel<<<(total + 255) / 256, 256, 0, st>>>(h->src_state, h->src_valid, h->src_size, h->src_prob, OM, OMpad, h->M, h->Tt,
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

}  // namespace

extern "C" {

int32_t jssp_abi_version(void) { return JSSP_ABI_VERSION; }

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
  if (cfg->plan_horizon <= 0 || cfg->delta_t <= 0 || cfg->num_jerk < 4 || cfg->num_width < 1 || cfg->num_width > JSSP_MAX_WIDTHS ||
      cfg->check_res < 1 || cfg->max_seeds < 1 || cfg->max_knots < 2 || cfg->max_obstacle_modes < 0 || cfg->max_total_steps < 1)
    return JSSP_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return JSSP_ERR_NO_DEVICE;
  }
  jssp_handle* h = new (std::nothrow) jssp_handle();
  if (!h) return JSSP_ERR_INVALID_ARG;
  h->cfg = *cfg;
  h->device = device;
  h->P = (int)(cfg->plan_horizon / cfg->delta_t) + 1;   // len(time_array), polyvt_sampling.py:44
  auto bail = [&](int code) { jssp_destroy(h); return code; };
#define CKC(call)                                                                     \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) { fprintf(stderr, "jssp_create: %s: %s\n", #call, cudaGetErrorString(e__)); return bail(JSSP_ERR_CUDA); } \
  } while (0)
  CKC(cudaSetDevice(device));
  CKC(cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  h->max_smem_optin -= 1024;   // static shared memory of the kernels + slack
  HostGrids g = make_grids(cfg->plan_horizon, cfg->highest_jerk, cfg->num_jerk, cfg->delta_t);
  std::vector<double> jneg, jpos;
  for (double j : g.js) { if (!(j > 0)) jneg.push_back(j); if (!(j < 0)) jpos.push_back(j); }   // polyvt_sampling.py:275,287
  CKC(upload(&h->d_time, g.time)); CKC(upload(&h->d_j1, g.j1)); CKC(upload(&h->d_t1, g.t1)); CKC(upload(&h->d_t2, g.t2));
  CKC(upload(&h->d_j3, g.j3)); CKC(upload(&h->d_t3, g.t3)); CKC(upload(&h->d_t4, g.t4));
  CKC(upload(&h->d_jneg, jneg)); CKC(upload(&h->d_jpos, jpos)); CKC(upload(&h->d_ts, g.ts));
  SeedGrids& sg = h->grids;
  sg.j1 = h->d_j1; sg.t1 = h->d_t1; sg.t2 = h->d_t2; sg.j3 = h->d_j3; sg.t3 = h->d_t3; sg.t4 = h->d_t4;
  sg.n_j1 = (int)g.j1.size(); sg.n_t1 = (int)g.t1.size(); sg.n_t2 = (int)g.t2.size(); sg.n_j3 = (int)g.j3.size();
  sg.n_t3 = (int)g.t3.size(); sg.n_t4 = (int)g.t4.size();
  sg.jneg = h->d_jneg; sg.jpos = h->d_jpos; sg.ts = h->d_ts;
  sg.n_jneg = (int)jneg.size(); sg.n_jpos = (int)jpos.size(); sg.n_ts = (int)g.ts.size();
  sg.total_time = cfg->plan_horizon;
  sg.acc_max = cfg->lon_accel_max - 1e-9; sg.vel_max = cfg->speed_max - 1e-9; sg.v_max = cfg->speed_max;
  sg.jmin5 = -cfg->highest_jerk * 0.3; sg.jmax5 = cfg->highest_jerk * 0.3;
  sg.lim2 = cfg->plan_horizon - 0.8 * 3; sg.lim3 = cfg->plan_horizon - 1.0 * 2; sg.lim4 = cfg->plan_horizon - 1.5;
  h->main_total = (uint32_t)((size_t)sg.n_j1 * sg.n_t1 * sg.n_t2 * sg.n_j3 * sg.n_t3 * sg.n_t4);
  h->stop_total = (uint32_t)((size_t)sg.n_jneg * sg.n_ts * sg.n_jpos * sg.n_ts);
  if ((int)g.time.size() != h->P) return bail(JSSP_ERR_INVALID_ARG);

  CKC(cudaMalloc(&h->d_seeds, (size_t)cfg->max_seeds * 10 * sizeof(double)));
  CKC(cudaMalloc(&h->d_list_main, (size_t)SORT_CAP * sizeof(uint32_t)));
  CKC(cudaMalloc(&h->d_list_stop, (size_t)SORT_CAP * sizeof(uint32_t)));
  h->kpad = (cfg->max_knots + 1) & ~1;   // capacity; the per-scenario stride is set in jssp_set_refline
  CKC(cudaMalloc(&h->d_knot, (size_t)cfg->max_knots * sizeof(double)));
  CKC(cudaMalloc(&h->d_coef, (size_t)8 * h->kpad * sizeof(double)));
  const size_t om = (size_t)(cfg->max_obstacle_modes > 0 ? cfg->max_obstacle_modes : 1), tt = (size_t)cfg->max_total_steps;
  CKC(cudaMalloc(&h->d_raw_state, om * tt * 3 * sizeof(double)));
  CKC(cudaMalloc(&h->d_raw_valid, om * tt));
  CKC(cudaMalloc(&h->d_raw_size, om * 2 * sizeof(double)));
  CKC(cudaMalloc(&h->d_raw_prob, om * sizeof(double)));
  CKC(cudaMalloc(&h->d_broad, ((om + 31) & ~(size_t)31) * tt * sizeof(float4)));
  CKC(cudaMalloc(&h->d_narrow, om * tt * sizeof(float4)));
  CKC(cudaMalloc(&h->d_obs64, om * tt * 4 * sizeof(double)));
  CKC(cudaMalloc(&h->d_ext64, om * 2 * sizeof(double)));
  CKC(cudaMalloc(&h->d_sint, ((om + 31) & ~(size_t)31) * tt * sizeof(float2)));
  CKC(cudaMalloc(&h->d_probf, om * sizeof(float)));
  CKC(cudaMalloc(&h->d_knot_xy, (size_t)cfg->max_knots * 2 * sizeof(double)));
  h->max_blocks = (int)(((size_t)cfg->max_seeds * JSSP_MAX_WIDTHS + SMALL_WARPS - 1) / SMALL_WARPS);   // warp-per-candidate grid
  CKC(cudaMalloc(&h->d_blk_cost, (size_t)h->max_blocks * sizeof(double)));
  CKC(cudaMalloc(&h->d_blk_idx, (size_t)h->max_blocks * sizeof(int)));
  CKC(cudaMalloc(&h->d_ticket, sizeof(unsigned int)));
  CKC(cudaMemset(h->d_ticket, 0, sizeof(unsigned int)));
  // one contiguous result block so that a cycle's results come back with a single copy:
  // [BestRecord, padded to 48 B][seed counts 16 B][selected trajectory P x 16 float64]
  static_assert(sizeof(BestRecord) <= 48, "result block layout");
  h->result_bytes = 64 + (size_t)h->P * JSSP_TRAJ_COLS * sizeof(double);
  CKC(cudaMalloc(&h->d_result, h->result_bytes));
  CKC(cudaMemset(h->d_result, 0, h->result_bytes));
  CKC(cudaMallocHost(&h->h_result, h->result_bytes));
  h->d_best = reinterpret_cast<BestRecord*>(h->d_result);
  h->d_counts = reinterpret_cast<int*>(h->d_result + 48);
  h->d_best_traj = reinterpret_cast<double*>(h->d_result + 64);
  h->h_best = reinterpret_cast<BestRecord*>(h->h_result);
  h->h_counts = reinterpret_cast<int*>(h->h_result + 48);
  h->h_best_traj = reinterpret_cast<double*>(h->h_result + 64);
  CKC(cudaFuncSetAttribute(evaluate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
  CKC(cudaFuncSetAttribute(evaluate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
  CKC(cudaFuncSetAttribute(evaluate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
  CKC(cudaFuncSetAttribute(evaluate_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
  CKC(cudaFuncSetAttribute(dump_states_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin));
  CKC(cudaDeviceSynchronize());
#undef CKC
  *out = h;
  return JSSP_OK;
}

int32_t jssp_destroy(jssp_handle* h) {
  if (!h) return JSSP_OK;
  cudaSetDevice(h->device);
  void* dev[] = {h->d_time, h->d_j1, h->d_t1, h->d_t2, h->d_j3, h->d_t3, h->d_t4, h->d_jneg, h->d_jpos, h->d_ts, h->d_seeds,
                 h->d_list_main, h->d_list_stop, h->d_knot, h->d_coef, h->d_raw_state, h->d_raw_valid, h->d_raw_size,
                 h->d_raw_prob, h->d_broad, h->d_narrow, h->d_obs64, h->d_ext64, h->d_sint, h->d_probf, h->d_knot_xy, h->d_blk_cost, h->d_blk_idx, h->d_ticket, h->d_result};
  for (void* p : dev) if (p) cudaFree(p);
  if (h->h_result) cudaFreeHost(h->h_result);
  delete h;
  return JSSP_OK;
}

int32_t jssp_set_refline(jssp_handle* h, const double* knot_s, const double* coef, int32_t K, void* stream) {
  if (!h || !knot_s || !coef || K < 2) return JSSP_ERR_INVALID_ARG;
  if (K > h->cfg.max_knots) return JSSP_ERR_CAPACITY;
  for (int i = 1; i < K; ++i) if (!(knot_s[i] > knot_s[i - 1])) return JSSP_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
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
  const size_t om = (size_t)O * M;
  if (om > 0) {
    CK(cudaMemcpyAsync(h->d_raw_state, state, om * Tt * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_raw_valid, valid, om * Tt, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_raw_size, size, (size_t)O * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_raw_prob, prob, om * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  h->src_state = h->d_raw_state; h->src_valid = h->d_raw_valid; h->src_size = h->d_raw_size; h->src_prob = h->d_raw_prob;
  return pack_if_dirty(h, st);
}

int32_t jssp_set_obstacles_dev(jssp_handle* h, const double* state_dev, const uint8_t* valid_dev, const double* size_dev,
                               const double* prob_dev, int32_t O, int32_t M, int32_t Tt, int32_t n_dict, void* stream) {
  if (!h) return JSSP_ERR_INVALID_ARG;
  if (O > 0 && (!state_dev || !valid_dev || !size_dev || !prob_dev)) return JSSP_ERR_INVALID_ARG;
  int rc = set_obstacles_common(h, O, M, Tt, n_dict);
  if (rc != JSSP_OK) return rc;
  CK(cudaSetDevice(h->device));
  h->src_state = state_dev; h->src_valid = valid_dev; h->src_size = size_dev; h->src_prob = prob_dev;
  return pack_if_dirty(h, (cudaStream_t)stream);
}

int32_t jssp_enumerate_seeds(jssp_handle* h, double s0, double v0, double a0, void* stream, int32_t* n_main, int32_t* n_stop) {
  if (!h) return JSSP_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  CK(cudaMemsetAsync(h->d_counts, 0, 4 * sizeof(int), st));
  SeedState ss{s0, v0, a0};
  seed_flag_main<<<(h->main_total + 255) / 256, 256, 0, st>>>(h->grids, ss, h->main_total, h->d_list_main, h->d_counts, SORT_CAP);
  seed_flag_stop<<<(h->stop_total + 255) / 256, 256, 0, st>>>(h->grids, ss, h->stop_total, h->d_list_stop, h->d_counts, SORT_CAP);
  seed_sort_emit<<<1, 1024, 0, st>>>(h->grids, ss, h->d_list_main, h->d_list_stop, h->d_counts, h->d_seeds, SORT_CAP, h->cfg.max_seeds);
  h->launches += 3;
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
  CK(cudaMemcpyAsync(h->h_result, h->d_result, want_traj ? h->result_bytes : 64, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  out->n_candidates = h->h_best->n_candidates;
  out->n_seeds = h->h_best->n_seeds;
  out->best_idx = (int32_t)h->h_best->idx;
  out->best_cost = h->h_best->cost;
  out->best_n_pts = h->h_best->n_pts;
  if (want_traj && out->best_idx >= 0) memcpy(out->best_traj_host, h->h_best_traj, (size_t)h->P * JSSP_TRAJ_COLS * sizeof(double));
  // enumerated seeds: surface an overflow detected on the device
  return (h->last.n_seeds_ptr && h->h_counts[3]) ? JSSP_ERR_SEED_OVERFLOW : JSSP_OK;
}

int32_t jssp_evaluate(jssp_handle* h, const jssp_cycle_in* in, void* stream, jssp_cycle_out* out) {
  if (!h || !in || !out) return JSSP_ERR_INVALID_ARG;
  if (h->K < 2 || !h->obs_set) return JSSP_ERR_NOT_READY;
  const bool ext = in->seeds_dev != nullptr;
  if (!ext && (in->flags & JSSP_FLAG_ENUMERATE)) {
    const int rc0 = jssp_enumerate_seeds(h, in->s0, in->v0, in->a0, stream, nullptr, nullptr);
    if (rc0 != JSSP_OK) return rc0;
  }
  if (!ext && !h->seeds_ready) return JSSP_ERR_NOT_READY;
  if (ext && (in->n_seeds < 0 || in->n_seeds > h->cfg.max_seeds)) return JSSP_ERR_CAPACITY;
  const int W = in->n_widths > 0 ? in->n_widths : h->cfg.num_width;
  if (W < 1 || W > JSSP_MAX_WIDTHS) return JSSP_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaSetDevice(h->device));
  int rc = pack_if_dirty(h, st);
  if (rc != JSSP_OK) return rc;

  EvalParams p{};
  const jssp_config& c = h->cfg;
  p.s0 = in->s0; p.v0 = in->v0; p.a0 = in->a0;
  // lateral targets: linspace(-(max_road_width - w)/2, +(max_road_width - w)/2, num_width) (jerk_space_sampling_planner.py:100-104)
  std::vector<double> tg;
  if (in->n_widths > 0) tg.assign(in->targets, in->targets + W);
  else if (W <= 1) tg.assign(1, 0.0);
  else { const double sw = c.max_road_width - c.ego_width; tg = np_linspace(-sw / 2, sw / 2, W); }
  for (int w = 0; w < W; ++w) quintic_coeffs(in->d0, in->d_d0, in->d_dd0, tg[w], 0.0, 0.0, c.plan_horizon, p.quintic[w]);
  p.W = W; p.P = h->P; p.now = in->time_step_now; p.tcap = in->final_time_step - in->time_step_now;
  p.n_dict = h->n_dict; p.check_res = c.check_res;
  p.kmax = c.max_curvature; p.ea = c.ego_length / 2.0; p.eb = c.ego_width / 2.0; p.p_min = c.p_min; p.w_risk = c.w_risk;
  p.vmax = c.speed_max; p.inv_vmax = 1.0 / c.speed_max; p.max_distance = c.max_distance; p.half_w = c.max_road_width / 2.0;
  p.thr_lon_a = c.thr_lon_accel; p.thr_lon_j = c.thr_lon_jerk; p.thr_lat_a = c.thr_lat_accel; p.thr_lat_j = c.thr_lat_jerk;
  memcpy(p.cw, c.cost_weights, sizeof(p.cw));
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
  p.blk_cost = h->d_blk_cost; p.blk_idx = h->d_blk_idx; p.ticket = h->d_ticket; p.best = h->d_best;
  p.best_traj = (in->flags & JSSP_FLAG_NO_EXPORT) ? nullptr : h->d_best_traj;

  const long long n_max = (long long)W * (ext ? in->n_seeds : c.max_seeds);
  // small cycles (the default sampling grid: ~10^3 candidates) run one warp per candidate: latency, not throughput, matters
  const SmemLayout Ls = eval_smem_layout(0, p.OMpad, 0, p.K, p.kpad, p.P, p.W);
  const size_t small_bytes = Ls.total + small_scratch_bytes(p.P);
  bool small = n_max <= 16384 && small_bytes <= (size_t)h->max_smem_optin;
  if (in->flags & JSSP_FLAG_FORCE_THREAD) small = false;
  if ((in->flags & JSSP_FLAG_FORCE_WARP) && small_bytes <= (size_t)h->max_smem_optin) small = true;
  int blocks = small ? (int)((n_max + SMALL_WARPS - 1) / SMALL_WARPS) : (int)((n_max + EVAL_THREADS - 1) / EVAL_THREADS);
  if (blocks < 1) blocks = 1;
  if (blocks > h->max_blocks) return JSSP_ERR_CAPACITY;
  if (small) { p.smem_obs = 0; p.win_rows = 0; p.list_cap = 0; evaluate_small_kernel<<<blocks, SMALL_WARPS * 32, small_bytes, st>>>(p); }
  else if (mode == 2) evaluate_kernel<2><<<blocks, EVAL_THREADS, L.total, st>>>(p);
  else if (mode == 1) evaluate_kernel<1><<<blocks, EVAL_THREADS, L.total, st>>>(p);
  else evaluate_kernel<0><<<blocks, EVAL_THREADS, L.total, st>>>(p);
  h->launches++;
  CK(cudaGetLastError());
  EvalParams q = p;
  q.smem_obs = 0; q.win_rows = 0; q.list_cap = 0;
  const SmemLayout L2 = eval_smem_layout(0, q.OMpad, 0, q.K, q.kpad, q.P, q.W);
  if (out->states_dev) {
    dump_states_kernel<<<blocks, EVAL_THREADS, L2.total, st>>>(q, out->states_dev);
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

int64_t jssp_launch_count(const jssp_handle* h) { return h ? h->launches : 0; }

int32_t jssp_measure_fma_peak(int32_t device, int32_t bits, double* tflops) {
  if (!tflops || (bits != 32 && bits != 64)) return JSSP_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return JSSP_ERR_NO_DEVICE; }
  return bits == 32 ? measure_fma<float>(device, tflops) : measure_fma<double>(device, tflops);
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
  predict_modes_kernel<<<(O * M + 127) / 128, 128, 0, (cudaStream_t)stream>>>(lanes_dev, L, P, lane_of_dev, s_start_dev, speed_dev, dt,
                                                                               O, M, Tt, state_dev, valid_dev);
  h->launches++;
  CK(cudaGetLastError());
  return JSSP_OK;
}

}  // extern "C"
