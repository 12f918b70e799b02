// jssp_fop.cuh - K6: FOP baseline planner through the same stages (frenet_optimal_planner.py)
#pragma once
#include "jssp_eval.cuh"
#include <cstddef>

namespace jssp {

// ------------------------------------------------------------------------------------------------------------------
// K6: the FOP baseline planner's plan cycle through the same stages (SURVEY.md section 8 row f3;
// frenet_optimal_planner.py:97-144,146-183,185-225,249-299,328-333).  Candidate n = (w * num_t + it) * num_speed + iv
// (the reference's loop order): lateral quintic to target w over horizon T_it, longitudinal QUARTIC from (s, s_d, s_dd) to
// (speed_iv, 0) over T_it, samples t = arange(0, T_it, dt) (per-candidate length), Frenet->Cartesian as for JSSP, feasibility
// = |s_dd| <= max_accel and |s_ddd| <= max_jerk at every sample (the curvature test is commented out in the reference),
// rectangle collision every 2nd sample (the handle is created with obstacle_inflation = 0, :280), cost = the JSSP terms over
// the candidate's own samples + w_time * (1 - (T - t_min) / (t_max - t_min)), best = LAST minimum (`min_cost >= cost`).
// One thread per candidate; QuarticPolynomial / the cost class are absent from the reference tree (builder-specified,
// oracle/fop_oracle.py).
// ------------------------------------------------------------------------------------------------------------------
struct FopParams {
  EvalParams base;            // constants, tables, obstacle pointers, outputs; W = num_width, P = longest sample count
  int num_t, num_speed;
  const double* horizon;      // [num_t]  T_it = linspace(min_t, max_t, num_t)
  const int* n_samples;       // [num_t]  len(arange(0, T_it, dt))
  const double* speed;        // [num_speed] target speeds = linspace(lowest, highest, num_speed)
  double d0, d_d0, d_dd0;
  double targets[JSSP_MAX_WIDTHS];
  double dt, max_accel, max_jerk, t_min, t_max;
};

struct QuarticLon {           // QuarticPolynomial(xs, vxs, axs, vxe, axe, T), closed form of the PythonRobotics 2x2 system
  double a0, a1, a2, a3, a4;
  __device__ __forceinline__ void init(double xs, double vxs, double axs, double vxe, double axe, double T) {
    a0 = xs; a1 = vxs; a2 = axs / 2.0;
    const double b0 = vxe - a1 - 2.0 * a2 * T, b1 = axe - 2.0 * a2;
    a3 = b0 / (T * T) - b1 / (3.0 * T);
    a4 = b1 / (4.0 * T * T) - b0 / (2.0 * T * T * T);
  }
  __device__ __forceinline__ void eval(double t, double& s, double& sd, double& sdd, double& sddd) const {
    const double t2 = t * t, t3 = t2 * t, t4 = t3 * t;
    s = a0 + a1 * t + a2 * t2 + a3 * t3 + a4 * t4;
    sd = a1 + 2 * a2 * t + 3 * a3 * t2 + 4 * a4 * t3;
    sdd = 2 * a2 + 6 * a3 * t + 12 * a4 * t2;
    sddd = 6 * a3 + 24 * a4 * t;
  }
};

struct FopSink : EvalSink {   // same collision / cost accumulation; accel and jerk limits instead of the curvature test
  double max_accel, max_jerk;
  __device__ __forceinline__ FopSink(const EvalParams& p_, const BlockTables& tb_, double ma, double mj) : EvalSink(p_, tb_), max_accel(ma), max_jerk(mj) {}
  __device__ __forceinline__ void sample(int i, double t, double s, double sd, double sdd, double sddd) {
    EvalSink::sample(i, t, s, sd, sdd, sddd);
    if (fabs(sdd) > max_accel || fabs(sddd) > max_jerk) curv_fail = true;                  // :211-216
  }
  __device__ __forceinline__ void segment(int i, double x, double y, double dx, double dy, double ds, double c, double sn) {
    collide(i, x, y, c, sn);
    pdx = dx; pdy = dy; pds = ds; pc = c; psn = sn;
  }
};

struct FopSmem { size_t off_knot, off_coef, off_time, off_quint, off_lat, off_latcost, total; };
__host__ __device__ inline FopSmem fop_smem_layout(int K, int kpad, int P, int W, int num_t) {
  FopSmem L;
  size_t o = 16;
  L.off_knot = o; o += (size_t)kpad * 8;
  L.off_coef = o; o += (size_t)8 * kpad * 8;
  L.off_time = o; o += (size_t)((P + 1) & ~1) * 8;
  L.off_quint = o; o += (size_t)W * num_t * 6 * 8;
  L.off_lat = o; o += (size_t)W * num_t * P * 4 * 8;
  L.off_latcost = o; o += (size_t)W * num_t * 4 * 8;
  L.total = o;
  return L;
}

// tables of a FOP block: spline, time samples i * dt, one lateral quintic per (width, horizon) sampled over its own length
__device__ __forceinline__ void fop_stage(const FopParams& fp, unsigned char* smem, BlockTables& tb, const double*& latcost, const double*& quint) {
  const EvalParams& p = fp.base;
  const FopSmem L = fop_smem_layout(p.K, p.kpad, p.P, p.W, fp.num_t);
  double* sm_knot = reinterpret_cast<double*>(smem + L.off_knot);
  double* sm_coef = reinterpret_cast<double*>(smem + L.off_coef);
  double* sm_time = reinterpret_cast<double*>(smem + L.off_time);
  double* sm_quint = reinterpret_cast<double*>(smem + L.off_quint);
  double* sm_lat = reinterpret_cast<double*>(smem + L.off_lat);
  double* sm_latcost = reinterpret_cast<double*>(smem + L.off_latcost);
  const int tid = threadIdx.x, nt = blockDim.x, P = p.P, nq = p.W * fp.num_t;
  for (int e = tid; e < p.K; e += nt) sm_knot[e] = p.knot[e];
  for (int e = tid; e < 8 * p.kpad; e += nt) sm_coef[e] = p.coef[e];
  for (int e = tid; e < P; e += nt) sm_time[e] = 0.0 + (double)e * fp.dt;                  // np.arange(0.0, T, dt)
  for (int e = tid; e < nq; e += nt) {
    const int w = e / fp.num_t, it = e - w * fp.num_t;
    quintic_coeffs(fp.d0, fp.d_d0, fp.d_dd0, fp.targets[w], 0.0, 0.0, fp.horizon[it], sm_quint + (size_t)e * 6);
  }
  __syncthreads();
  for (int e = tid; e < nq * P; e += nt) {
    const int q = e / P, i = e - q * P;
    double* o = sm_lat + (size_t)e * 4;
    if (i < fp.n_samples[q % fp.num_t]) quintic_eval(sm_quint + (size_t)q * 6, sm_time[i], o[0], o[1], o[2], o[3]);
    else { o[0] = o[1] = o[2] = o[3] = 0.0; }
  }
  __syncthreads();
  for (int q = tid >> 5; q < nq; q += nt >> 5) {           // lateral cost sums over the candidate's own samples (stage_tables' order)
    const int n = fp.n_samples[q % fp.num_t];
    double s0 = 0, s1 = 0, s2 = 0;
    for (int i = tid & 31; i < n; i += 32) {
      const double* o = sm_lat + ((size_t)q * P + i) * 4;
      const double e1 = fmax(fabs(o[2]) - p.thr_lat_a, 0.0), e2 = fmax(fabs(o[3]) - p.thr_lat_j, 0.0);
      s0 = s0 + o[0] * o[0]; s1 = s1 + e1 * e1; s2 = s2 + e2 * e2;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s0 = s0 + __shfl_down_sync(0xffffffffu, s0, off);
      s1 = s1 + __shfl_down_sync(0xffffffffu, s1, off);
      s2 = s2 + __shfl_down_sync(0xffffffffu, s2, off);
    }
    if ((tid & 31) == 0) { sm_latcost[q * 4] = s0; sm_latcost[q * 4 + 1] = s1; sm_latcost[q * 4 + 2] = s2; }
  }
  __syncthreads();
  tb.broad = p.broad; tb.broad_row0 = p.now; tb.broad_stride = p.check_res;
  tb.sp.knot = sm_knot; tb.sp.coef = sm_coef; tb.sp.K = p.K; tb.sp.kpad = p.kpad;
  tb.sp.inv_h = (double)(p.K - 1) / sm_knot[p.K - 1];
  tb.time = sm_time; tb.lat = sm_lat;
  tb.list = nullptr; tb.row_off = nullptr; tb.row_cnt = nullptr; tb.use_lists = false;
  latcost = sm_latcost; quint = sm_quint;
}

// the lateral start state of the next cycle lives next to the EvalParams record inside the scenario's FopParams
__device__ __forceinline__ void fop_store_lateral_start(EvalParams* self, double d, double d_d, double d_dd) {
  FopParams* f = reinterpret_cast<FopParams*>(self);           // `base` is the first member
  f->d0 = d; f->d_d0 = d_d; f->d_dd0 = d_dd;
}

__device__ __forceinline__ void evaluate_fop_body(const FopParams& fp, unsigned char* smem, double* red_cost, int* red_idx) {
  const EvalParams& p = fp.base;
  const int N = p.W * fp.num_t * fp.num_speed;
  BlockTables tb;
  const double *latcost, *quint;
  fop_stage(fp, smem, tb, latcost, quint);
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double my_cost = INFINITY;
  int my_key = -1;                                          // N - 1 - n: the LAST minimum wins (:328-333)
  if (n < N) {
    const int iv = n % fp.num_speed, q = n / fp.num_speed, it = q % fp.num_t;
    const int Pc = fp.n_samples[it];
    const double T = fp.horizon[it];
    QuarticLon lon;
    lon.init(p.s0, p.v0, p.a0, fp.speed[iv], 0.0, T);                                      // :125
    FopSink sink(p, tb, fp.max_accel, fp.max_jerk);
    const int seg0 = spline_bisect(tb.sp, fmin(fmax(p.s0, tb.sp.knot[0]), tb.sp.knot[p.K - 1]));
    walk_core(tb, lon, tb.lat + (size_t)q * p.P * 4, Pc, seg0, sink);
    const double Pd = (double)Pc;
    const double* lc = latcost + q * 4;
    double cost = p.cw[1] * (1.0 - (sink.s_last - sink.s_first) / p.max_distance);
    cost = cost + p.cw[2] * (lc[0] / Pd / (p.half_w * p.half_w));
    cost = cost + p.cw[3] * (lc[1] / Pd);
    cost = cost + p.cw[4] * (lc[2] / Pd);
    cost = cost + p.cw[5] * (sink.sum_a / Pd);
    cost = cost + p.cw[6] * (sink.sum_j / Pd);
    cost = cost + p.cw[7] * (sink.sum_v / Pd);
    if (fp.t_max > fp.t_min) cost = cost + p.cw[0] * (1.0 - (T - fp.t_min) / (fp.t_max - fp.t_min));
    if (p.feasible) p.feasible[n] = sink.curv_fail ? 0 : 1;
    if (p.collision) p.collision[n] = sink.collided ? 1 : 0;
    if (p.cost) p.cost[n] = (float)cost;
    if (!sink.curv_fail && !sink.collided && isfinite(cost)) { my_cost = cost; my_key = N - 1 - n; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oc = __shfl_down_sync(0xffffffffu, my_cost, o);
    const int oi = __shfl_down_sync(0xffffffffu, my_key, o);
    if (lex_less(oc, oi, my_cost, my_key)) { my_cost = oc; my_key = oi; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red_cost[wid] = my_cost; red_idx[wid] = my_key; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
      if (lex_less(red_cost[k], red_idx[k], my_cost, my_key)) { my_cost = red_cost[k]; my_key = red_idx[k]; }
    p.blk_cost[blockIdx.x] = my_cost;
    p.blk_idx[blockIdx.x] = my_key;
  }
}

__global__ void __launch_bounds__(EVAL_THREADS, EVAL_MIN_BLOCKS) evaluate_fop_kernel(const __grid_constant__ FopParams fp) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  evaluate_fop_body(fp, smem, red_cost, red_idx);
}

__device__ __forceinline__ void load_fop_slot(FopParams* dst, const FopParams* src) {
  static_assert(sizeof(FopParams) % 8 == 0 && offsetof(FopParams, base) == 0, "FopParams is copied as 8-byte words, base first");
  const unsigned long long* s = reinterpret_cast<const unsigned long long*>(src);
  unsigned long long* d = reinterpret_cast<unsigned long long*>(dst);
  for (int i = threadIdx.x; i < (int)(sizeof(FopParams) / 8); i += blockDim.x) d[i] = s[i];
  __syncthreads();
}

// batched replay with the FOP baseline: blockIdx.y = scenario
__global__ void __launch_bounds__(EVAL_THREADS, EVAL_MIN_BLOCKS) evaluate_fop_batch_kernel(const FopParams* __restrict__ fslots) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(16) FopParams sfp;
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  pdl_launch_dependents(); pdl_wait();
  if (!fslots[blockIdx.y].base.batch->live) return;
  load_fop_slot(&sfp, fslots + blockIdx.y);
  evaluate_fop_body(sfp, smem, red_cost, red_idx);
}

// one block: reduce the per-block winners, write the selection record and export the selected trajectory; returns the
// selected candidate (-1 = none) to every thread
__device__ __forceinline__ int fop_select_export_body(const FopParams& fp, int n_blocks, unsigned char* smem, double* red_cost, int* red_idx,
                                                      int* sm_best_p, int* sm_npts_p, double (*ex)[EXPORT_MAX_P]) {
  int& sm_best = *sm_best_p;
  int& sm_npts = *sm_npts_p;
  const EvalParams& p = fp.base;
  const int N = p.W * fp.num_t * fp.num_speed;
  double my_cost = INFINITY;
  int my_key = -1;
  for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
    const double c = p.blk_cost[b];
    const int k = p.blk_idx[b];
    if (lex_less(c, k, my_cost, my_key)) { my_cost = c; my_key = k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oc = __shfl_down_sync(0xffffffffu, my_cost, o);
    const int oi = __shfl_down_sync(0xffffffffu, my_key, o);
    if (lex_less(oc, oi, my_cost, my_key)) { my_cost = oc; my_key = oi; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red_cost[wid] = my_cost; red_idx[wid] = my_key; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
      if (lex_less(red_cost[k], red_idx[k], my_cost, my_key)) { my_cost = red_cost[k]; my_key = red_idx[k]; }
    const int best = my_key >= 0 ? N - 1 - my_key : -1;
    BestRecord r;
    r.key = orderable_bits(best >= 0 ? my_cost : INFINITY);
    r.idx = best; r.cost = my_cost; r.n_pts = 0; r.n_candidates = N; r.n_seeds = fp.num_t * fp.num_speed; r.pad = 0;
    *p.best = r;
    sm_best = best;
  }
  __syncthreads();
  const int best = sm_best;
  if (best < 0 || p.best_traj == nullptr || p.P > EXPORT_MAX_P) return best;
  BlockTables tb;
  const double *latcost, *quint;
  fop_stage(fp, smem, tb, latcost, quint);
  const int iv = best % fp.num_speed, q = best / fp.num_speed, it = q % fp.num_t;
  const int Pc = fp.n_samples[it], i = threadIdx.x;
  if (i == 0) sm_npts = Pc;
  __syncthreads();
  double row[JSSP_TRAJ_COLS];
  const double nanv = nan("");
#pragma unroll
  for (int c = 0; c < JSSP_TRAJ_COLS; ++c) row[c] = nanv;
  const double* lat = tb.lat + (size_t)q * p.P * 4;
  if (i < Pc) {
    QuarticLon lon;
    lon.init(p.s0, p.v0, p.a0, fp.speed[iv], 0.0, fp.horizon[it]);
    lon.eval(tb.time[i], row[1], row[2], row[3], row[4]);
    row[0] = tb.time[i];
    row[5] = lat[i * 4]; row[6] = lat[i * 4 + 1]; row[7] = lat[i * 4 + 2]; row[8] = lat[i * 4 + 3];
    if (!spline_in_range(tb.sp, row[1])) atomicMin(&sm_npts, i);
  }
  // rows beyond the candidate's own length stay NaN; the exported block always has P rows
  export_tail(tb, row, p.P, i < Pc ? lat[i * 4] : 0.0, fp.dt, p.best_traj, &sm_npts, p.best, ex);
  return best;
}

__global__ void __launch_bounds__(EVAL_THREADS) fop_select_export_kernel(const __grid_constant__ FopParams fp, int n_blocks) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  __shared__ int sm_best, sm_npts;
  __shared__ double ex[6][EXPORT_MAX_P];
  fop_select_export_body(fp, n_blocks, smem, red_cost, red_idx, &sm_best, &sm_npts, ex);
}

// batched replay: one block per scenario selects, exports and hands off (blockIdx.x = scenario)
__global__ void __launch_bounds__(EVAL_THREADS) fop_select_export_batch_kernel(const FopParams* __restrict__ fslots, int n_blocks) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(16) FopParams sfp;
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  __shared__ int sm_best, sm_npts;
  __shared__ double ex[6][EXPORT_MAX_P];
  pdl_launch_dependents(); pdl_wait();
  if (!fslots[blockIdx.x].base.batch->live) return;
  load_fop_slot(&sfp, fslots + blockIdx.x);
  const int best = fop_select_export_body(sfp, n_blocks, smem, red_cost, red_idx, &sm_best, &sm_npts, ex);
  batch_handoff(sfp.base, best, sfp.base.W * sfp.num_t * sfp.num_speed, &sm_npts);
}

}  // namespace jssp
