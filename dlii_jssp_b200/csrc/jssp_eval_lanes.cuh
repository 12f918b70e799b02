// jssp_eval_lanes.cuh - K2c: ONE WARP PER SEED, LANES OVER THE SAMPLES - the evaluation without any serial walk.
//
// A candidate's checks only couple neighbouring samples (a chord needs two points, the curvature test two chords), so nothing
// forces one thread to walk the horizon.  Here a warp takes a seed; its lanes own the samples i = lane, lane + 32, ...:
//   P1  (once per seed)       unrolled s, s_d, s_dd, s_ddd from the seed's segment table (polyvt_sampling.py:311-385), the cost
//                             terms, the reference point and unit normal (jerk_space_sampling_planner.py:131-145) -> shared memory;
//                             three lanes then add the cost terms strictly in time order (the oracle's order: costs bit-identical)
//   S1  (per lateral target)  lane i: its point, the next point, the chord, its length and inverse length -> shared memory
//   S2                        lane i: curvature of the chord pair (i, i + 1)  (:155-180; the guard-banded test of the other kernels)
//   S3                        lane r: collision row r (sample r * check_res) against the row's obstacle states (:233-286)
//   S4                        ballots / ordered sums -> masks, risk, cost, running argmin (lane 0)
// Every stage is a handful of independent instructions per lane; a warp carries no state from sample to sample, needs few
// registers (so twice as many warps are resident as for the walking kernels) and a scenario's or cycle's seeds spread over as
// many warps as the GPU holds.  What the serial walks get for free is lost: all collision rows of a candidate are tested even if
// an early one hits (select-only evaluation still skips the rows of a candidate that failed the curvature test).
#pragma once
#include "jssp_eval_tiled.cuh"

namespace jssp {

#ifndef JSSP_LANES_THREADS
#define JSSP_LANES_THREADS 512
#endif
constexpr int LANES_THREADS = JSSP_LANES_THREADS;
constexpr int LANES_SEG = 4;       // doubles per sample in the chord table: dx, dy, ds, 1/ds

// per-warp scratch in doubles: segment table | reference points [P][4] | max(cost terms [P][3], chord table [P][6]) | row results [R][2]
__host__ __device__ inline size_t lanes_warp_doubles(int P, int rows) { return (TILE_PROF + (size_t)P * 4 + (size_t)P * LANES_SEG + (size_t)rows * 2 + 1) & ~(size_t)1; }
__host__ __device__ inline size_t lanes_scratch_bytes(int P, int rows, int threads) { return (size_t)(threads / 32) * lanes_warp_doubles(P, rows) * 8; }

// One seed, all of its lateral targets, by one warp.  Returns through best_cost / best_idx the warp's running argmin (lane 0).
__device__ __forceinline__ void lanes_eval_seed(const EvalParams& p, const BlockTables& tb, const double* latcost, double* wscr, int k, int Ns,
                                                int n_rows, double& best_cost, int& best_idx) {
  const int P = p.P, W = p.W, cr = p.check_res;
  const int lane = threadIdx.x & 31;
  double* prof = wscr;
  double* base = prof + TILE_PROF;                       // [P][4] ix, iy, u, v
  double* terms = base + (size_t)P * 4;                  // [P][3] during P1 ...
  double* seg = terms;                                   // ... then [P][4] dx, dy, ds, 1/ds per lateral target
  double* rowres = seg + (size_t)P * LANES_SEG;          // [rows][2] collided flag, risk
  if (lane == 0) tile_profile_store(prof, p.seeds + (size_t)k * 10, p.s0, p.v0, p.a0);
  __syncwarp();
  // -- P1 ---------------------------------------------------------------------------------------------------------------
  int n_pts = P, kseg = 0;
  double s_first = 0.0, s_last = 0.0;
  for (int b0 = 0; b0 < P; b0 += 32) {
    const int i = b0 + lane;
    bool off = false;
    if (i < P) {
      double s, sd, sdd, sddd;
      tile_lon_eval(prof, kseg, tb.time[i], s, sd, sdd, sddd);
      const double ea = fmax(fabs(sdd) - p.thr_lon_a, 0.0);
      const double ej = fmax(fabs(sddd) - p.thr_lon_j, 0.0);
      const double ev = (p.vmax - sd) * p.inv_vmax;
      terms[i * 3] = ea * ea; terms[i * 3 + 1] = ej * ej; terms[i * 3 + 2] = ev * ev;
      if (i == 0) s_first = s;
      if (i == P - 1) s_last = s;
      off = !spline_in_range(tb.sp, s);
      if (!off) {
        double ix, iy, u, v;
        frenet_base(tb.sp, spline_walk(tb.sp, s, 0), s, ix, iy, u, v);
        double2* e = reinterpret_cast<double2*>(base + (size_t)i * 4);
        e[0] = make_double2(ix, iy); e[1] = make_double2(u, v);
      }
    }
    const unsigned int m = __ballot_sync(0xffffffffu, off);
    if (m != 0u && n_pts == P) n_pts = b0 + __ffs(m) - 1;      // first sample off the reference line (jssp.py:138-139)
  }
  s_first = __shfl_sync(0xffffffffu, s_first, 0);
  s_last = __shfl_sync(0xffffffffu, s_last, (P - 1) & 31);
  __syncwarp();
  // cost sums: lanes 0, 1, 2 add one term each over the samples, strictly in time order
  double acc = 0.0;
  if (lane < 3) for (int i = 0; i < P; ++i) acc = acc + terms[i * 3 + lane];
  const double sum_a = __shfl_sync(0xffffffffu, acc, 0), sum_j = __shfl_sync(0xffffffffu, acc, 1), sum_v = __shfl_sync(0xffffffffu, acc, 2);
  __syncwarp();                                          // the terms are dead: the chord table takes their place
  // -- per lateral target ---------------------------------------------------------------------------------------------------
  for (int w = 0; w < W; ++w) {
    const double* lat = tb.lat + (size_t)w * P * 4;
    // S1: points and chords
    for (int i = lane; i < n_pts; i += 32) {
      double x, y;
      {
        const double2 a = *reinterpret_cast<const double2*>(base + (size_t)i * 4), b = *reinterpret_cast<const double2*>(base + (size_t)i * 4 + 2);
        const double d = lat[i * 4];
        x = a.x - d * b.x; y = a.y + d * b.y;
      }
      double dx = 0.0, dy = 0.0, ds = 0.0, inv = 0.0;
      if (i + 1 < n_pts) {
        const double2 a = *reinterpret_cast<const double2*>(base + (size_t)(i + 1) * 4), b = *reinterpret_cast<const double2*>(base + (size_t)(i + 1) * 4 + 2);
        const double d = lat[(i + 1) * 4];
        dx = (a.x - d * b.x) - x; dy = (a.y + d * b.y) - y;
        const double q = dx * dx + dy * dy;
        if (q > 1e-200 && q < 1e200) { inv = rsqrt_fast(q); ds = q * inv; }
        else { ds = hypot(dx, dy); inv = ds > 0.0 ? 1.0 / ds : 0.0; }
      }
      double2* e = reinterpret_cast<double2*>(seg + (size_t)i * LANES_SEG);
      e[0] = make_double2(dx, dy); e[1] = make_double2(ds, inv);
    }
    __syncwarp();
    // S2: curvature of the chord pairs (i, i + 1), i + 2 <= n_pts - 1
    bool fail = false;
    for (int i = lane; i + 2 < n_pts; i += 32) {
      const double* a = seg + (size_t)i * LANES_SEG;
      const double* b = a + LANES_SEG;
      const double pc = a[2] > 0.0 ? a[0] * a[3] : 1.0, psn = a[2] > 0.0 ? a[1] * a[3] : 0.0;
      const double c = b[2] > 0.0 ? b[0] * b[3] : 1.0, sn = b[2] > 0.0 ? b[1] * b[3] : 0.0;
      fail |= curvature_fails(p.kmax, a[0], a[1], a[2], pc, psn, b[0], b[1], b[2], c, sn);
    }
    fail = __any_sync(0xffffffffu, fail);
    // S3: collision rows (select-only evaluation: a candidate that failed the curvature test is out already)
    bool collided = false;
    double risk = 0.0;
    if (!(p.select_only && fail)) {
      if (n_pts == 1 && p.n_dict > 0 && p.tcap >= 1) collided = true;            // bare except -> collision (:250-254)
      bool any_risk = false;
      if (n_pts >= 2 && p.n_dict > 0) {
        for (int r = lane; r < n_rows; r += 32) {
          const int i = r * cr;
          bool hit = false;
          double rk = 0.0;
          if (i < n_pts) {
            const double* h = seg + (size_t)(i + 1 < n_pts ? i : i - 1) * LANES_SEG;     // the last point reuses the previous heading (:158-160)
            const double c = h[2] > 0.0 ? h[0] * h[3] : 1.0, sn = h[2] > 0.0 ? h[1] * h[3] : 0.0;
            const double2 a = *reinterpret_cast<const double2*>(base + (size_t)i * 4), b = *reinterpret_cast<const double2*>(base + (size_t)i * 4 + 2);
            const double d = lat[i * 4];
            collide_row(p, tb, r, i, a.x - d * b.x, a.y + d * b.y, c, sn, hit, rk);
          }
          rowres[r * 2] = hit ? 1.0 : 0.0; rowres[r * 2 + 1] = rk;
          collided |= hit;
          any_risk |= rk != 0.0;
        }
        collided = __any_sync(0xffffffffu, collided);
        any_risk = __any_sync(0xffffffffu, any_risk);
        __syncwarp();
        if (any_risk && lane == 0)                                               // soft hits: rows in time order, up to the first veto
          for (int r = 0; r < n_rows; ++r) { risk = risk + rowres[r * 2 + 1]; if (rowres[r * 2] != 0.0) break; }
      }
    }
    // S4: cost, outputs, running argmin
    if (lane == 0) {
      const int n = w * Ns + k;                // candidate index: width-major (jerk_space_sampling_planner.py:106,124-125)
      const double cost = candidate_cost(p, sum_a, sum_j, sum_v, s_first, s_last, risk, latcost + w * 4);
      if (p.feasible) p.feasible[n] = fail ? 0 : 1;
      if (p.collision) p.collision[n] = collided ? 1 : 0;
      if (p.cost) p.cost[n] = (float)cost;
      if (!fail && !collided && isfinite(cost) && lex_less(cost, n, best_cost, best_idx)) { best_cost = cost; best_idx = n; }
    }
    __syncwarp();                                        // the chord table is rewritten for the next lateral target
  }
}

// arc length of a seed at the collision rows -> the block's [row_min, row_max] (list building, phase A): lane r takes row r
__device__ __forceinline__ void lanes_row_ranges(const EvalParams& p, const BlockTables& tb, const ListsView& v, double* prof, int k) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) tile_profile_store(prof, p.seeds + (size_t)k * 10, p.s0, p.v0, p.a0);
  __syncwarp();
  const double s_end = tb.sp.knot[p.K - 1];
  int kseg = 0;
  for (int r = lane; r < p.win_rows; r += 32) {
    double s, sd, sdd, sddd;
    tile_lon_eval(prof, kseg, tb.time[r * p.check_res], s, sd, sdd, sddd);
    s = fmin(fmax(s, 0.0), s_end);                      // off-line samples are never collision-checked
    atomicMin(&v.row_min[r], __float_as_uint(__double2float_rd(s)));       // s >= 0: the bit patterns order like the values
    atomicMax(&v.row_max[r], __float_as_uint(__double2float_ru(s)));
  }
  __syncwarp();
}

template <int kMode>
__device__ __forceinline__ void evaluate_lanes_body(const EvalParams& p, unsigned char* smem, double* red_cost, int* red_idx, bool* is_last,
                                                    int* sm_npts) {
  const int Ns = resolve_n_seeds(p);
  const int N = Ns * p.W;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int nblk = blk_count(p), blk = blk_index(p);
  // contiguous seed range of the block (coherent ranges keep the culled lists short when the seeds come in loop order)
  const int per = (Ns + nblk - 1) / nblk;
  const int k_lo = min(blk * per, Ns), k_hi = min(k_lo + per, Ns);
  double my_cost = INFINITY;
  int my_idx = -1;
  BlockTables tb;
  const double* latcost;
  const bool staged = k_lo < k_hi;                     // block-uniform
  if (staged) {
    stage_tables<kMode>(p, smem, tb, latcost);
    const SmemLayout L = eval_smem_layout(p.win_rows, p.OMpad, kMode, p.K, p.kpad, p.P, p.W, p.list_cap);
    const int n_rows = p.tcap > 0 ? (min(p.P, p.tcap) + p.check_res - 1) / p.check_res : 0;
    double* wscr = reinterpret_cast<double*>(smem + ((L.total + 15) & ~(size_t)15)) + (size_t)wid * lanes_warp_doubles(p.P, (p.P + p.check_res - 1) / p.check_res);
    if (kMode == 2) {
      const ListsView v = lists_view(p, smem);
      lists_begin(p, v);
      for (int k = k_lo + wid; k < k_hi; k += nwarps) lanes_row_ranges(p, tb, v, wscr, k);
      lists_compact(p, v, tb);
    }
    for (int k = k_lo + wid; k < k_hi; k += nwarps) lanes_eval_seed(p, tb, latcost, wscr, k, Ns, n_rows, my_cost, my_idx);
  }
  (void)lane;
  __syncthreads();
  select_and_export<kMode>(p, smem, tb, latcost, staged, my_cost, my_idx, N, Ns, red_cost, red_idx, is_last, sm_npts);
}

template <int kMode>
__global__ void __launch_bounds__(LANES_THREADS, 2) evaluate_lanes_kernel(const __grid_constant__ EvalParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red_cost[LANES_THREADS / 32];
  __shared__ int red_idx[LANES_THREADS / 32];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  evaluate_lanes_body<kMode>(p, smem, red_cost, red_idx, &is_last, &sm_npts);
}

constexpr int LANES_BATCH_THREADS = 256;
template <int kMode>
__global__ void __launch_bounds__(LANES_BATCH_THREADS, 4) evaluate_lanes_batch_kernel(const EvalParams* __restrict__ slots) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(16) EvalParams sp;
  __shared__ double red_cost[LANES_BATCH_THREADS / 32];
  __shared__ int red_idx[LANES_BATCH_THREADS / 32];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  pdl_launch_dependents(); pdl_wait();
  if (!slot_live(slots, blockIdx.x)) return;
  load_slot(&sp, slots + blockIdx.x);
  evaluate_lanes_body<kMode>(sp, smem, red_cost, red_idx, &is_last, &sm_npts);
}

}  // namespace jssp
