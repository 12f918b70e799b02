// jssp_eval_small.cuh - K2b: one warp per candidate (small cycles) and its batch form
#pragma once
#include "jssp_eval.cuh"

namespace jssp {

// ------------------------------------------------------------------------------------------------------------------
// K2b: the same evaluation with ONE WARP PER CANDIDATE for small cycles (the default grid yields ~10^3 candidates: with a
// thread per candidate the kernel's duration is one thread walking 51 samples serially, ~100 us).  Lanes own samples
// i = lane, lane + 32, ...; positions, segments and per-sample cost terms go through shared memory, curvature is tested in
// parallel over segments, each collision row tests 32 obstacle states at once (ballot), and lane 0 adds the cost terms
// in time order so that the cost is bit-identical to the thread-per-candidate kernel.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SMALL_WARPS = 8;                 // candidates per block
constexpr int SMALL_ROW = 10;                  // doubles per sample in the warp scratch: x, y | dx, dy, ds, c, sn | ea2, ej2, ev2
__host__ __device__ inline size_t small_scratch_bytes(int P) { return (size_t)SMALL_WARPS * ((size_t)P * SMALL_ROW + 2) * 8; }

__device__ __forceinline__ void evaluate_small_body(const EvalParams& p, unsigned char* smem, double* red_cost, int* red_idx,
                                                    bool* is_last_p, int* sm_npts_p) {
  const int Ns = resolve_n_seeds(p);
  const int N = Ns * p.W;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int P = p.P;
  double my_cost = INFINITY;
  int my_idx = -1;
  BlockTables tb;
  const double* latcost;
  const bool staged = blk_index(p) * SMALL_WARPS < N;   // block-uniform
  if (staged) {
    stage_tables<0>(p, smem, tb, latcost);
    const SmemLayout L = eval_smem_layout(0, p.OMpad, 0, p.K, p.kpad, p.P, p.W);
    double* scr = reinterpret_cast<double*>(smem + L.total) + (size_t)wid * ((size_t)P * SMALL_ROW + 2);
    double* ends = scr + (size_t)P * SMALL_ROW;        // s at the first / last sample
    // grid-stride over groups of SMALL_WARPS candidates: the candidate count is only known on the device, the grid is sized
    // from a hint and stays small (no thousands of empty blocks in front of a ~100-block cycle)
    for (int n = blk_index(p) * SMALL_WARPS + wid; n < N; n += blk_count(p) * SMALL_WARPS) {     // warp-uniform
      const int w = n / Ns, k = n - w * Ns;
      const double* lat = tb.lat + (size_t)w * P * 4;
      LonProfile lp;
      lon_profile_init(lp, p.seeds + (size_t)k * 10, p.s0, p.v0, p.a0);
      // 1) samples: Frenet state, cost terms, position; n_pts = first sample off the reference line
      int n_pts = P;
      for (int base = 0; base < P; base += 32) {
        const int i = base + lane;
        bool off = false;
        if (i < P) {
          double s, sd, sdd, sddd;
          lon_eval(lp, p.s0, p.v0, p.a0, tb.time[i], s, sd, sdd, sddd);
          const double ea = fmax(fabs(sdd) - p.thr_lon_a, 0.0), ej = fmax(fabs(sddd) - p.thr_lon_j, 0.0), ev = (p.vmax - sd) * p.inv_vmax;
          double* r = scr + (size_t)i * SMALL_ROW;
          r[7] = ea * ea; r[8] = ej * ej; r[9] = ev * ev;
          if (i == 0) ends[0] = s;
          if (i == P - 1) ends[1] = s;
          off = !spline_in_range(tb.sp, s);
          if (!off) frenet_point(tb.sp, spline_walk(tb.sp, s, 0), s, lat[i * 4], r[0], r[1]);
        }
        const unsigned int m = __ballot_sync(0xffffffffu, off);
        if (m != 0u && n_pts == P) n_pts = base + __ffs(m) - 1;
      }
      __syncwarp();
      // 2) forward differences p_{i+1} - p_i: length and unit vector (same expressions as walk_candidate)
      for (int base = 0; base + 1 < n_pts; base += 32) {
        const int i = base + lane;
        if (i + 1 < n_pts) {
          double* r = scr + (size_t)i * SMALL_ROW;
          const double dx = r[SMALL_ROW] - r[0], dy = r[SMALL_ROW + 1] - r[1];
          const double q = dx * dx + dy * dy;
          double ds, c = 1.0, sn = 0.0;
          if (q > 1e-200 && q < 1e200) { const double inv = rsqrt_fast(q); ds = q * inv; c = dx * inv; sn = dy * inv; }
          else { ds = hypot(dx, dy); if (ds > 0.0) { c = dx / ds; sn = dy / ds; } }
          r[2] = dx; r[3] = dy; r[4] = ds; r[5] = c; r[6] = sn;
        }
      }
      __syncwarp();
      // 3) curvature of every interior segment pair
      bool fail = false;
      for (int base = 0; base + 2 < n_pts; base += 32) {
        const int i = base + lane + 1;                 // segment i against segment i - 1
        if (i + 1 < n_pts) {
          const double* a = scr + (size_t)(i - 1) * SMALL_ROW;
          const double* b = scr + (size_t)i * SMALL_ROW;
          fail |= curvature_fails(p.kmax, a[2], a[3], a[4], a[5], a[6], b[2], b[3], b[4], b[5], b[6]);
        }
      }
      fail = __any_sync(0xffffffffu, fail);
      // 4) collision rows, 32 obstacle states at a time
      bool collided = false;
      double risk = 0.0;
      if (p.n_dict > 0) {
        if (n_pts == 1 && p.tcap >= 1) collided = true;                 // bare except -> collision (:250-254)
        const int lim = n_pts >= 2 ? min(n_pts, p.tcap) : 0;
        EvalSink probe(p, tb);                                          // for narrow_hit()
        for (int i = 0; i < lim && !collided; i += p.check_res) {
          const int abs_step = p.now + i;
          if (abs_step >= p.Tt) break;
          const double* r = scr + (size_t)i * SMALL_ROW;
          const double* h = scr + (size_t)(i == n_pts - 1 ? i - 1 : i) * SMALL_ROW;   // last point reuses the previous heading
          const double x = r[0], y = r[1], c = h[5], sn = h[6];
          const float fx = (float)(x - p.ox), fy = (float)(y - p.oy);
          const float m2x = -2.0f * fx, m2y = -2.0f * fy, ne2 = -fmaf(fx, fx, fy * fy);
          const float4* row = p.broad + (size_t)abs_step * p.OMpad;
          for (int base = 0; base < p.OMpad; base += 32) {
            const float4 b = __ldg(row + base + lane);
            bool hit = fmaf(b.x, m2x, fmaf(b.y, m2y, b.z)) <= ne2;
            if (hit) hit = narrow_hit(p, abs_step, base + lane, b.x - fx, b.y - fy, x, y, c, sn);
            const bool hard = hit && (double)b.w >= p.p_min;
            if (__any_sync(0xffffffffu, hard)) { collided = true; break; }
            unsigned int soft = __ballot_sync(0xffffffffu, hit);
            while (soft) {                                              // risk in obstacle order, like the serial walk
              const int j = __ffs(soft) - 1;
              soft &= soft - 1u;
              risk = risk + (double)__shfl_sync(0xffffffffu, b.w, j);
            }
          }
        }
      }
      // 5) cost: lane 0 adds the per-sample terms in time order
      if (lane == 0) {
        double sum_a = 0, sum_j = 0, sum_v = 0;
        for (int i = 0; i < P; ++i) {
          const double* r = scr + (size_t)i * SMALL_ROW;
          sum_a = sum_a + r[7]; sum_j = sum_j + r[8]; sum_v = sum_v + r[9];
        }
        const double cost = candidate_cost(p, sum_a, sum_j, sum_v, ends[0], ends[1], risk, latcost + w * 4);
        if (p.feasible) p.feasible[n] = fail ? 0 : 1;
        if (p.collision) p.collision[n] = collided ? 1 : 0;
        if (p.cost) p.cost[n] = (float)cost;
        if (!fail && !collided && isfinite(cost) && lex_less(cost, n, my_cost, my_idx)) { my_cost = cost; my_idx = n; }
      }
      __syncwarp();                                    // the warp's scratch is reused by its next candidate
    }
  }
  select_and_export<0>(p, smem, tb, latcost, staged, my_cost, my_idx, N, Ns, red_cost, red_idx, is_last_p, sm_npts_p);
}

__global__ void __launch_bounds__(SMALL_WARPS * 32) evaluate_small_kernel(const __grid_constant__ EvalParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red_cost[SMALL_WARPS];
  __shared__ int red_idx[SMALL_WARPS];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  pdl_wait();          // one-call plan cycles chain this kernel to the seed enumeration (no-op for ordinary launches)
  evaluate_small_body(p, smem, red_cost, red_idx, &is_last, &sm_npts);
}

__global__ void __launch_bounds__(SMALL_WARPS * 32) evaluate_small_batch_kernel(const EvalParams* __restrict__ slots) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(16) EvalParams sp;
  __shared__ double red_cost[SMALL_WARPS];
  __shared__ int red_idx[SMALL_WARPS];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  pdl_launch_dependents(); pdl_wait();
  if (!slot_live(slots, blockIdx.x)) return;
  load_slot(&sp, slots + blockIdx.x);
  evaluate_small_body(sp, smem, red_cost, red_idx, &is_last, &sm_npts);
}

}  // namespace jssp
