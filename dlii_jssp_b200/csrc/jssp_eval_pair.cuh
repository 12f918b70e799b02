// jssp_eval_pair.cuh - K2 (large cycles): the PAIRED walk - the tiled two-phase walk with the candidate's serial work split
// by FUNCTION over two warps.
//
// A warp pair shares floor(32 / W) seeds.  Both warps run phase 1 of a tile (the seed-only quantities of the tile's samples,
// half of the samples each, into the pair's shared-memory tile, see jssp_eval_tiled.cuh).  In phase 2 lane (g, j) of BOTH warps
// stands for the same candidate (seed g, lateral target j), but
//   the CURVATURE warp  forms the candidate's points and forward differences and runs the curvature test of every segment
//                       (jerk_space_sampling_planner.py:155-180),
//   the COLLISION warp  forms the same points and runs the collision rows of every check_res-th sample (:233-286), adds the
//                       seed's longitudinal cost terms in time order and, at the end, owns the candidate's result.
// A pair is synchronised by its own named barrier (bar.sync id, 64) twice per tile; pairs are decoupled from each other.
// Why: the tiled walk is bound by the latency of one lane's dependent chain (3.5 warps per scheduler, one issue per ~7.7
// cycles).  Splitting by function halves the chain, doubles the warps per SM, removes the divergence between the two code paths
// inside a warp, and - unlike a split of the horizon - keeps the early stop of a collided candidate.  Arithmetic, operand order
// and decisions are the tiled walk's own functions (TileSink), so masks, float64 costs and the selection are identical.
#pragma once
#include "jssp_eval_tiled.cuh"

namespace jssp {

#ifndef JSSP_PAIR_THREADS
#define JSSP_PAIR_THREADS 448
#endif
#ifndef JSSP_PAIR_MIN_BLOCKS
#define JSSP_PAIR_MIN_BLOCKS 2
#endif
constexpr int PAIR_THREADS = JSSP_PAIR_THREADS;          // 7 pairs; with 2 blocks per SM every SM holds 28 warps
constexpr int PAIR_MIN_BLOCKS = JSSP_PAIR_MIN_BLOCKS;
constexpr int PAIR_XCH = 32;                             // doubles per pair after the tile: 32 curvature flags (int) + 32 veto flags (int)

__host__ __device__ inline int pair_doubles(int W, int rounds) { return tile_geom(W, rounds).warp_doubles + PAIR_XCH; }
__host__ __device__ inline size_t pair_scratch_bytes(int W, int threads, int rounds) { return (size_t)(threads / 64) * pair_doubles(W, rounds) * 8; }
__host__ __device__ inline int pair_seeds_per_block(int W, int threads) { return (threads / 64) * (32 / W); }

__device__ __forceinline__ void pair_sync(int id) {
  __syncwarp();
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

template <int kMode>
__device__ __forceinline__ void evaluate_pair_body(const EvalParams& p, unsigned char* smem, double* red_cost, int* red_idx,
                                                   bool* is_last, int* sm_npts) {
  const int Ns = resolve_n_seeds(p);
  const int W = p.W, P = p.P;
  const int N = Ns * W;
  const TileGeom tg = tile_geom(W, p.tile_rounds);
  const int gpw = tg.gpw, TL = tg.TL;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int pair = wid >> 1;
  // the collision warp of pair q sits at warp 2q or 2q + 1 so that the (heavier) curvature warps cover all four schedulers
  const bool roleB = (wid & 1) != ((wid >> 2) & 1);
  const int g = lane / W, j = lane - g * W;
  const int spb = (blockDim.x >> 6) * gpw;       // seeds per block
  const int k = blk_index(p) * spb + pair * gpw + g;
  const bool active = g < gpw && k < Ns;
  double my_cost = INFINITY;
  int my_idx = -1;
  BlockTables tb;
  const double* latcost;
  const bool staged = blk_index(p) * spb < Ns;   // block-uniform
  PHASE(p, 0);
  if (staged) {
    stage_tables<kMode>(p, smem, tb, latcost);
    PHASE(p, 1);
    const SmemLayout L = eval_smem_layout(p.win_rows, p.OMpad, kMode, p.K, p.kpad, p.P, p.W, p.list_cap);
    double* pscr = reinterpret_cast<double*>(smem + ((L.total + 15) & ~(size_t)15)) + (size_t)pair * (tg.warp_doubles + PAIR_XCH);
    const int gq = active ? g : 0;
    double* gbase = pscr + (size_t)gq * tg.grp_stride;
    double* gterms = pscr + tg.off_terms + (size_t)gq * TL * 3;
    double* prof = pscr + tg.off_prof + (size_t)gq * TILE_PROF;
    int* xflags = reinterpret_cast<int*>(pscr + tg.warp_doubles);      // [0, 32) curvature result, [32, 64) select-only veto
    if (!roleB) {
      if (active && j == 0) tile_profile_store(prof, p.seeds + (size_t)k * 10, p.s0, p.v0, p.a0);
      xflags[32 + lane] = 0;
    }
    __syncthreads();
    if (kMode == 2) { TileSAt s_at{prof}; build_lists(p, smem, tb, s_at, active, j, W, roleB ? 1 : 0, 2); }
    PHASE(p, 2);
    TileSink sink(p, tb);
    const double* lat = tb.lat + (size_t)j * P * 4;
    const int n_tiles = (P + TL - 1) / TL;
    const int barid = 1 + pair;
    int kseg = 0, n_pts = P;
    double sa = 0, sj = 0, sv = 0;                 // collision warp, lane 0 of a group: the longitudinal cost sums in time order
    double xp = 0, yp = 0;
    bool walking = active;
    bool finish_ok = true;                         // collision warp: the walk reached the end of the path (finish() is due)
    for (int t = 0; t < n_tiles; ++t) {
      const int base = t * TL;
      const int len = min(TL, P - base);
      // phase 1: the two warps share the tile's samples of every seed
      if (active) {
        for (int il = j + (roleB ? W : 0); il < len; il += 2 * W) {
          const int i = base + il;
          double s, sd, sdd, sddd;
          tile_lon_eval(prof, kseg, tb.time[i], s, sd, sdd, sddd);
          const double ea = fmax(fabs(sdd) - p.thr_lon_a, 0.0);
          const double ej = fmax(fabs(sddd) - p.thr_lon_j, 0.0);
          const double ev = (p.vmax - sd) * p.inv_vmax;
          double* tm = gterms + il * 3;
          tm[0] = ea * ea; tm[1] = ej * ej; tm[2] = ev * ev;
          if (i == 0) prof[28] = s;
          if (i == P - 1) prof[29] = s;
          double2* e = reinterpret_cast<double2*>(gbase + il * 4);
          if (spline_in_range(tb.sp, s)) {
            double ix, iy, u, v;
            frenet_base_fma(tb.sp, spline_walk(tb.sp, s, 0), s, ix, iy, u, v);
            e[0] = make_double2(ix, iy); e[1] = make_double2(u, v);
          } else {
            e[0] = make_double2(nan(""), 0.0);     // off the reference line: the Cartesian path ends here (jssp.py:138-139)
          }
        }
      }
      pair_sync(barid);
      if (walking && p.select_only && xflags[32 + lane]) { walking = false; finish_ok = false; }   // vetoed by the partner
      if (roleB) {
        // ---- collision warp: cost sums (one lane per seed, strictly in time order) and the collision rows
        if (active && j == 0)
          for (int il = 0; il < len; ++il) { sa = sa + gterms[il * 3]; sj = sj + gterms[il * 3 + 1]; sv = sv + gterms[il * 3 + 2]; }
        if (walking) {
          for (int il = 0; il < len; ++il) {
            const int i = base + il;
            const double2* e = reinterpret_cast<const double2*>(gbase + il * 4);
            const double2 pxy = e[0];
            if (!(pxy.x == pxy.x)) { n_pts = i; walking = false; break; }
            const double2 uv = e[1];
            const double d = lat[i * 4];
            const double x = fma(-d, uv.x, pxy.x), y = fma(d, uv.y, pxy.y);
            if (i > 0) {
              const double dx = x - xp, dy = y - yp;               // segment i - 1: heading of sample i - 1
              if (i - 1 == sink.next_check) {
                const float dxf = (float)dx, dyf = (float)dy;
                sink.collide(i - 1, xp, yp, dx, dy, fmaf(dxf, dxf, dyf * dyf));
                if (sink.collided) {                               // nothing is left to do for this candidate's collision rows
                  walking = false; finish_ok = false;
                  if (p.select_only) xflags[32 + lane] = 1;
                  break;
                }
              }
              sink.pdx = dx; sink.pdy = dy;
            }
            xp = x; yp = y;
          }
        }
      } else {
        // ---- curvature warp: two segments per step where the tile has them (two independent dependency chains)
        if (walking) {
          int il = 0;
          while (il < len) {
            const int i = base + il;
            const double2* e = reinterpret_cast<const double2*>(gbase + il * 4);
            const double2 pxy = e[0];
            if (!(pxy.x == pxy.x)) { walking = false; break; }
            const double2 uv = e[1];
            const double d = lat[i * 4];
            const double x = fma(-d, uv.x, pxy.x), y = fma(d, uv.y, pxy.y);
            if (i == 0) { xp = x; yp = y; ++il; continue; }
            bool two = il + 1 < len;
            double x1 = 0, y1 = 0;
            if (two) {
              const double2 qxy = e[2];
              const double2 uv1 = e[3];
              const double d1 = lat[(i + 1) * 4];
              x1 = fma(-d1, uv1.x, qxy.x); y1 = fma(d1, uv1.y, qxy.y);
              two = qxy.x == qxy.x;                // the path ends at the next point: it is met again, alone, in the next step
            }
            if (two) {
              sink.curv2(i - 1, x - xp, y - yp, x1 - x, y1 - y);
              xp = x1; yp = y1; il += 2;
            } else {
              sink.curv1(i - 1, x - xp, y - yp);
              xp = x; yp = y; ++il;
            }
            if (p.select_only && sink.curv_fail) { xflags[32 + lane] = 1; walking = false; break; }
          }
        }
      }
      pair_sync(barid);                            // the tile is rewritten by the next phase 1
    }
    // ---- the collision warp owns the result: last point's collision row, the partner's curvature flag, cost, outputs
    if (roleB) {
      if (active && finish_ok) {
        const float dxf = (float)sink.pdx, dyf = (float)sink.pdy;
        sink.pqf = fmaf(dxf, dxf, dyf * dyf);
        sink.finish(n_pts, xp, yp);
      }
      if (active && j == 0) { gbase[0] = sa; gbase[1] = sj; gbase[2] = sv; }
    } else {
      xflags[lane] = sink.curv_fail ? 1 : 0;
    }
    pair_sync(barid);
    PHASE(p, 12);
    if (roleB && active) {
      sink.curv_fail = xflags[lane] != 0;
      sink.sum_a = gbase[0]; sink.sum_j = gbase[1]; sink.sum_v = gbase[2];
      sink.s_first = prof[28]; sink.s_last = prof[29];
      const int n = j * Ns + k;                    // candidate index: width-major (jerk_space_sampling_planner.py:106,124-125)
      const double cost = candidate_cost(p, sink, latcost + j * 4);
      if (p.feasible) p.feasible[n] = sink.curv_fail ? 0 : 1;
      if (p.collision) p.collision[n] = sink.collided ? 1 : 0;
      if (p.cost) p.cost[n] = (float)cost;
      if (!sink.curv_fail && !sink.collided && isfinite(cost)) { my_cost = cost; my_idx = n; }
    }
  }
  PHASE(p, 13);
  __syncthreads();
  PHASE(p, 3);
  select_and_export<kMode>(p, smem, tb, latcost, staged, my_cost, my_idx, N, Ns, red_cost, red_idx, is_last, sm_npts);
  PHASE(p, 4);
}

template <int kMode>
__global__ void __launch_bounds__(PAIR_THREADS, PAIR_MIN_BLOCKS) evaluate_pair_kernel(const __grid_constant__ EvalParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red_cost[PAIR_THREADS / 32];
  __shared__ int red_idx[PAIR_THREADS / 32];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  evaluate_pair_body<kMode>(p, smem, red_cost, red_idx, &is_last, &sm_npts);
}

}  // namespace jssp
