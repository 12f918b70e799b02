// jssp_replay.cuh - K7: the PERSISTENT replay kernel.  One CTA owns one scenario and loops over its plan cycles on the device:
//
//     seed enumeration (polyvt_sampling.py:167-305)  ->  tiled evaluation of every candidate (polyvt_sampling.py:311-385,
//     jerk_space_sampling_planner.py:131-294)  ->  selection (:326-327)  ->  export of the selected trajectory  ->  hand-off
//     (env.py:24-30, intersection_planner.py:312, controller.py:303, 338-380)  ->  next cycle
//
// with block barriers between the stages and NO kernel boundary, no lock-step between scenarios (a scenario never waits for
// the slowest one of the batch), no empty blocks and no launch per cycle: a whole sweep is ONE launch, the hardware block
// scheduler deals waiting scenarios to the SMs as running ones finish.  It replaces the two-launches-per-cycle lock-step
// pipeline (seed_enum_kernel<true> + evaluate_*_batch_kernel) of the driver loop code/__main__.py:124-167.
//
// The stages are the same device functions as the stand-alone kernels (seed_enum_body, walk_tiles, select_and_export,
// export_block, batch_handoff), so a replay is identical, cycle by cycle, to the lock-step replay and to the one-scenario-at-a-
// time replay (tests/test_gpu_batch.py, tests/test_gpu_replay.py).
//
// Shared memory: the reference line (knots + coefficients, TMA bulk copies, once per scenario) and the time grid stay
// resident; the enumeration's survivor lists and the evaluation's tables (lateral samples, obstacle window of the cycle, the
// warps' tiles) share one region.
#pragma once
#include "jssp_eval_tiled.cuh"
#include "jssp_bestfirst.cuh"

namespace jssp {

constexpr int REPLAY_MAX_THREADS = 512;

#ifdef JSSP_PROFILE_PHASES
// profiling builds only (tools/replay_phase_probe.py): nanoseconds scenario 0 spent in each stage, summed over its cycles:
// 0 slot load, 1 seed enumeration, 2 staging, 3 tiled evaluation, 4 selection + export + hand-off, 5 cycles, 6 passes
__device__ unsigned long long g_replay_ns[8];
__device__ unsigned long long g_bestfirst_ns[8];      // best-first cycles of scenario 0: ns in the keys stage, rounds, checks of warp 0
#define RSTAMP(k) do { __syncthreads(); if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); \
                        g_replay_ns[k] += t_ - t_prev; t_prev = t_; } } while (0)
#else
#define RSTAMP(k)
#endif

struct ReplaySmem {
  size_t off_knot, off_coef, off_time, off_union;       // resident part
  size_t off_lat, off_latcost, off_window, off_tile;    // evaluation view of the union
  BestFirstSmem bf;                                     // best-first evaluation: replaces the tiles (offsets from the start of the block's shared memory)
  size_t total;
};
// sized by the host for the LARGEST scenario of the batch (K, kpad, OMpad, rows) and evaluated by a CTA for its own scenario:
// a scenario's view never exceeds the launch's
__host__ __device__ inline ReplaySmem replay_smem_layout(int kpad, int P, int W, int win_rows, int OMpad, size_t seed_bytes, int threads,
                                                         int tile_rounds = 0, int bf_keys = 0) {
  ReplaySmem L;
  size_t o = 16;                                         // mbarrier
  L.off_knot = o; o += (size_t)kpad * 8;
  L.off_coef = o; o += (size_t)8 * kpad * 8;
  L.off_time = o; o += (size_t)((P + 1) & ~1) * 8;
  o = (o + 15) & ~(size_t)15;
  L.off_union = o;
  size_t e = o;
  L.off_lat = e; e += (size_t)W * P * 4 * 8;
  L.off_latcost = e; e += (size_t)W * 4 * 8;
  L.off_window = e; e += (size_t)win_rows * OMpad * 16;
  e = (e + 15) & ~(size_t)15;
  L.off_tile = e;
  L.bf = BestFirstSmem{0, 0, 0, 0};
  if (bf_keys > 0) L.bf = bestfirst_layout(e, P, bf_keys, threads);          // shares the region with the tiles (a cycle uses one of them)
  if (tile_rounds >= 0) e += tile_scratch_bytes(W, threads, tile_rounds);      // tile_rounds < 0: best-first only, no tiles
  if (L.bf.total > e) e = L.bf.total;
  const size_t s = o + ((seed_bytes + 15) & ~(size_t)15);
  L.total = e > s ? e : s;
  return L;
}

// Selection, export and hand-off of a cycle whose candidates were all evaluated by THIS block: block argmin of (cost, index) by
// warp shuffles and one value per warp through shared memory - no per-block records in global memory, no ticket, no fence (the
// multi-block form is select_and_export) - then the same export_block / batch_handoff as everywhere else.
__device__ __forceinline__ void replay_select_export(const EvalParams& p, BlockTables& tb, double my_cost, int my_idx, int N, int Ns,
                                                     double* red_cost, int* red_idx, int* sm_npts) {
  const int nwarps = blockDim.x >> 5, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oc = __shfl_down_sync(0xffffffffu, my_cost, o);
    const int oi = __shfl_down_sync(0xffffffffu, my_idx, o);
    if (lex_less(oc, oi, my_cost, my_idx)) { my_cost = oc; my_idx = oi; }
  }
  if (lane == 0) { red_cost[wid] = my_cost; red_idx[wid] = my_idx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < nwarps; ++q)
      if (lex_less(red_cost[q], red_idx[q], my_cost, my_idx)) { my_cost = red_cost[q]; my_idx = red_idx[q]; }
    BestRecord r;
    r.key = orderable_bits(my_idx >= 0 ? my_cost : INFINITY);
    r.idx = my_idx;
    r.cost = my_cost;
    r.n_pts = 0;
    r.n_candidates = N;
    r.n_seeds = Ns;
    r.pad = p.n_seeds_ptr ? p.n_seeds_ptr[3] : 0;   // seed-enumeration overflow flag, reported with the selection
    *p.best = r;
    red_idx[0] = my_idx;
  }
  __syncthreads();
  const int best = red_idx[0];
  if (best >= 0) export_block(p, tb, best, p.best_traj, sm_npts);
  batch_handoff(p, best, N, sm_npts);
}

// kBestFirstOnly: every cycle of the launch is selected best-first (the host has checked the conditions) and the exhaustive tiled
// passes are not compiled in - the kernel then fits more threads per CTA (768) or more CTAs per SM.
template <int kMaxThreads, bool kBestFirstOnly = false>
__global__ void __launch_bounds__(kMaxThreads, 1) replay_persistent_kernel(SeedGrids g, const EvalParams* __restrict__ slots, SeedWork wk,
                                                                                   int n_cycles, int max_rows, int max_OMpad, int max_kpad, int tile_rounds, int bf_keys) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(16) EvalParams sp;
  __shared__ double red_cost[kMaxThreads / 32];
  __shared__ int red_idx[kMaxThreads / 32];
  __shared__ int sm_npts;
  const int sc = blockIdx.x;
  if (!slot_live(slots, sc)) return;                     // block-uniform
  const int tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, wid = tid >> 5, nwarps = nt >> 5;
  load_slot(&sp, slots + sc);
  // the launch's layout (the union must start where the largest scenario's resident part ends)
  const ReplaySmem L = replay_smem_layout(max_kpad, sp.P, sp.W, max_rows, max_OMpad, seed_smem_bytes(wk), nt, tile_rounds, bf_keys);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  double* sm_knot = reinterpret_cast<double*>(smem + L.off_knot);
  double* sm_coef = reinterpret_cast<double*>(smem + L.off_coef);
  double* sm_time = reinterpret_cast<double*>(smem + L.off_time);
  double* sm_lat = reinterpret_cast<double*>(smem + L.off_lat);
  double* sm_latcost = reinterpret_cast<double*>(smem + L.off_latcost);
  float4* sm_win = reinterpret_cast<float4*>(smem + L.off_window);
  double* sm_tile = reinterpret_cast<double*>(smem + L.off_tile);
  // reference line + time grid: once per scenario, through the TMA engine
  {
    const uint32_t knot_bytes = (uint32_t)sp.kpad * 8u, coef_bytes = (uint32_t)sp.kpad * 64u, time_bytes = (uint32_t)((sp.P + 1) & ~1) * 8u;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_arrive_expect_tx(bar, knot_bytes + coef_bytes + time_bytes);
      bulk_copy_g2s(sm_knot, sp.knot, knot_bytes, bar);
      bulk_copy_g2s(sm_coef, sp.coef, coef_bytes, bar);
      bulk_copy_g2s(sm_time, sp.time, time_bytes, bar);
    }
    mbar_wait(bar, 0);
  }
  const int W = sp.W;
  const TileGeom tg = tile_geom(W, tile_rounds);
  const int g_ = lane / W, j = lane - g_ * W;
  const int spb = nwarps * tg.gpw;                        // seeds per pass of the block
  double* wscr = sm_tile + (size_t)wid * tg.warp_doubles;
  const int gg = g_ < tg.gpw ? g_ : 0;
  double* gbase = wscr + (size_t)gg * tg.grp_stride;
  double* gterms = wscr + tg.off_terms + (size_t)gg * tg.TL * 3;
  double* prof = wscr + tg.off_prof + (size_t)gg * TILE_PROF;
  double* seeds = wk.seeds + (size_t)sc * wk.cap_seeds * 10;
  int* counts = wk.counts + (size_t)sc * 4;

#ifdef JSSP_PROFILE_PHASES
  unsigned long long t_prev = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) { for (int q = 0; q < 8; ++q) g_replay_ns[q] = 0; for (int q = 0; q < 8; ++q) g_bestfirst_ns[q] = 0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_prev)); }
#endif
  for (int c = 0; c < n_cycles; ++c) {
    if (c > 0) load_slot(&sp, slots + sc);                // the hand-off wrote the next cycle's inputs into the slot record
    if (!sp.batch->live) break;                           // block-uniform
    RSTAMP(0);
    // -- K1: seed enumeration by the whole CTA (survivor lists in the union region) --------------------------------------
    {
      SeedBlockSync sy;
      const SeedState st{sp.s0, sp.v0, sp.a0};
      seed_enum_body(g, st, wk, sy, smem + L.off_union, seeds, counts);
    }
    __syncthreads();                                      // seeds + counts (global) are visible to the block; the lists are dead
    RSTAMP(1);
    // -- K2: evaluation ---------------------------------------------------------------------------------------------------
    const int Ns = min(__ldcg(counts + 2), sp.seed_cap);
    const int N = Ns * W;
    // best-first selection (JSSP_FLAG_BEST_FIRST): nobody reads per-candidate results and the risk term can only add to a cost
    const bool best_first = kBestFirstOnly || (bf_keys > 0 && N <= bf_keys && sp.select_only && !sp.feasible && !sp.collision && !sp.cost && sp.w_risk >= 0.0);
    // obstacle window of the cycle: rows now, now + check_res, ... of the broad-phase table (absent rows = far away); transposed
    // ([OMpad][win_rows]) for the best-first checks, whose lanes own consecutive rows
    {
      const int n_el = sp.win_rows * sp.OMpad;
      for (int e = tid; e < n_el; e += nt) {
        const int r = e / sp.OMpad, o = e - r * sp.OMpad;
        const int step = sp.now + r * sp.check_res;
        sm_win[best_first ? o * sp.win_rows + r : e] = step < sp.Tt ? sp.broad[(size_t)step * sp.OMpad + o] : make_float4(0.f, 0.f, FAR_AWAY, 0.f);
      }
    }
    stage_lateral(sp, sm_time, sm_lat, sm_latcost);       // ends with a block barrier
    BlockTables tb;
    tb.broad = sp.win_rows > 0 ? sm_win : sp.broad;
    tb.broad_row0 = sp.win_rows > 0 ? 0 : sp.now; tb.broad_stride = sp.win_rows > 0 ? 1 : sp.check_res;
    tb.sp.knot = sm_knot; tb.sp.coef = sm_coef; tb.sp.K = sp.K; tb.sp.kpad = sp.kpad;
    tb.sp.inv_h = (double)(sp.K - 1) / sm_knot[sp.K - 1];
    tb.time = sm_time; tb.lat = sm_lat;
    tb.list = nullptr; tb.row_off = nullptr; tb.row_cnt = nullptr; tb.use_lists = false;
    const double* latcost = sm_latcost;
    double my_cost = INFINITY;
    int my_idx = -1;
    RSTAMP(2);
    if (best_first) {
      double bc; int bi;
#ifdef JSSP_PROFILE_PHASES
      bestfirst_select(sp, tb, latcost, seeds, Ns, sm_win, sp.win_rows, smem, L.bf, bc, bi, blockIdx.x == 0 ? g_bestfirst_ns : nullptr);
#else
      bestfirst_select(sp, tb, latcost, seeds, Ns, sm_win, sp.win_rows, smem, L.bf, bc, bi);
#endif
      if (tid == 0) { my_cost = bc; my_idx = bi; }
    }
    if (!kBestFirstOnly)
    for (int base = best_first ? Ns : 0; base < Ns; base += spb) {           // block-uniform trip count
      const int k = base + wid * tg.gpw + g_;
      const bool active = g_ < tg.gpw && k < Ns;
      if (active && j == 0) tile_profile_store(prof, seeds + (size_t)k * 10, sp.s0, sp.v0, sp.a0);
      __syncwarp();
      TileSink sink(sp, tb);
      walk_tiles<false>(sp, tb, tg, gbase, gterms, prof, active, j, W, sink, TileChunk{});
      if (active) {
        const int n = j * Ns + k;                          // width-major candidate index (jerk_space_sampling_planner.py:106,124-125)
        const double cost = candidate_cost(sp, sink, latcost + j * 4);
        if (sp.feasible) sp.feasible[n] = sink.curv_fail ? 0 : 1;
        if (sp.collision) sp.collision[n] = sink.collided ? 1 : 0;
        if (sp.cost) sp.cost[n] = (float)cost;
        if (!sink.curv_fail && !sink.collided && isfinite(cost) && lex_less(cost, n, my_cost, my_idx)) { my_cost = cost; my_idx = n; }
      }
    }
    __syncthreads();
    RSTAMP(3);
#ifdef JSSP_PROFILE_PHASES
    if (blockIdx.x == 0 && threadIdx.x == 0) { g_replay_ns[5] += 1; g_replay_ns[6] += (Ns + spb - 1) / spb; g_replay_ns[7] += Ns; }
#endif
    // -- selection, export, hand-off ------------------------------------------------------------------------------------------
    replay_select_export(sp, tb, my_cost, my_idx, N, Ns, red_cost, red_idx, &sm_npts);
    __syncthreads();                                      // the hand-off's writes (slot record, replay state) are visible
    RSTAMP(4);
  }
}

}  // namespace jssp
