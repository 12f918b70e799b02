// jssp_bestfirst.cuh - BEST-FIRST selection of a select-only plan cycle by ONE block (used by the persistent replay kernel).
//
// The reference filters and then scores the survivors (check_constraints_curvature -> check_collisions -> calc_all_trajs_cost ->
// PriorityQueue argmin, jerk_space_sampling_planner.py:311-327); the replay only consumes the argmin.  Every cost term except the
// soft-collision risk is a function of the seed and of the lateral target alone - no Frenet->Cartesian walk is needed for it - and
// the risk term only ADDS (w_risk >= 0, risk >= 0).  So:
//
//   1. bounds     every candidate's cost without the risk term, from partial sums (4 lanes per seed split the horizon: no serial
//                 walk), minus a 1e-9 margin and rounded DOWN to float32: a key that is <= the candidate's exact cost;
//   2. checks     the keys are dealt to the warps; every warp, on its own, takes the unchecked candidate with the smallest key of
//                 ITS slice and CHECKS it: one warp per candidate, lanes over the samples - positions, curvature of every segment
//                 pair, every collision row, all in parallel - with exactly the arithmetic of the tiled walk (tile_lon_eval,
//                 frenet_base_fma, TileSink::curvature_classify, the collision row of TileSink::collide), then, for a survivor,
//                 the EXACT cost: the per-sample terms added strictly in time order, as everywhere else.  The best exact cost found
//                 so far is shared by the warps of the block (a 64-bit atomic minimum in shared memory);
//   3. stop       a warp stops when the smallest key left in its slice exceeds that cost: none of its candidates can win any more
//                 (keys are lower bounds, the shared cost only falls).  No barrier until every warp has stopped.
//
// The selected candidate, its cost and the exported trajectory are those of the exhaustive evaluation (the same lexicographic
// (cost, index) minimum over the same survivors); candidates that cannot win are never walked.  When nothing survives, every
// candidate ends up checked.
#pragma once
#include "jssp_eval_tiled.cuh"

namespace jssp {

constexpr int BF_LANES_PER_SEED = 4;

__host__ __device__ inline int bestfirst_warp_doubles(int P) { return (TILE_PROF + 5 * P + 1) & ~1; }   // segment table | xy[P][2] | terms[P][3]
struct BestFirstSmem { size_t off_keys, off_red, off_warp, total; };
__host__ __device__ inline BestFirstSmem bestfirst_layout(size_t base, int P, int keys_cap, int threads) {
  BestFirstSmem L;
  const int nw = threads / 32;
  size_t o = (base + 15) & ~(size_t)15;
  L.off_keys = o; o += (size_t)((keys_cap + nw - 1) / nw) * nw * 8;     // one contiguous slice per warp
  L.off_red = o; o += (size_t)2 * nw * 24;                             // two parities x (next key, best cost, best index)
  L.off_warp = o; o += (size_t)nw * bestfirst_warp_doubles(P) * 8;
  L.total = o;
  return L;
}

__device__ __forceinline__ unsigned int bf_key32(double lb) {
  if (!(lb == lb)) return 0u;                                           // NaN bound: checked first (it cannot survive, but never skip it)
  const float f = __double2float_rd(lb - (1e-9 + 1e-9 * fabs(lb)));
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float bf_key32_value(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ unsigned long long bf_warp_min(unsigned long long v) {
  const unsigned int hi = (unsigned int)(v >> 32);
  const unsigned int mh = __reduce_min_sync(0xffffffffu, hi);
  const unsigned int lo = hi == mh ? (unsigned int)v : 0xffffffffu;
  const unsigned int ml = __reduce_min_sync(0xffffffffu, lo);
  return ((unsigned long long)mh << 32) | ml;
}

// step 1: the keys of all N = Ns * W candidates; key of candidate n sits in the slice of warp n % nwarps at n / nwarps
__device__ __forceinline__ void bestfirst_keys(const EvalParams& p, const BlockTables& tb, const double* latcost, const double* seeds, int Ns,
                                               unsigned long long* keys, int slice) {
  constexpr int L = BF_LANES_PER_SEED;
  const int tid = threadIdx.x, nt = blockDim.x, nwarps = nt >> 5;
  const int N = Ns * p.W;
  for (int e = tid; e < slice * nwarps; e += nt) keys[e] = ~0ull;
  __syncthreads();
  const int q = tid & (L - 1);
  // the bound is an approximation anyway (margin in bf_key32): reciprocals instead of candidate_cost's float64 divisions
  const double invP = 1.0 / (double)p.P, inv_md = 1.0 / p.max_distance, inv_hw2 = 1.0 / (p.half_w * p.half_w);
  for (int k0 = 0; k0 < Ns; k0 += nt / L) {                              // block-uniform trip count
    const int k = k0 + tid / L;
    double sa = 0, sj = 0, sv = 0, sf = 0, sl = 0;
    if (k < Ns) {
      double seed[10];
      load_seed_row(seed, seeds + (size_t)k * 10);
      // speed and acceleration at the segment ends (the expressions of lon_profile_init); the arc length enters the bound only as the
      // displacement s(T) - s(0)
      const double j1 = seed[0], t1 = seed[1], j2 = seed[2], t2 = seed[3], j3 = seed[4], t3 = seed[5], j4 = seed[6], t4 = seed[7], j5 = seed[8];
      const double v1 = vt_by_jt(p.v0, p.a0, j1, t1), a1 = at_by_jt(p.a0, j1, t1);
      const double v2 = vt_by_jt(v1, a1, j2, t2), a2 = at_by_jt(a1, j2, t2);
      const double v3 = vt_by_jt(v2, a2, j3, t3), a3 = at_by_jt(a2, j3, t3);
      const double v4 = vt_by_jt(v3, a3, j4, t4), a4 = at_by_jt(a3, j4, t4);
      const double b1 = t1 - 1e-9, b2 = t1 + t2 + 1e-9, b3 = t1 + t2 + t3 + 1e-9, b4 = t1 + t2 + t3 + t4 + 1e-9;
      // this lane's samples q, q + L, ... segment by segment (lon_eval's branch chain as five loops: speed, acceleration and jerk of a
      // segment stay in registers, the jerk term is a constant of the segment)
      int i = q;
#define BF_SEGMENT(VB, AB, JB, OFF, COND)                                                                                  \
      {                                                                                                                  \
        const double ejs_ = fmax(fabs(JB) - p.thr_lon_j, 0.0), ej2_ = ejs_ * ejs_;                                       \
        while (i < p.P && (COND)) {                                                                                     \
          const double dt_ = tb.time[i] - (OFF);                                                                        \
          const double sd_ = vt_by_jt(VB, AB, JB, dt_), sdd_ = at_by_jt(AB, JB, dt_);                                   \
          const double ea_ = fmax(fabs(sdd_) - p.thr_lon_a, 0.0), ev_ = (p.vmax - sd_) * p.inv_vmax;                    \
          sa = sa + ea_ * ea_; sj = sj + ej2_; sv = sv + ev_ * ev_;                                                      \
          i += L;                                                                                                       \
        }                                                                                                               \
      }
      BF_SEGMENT(p.v0, p.a0, j1, 0.0, tb.time[i] < b1)
      BF_SEGMENT(v1, a1, j2, t1, tb.time[i] < b2)
      BF_SEGMENT(v2, a2, j3, t1 + t2, tb.time[i] < b3)
      BF_SEGMENT(v3, a3, j4, t1 + t2 + t3, tb.time[i] < b4)
      BF_SEGMENT(v4, a4, j5, t1 + t2 + t3 + t4, true)
#undef BF_SEGMENT
      // s(T) - s(0): the sum of the five segments' displacements (the arc length is continuous across the segment boundaries, so the
      // branch the reference takes at a boundary does not matter to the bound); one lane per seed
      if (q == 0) {
        const double t5 = tb.time[p.P - 1] - t1 - t2 - t3 - t4;
        if (t5 >= 0.0) {
          sl = (p.v0 * t1 + 0.5 * p.a0 * t1 * t1 + (1 / 6.0) * j1 * t1 * t1 * t1) + (v1 * t2 + 0.5 * a1 * t2 * t2 + (1 / 6.0) * j2 * t2 * t2 * t2) +
               (v2 * t3 + 0.5 * a2 * t3 * t3 + (1 / 6.0) * j3 * t3 * t3 * t3) + (v3 * t4 + 0.5 * a3 * t4 * t4 + (1 / 6.0) * j4 * t4 * t4 * t4) +
               (v4 * t5 + 0.5 * a4 * t5 * t5 + (1 / 6.0) * j5 * t5 * t5 * t5);
        } else {                                       // segments longer than the horizon: the reference's own expression
          LonProfile lp;
          lon_profile_init(lp, seed, p.s0, p.v0, p.a0);
          double d1, d2, d3;
          lon_eval(lp, p.s0, p.v0, p.a0, tb.time[p.P - 1], sl, d1, d2, d3);
          sl = sl - p.s0;
        }
      }
    }
#pragma unroll
    for (int o = 1; o < L; o <<= 1) {
      sa = sa + __shfl_xor_sync(0xffffffffu, sa, o); sj = sj + __shfl_xor_sync(0xffffffffu, sj, o); sv = sv + __shfl_xor_sync(0xffffffffu, sv, o);
      sf = sf + __shfl_xor_sync(0xffffffffu, sf, o); sl = sl + __shfl_xor_sync(0xffffffffu, sl, o);
    }
    if (k < Ns) {
      for (int j = q; j < p.W; j += L) {
        const int n = j * Ns + k;                                        // width-major candidate index (jerk_space_sampling_planner.py:106,124-125)
        const double* lc = latcost + j * 4;
        const double lb = p.cw[1] * (1.0 - (sl - sf) * inv_md) +
                          (p.cw[2] * lc[0] * inv_hw2 + p.cw[3] * lc[1] + p.cw[4] * lc[2] + p.cw[5] * sa + p.cw[6] * sj + p.cw[7] * sv) * invP;
        keys[(size_t)(n % nwarps) * slice + n / nwarps] = ((unsigned long long)bf_key32(lb) << 32) | (unsigned int)n;
      }
    }
  }
  (void)N;
  __syncthreads();
}

// Collision row `r` of sample i for the lane that owns it, on the TRANSPOSED window winT[om][RP] (lanes own consecutive rows: the
// 16-byte entries of one obstacle state are contiguous across the lanes - no bank conflicts).  Same tests, order and semantics as
// the full-row path of TileSink::collide.
__device__ __forceinline__ void bf_collide(const EvalParams& p, const TileSink& sink, const float4* winT, int RP, int r, int i, double x, double y,
                                           double hdx, double hdy, float hqf, bool& collided, double& risk) {
  if (collided || p.n_dict <= 0 || i >= p.tcap) return;
  const int abs_step = p.now + i;
  if (abs_step >= p.Tt) return;
  float cf = 1.0f, snf = 0.0f;
  if (hqf > 1e-16f && hqf < 1e30f) { const float inv = rsqrtf(hqf); cf = (float)hdx * inv; snf = (float)hdy * inv; }
  else { const double ds = hypot(hdx, hdy); if (ds > 0.0) { cf = (float)(hdx / ds); snf = (float)(hdy / ds); } }
  const float fx = (float)(x - p.ox), fy = (float)(y - p.oy);
  const float m2x = -2.0f * fx, m2y = -2.0f * fy, ne2 = -fmaf(fx, fx, fy * fy);
  const float4* col = winT + r;
  for (int base = 0; base < p.OM; base += 32) {
    unsigned int mask = 0u;
    const int nj = min(32, p.OM - base);
#pragma unroll 4
    for (int q = 0; q < nj; ++q) {
      const float4 b = col[(size_t)(base + q) * RP];
      mask |= (fmaf(b.x, m2x, fmaf(b.y, m2y, b.z)) <= ne2) ? (1u << q) : 0u;
    }
    while (mask) {      // rare: bounding circles overlap -> separating-axis test
      const int om = base + __ffs(mask) - 1;
      mask &= mask - 1u;
      const float4 b = col[(size_t)om * RP];
      if (sink.narrow_hit_lazy(abs_step, om, b.x - fx, b.y - fy, x, y, cf, snf, hdx, hdy)) {
        if ((double)b.w >= p.p_min) { collided = true; return; }
        risk = risk + (double)b.w;
      }
    }
  }
}

// step 2: one warp checks candidate n (all lanes call).  Returns true for a survivor and its exact cost (warp-uniform).
__device__ __forceinline__ bool bestfirst_check(const EvalParams& p, const BlockTables& tb, const double* latcost, const double* seeds, int Ns,
                                                const float4* winT, int RP, int n, double* wscr, double& cost_out,
                                                unsigned long long* prof_ns = nullptr) {
  const unsigned int full = 0xffffffffu;
#ifdef JSSP_PROFILE_PHASES
  unsigned long long tq_ = 0;
  if (prof_ns && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tq_));
#define BF_STAMP(k) do { __syncwarp(); if (prof_ns && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); prof_ns[k] += t_ - tq_; tq_ = t_; } } while (0)
#else
#define BF_STAMP(k)
#endif
  const int lane = threadIdx.x & 31, P = p.P;
  const int j = n / Ns, k = n - j * Ns;
  double* prof = wscr;
  double2* xy = reinterpret_cast<double2*>(wscr + TILE_PROF);
  double* terms = wscr + TILE_PROF + 2 * P;
  const double* lat = tb.lat + (size_t)j * P * 4;
  if (lane == 0) tile_profile_store(prof, seeds + (size_t)k * 10, p.s0, p.v0, p.a0);
  __syncwarp();
  BF_STAMP(4);
  // Stage h = samples [32h, 32h + 32): Frenet state, cost terms, position (the arithmetic of the tiled walk's phase 1 and of its
  // first phase-2 step), then the segments whose points are now known: lane l owns the segment p_{i+1} - p_i with i = 32h - 1 + l -
  // curvature against the segment before it, the collision row of sample i and, for the last segment, the row of the last point
  // (which reuses the segment's heading, jssp.py:158-160).  A candidate vetoed in one stage never runs the next.
  int n_pts = P, kseg = 0;
  bool fail = false, coll = false;
  double risk = 0.0;
  for (int base = 0; base < P; base += 32) {
    {
      const int i = base + lane;
      bool off = false;
      if (i < P) {
        double s, sd, sdd, sddd;
        tile_lon_eval(prof, kseg, tb.time[i], s, sd, sdd, sddd);
        const double ea = fmax(fabs(sdd) - p.thr_lon_a, 0.0);
        const double ej = fmax(fabs(sddd) - p.thr_lon_j, 0.0);
        const double ev = (p.vmax - sd) * p.inv_vmax;
        double* tm = terms + i * 3;
        tm[0] = ea * ea; tm[1] = ej * ej; tm[2] = ev * ev;
        if (i == 0) prof[28] = s;
        if (i == P - 1) prof[29] = s;
        if (n_pts == P) {
          if (spline_in_range(tb.sp, s)) {
            double ix, iy, u, v;
            frenet_base_fma(tb.sp, spline_walk(tb.sp, s, 0), s, ix, iy, u, v);
            const double d = lat[i * 4];
            xy[i] = make_double2(fma(-d, u, ix), fma(d, v, iy));
          } else {
            off = true;                                                  // the Cartesian path ends here (jssp.py:138-139)
          }
        }
      }
      const unsigned int m = __ballot_sync(full, off);
      if (m != 0u && n_pts == P) n_pts = base + __ffs(m) - 1;
    }
    __syncwarp();
    BF_STAMP(5);
    if (base == 0 && n_pts == 1 && p.n_dict > 0 && p.tcap >= 1) coll = true;     // bare except -> collision (:250-254)
    const int i = base - 1 + lane, known = min(n_pts, base + 32);                // points [0, known) exist
    if (i >= 0 && i + 1 < known) {
      TileSink sink(p, tb);
      const double2 p0 = xy[i], p1 = xy[i + 1];
      const double bx = p1.x - p0.x, by = p1.y - p0.y;
      float bq, bds;
      TileSink::seg_lengths(bx, by, bq, bds);
      if (i >= 1) {
        const double2 pm = xy[i - 1];
        const double ax = p0.x - pm.x, ay = p0.y - pm.y;
        float aq, ads;
        TileSink::seg_lengths(ax, ay, aq, ads);
        int r = sink.curvature_classify(ax, ay, aq, ads, bx, by, bq, bds);
        if (r == 2) r = TileSink::curvature_redecide(p.kmax, ax, ay, bx, by) ? 1 : 0;
        fail = fail || r != 0;
      }
      const int cr = p.check_res;
      if (i % cr == 0) bf_collide(p, sink, winT, RP, i / cr, i, p0.x, p0.y, bx, by, bq, coll, risk);
      if (i == n_pts - 2 && (n_pts - 1) % cr == 0) bf_collide(p, sink, winT, RP, (n_pts - 1) / cr, n_pts - 1, p1.x, p1.y, bx, by, bq, coll, risk);
    }
    BF_STAMP(6);
    if (__any_sync(full, fail || coll)) return false;
  }
  if (__any_sync(full, risk != 0.0)) {
    // soft hits (obstacle modes below p_min): their probabilities are added in (row, list) order, as the serial walk does - rare
    risk = 0.0;
    if (lane == 0) {
      TileSink sink(p, tb);
      const int cr = p.check_res;
      bool c2 = false;
      for (int i = 0; i + 1 < n_pts; ++i) {
        const double2 p0 = xy[i], p1 = xy[i + 1];
        const double bx = p1.x - p0.x, by = p1.y - p0.y;
        float bq, bds;
        TileSink::seg_lengths(bx, by, bq, bds);
        if (i % cr == 0) bf_collide(p, sink, winT, RP, i / cr, i, p0.x, p0.y, bx, by, bq, c2, risk);
        if (i == n_pts - 2 && (n_pts - 1) % cr == 0) bf_collide(p, sink, winT, RP, (n_pts - 1) / cr, n_pts - 1, p1.x, p1.y, bx, by, bq, c2, risk);
      }
    }
    risk = __shfl_sync(full, risk, 0);
  }
  // exact cost: three lanes add one chain each, strictly in time order
  double acc = 0.0;
  if (lane < 3) {
    const double* tm = terms + lane;
    int i = 0;
    for (; i + 4 <= P; i += 4) {                  // the loads of four samples in flight, the additions in order
      const double t0 = tm[i * 3], t1 = tm[i * 3 + 3], t2 = tm[i * 3 + 6], t3 = tm[i * 3 + 9];
      acc = acc + t0; acc = acc + t1; acc = acc + t2; acc = acc + t3;
    }
    for (; i < P; ++i) acc = acc + tm[i * 3];
  }
  const double sa = __shfl_sync(full, acc, 0), sj = __shfl_sync(full, acc, 1), sv = __shfl_sync(full, acc, 2);
  const double cost = candidate_cost(p, sa, sj, sv, prof[28], prof[29], risk, latcost + j * 4);
  cost_out = cost;
  BF_STAMP(7);
  return isfinite(cost);
}

// The whole selection (all threads of the block call; ends with block-uniform results in best_cost / best_idx).  `winT` = the cycle's
// obstacle window, transposed ([OMpad][RP] entries; see bf_collide).
__device__ __forceinline__ void bestfirst_select(const EvalParams& p, const BlockTables& tb, const double* latcost, const double* seeds, int Ns,
                                                 const float4* winT, int RP, unsigned char* smem_bf, const BestFirstSmem& L, double& best_cost,
                                                 int& best_idx, unsigned long long* prof_ns = nullptr) {
  const int nt = blockDim.x, nwarps = nt >> 5, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int N = Ns * p.W;
  const int slice = (N + nwarps - 1) / nwarps;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_bf + L.off_keys);
  unsigned long long* shared_best = reinterpret_cast<unsigned long long*>(smem_bf + L.off_red);   // orderable bits of the best exact cost so far
  double* red_cost = reinterpret_cast<double*>(shared_best + 2);                                   // [nwarps]
  int* red_idx = reinterpret_cast<int*>(red_cost + nwarps);                                        // [nwarps]
  double* wscr = reinterpret_cast<double*>(smem_bf + L.off_warp) + (size_t)wid * bestfirst_warp_doubles(p.P);
  best_cost = INFINITY; best_idx = -1;
  if (N <= 0) return;
  unsigned long long t0_ = 0;
  if (prof_ns && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0_));
  if (threadIdx.x == 0) *shared_best = orderable_bits(INFINITY);
  bestfirst_keys(p, tb, latcost, seeds, Ns, keys, slice);          // ends with a block barrier
  if (prof_ns && threadIdx.x == 0) { unsigned long long t1_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1_)); prof_ns[0] += t1_ - t0_; }
  unsigned long long* mine = keys + (size_t)wid * slice;
  double wcost = INFINITY;                     // this warp's best survivor
  int widx = -1;
  for (;;) {
    unsigned long long pick = ~0ull;
    for (int m = lane; m < slice; m += 32) { const unsigned long long v = mine[m]; pick = v < pick ? v : pick; }
    pick = bf_warp_min(pick);
    if (pick == ~0ull) break;                                                       // the slice is exhausted
    const unsigned int k32 = (unsigned int)(pick >> 32);
    const unsigned long long seen = *reinterpret_cast<volatile unsigned long long*>(shared_best);
    if (k32 != 0u && orderable_bits((double)bf_key32_value(k32)) > seen) break;     // no candidate of this slice can cost less any more
    const int n = (int)(unsigned int)pick;
    double c;
    const bool ok = bestfirst_check(p, tb, latcost, seeds, Ns, winT, RP, n, wscr, c, prof_ns);
    if (prof_ns && threadIdx.x == 0) { prof_ns[1] += 1; prof_ns[2] += 1; }
    if (ok) {
      if (lex_less(c, n, wcost, widx)) { wcost = c; widx = n; }
      if (lane == 0) atomicMin(shared_best, (unsigned long long)orderable_bits(c));
    }
    __syncwarp();
    if (lane == 0) mine[n / nwarps] = ~0ull;
    __syncwarp();
  }
  if (lane == 0) { red_cost[wid] = wcost; red_idx[wid] = widx; }
  __syncthreads();
  double bc = INFINITY;
  int bi = -1;
  for (int w = 0; w < nwarps; ++w) {
    const double c = red_cost[w];
    const int ix = red_idx[w];
    if (lex_less(c, ix, bc, bi)) { bc = c; bi = ix; }
  }
  best_cost = bc; best_idx = bi;
}

}  // namespace jssp
