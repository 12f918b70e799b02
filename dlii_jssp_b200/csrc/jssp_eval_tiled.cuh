// jssp_eval_tiled.cuh - K2 (default for large cycles and inside the persistent replay kernel): the TILED two-phase walk.
//
// The W candidates that share a longitudinal seed (one per lateral target) sit in W consecutive lanes, a warp holds
// floor(32 / W) seeds, and the horizon is processed in tiles of TL = R * W samples.  Per tile:
//   phase 1  (parallel over samples, no dependent chain): lane j of a seed's group evaluates the SEED-ONLY quantities of the
//            samples j, j + W, ... of the tile - unrolled s, s_d, s_dd, s_ddd from the seed's segment table
//            (polyvt_sampling.py:311-385), the three longitudinal cost terms, the spline segment, the reference point and its
//            unit normal (jerk_space_sampling_planner.py:131-145) - and stores (ix, iy, u, v) in the warp's own shared-memory
//            tile; the cost terms stay in per-lane partial sums;
//   phase 2  (one lane per candidate): the lane adds its own lateral offset, forms the forward difference, its length and unit
//            vector, tests the curvature (:155-180) and, on every check_res-th sample, the collision row (:233-286).
// Only __syncwarp() separates the phases: warps are decoupled, one warp's phase 1 (throughput) overlaps another's phase 2
// (latency).  Compared with the shuffle form (walk_group) nothing seed-only sits on the candidate's serial chain any more and
// the 15 SHFL per sample are gone.  Seed-only arithmetic keeps the reference's unfused operand order and the cost terms are
// added strictly in time order (one lane per seed, from the tile), so s, s_d, s_dd, s_ddd and the COSTS are bit-identical to the
// other kernels' and to the oracle's sequential sums - near-ties select the same candidate.  The spline evaluation and phase 2
// use explicit fused multiply-adds and float32 with float64 guard-band re-decisions (see TileSink): masks unchanged.
#pragma once
#include "jssp_eval.cuh"

namespace jssp {

constexpr int TILE_PROF = 32;      // doubles per seed in the segment table: 5 x (s, v, a, j) | t1..t4 | 4 bounds | s(0), s(P-1) | pad

struct TileGeom {
  int gpw;          // seeds (groups) per warp
  int R;            // phase-1 rounds per tile
  int TL;           // samples per tile = R * W
  int grp_stride;   // doubles between the tiles of two groups: TL entries of 4 doubles + two carry entries + 2 (bank spread)
  int off_terms;    // doubles from the warp's scratch to its cost-term table [gpw][TL][3]
  int off_mask;     // ... to its collision masks [gpw][TL] (64-bit)
  int off_prof;     // ... to its segment tables [gpw][TILE_PROF]
  int warp_doubles; // doubles of scratch per warp
};
constexpr int TILE_MAX_ROUNDS = 4;
__host__ __device__ inline int tile_rounds_wanted(int W) { const int r = (16 + W - 1) / W; return r < TILE_MAX_ROUNDS ? r : TILE_MAX_ROUNDS; }   // ~16 samples per tile
__host__ __device__ inline TileGeom tile_geom(int W, int rounds = 0) {
  TileGeom g;
  g.gpw = 32 / W;
  const int r_want = tile_rounds_wanted(W), r_cap = 80 / (g.gpw * W);   // default: at most ~80 entries per warp
  g.R = r_want < r_cap ? r_want : r_cap;
  if (rounds > 0) g.R = rounds;                                        // caller's choice (shared-memory budget of the launch)
  if (g.R < 1) g.R = 1;
  g.TL = g.R * W;
  g.grp_stride = g.TL * 4 + 8 + 2;
  g.off_terms = g.gpw * g.grp_stride;
  g.off_mask = g.off_terms + g.gpw * g.TL * 3;
  g.off_prof = g.off_mask + g.gpw * g.TL;
  g.off_prof = (g.off_prof + 1) & ~1;
  g.warp_doubles = g.off_prof + g.gpw * TILE_PROF;
  return g;
}
// scratch of a block behind its tables: the warps' tiles, the circle-enlargement table adj[OMpad] (float) and the candidates'
// lateral offsets latf[W][P] (float2: d_i, d_i - d_{i-1})
__host__ __device__ inline size_t tile_scratch_bytes(int W, int threads, int rounds = 0, int OMpad = 0, int P = 0) {
  return (size_t)(threads / 32) * tile_geom(W, rounds).warp_doubles * 8 + (((size_t)OMpad * 4 + 15) & ~(size_t)15) + (size_t)W * P * 8;
}

// segment table of one seed: the start state and jerk of each of the five segments (the same expressions as lon_profile_init:
// bit-identical to the reference's segment end states), the durations and the asymmetric-epsilon boundaries of
// get_one_frenet_trajectory's branch chain (polyvt_sampling.py:334-374)
__device__ __forceinline__ void tile_profile_store(double* prof, const double* seed_row, double s0, double v0, double a0) {
  LonProfile lp;
  double seed[10];
  load_seed_row(seed, seed_row);
  lon_profile_init(lp, seed, s0, v0, a0);
  prof[0] = s0; prof[1] = v0; prof[2] = a0; prof[3] = lp.j1;
  prof[4] = lp.s1; prof[5] = lp.v1; prof[6] = lp.a1; prof[7] = lp.j2;
  prof[8] = lp.s2; prof[9] = lp.v2; prof[10] = lp.a2; prof[11] = lp.j3;
  prof[12] = lp.s3; prof[13] = lp.v3; prof[14] = lp.a3; prof[15] = lp.j4;
  prof[16] = lp.s4; prof[17] = lp.v4; prof[18] = lp.a4; prof[19] = lp.j5;
  prof[20] = lp.t1; prof[21] = lp.t2; prof[22] = lp.t3; prof[23] = lp.t4;
  prof[24] = lp.t1 - 1e-9; prof[25] = lp.t1 + lp.t2 + 1e-9; prof[26] = lp.t1 + lp.t2 + lp.t3 + 1e-9;
  prof[27] = lp.t1 + lp.t2 + lp.t3 + lp.t4 + 1e-9;
}

// s, s_d, s_dd, s_ddd at time t from the segment table; `k` = segment of the previous (earlier) sample of this lane, or 0
__device__ __forceinline__ void tile_lon_eval(const double* prof, int& k, double t, double& s, double& sd, double& sdd, double& sddd) {
  while (k < 4 && !(t < prof[24 + k])) ++k;
  const double* q = prof + 4 * k;
  const double sb = q[0], vb = q[1], ab = q[2], jb = q[3];
  double dt = t;
  if (k >= 1) dt = dt - prof[20];
  if (k >= 2) dt = dt - prof[21];
  if (k >= 3) dt = dt - prof[22];
  if (k >= 4) dt = dt - prof[23];
  s = st_by_jt(sb, vb, ab, jb, dt);
  sd = vt_by_jt(vb, ab, jb, dt);
  sdd = at_by_jt(ab, jb, dt);
  sddd = jb;
}

// the seed-only part of frenet_point with fused multiply-adds: reference point (ix, iy) and u = y'/|r'|, v = x'/|r'|
__device__ __forceinline__ void frenet_base_fma(const SplineView& sp, int seg, double s, double& ix, double& iy, double& u, double& v) {
  const double dx = s - sp.knot[seg];
  const double* c = sp.coef + seg;
  const int kp = sp.kpad;
  const double c1 = c[kp], c2 = c[2 * kp], c3 = c[3 * kp], c5 = c[5 * kp], c6 = c[6 * kp], c7 = c[7 * kp];
  ix = fma(fma(fma(c3, dx, c2), dx, c1), dx, c[0]);
  iy = fma(fma(fma(c7, dx, c6), dx, c5), dx, c[4 * kp]);
  const double tx = fma(fma(3.0 * c3, dx, 2.0 * c2), dx, c1);
  const double ty = fma(fma(3.0 * c7, dx, 2.0 * c6), dx, c5);
  const double inv = rsqrt_fast(fma(tx, tx, ty * ty));
  u = ty * inv;
  v = tx * inv;
}

// arc length of the seed at time t (list building: the block's range per collision row)
struct TileSAt {
  const double* prof;
  __device__ __forceinline__ double operator()(double t) const {
    int k = 0;
    double s, sd, sdd, sddd;
    tile_lon_eval(prof, k, t, s, sd, sdd, sddd);
    return s;
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// Phase-2 arithmetic.  The curvature test needs the angle between consecutive chord vectors, i.e. SECOND differences of the
// positions: from absolute positions that takes float64.  Here the differences are formed where they are exact and the rest
// is float32:
//   phase 1a (per seed, float64)  reference point and unit normal (ix, iy, u, v) of every sample, as before;
//   phase 1b (per seed)           per segment (sample i-1 -> i): the DIFFERENCES dix, diy, du, dv computed in float64 and rounded
//                                 to float32 (relative error 6e-8 of the difference itself), plus u, v and the reference point
//                                 relative to the origin of the collision frame at sample i: one 32-byte record;
//   phase 2 (per candidate, float32 only)  chord  dx = dix - (dd * u1 + d0 * du),  dy = diy + (dd * v1 + d0 * dv)  with the lateral
//                                 offset d0 and its difference dd (float64 difference rounded) - no cancellation, relative error
//                                 ~1e-6 of |chord|; length by MUFU.RSQ; cross / dot of consecutive chords; the comparison
//                                 |cross| > sin(kmax * |p|) * |p| * |d|  (<=> |angle| > phi for |angle| < pi/2).
//   The decision is taken in float32 only outside a guard band |(|cross| - rhs)| > 6e-6 |p||d| + 1e-5 rhs (the float32 error of the
//   cross product is <= ~3e-6 |p||d|); inside it, for chords shorter than 1e-8 m, for phi >= 0.49 and when the un-unwrapped heading
//   difference may jump by 2*pi, the three points are recomputed FROM SCRATCH in float64 with the other kernels' arithmetic
//   (tile_lon_eval + frenet_point) and the oracle's own expression decides (out of line, rare).  Masks therefore equal the
//   float64 evaluation except on measure-zero inputs.
// Collision rows: positions relative to the collision frame's origin in float32 (as the broad / narrow phases always were); the
// bounding-circle test runs ONCE PER SEED in phase 1.5 (circles enlarged by the lateral reach dmax) into a 64-bit mask per (seed,
// row); phase 2 visits the set bits with the candidate's own circle test and the separating-axis classification; inside the
// SAT guard band the pose is recomputed in float64 from scratch.
// ---------------------------------------------------------------------------------------------------------------------
struct SegRec { float dix, diy, du, dv, u1, v1, fx1, fy1; };      // segment ending at a sample: differences + the end point's frame
static_assert(sizeof(SegRec) == 32, "one tile entry");

__device__ __forceinline__ float circle_enlarge(const EvalParams& p, int om, float dmax) {
  // pack_obstacles_kernel: R = r_ego + r_obs + BROAD_PAD, z = |o'|^2 - (R^2 * 1.0001 + BROAD_CANCEL);  (R + dmax)^2 = R^2 + 2 R dmax + dmax^2
  const double a2 = p.ext64[2 * om], b2 = p.ext64[2 * om + 1];
  const float r = (float)(sqrt(p.ea * p.ea + p.eb * p.eb) + sqrt(a2 * a2 + b2 * b2)) + 0.05f;
  return (2.0f * r * dmax + dmax * dmax) * 1.0002f + 1e-4f;
}
// adj[om] for every obstacle mode of the cycle (block-wide; the caller synchronises)
__device__ __forceinline__ void fill_circle_adj(const EvalParams& p, float dmax, float* adj) {
  for (int om = threadIdx.x; om < p.OMpad; om += blockDim.x) adj[om] = om < p.OM ? circle_enlarge(p, om, dmax) : 0.f;
}
// latf[w][i] = (d_i, d_i - d_{i-1}) in float32, the difference taken in float64 (block-wide; the caller synchronises)
__device__ __forceinline__ void fill_lateral_f32(const double* sm_lat, int W, int P, float2* latf) {
  for (int e = threadIdx.x; e < W * P; e += blockDim.x) {
    const int i = e % P;
    const double d = sm_lat[(size_t)e * 4];
    latf[e] = make_float2((float)d, i >= 1 ? (float)(d - sm_lat[(size_t)(e - 1) * 4]) : 0.f);
  }
}

struct TileSink {
  const EvalParams& p;
  const BlockTables& tb;
  const double* prof;                                    // the seed's segment table   } for the float64 re-decisions
  const double* lat64;                                   // the candidate's lateral samples [P][4] }
  double sum_a = 0, sum_j = 0, sum_v = 0, s_first = 0, s_last = 0, risk = 0;
  float pdx = 0.f, pdy = 0.f, pq = 0.f, pds = 0.f;       // previous chord, its squared length and length
  float px = 0.f, py = 0.f;                             // the candidate at the previous sample (collision frame)
  float kmaxf;
  bool curv_fail = false, collided = false;
  __device__ __forceinline__ TileSink(const EvalParams& p_, const BlockTables& tb_, const double* prof_, const double* lat_)
      : p(p_), tb(tb_), prof(prof_), lat64(lat_), kmaxf((float)p_.kmax) {}
  __device__ __forceinline__ bool vetoed() const { return p.select_only && (curv_fail || collided); }

  // float64 from scratch, with the arithmetic of the other kernels (lon_eval + frenet_point): sample i of this candidate
  __device__ __noinline__ static void point_exact(const BlockTables& tb, const double* prof, const double* lat64, int i, double& x, double& y) {
    int k = 0;
    double s, sd, sdd, sddd;
    tile_lon_eval(prof, k, tb.time[i], s, sd, sdd, sddd);
    frenet_point(tb.sp, spline_walk(tb.sp, s, 0), s, lat64[i * 4], x, y);
  }
  // the float64 decision for the chords (sg-1 -> sg) and (sg -> sg+1), exactly as the other kernels' guard-band path:
  // c = (yaw_sg - yaw_{sg-1}) / ds_{sg-1} without unwrapping (jerk_space_sampling_planner.py:155-163, :175)
  __device__ __noinline__ static bool curvature_exact_at(const EvalParams& p, const BlockTables& tb, const double* prof, const double* lat64, int sg) {
    double x0, y0, x1, y1, x2, y2;
    point_exact(tb, prof, lat64, sg - 1, x0, y0);
    point_exact(tb, prof, lat64, sg, x1, y1);
    point_exact(tb, prof, lat64, sg + 1, x2, y2);
    const double pdx = x1 - x0, pdy = y1 - y0, dx = x2 - x1, dy = y2 - y1;
    if (dx == 0.0 && dy == 0.0 && pdx == 0.0 && pdy == 0.0) return false;      // standstill: both headings atan2(0, 0) = 0; 0 / 0 = nan passes
    const double q = pdx * pdx + pdy * pdy;
    const double pds = (q > 1e-200 && q < 1e200) ? q * rsqrt_fast(q) : hypot(pdx, pdy);
    const double dyaw = heading_of(dx, dy) - heading_of(pdx, pdy);
    return fabs(dyaw) > p.kmax * pds;                  // pds == 0: inf fails (dyaw != 0), nan passes (dyaw == 0)
  }

  // chord (dx, dy), squared length q, length ds follows the stored previous chord; sg = the sample where the two chords meet
  __device__ __forceinline__ bool curvature_fails_fast(int sg, float dx, float dy, float q, float ds) const {
    const float phi = kmaxf * pds;
    const float cross = fmaf(pdx, dy, -(pdy * dx)), dot = fmaf(pdx, dx, pdy * dy);
    const float p2 = phi * phi;
    const float sp = phi * fmaf(p2, fmaf(p2, fmaf(p2, -1.0f / 5040.0f, 1.0f / 120.0f), -1.0f / 6.0f), 1.0f);
    const float len2 = pds * ds;
    const float lhs = fabsf(cross), rhs = sp * len2;
    const bool ranges = (pq > 1e-16f) & (pq < 1e30f) & (q > 1e-16f) & (q < 1e30f) & (phi < 0.49f);
    // The un-unwrapped heading difference (atan2 of each chord) jumps by 2*pi when the heading crosses +-pi, i.e. when dy changes
    // sign while the chords point into the left half plane; then |dyaw| = 2*pi - |angle| >= pi > phi: the candidate fails.  With
    // a sign change and one chord not clearly in the right half plane the angle between the chords is >= ~pi/2 - 1e-3 > phi: it
    // fails as well.  Only when a sign of dy is itself uncertain in float32 (|dy| <= 1e-5 |dx|) does float64 have to look.
    const bool right_clear = (dx > 1e-3f * ds) & (pdx > 1e-3f * pds);
    const bool tiny_dy = (fabsf(dy) <= 1e-5f * fabsf(dx)) | (fabsf(pdy) <= 1e-5f * fabsf(pdx));
    const bool sdiff = signbit(dy) != signbit(pdy);
    const bool obtuse = !(dot > 0.0f);                                                   // |angle| >= pi/2 > phi
    const bool clear = fabsf(lhs - rhs) > fmaf(6e-6f, len2, 1e-5f * rhs);
    if (ranges & (right_clear | !tiny_dy)) {
      if (sdiff & !right_clear) return true;
      if (obtuse | clear) return obtuse | (lhs > rhs);
    }
    return curvature_exact_at(p, tb, prof, lat64, sg);
  }

  // separating-axis test at sample i with the float32 pose; inside the guard band the pose (position of sample i, heading of the
  // chord hs -> hs + 1) is recomputed in float64 from scratch
  __device__ __noinline__ static bool narrow_hit_at(const EvalParams& p, const BlockTables& tb, const double* prof, const double* lat64,
                                                    int abs_step, int om, float dxo, float dyo, float cf, float snf, int i, int hs) {
    const size_t gi = (size_t)abs_step * p.OM + om;
    const float4 nb = __ldg(p.narrow + gi);
    const int cls = sat_classify_f32(dxo, dyo, cf, snf, (float)p.ea, (float)p.eb, nb.x, nb.y, nb.z, nb.w, SAT_EPS);
    if (cls != 0) return cls < 0;
    double x, y, hx0, hy0, hx1, hy1;
    point_exact(tb, prof, lat64, i, x, y);
    point_exact(tb, prof, lat64, hs, hx0, hy0);
    point_exact(tb, prof, lat64, hs + 1, hx1, hy1);
    const double hdx = hx1 - hx0, hdy = hy1 - hy0;
    const double q = hdx * hdx + hdy * hdy;
    double c = 1.0, sn = 0.0;
    if (q > 1e-200 && q < 1e200) { const double inv = rsqrt_fast(q); c = hdx * inv; sn = hdy * inv; }
    else { const double ds = hypot(hdx, hdy); if (ds > 0.0) { c = hdx / ds; sn = hdy / ds; } }
    const double* o = p.obs64 + gi * 4;
    return sat_intersects_f64(o[0] - x, o[1] - y, c, sn, p.ea, p.eb, o[2], o[3], p.ext64[2 * om], p.ext64[2 * om + 1]);
  }

  // one obstacle state whose enlarged circle the seed's reference point touched (b = its table entry): the candidate's own
  // circle test, then the separating-axis test.  Returns true when the candidate is vetoed.
  __device__ __forceinline__ bool visit(int abs_step, int om, float4 b, float fx, float fy, float ne2, float cf, float snf, int i, int hs) {
    if (!(fmaf(b.x, -2.0f * fx, fmaf(b.y, -2.0f * fy, b.z)) <= ne2)) return false;
    if (!narrow_hit_at(p, tb, prof, lat64, abs_step, om, b.x - fx, b.y - fy, cf, snf, i, hs)) return false;
    const float pr = tb.use_lists ? __ldg(p.probf + om) : b.w;
    if ((double)pr >= p.p_min) { collided = true; return true; }
    risk = risk + (double)pr;
    return false;
  }

  // collision row r of sample i: pose (fx, fy) in the collision frame, heading of the chord hs -> hs + 1 given as (hdx, hdy, hq);
  // `mask` = the seed's hits of this row (bit k = entry k of the row's list / table row)
  __device__ __forceinline__ void collide(int i, int r, float fx, float fy, float hdx, float hdy, float hq, int hs, unsigned long long mask) {
    if (collided || p.n_dict <= 0 || i >= p.tcap) return;
    const int abs_step = p.now + i;
    if (abs_step >= p.Tt) return;
    const int cnt = tb.use_lists ? tb.row_cnt[r] : p.OM;
    const float4* e = tb.use_lists ? tb.list + tb.row_off[r] : tb.broad + (size_t)(tb.broad_row0 + r * tb.broad_stride) * p.OMpad;
    const bool by_mask = tb.masks && cnt <= 64;
    if (by_mask && mask == 0ull) return;                // the usual case: nothing near any candidate of this seed
    float cf = 1.0f, snf = 0.0f;                        // zero-length chord: atan2(0, 0) = 0
    if (hq > 1e-30f) { const float inv = rsqrtf(hq); cf = hdx * inv; snf = hdy * inv; }
    else if (hdx != 0.f || hdy != 0.f) {                // too short for float32 squares: normalise after scaling
      const float sc = 1.0f / fmaxf(fabsf(hdx), fabsf(hdy)), ax = hdx * sc, ay = hdy * sc, inv = rsqrtf(fmaf(ax, ax, ay * ay));
      cf = ax * inv; snf = ay * inv;
    }
    const float ne2 = -fmaf(fx, fx, fy * fy);
    if (by_mask) {
      while (mask) {
        const int k = __ffsll((long long)mask) - 1;
        mask &= mask - 1ull;
        const float4 b = e[k];
        if (visit(abs_step, tb.use_lists ? __float_as_int(b.w) : k, b, fx, fy, ne2, cf, snf, i, hs)) return;
      }
      return;
    }
    // no seed-level masks (table read from global memory, or more than 64 entries in the row): every entry of the row
    for (int k = 0; k < cnt; ++k) {
      const float4 b = e[k];
      if (visit(abs_step, tb.use_lists ? __float_as_int(b.w) : k, b, fx, fy, ne2, cf, snf, i, hs)) return;
    }
  }
};
__device__ __forceinline__ double candidate_cost(const EvalParams& p, const TileSink& k, const double* latcost) {
  return candidate_cost(p, k.sum_a, k.sum_j, k.sum_v, k.s_first, k.s_last, k.risk, latcost);
}

// max |d| over the cycle's lateral samples (block-wide; ends with a block barrier): the radius by which the seed-level circles grow
__device__ __forceinline__ float block_lateral_reach(const double* sm_lat, int W, int P) {
  __shared__ unsigned int sm_bits_[1];
  unsigned int* sm_bits = sm_bits_;
  if (threadIdx.x == 0) *sm_bits = 0u;
  __syncthreads();
  float m = 0.f;
  for (int e = threadIdx.x; e < W * P; e += blockDim.x) m = fmaxf(m, __double2float_ru(fabs(sm_lat[(size_t)e * 4])));
  m = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m)));      // non-negative floats order like their bits
  if ((threadIdx.x & 31) == 0) atomicMax(sm_bits, __float_as_uint(m));
  __syncthreads();
  return __uint_as_float(*sm_bits) * 1.0001f + 1e-4f;
}

// The tiled walk of the candidate in lane (g, j): g = group (seed) within the warp, j = lateral target.  `gbase` / `gterms` /
// `gmask` = this group's tile (points -> segment records, cost terms, collision masks), `prof` = its segment table (already
// stored), all in the warp's scratch; latf = the block's float32 lateral table.  Inactive lanes (no seed) only keep the
// warp-level barriers company.  On return `sink` holds the masks, the risk term and the longitudinal cost sums.
__device__ __forceinline__ void walk_tiles(const EvalParams& p, const BlockTables& tb, const TileGeom& tg, double* gbase, double* gterms,
                                           unsigned long long* gmask, double* prof, const float2* latf, bool active, int j, int W,
                                           TileSink& sink) {
  const int P = p.P, TL = tg.TL, cr = p.check_res;
  const float2* mylat = latf + (size_t)j * P;
  SegRec* rec = reinterpret_cast<SegRec*>(gbase);                  // the records replace the float64 points in place
  double* carry = gbase + (size_t)TL * 4;                           // two entries: the last float64 point of the previous tile (ping-pong)
  int* np_s = reinterpret_cast<int*>(prof + 30);       // first sample off the reference line (P if none), filled by phase 1
  int k = 0;
  double sa = 0, sj = 0, sv = 0;                 // lane 0 of the group: the longitudinal cost sums, added strictly in time order
  bool walking = active;
  const bool rows_on = p.n_dict > 0 && tb.masks;
  unsigned long long pmask = 0ull;               // collision mask of the previous sample
  int n_pts = P;
  int next_check = 0, row = 0;                   // next sample index with a collision row and that row (i % check_res == 0 without dividing)
  const int lane = threadIdx.x & 31;
  const unsigned int grp_lanes = (W >= 32 ? 0xffffffffu : ((1u << W) - 1u)) << (lane - j);
  const float oxf = 0.f;
  (void)oxf;
  if (active && j == 0) *np_s = P;
  __syncwarp();
  int tile = 0;
  for (int base = 0; base < P; base += TL, ++tile) {
    const int len = min(TL, P - base);
    // phase 1a: seed-only float64 quantities of the samples this lane owns in the tile
    if (active) {
      for (int il = j; il < len; il += W) {
        const int i = base + il;
        double s, sd, sdd, sddd;
        tile_lon_eval(prof, k, tb.time[i], s, sd, sdd, sddd);
        const double ea = fmax(fabs(sdd) - p.thr_lon_a, 0.0);
        const double ej = fmax(fabs(sddd) - p.thr_lon_j, 0.0);
        const double ev = (p.vmax - sd) * p.inv_vmax;
        double* tm = gterms + il * 3;
        tm[0] = ea * ea; tm[1] = ej * ej; tm[2] = ev * ev;
        if (i == 0) prof[28] = s;
        if (i == P - 1) prof[29] = s;
        gmask[il] = 0ull;
        double2* e = reinterpret_cast<double2*>(gbase + il * 4);
        if (spline_in_range(tb.sp, s)) {
          double ix, iy, u, v;
          frenet_base_fma(tb.sp, spline_walk(tb.sp, s, 0), s, ix, iy, u, v);
          e[0] = make_double2(ix, iy); e[1] = make_double2(u, v);
        } else {
          atomicMin(np_s, i);                       // off the reference line: the Cartesian path ends here (jssp.py:138-139)
          e[0] = make_double2(0.0, 0.0); e[1] = make_double2(0.0, 0.0);
        }
      }
    }
    __syncwarp();
    n_pts = active ? *np_s : 0;
    const int live = min(len, n_pts - base);          // samples of the tile that are on the Cartesian path (<= 0: none)
    // the cost sums are sequential float64 sums over the samples (the oracle's order: costs bit-identical to the other kernels):
    // one lane per seed adds the tile's terms - three short independent chains
    if (active && j == 0) {
      for (int il = 0; il < len; ++il) { sa = sa + gterms[il * 3]; sj = sj + gterms[il * 3 + 1]; sv = sv + gterms[il * 3 + 2]; }
    }
    // phase 1b: the points become segment records, last round first (a record overwrites the point of its own sample; the start
    // point of a segment is read before the round that overwrites it; the tile's last point is kept for the next tile)
    const double* carry_old = carry + ((tile & 1) ? 4 : 0);
    double* carry_new = carry + ((tile & 1) ? 0 : 4);
    for (int r = tg.R - 1; r >= 0; --r) {
      const int il = j + W * r;
      const bool mine = active && il < len;
      double2 a0 = make_double2(0.0, 0.0), a1 = a0, b0 = a0, b1 = a0;
      if (mine) {
        const double2* eb = reinterpret_cast<const double2*>(gbase + il * 4);
        b0 = eb[0]; b1 = eb[1];
        const double2* ea = il >= 1 ? reinterpret_cast<const double2*>(gbase + (il - 1) * 4) : reinterpret_cast<const double2*>(carry_old);
        if (base + il >= 1) { a0 = ea[0]; a1 = ea[1]; } else { a0 = b0; a1 = b1; }
      }
      __syncwarp();
      if (mine) {
        SegRec q;
        q.dix = (float)(b0.x - a0.x); q.diy = (float)(b0.y - a0.y); q.du = (float)(b1.x - a1.x); q.dv = (float)(b1.y - a1.y);
        q.u1 = (float)b1.x; q.v1 = (float)b1.y; q.fx1 = (float)(b0.x - p.ox); q.fy1 = (float)(b0.y - p.oy);
        rec[il] = q;
        if (il == len - 1) { double2* c = reinterpret_cast<double2*>(carry_new); c[0] = b0; c[1] = b1; }
      }
    }
    __syncwarp();
    // phase 1.5: bounding circles once per seed - lane j of the group takes entries j, j + W, ... of every collision row of the tile.
    // Skipped once no candidate of the seed can use the rows any more (all collided / vetoed / off the path).
    const bool wants = walking && !sink.collided;
    const bool grp_wants = (__ballot_sync(0xffffffffu, wants) & grp_lanes) != 0u;
    if (rows_on && active && grp_wants) {
      const int il0 = (cr - base % cr) % cr;
      int r = (base + il0) / cr;
      for (int il = il0; il < live; il += cr, ++r) {
        const int i = base + il, abs_step = p.now + i;
        if (i >= p.tcap || abs_step >= p.Tt) break;
        const float fx = rec[il].fx1, fy = rec[il].fy1;
        const float m2x = -2.0f * fx, m2y = -2.0f * fy, ne2 = -fmaf(fx, fx, fy * fy);
        const int cnt = tb.use_lists ? tb.row_cnt[r] : p.OM;
        const float4* e = tb.use_lists ? tb.list + tb.row_off[r] : tb.broad + (size_t)(tb.broad_row0 + r * tb.broad_stride) * p.OMpad;
        unsigned int lo = 0u, hi = 0u;
        if (cnt <= 64) {
          for (int q = j; q < cnt; q += W) {
            const float4 b = e[q];
            const float z = b.z - tb.adj[tb.use_lists ? __float_as_int(b.w) : q];
            const bool hit = fmaf(b.x, m2x, fmaf(b.y, m2y, z)) <= ne2;
            if (q < 32) lo |= hit ? (1u << q) : 0u; else hi |= hit ? (1u << (q - 32)) : 0u;
          }
        } else {
          lo = hi = 0xffffffffu;                      // longer rows: phase 2 runs the full loop over the row (see collide)
        }
        if (lo | hi) atomicOr(gmask + il, ((unsigned long long)hi << 32) | lo);
      }
    }
    __syncwarp();
    // phase 2: this lane's candidate, float32
    if (walking) {
      for (int il = 0; il < live; ++il) {
        const int i = base + il;
        const SegRec q = rec[il];
        const float2 ld = mylat[i];                  // d_i, d_i - d_{i-1}
        const float x1 = fmaf(-ld.x, q.u1, q.fx1), y1 = fmaf(ld.x, q.v1, q.fy1);
        if (i >= 1) {
          const int sg = i - 1;                     // the chord sg -> i
          const float d0 = ld.x - ld.y;
          const float dx = q.dix - fmaf(ld.y, q.u1, d0 * q.du), dy = q.diy + fmaf(ld.y, q.v1, d0 * q.dv);
          const float qq = fmaf(dx, dx, dy * dy);
          const float ds = (qq > 1e-16f && qq < 1e30f) ? qq * rsqrtf(qq) : 0.f;
          if (sg >= 1 && sink.curvature_fails_fast(sg, dx, dy, qq, ds)) sink.curv_fail = true;
          if (sg == next_check) {                   // collision row of sample sg: its pose, the heading of this chord
            sink.collide(sg, row, sink.px, sink.py, dx, dy, qq, sg, pmask);
            next_check += cr; ++row;
          }
          sink.pdx = dx; sink.pdy = dy; sink.pq = qq; sink.pds = ds;
          if (sink.vetoed()) { walking = false; break; }     // select-only evaluation: this candidate can no longer be selected
        }
        sink.px = x1; sink.py = y1; pmask = gmask[il];
      }
      if (live < len) walking = false;             // the path ended inside this tile (or before it)
    }
    __syncwarp();                                  // the tile is rewritten by the next phase 1
  }
  if (active && !sink.vetoed()) {
    if (n_pts >= 2) {                              // the last point reuses the previous heading (:158-160)
      if (n_pts - 1 == next_check) sink.collide(n_pts - 1, row, sink.px, sink.py, sink.pdx, sink.pdy, sink.pq, n_pts - 2, pmask);
    } else if (n_pts == 1 && p.n_dict > 0 && p.tcap >= 1) sink.collided = true;   // bare except -> collision (:250-254)
  }
  if (active && j == 0) { gbase[0] = sa; gbase[1] = sj; gbase[2] = sv; }
  __syncwarp();
  if (active) {
    sink.sum_a = gbase[0]; sink.sum_j = gbase[1]; sink.sum_v = gbase[2];
    sink.s_first = prof[28]; sink.s_last = prof[29];
  }
  __syncwarp();
}

template <int kMode>
__device__ __forceinline__ void evaluate_tiled_body(const EvalParams& p, unsigned char* smem, double* red_cost, int* red_idx,
                                                    bool* is_last, int* sm_npts) {
  const int Ns = resolve_n_seeds(p);
  const int W = p.W;
  const int N = Ns * W;
  const TileGeom tg = tile_geom(W, p.tile_rounds);
  const int gpw = tg.gpw;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int g = lane / W, j = lane - g * W;
  const int spb = (blockDim.x >> 5) * gpw;     // seeds per block
  const int k = blk_index(p) * spb + wid * gpw + g;
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
    double* wscr = reinterpret_cast<double*>(smem + ((L.total + 15) & ~(size_t)15)) + (size_t)wid * tg.warp_doubles;
    const int gg = g < gpw ? g : 0;
    double* gbase = wscr + (size_t)gg * tg.grp_stride;
    double* gterms = wscr + tg.off_terms + (size_t)gg * tg.TL * 3;
    double* prof = wscr + tg.off_prof + (size_t)gg * TILE_PROF;
    unsigned long long* gmask = reinterpret_cast<unsigned long long*>(wscr + tg.off_mask) + (size_t)gg * tg.TL;
    if (active && j == 0) tile_profile_store(prof, p.seeds + (size_t)k * 10, p.s0, p.v0, p.a0);
    __syncwarp();
    // seed-level bounding circles: circles enlarged by the lateral reach of the cycle, z' = z - adj[om]
    const float dmax = block_lateral_reach(tb.lat, W, p.P);
    float* sm_adj = reinterpret_cast<float*>(reinterpret_cast<double*>(smem + ((L.total + 15) & ~(size_t)15)) + (size_t)(blockDim.x >> 5) * tg.warp_doubles);
    float2* sm_latf = reinterpret_cast<float2*>(sm_adj + ((p.OMpad + 3) & ~3));
    fill_circle_adj(p, dmax, sm_adj);
    fill_lateral_f32(tb.lat, W, p.P, sm_latf);
    tb.adj = sm_adj;
    if (kMode == 2) {
      TileSAt s_at{prof};
      build_lists(p, smem, tb, s_at, active, j, W);      // ends with a block barrier
      tb.masks = tb.use_lists && !(p.tile_flags & 1);
    } else {
      __syncthreads();
      tb.masks = kMode == 1 && !(p.tile_flags & 1);
    }
    PHASE(p, 2);
    TileSink sink(p, tb, prof, tb.lat + (size_t)j * p.P * 4);
    walk_tiles(p, tb, tg, gbase, gterms, gmask, prof, sm_latf, active, j, W, sink);
    if (active) {
      const int n = j * Ns + k;                // candidate index: width-major (jerk_space_sampling_planner.py:106,124-125)
      const double cost = candidate_cost(p, sink, latcost + j * 4);
      if (p.feasible) p.feasible[n] = sink.curv_fail ? 0 : 1;
      if (p.collision) p.collision[n] = sink.collided ? 1 : 0;
      if (p.cost) p.cost[n] = (float)cost;
      if (!sink.curv_fail && !sink.collided && isfinite(cost)) { my_cost = cost; my_idx = n; }
    }
  }
  __syncthreads();
  PHASE(p, 3);
  select_and_export<kMode>(p, smem, tb, latcost, staged, my_cost, my_idx, N, Ns, red_cost, red_idx, is_last, sm_npts);
  PHASE(p, 4);
}

template <int kMode>
__global__ void __launch_bounds__(EVAL_THREADS, EVAL_MIN_BLOCKS) evaluate_tiled_kernel(const __grid_constant__ EvalParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  evaluate_tiled_body<kMode>(p, smem, red_cost, red_idx, &is_last, &sm_npts);
}

template <int kMode>
__global__ void __launch_bounds__(BATCH_EVAL_THREADS, BATCH_MIN_BLOCKS) evaluate_tiled_batch_kernel(const EvalParams* __restrict__ slots) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(16) EvalParams sp;
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  pdl_launch_dependents(); pdl_wait();
  if (!slot_live(slots, blockIdx.x)) return;
  load_slot(&sp, slots + blockIdx.x);
  evaluate_tiled_body<kMode>(sp, smem, red_cost, red_idx, &is_last, &sm_npts);
}

}  // namespace jssp
