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
  int gpw;          // lane groups per warp (W lanes each)
  int H;            // chunks of the horizon per candidate (1: a group walks the whole horizon; see TileChunk)
  int spw;          // seeds per warp = gpw / H
  int R;            // phase-1 rounds per tile
  int TL;           // samples per tile = R * W
  int grp_stride;   // doubles between the tiles of two groups (TL * 4 + 2: neighbouring groups land 16 B apart in the banks)
  int off_terms;    // doubles from the warp's scratch to its cost-term table [gpw][TL][3]
  int off_prof;     // ... to its segment tables [gpw][TILE_PROF]
  int own_max;      // chunked: most samples whose cost terms one chunk owns
  int off_lterms;   // chunked: ... to the cost terms of the chunks after the first [spw * (H - 1)][own_max][3]
  int warp_doubles; // doubles of scratch per warp
};
__host__ __device__ inline int tile_rounds_wanted(int W) { return (16 + W - 1) / W; }     // ~16 samples per tile
__host__ __device__ inline TileGeom tile_geom(int W, int rounds = 0, int H = 1, int P = 0) {
  TileGeom g;
  g.gpw = 32 / W;
  g.H = H < 1 ? 1 : (H > g.gpw ? g.gpw : H);
  g.spw = g.gpw / g.H;
  const int r_want = tile_rounds_wanted(W), r_cap = 80 / (g.gpw * W);   // default: at most ~80 entries (~6 KB) per warp
  g.R = r_want < r_cap ? r_want : r_cap;
  if (rounds > 0) g.R = rounds;                                        // caller's choice (shared-memory budget of the launch)
  if (g.R < 1) g.R = 1;
  g.TL = g.R * W;
  g.grp_stride = g.TL * 4 + 2;
  g.off_terms = g.gpw * g.grp_stride;
  g.off_prof = g.off_terms + g.gpw * g.TL * 3;
  g.off_prof = (g.off_prof + 1) & ~1;
  g.off_lterms = g.off_prof + g.gpw * TILE_PROF;
  g.own_max = g.H > 1 ? (P - 1 + g.H - 1) / g.H + 1 : 0;
  g.warp_doubles = g.off_lterms + g.spw * (g.H - 1) * g.own_max * 3;
  g.warp_doubles = (g.warp_doubles + 1) & ~1;
  return g;
}
__host__ __device__ inline size_t tile_scratch_bytes(int W, int threads, int rounds = 0, int H = 1, int P = 0) {
  return (size_t)(threads / 32) * tile_geom(W, rounds, H, P).warp_doubles * 8;
}

// CHUNKED walk (batched replay with few scenarios per GPU: the plan cycle is a latency chain, the SMs are half empty): the H lane
// groups of a seed split the horizon.  Chunk h tests the segments [a_h, a_{h+1}) with a_h = h * (P - 1) / H, i.e. it walks the
// points [a_h - 1, a_{h+1}]: the segment before the chunk only PRIMES the curvature test (it is tested by the chunk before).
// The cost terms of the samples [a_h, a_{h+1}) (+ the last sample for the last chunk) belong to chunk h.  After the walk the
// chunks are combined IN TIME ORDER, with the serial walk's semantics: masks are OR-ed up to the chunk in which the path left
// the reference line, risk stops at the first hard collision, and the cost sums continue the previous chunk's sums term by
// term - so masks, sums and therefore the selection are those of the serial walk (the risk partial sums of the chunks are
// added chunk by chunk: exact for probabilities that are float32 values of comparable magnitude, as the tables hold them).
struct TileChunk {
  int lo, hi;             // first / last point walked
  int own_lo, own_hi;     // cost terms of the samples [own_lo, own_hi)
  bool first, last;
  double* lterms;         // chunks after the first: their cost terms [own_hi - own_lo][3]
};
__device__ __forceinline__ TileChunk tile_chunk(int h, int H, int P, double* lterms) {
  TileChunk c;
  const int a = h * (P - 1) / H, b = (h + 1) * (P - 1) / H;
  c.first = h == 0; c.last = h == H - 1;
  c.lo = a > 0 ? a - 1 : 0; c.hi = b;
  c.own_lo = a; c.own_hi = c.last ? P : b;
  c.lterms = lterms;
  return c;
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
// Phase-2 arithmetic: float64 only where cancellation needs it, float32 for the rest, float64 re-decision inside guard bands.
//   float64: the candidate's position (2 FMA), the forward difference (2 ADD), and cross / dot of two consecutive difference
//            vectors (2 MUL + 2 FMA) - these carry the angle between the segments exactly, at any speed;
//   float32: segment lengths (MUFU.RSQ + one Newton step), sin(kmax * ds) by its series, the comparison
//            |cross| > sin(phi) * |p| * |d|  (<=> |angle| > phi for |angle| < pi/2), the unit heading for the collision rows;
//   float64 re-decision (the oracle's own expression, atan2 of both differences): inside a 4e-6 relative band around the
//            threshold, for segments shorter than 1e-8 m, for phi >= 0.49 and when the un-unwrapped heading difference jumps
//            by 2*pi (exact signs).  Masks therefore equal the float64 evaluation except on measure-zero inputs.
// ---------------------------------------------------------------------------------------------------------------------
struct TileSink {
  const EvalParams& p;
  const BlockTables& tb;
  double sum_a = 0, sum_j = 0, sum_v = 0, s_first = 0, s_last = 0, risk = 0;
  double pdx = 0, pdy = 0;                              // previous segment: float64 difference vector
  float pqf = 0.f, pdsf = 0.f;                          // its squared length and length (float32)
  int next_check = 0, row = 0;                          // next sample index to collision-check and its window row
  bool curv_fail = false, collided = false;
  __device__ __forceinline__ TileSink(const EvalParams& p_, const BlockTables& tb_) : p(p_), tb(tb_) {}
  __device__ __forceinline__ bool vetoed() const { return p.select_only && (curv_fail || collided); }

  // the float64 decision, exactly as the other kernels' guard-band path: c = (yaw_i - yaw_{i-1}) / ds_{i-1} without unwrapping
  __device__ __noinline__ static bool curvature_exact(double kmax, double pdx, double pdy, double dx, double dy) {
    const double q = pdx * pdx + pdy * pdy;
    const double pds = (q > 1e-200 && q < 1e200) ? q * rsqrt_fast(q) : hypot(pdx, pdy);
    const double dyaw = heading_of(dx, dy) - heading_of(pdx, pdy);
    return fabs(dyaw) > kmax * pds;                    // pds == 0: inf fails (dyaw != 0), nan passes (dyaw == 0)
  }

  // segment (dx, dy) [float64] with squared length qf / length dsf [float32] follows the stored previous segment
  __device__ __forceinline__ bool curvature_fails_fast(double dx, double dy, float qf, float dsf) const {
    const float phif = (float)p.kmax * pdsf;
    bool exact = !(pqf > 1e-16f && pqf < 1e30f) || !(qf > 1e-16f && qf < 1e30f) || !(phif < 0.49f);
    if (!exact) {
      const double cross = fma(pdx, dy, -(pdy * dx)), dot = fma(pdx, dx, pdy * dy);
      const bool nb = signbit(dy);
      if ((signbit(pdy) != nb) && (nb ? (cross > 0.0) : (cross < 0.0))) exact = true;      // the difference jumps by 2*pi: rare
      else if (!(dot > 0.0)) return true;                                                     // |angle| >= pi/2 > phi
      else {
        const float p2 = phif * phif;
        const float spf = phif * fmaf(p2, fmaf(p2, fmaf(p2, -1.0f / 5040.0f, 1.0f / 120.0f), -1.0f / 6.0f), 1.0f);
        const float lhs = (float)fabs(cross), rhs = spf * pdsf * dsf;
        if (fabsf(lhs - rhs) > 4e-6f * fmaxf(lhs, rhs)) return lhs > rhs;
        exact = true;
      }
    } else if (dx == 0.0 && dy == 0.0 && pdx == 0.0 && pdy == 0.0) {
      return false;                                     // standstill: both headings are atan2(0, 0) = 0, 0 / 0 = nan passes
    }
    return curvature_exact(p.kmax, pdx, pdy, dx, dy);
  }

  // separating-axis test with the float32 heading (cf, snf); inside the guard band the float64 heading is derived from the
  // float64 difference vector (hdx, hdy) of the heading segment and the pair is re-decided in float64
  __device__ __forceinline__ bool narrow_hit_lazy(int abs_step, int om, float dxo, float dyo, double x, double y, float cf, float snf,
                                                  double hdx, double hdy) const {
    const size_t gi = (size_t)abs_step * p.OM + om;
    const float4 nb = __ldg(p.narrow + gi);
    const int cls = sat_classify_f32(dxo, dyo, cf, snf, (float)p.ea, (float)p.eb, nb.x, nb.y, nb.z, nb.w, SAT_EPS);
    if (cls != 0) return cls < 0;
    const double q = hdx * hdx + hdy * hdy;
    double c = 1.0, sn = 0.0;
    if (q > 1e-200 && q < 1e200) { const double inv = rsqrt_fast(q); c = hdx * inv; sn = hdy * inv; }
    else { const double ds = hypot(hdx, hdy); if (ds > 0.0) { c = hdx / ds; sn = hdy / ds; } }
    const double* o = p.obs64 + gi * 4;
    return sat_intersects_f64(o[0] - x, o[1] - y, c, sn, p.ea, p.eb, o[2], o[3], p.ext64[2 * om], p.ext64[2 * om + 1]);
  }

  // collision row of sample i (pose (x, y), heading of the segment (hdx, hdy)) - same semantics as collide_row
  __device__ __forceinline__ void collide(int i, double x, double y, double hdx, double hdy, float hqf) {
#ifdef JSSP_KO_COLL
    return;
#endif
    if (i != next_check) return;                        // i % check_res == 0 without the division
    next_check += p.check_res;
    const int r = row++;
    if (collided || p.n_dict <= 0 || i >= p.tcap) return;
    const int abs_step = p.now + i;
    if (abs_step >= p.Tt) return;
    float cf = 1.0f, snf = 0.0f;
    if (hqf > 1e-16f && hqf < 1e30f) { const float inv = rsqrtf(hqf); cf = (float)hdx * inv; snf = (float)hdy * inv; }
    else { const double ds = hypot(hdx, hdy); if (ds > 0.0) { cf = (float)(hdx / ds); snf = (float)(hdy / ds); } }
    const float fx = (float)(x - p.ox), fy = (float)(y - p.oy);
    const float m2x = -2.0f * fx, m2y = -2.0f * fy, ne2 = -fmaf(fx, fx, fy * fy);
    if (tb.use_lists) {      // block-level culled rows: a handful of entries instead of OMpad
      const int cnt = tb.row_cnt[r];
      const float4* e = tb.list + tb.row_off[r];
      for (int k = 0; k < cnt; ++k) {
        const float4 b = e[k];
        if (fmaf(b.x, m2x, fmaf(b.y, m2y, b.z)) <= ne2) {
          const int om = __float_as_int(b.w);
          if (narrow_hit_lazy(abs_step, om, b.x - fx, b.y - fy, x, y, cf, snf, hdx, hdy)) {
            const float pr = __ldg(p.probf + om);
            if ((double)pr >= p.p_min) { collided = true; return; }
            risk = risk + (double)pr;
          }
        }
      }
      return;
    }
    const float4* rowp = tb.broad + (size_t)(tb.broad_row0 + r * tb.broad_stride) * p.OMpad;
    for (int base = 0; base < p.OM; base += 32) {
      unsigned int mask = 0u;
      const int nj = min(32, p.OM - base);
#pragma unroll 4
      for (int q = 0; q < nj; ++q) {
        const float4 b = rowp[base + q];
        mask |= (fmaf(b.x, m2x, fmaf(b.y, m2y, b.z)) <= ne2) ? (1u << q) : 0u;
      }
      while (mask) {      // rare: bounding circles overlap -> separating-axis test
        const int om = base + __ffs(mask) - 1;
        mask &= mask - 1u;
        const float4 b = rowp[om];
        if (narrow_hit_lazy(abs_step, om, b.x - fx, b.y - fy, x, y, cf, snf, hdx, hdy)) {
          if ((double)b.w >= p.p_min) { collided = true; return; }
          risk = risk + (double)b.w;
        }
      }
    }
  }

  // The same decision as curvature_fails_fast in straight-line form (no branch on the common path, so that the tests of two
  // consecutive segments overlap in the pipeline): 0 passes, 1 fails, 2 = re-decide with curvature_redecide.
  __device__ __forceinline__ int curvature_classify(double adx, double ady, float aqf, float adsf, double dx, double dy, float qf, float dsf) const {
#ifdef JSSP_KO_CURV
    return 0;
#endif
    const float phif = (float)p.kmax * adsf;
    const bool rng = (aqf > 1e-16f && aqf < 1e30f) && (qf > 1e-16f && qf < 1e30f) && (phif < 0.49f);
    const double cross = fma(adx, dy, -(ady * dx)), dot = fma(adx, dx, ady * dy);
    const bool nb = signbit(dy);
    const bool jump = (signbit(ady) != nb) && (nb ? (cross > 0.0) : (cross < 0.0));     // the difference jumps by 2*pi: rare
    const float p2 = phif * phif;
    const float spf = phif * fmaf(p2, fmaf(p2, fmaf(p2, -1.0f / 5040.0f, 1.0f / 120.0f), -1.0f / 6.0f), 1.0f);
    const float lhs = (float)fabs(cross), rhs = spf * adsf * dsf;
    const bool clear = fabsf(lhs - rhs) > 4e-6f * fmaxf(lhs, rhs);
    const bool obtuse = !(dot > 0.0);                                                   // |angle| >= pi/2 > phi
    return (!rng || jump) ? 2 : (obtuse ? 1 : (clear ? (lhs > rhs ? 1 : 0) : 2));
  }
  __device__ __noinline__ static bool curvature_redecide(double kmax, double adx, double ady, double dx, double dy) {
    if (dx == 0.0 && dy == 0.0 && adx == 0.0 && ady == 0.0) return false;   // standstill: both headings are atan2(0, 0) = 0, 0 / 0 = nan passes
    return curvature_exact(kmax, adx, ady, dx, dy);
  }
  static __device__ __forceinline__ void seg_lengths(double dx, double dy, float& qf, float& dsf) {
    const float dxf = (float)dx, dyf = (float)dy;
    qf = fmaf(dxf, dxf, dyf * dyf);
    float inv = rsqrtf(qf);
    inv = inv * fmaf(-0.5f * qf, inv * inv, 1.5f);
    dsf = (qf > 1e-16f && qf < 1e30f) ? qf * inv : 0.f;
  }

  // two consecutive segments at once: i (pose (x0, y0), difference A) and i + 1 (pose (x1, y1), difference B) - the same
  // decisions, flags and risk terms, in the same order, as segment(i, ...) followed by segment(i + 1, ...)
  __device__ __forceinline__ void segment2(int i, double x0, double y0, double ax, double ay, double x1, double y1, double bx, double by) {
    float aq, ads, bq, bds;
    seg_lengths(ax, ay, aq, ads);
    seg_lengths(bx, by, bq, bds);
    int ra = curvature_classify(pdx, pdy, pqf, pdsf, ax, ay, aq, ads);
    int rb = curvature_classify(ax, ay, aq, ads, bx, by, bq, bds);
    if (i < 1) ra = 0;
    if ((ra | rb) & 2) {                                                    // rare: guard band, tiny segments, heading wrap
      if (ra == 2) ra = curvature_redecide(p.kmax, pdx, pdy, ax, ay) ? 1 : 0;
      if (rb == 2) rb = curvature_redecide(p.kmax, ax, ay, bx, by) ? 1 : 0;
    }
    if (ra | rb) curv_fail = true;
    collide(i, x0, y0, ax, ay, aq);
    collide(i + 1, x1, y1, bx, by, bq);
    pdx = bx; pdy = by; pqf = bq; pdsf = bds;
  }

  // segment i = p_{i+1} - p_i arrives: curvature against segment i - 1, collision row of sample i, shift
  __device__ __forceinline__ void segment(int i, double x, double y, double dx, double dy) {
    const float dxf = (float)dx, dyf = (float)dy;
    const float qf = fmaf(dxf, dxf, dyf * dyf);
    float dsf = 0.f;
    if (qf > 1e-16f && qf < 1e30f) { float inv = rsqrtf(qf); inv = inv * fmaf(-0.5f * qf, inv * inv, 1.5f); dsf = qf * inv; }
    if (i >= 1 && curvature_fails_fast(dx, dy, qf, dsf)) curv_fail = true;
    collide(i, x, y, dx, dy, qf);
    pdx = dx; pdy = dy; pqf = qf; pdsf = dsf;
  }
  // chunked walk: the chunk's first tested segment is `a`; (dx, dy) = segment a - 1, tested by the chunk before
  __device__ __forceinline__ void start_at(int a) { next_check = (a + p.check_res - 1) / p.check_res * p.check_res; row = next_check / p.check_res; }
  __device__ __forceinline__ void prime(double dx, double dy) {
    const float dxf = (float)dx, dyf = (float)dy;
    const float qf = fmaf(dxf, dxf, dyf * dyf);
    float dsf = 0.f;
    if (qf > 1e-16f && qf < 1e30f) { float inv = rsqrtf(qf); inv = inv * fmaf(-0.5f * qf, inv * inv, 1.5f); dsf = qf * inv; }
    pdx = dx; pdy = dy; pqf = qf; pdsf = dsf;
  }
  __device__ __forceinline__ void finish(int n_pts, double x, double y) {
    if (n_pts >= 2) collide(n_pts - 1, x, y, pdx, pdy, pqf);                // last point reuses the previous heading (:158-160)
    else if (n_pts == 1 && p.n_dict > 0 && p.tcap >= 1) collided = true;    // bare except -> collision (:250-254)
  }
};
__device__ __forceinline__ double candidate_cost(const EvalParams& p, const TileSink& k, const double* latcost) {
  return candidate_cost(p, k.sum_a, k.sum_j, k.sum_v, k.s_first, k.s_last, k.risk, latcost);
}

// The tiled walk of the candidate in lane (g, j): g = group (seed) within the warp, j = lateral target.  `gbase` = this group's
// tile, `prof` = this group's segment table (already stored), both in the warp's scratch.  Inactive lanes (no seed) only keep
// the warp-level barriers company.  On return `sink` holds the masks, the risk term and the longitudinal cost sums.
// kChunked: the group walks the chunk `ck` of the horizon only and the chunks of a candidate are combined at the end (see
// TileChunk); the combined result is valid in the lanes of the candidate's FIRST chunk.
template <bool kChunked>
__device__ __forceinline__ void walk_tiles(const EvalParams& p, const BlockTables& tb, const TileGeom& tg, double* gbase, double* gterms,
                                           double* prof, bool active, int j, int W, TileSink& sink, const TileChunk& ck) {
  const int P = p.P, TL = tg.TL;
  const double* lat = tb.lat + (size_t)j * P * 4;
  const int lo = kChunked ? ck.lo : 0, hi = kChunked ? ck.hi : P - 1;
  // every lane of the warp runs the same number of tiles (the chunks differ by a point or two): the barriers stay warp-wide
  const int n_tiles = ((kChunked ? tg.own_max + 1 : P) + TL - 1) / TL;
  int k = 0, n_pts = P;
  bool ended = false;                            // the path left the reference line at point n_pts
  double sa = 0, sj = 0, sv = 0;                 // lane 0 of the group: the longitudinal cost sums, added strictly in time order
  double xp = 0, yp = 0;
  bool walking = active;
  for (int t = 0; t < n_tiles; ++t) {
    const int base = lo + t * TL;
    const int len = min(TL, hi + 1 - base);      // <= 0: the chunk is done
    // phase 1: seed-only quantities of the samples this lane owns in the tile
    if (active) {
      for (int il = j; il < len; il += W) {
        const int i = base + il;
        double s, sd, sdd, sddd;
        tile_lon_eval(prof, k, tb.time[i], s, sd, sdd, sddd);
        const double ea = fmax(fabs(sdd) - p.thr_lon_a, 0.0);
        const double ej = fmax(fabs(sddd) - p.thr_lon_j, 0.0);
        const double ev = (p.vmax - sd) * p.inv_vmax;
        if (kChunked && !ck.first) {               // later chunks keep their terms until the earlier sums are final
          if (i >= ck.own_lo && i < ck.own_hi) { double* tm = ck.lterms + (i - ck.own_lo) * 3; tm[0] = ea * ea; tm[1] = ej * ej; tm[2] = ev * ev; }
        } else {
          double* tm = gterms + il * 3;
          tm[0] = ea * ea; tm[1] = ej * ej; tm[2] = ev * ev;
        }
        if (i == 0) prof[28] = s;
        if (i == P - 1) prof[29] = s;
        double2* e = reinterpret_cast<double2*>(gbase + il * 4);
        if (spline_in_range(tb.sp, s)) {
          double ix, iy, u, v;
          frenet_base_fma(tb.sp, spline_walk(tb.sp, s, 0), s, ix, iy, u, v);
          e[0] = make_double2(ix, iy); e[1] = make_double2(u, v);
        } else {
          e[0] = make_double2(nan(""), 0.0);      // off the reference line: the Cartesian path ends here (jssp.py:138-139)
        }
      }
    }
    __syncwarp();
    if (t == 1) PHASE(p, 14);
    // the cost sums are sequential float64 sums over the samples (the oracle's order: costs bit-identical to the other kernels):
    // one lane per seed adds the tile's terms - three short independent chains
    if (active && j == 0 && (!kChunked || ck.first)) {
      const int own = kChunked ? min(len, ck.own_hi - base) : len;
      for (int il = 0; il < own; ++il) { sa = sa + gterms[il * 3]; sj = sj + gterms[il * 3 + 1]; sv = sv + gterms[il * 3 + 2]; }
    }
    // phase 2: this lane's candidate - two samples per step where the tile has them (two independent dependency chains)
    if (walking) {
      int il = 0;
      while (il < len) {
        const int i = base + il;
        const double2* e = reinterpret_cast<const double2*>(gbase + il * 4);
        const double2 pxy = e[0];
        if (!(pxy.x == pxy.x)) { n_pts = i; ended = true; walking = false; break; }
        const double2 uv = e[1];
        const double d = lat[i * 4];
        const double x = fma(-d, uv.x, pxy.x), y = fma(d, uv.y, pxy.y);
        if (i == lo) { xp = x; yp = y; ++il; continue; }
        if (kChunked && !ck.first && i == lo + 1) {
          sink.prime(x - xp, y - yp);            // the segment before the chunk: tested by the chunk before
          xp = x; yp = y; ++il; continue;
        }
        bool two = il + 1 < len;
        double x1 = 0, y1 = 0;
        if (two) {
          const double2 qxy = e[2];
          const double2 uv1 = e[3];
          const double d1 = lat[(i + 1) * 4];
          x1 = fma(-d1, uv1.x, qxy.x); y1 = fma(d1, uv1.y, qxy.y);
          two = qxy.x == qxy.x;                  // the path ends at the next point: it is met again, alone, in the next step
        }
        if (two) {
          sink.segment2(i - 1, xp, yp, x - xp, y - yp, x, y, x1 - x, y1 - y);
          xp = x1; yp = y1; il += 2;
        } else {
          sink.segment(i - 1, xp, yp, x - xp, y - yp);
          xp = x; yp = y; ++il;
        }
        if (sink.vetoed()) { walking = false; break; }     // select-only evaluation: this candidate can no longer be selected
      }
    }
    __syncwarp();                                  // the tile is rewritten by the next phase 1
    if (t == 0) PHASE(p, 8);                       // profiling builds: progress of the block's first warp through the tiles
    if (t == 1) PHASE(p, 9);
    if (t == n_tiles / 2) PHASE(p, 10);
    if (t == n_tiles - 1) PHASE(p, 11);
  }
  if (!kChunked) {
    if (active && !sink.vetoed()) sink.finish(n_pts, xp, yp);
    if (active && j == 0) { gbase[0] = sa; gbase[1] = sj; gbase[2] = sv; }
    __syncwarp();
    if (active) {
      sink.sum_a = gbase[0]; sink.sum_j = gbase[1]; sink.sum_v = gbase[2];
      sink.s_first = prof[28]; sink.s_last = prof[29];
    }
    __syncwarp();
    return;
  }
  // ---- chunked: the last point's collision row belongs to the chunk that saw the path end (an end within the chunk's first two
  // points was also seen - and is owned - by the chunk before) or, on a complete path, to the last chunk
  const bool owns_end = ended ? (ck.first || n_pts >= lo + 2) : ck.last;
  if (active && owns_end && !sink.vetoed()) sink.finish(n_pts, xp, yp);
  const unsigned int full = 0xffffffffu;
  const int H = tg.H, lane = threadIdx.x & 31;
  const int q = lane / W, g = q / H, h = q - g * H;
  const int fl = (sink.curv_fail ? 1 : 0) | (sink.collided ? 2 : 0) | (ended ? 4 : 0);
  const double rk = sink.risk;
  bool curv = false, coll = false, over = false;
  double risk = 0.0;
  for (int h2 = 0; h2 < H; ++h2) {                 // in time order, with the serial walk's semantics
    const int src = ((g * H + h2) * W + j) & 31;
    const int f2 = __shfl_sync(full, fl, src);
    const double r2 = __shfl_sync(full, rk, src);
    if (!over) {
      curv = curv || (f2 & 1);
      if (!coll) { risk = risk + r2; coll = (f2 & 2) != 0; }   // no risk is collected after a hard collision
      over = (f2 & 4) != 0;                         // the path ended in this chunk: the later chunks walked points that do not exist
    }
  }
  for (int h2 = 1; h2 < H; ++h2) {                 // cost sums: chunk h2 continues the sums of chunk h2 - 1 term by term
    const int src = ((g * H + h2 - 1) * W) & 31;
    const double a2 = __shfl_sync(full, sa, src), j2 = __shfl_sync(full, sj, src), v2 = __shfl_sync(full, sv, src);
    if (active && j == 0 && h == h2) {
      sa = a2; sj = j2; sv = v2;
      const int own = ck.own_hi - ck.own_lo;
      for (int m = 0; m < own; ++m) { sa = sa + ck.lterms[m * 3]; sj = sj + ck.lterms[m * 3 + 1]; sv = sv + ck.lterms[m * 3 + 2]; }
    }
  }
  const int srcl = ((g * H + H - 1) * W) & 31;
  const double ta = __shfl_sync(full, sa, srcl), tj = __shfl_sync(full, sj, srcl), tv = __shfl_sync(full, sv, srcl);
  if (active) {
    sink.curv_fail = curv; sink.collided = coll; sink.risk = risk;
    sink.sum_a = ta; sink.sum_j = tj; sink.sum_v = tv;
    sink.s_first = (prof - h * TILE_PROF)[28]; sink.s_last = (prof + (H - 1 - h) * TILE_PROF)[29];
  }
  __syncwarp();
}

template <int kMode, bool kChunked = false>
__device__ __forceinline__ void evaluate_tiled_body(const EvalParams& p, unsigned char* smem, double* red_cost, int* red_idx,
                                                    bool* is_last, int* sm_npts) {
  static_assert(!(kChunked && kMode == 2), "the chunked walk has no block-level lists");
  const int Ns = resolve_n_seeds(p);
  const int W = p.W;
  const int N = Ns * W;
  const TileGeom tg = tile_geom(W, p.tile_rounds, kChunked ? p.tile_chunks : 1, p.P);
  const int gpw = tg.gpw, H = tg.H, spw = tg.spw;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int q = lane / W, j = lane - q * W;    // lane group and lateral target
  const int g = q / H, h = q - g * H;          // seed within the warp and chunk of the horizon (H = 1: g = q, h = 0)
  const int spb = (blockDim.x >> 5) * spw;     // seeds per block
  const int k = blk_index(p) * spb + wid * spw + g;
  const bool active = q < gpw && g < spw && k < Ns;
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
    const int gq = active ? q : 0;
    double* gbase = wscr + (size_t)gq * tg.grp_stride;
    double* gterms = wscr + tg.off_terms + (size_t)gq * tg.TL * 3;
    double* prof = wscr + tg.off_prof + (size_t)gq * TILE_PROF;
    if (active && j == 0) tile_profile_store(prof, p.seeds + (size_t)k * 10, p.s0, p.v0, p.a0);
    __syncwarp();
    if (kMode == 2) { TileSAt s_at{prof}; build_lists(p, smem, tb, s_at, active, j, W); }
    PHASE(p, 2);
    TileSink sink(p, tb);
    TileChunk ck{};
    if (kChunked) {
      ck = tile_chunk(active ? h : 0, H, p.P, (active && h > 0) ? wscr + tg.off_lterms + (size_t)(g * (H - 1) + h - 1) * tg.own_max * 3 : nullptr);
      sink.start_at(ck.own_lo);
    }
    walk_tiles<kChunked>(p, tb, tg, gbase, gterms, prof, active, j, W, sink, ck);
    PHASE(p, 12);
    if (active && h == 0) {
      const int n = j * Ns + k;                // candidate index: width-major (jerk_space_sampling_planner.py:106,124-125)
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

// The chunked form for few scenarios per GPU (latency of the plan cycle): `chunks` lane groups per seed split the horizon.
template <int kMode>
__global__ void __launch_bounds__(BATCH_EVAL_THREADS, CHUNK_MIN_BLOCKS) evaluate_tiled_chunked_batch_kernel(const EvalParams* __restrict__ slots, int chunks) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(16) EvalParams sp;
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  pdl_launch_dependents(); pdl_wait();
  if (!slot_live(slots, blockIdx.x)) return;
  load_slot(&sp, slots + blockIdx.x);
  if (threadIdx.x == 0) sp.tile_chunks = chunks;
  __syncthreads();
  evaluate_tiled_body<kMode, true>(sp, smem, red_cost, red_idx, &is_last, &sm_npts);
}

}  // namespace jssp
