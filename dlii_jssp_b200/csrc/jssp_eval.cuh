// jssp_eval.cuh - K2: candidate evaluation - parameters, walks, sinks, staging, culled lists, export, batched-replay hand-off, selection, the thread / group kernels and their batch forms
#pragma once
#include "jssp_common.cuh"
#include "jssp_seed.cuh"
#include "jssp_split.cuh"

namespace jssp {

// ------------------------------------------------------------------------------------------------------------------
// K2: candidate evaluation
// ------------------------------------------------------------------------------------------------------------------
struct EvalParams {
  // cycle
  double s0, v0, a0;
  double quintic[JSSP_MAX_WIDTHS][6];
  int W, P, now, tcap;            // tcap = final_time_step - time_step_now
  int n_dict, check_res;
  // constants
  double kmax, ea, eb, p_min, w_risk, vmax, inv_vmax, max_distance, half_w, thr_lon_a, thr_lon_j, thr_lat_a, thr_lat_j;
  double cw[8];
  // tables
  const double* time;
  const double* knot;
  const double* coef;
  int K, kpad;
  double ox, oy;
  const double* seeds;
  const int* n_seeds_ptr;         // device count (enumerated seeds) or NULL
  int n_seeds_fixed, seed_cap;
  // obstacles
  const float4* broad;
  const float4* narrow;
  const double* obs64;
  const double* ext64;
  int OM, OMpad, Tt, win_rows, smem_obs;
  const float2* sint;             // [Tt][OMpad] reachable arc-length interval of every obstacle state (block-level culling)
  const float* probf;             // [OM]
  int list_cap;                   // entries of the per-block compacted obstacle lists (mode 2)
  int tile_rounds;                // tiled kernel: phase-1 rounds per tile (0 = default), chosen by the host for its shared-memory budget
  int tile_chunks;                // chunked tiled kernel: lane groups per seed that split the horizon (0 / 1 = none)
  int select_only;                // no per-candidate outputs are wanted: a candidate's walk may stop at its first veto (the
                                  // reference filters the same way: check_constraints_curvature -> check_collisions -> cost on
                                  // the survivors only, has_collision returns at the first hit; jerk_space_sampling_planner.py:311-319)
  // outputs
  uint8_t* feasible;
  uint8_t* collision;
  float* cost;
  double* blk_cost;
  int* blk_idx;
  unsigned int* ticket;
  BestRecord* best;
  double* best_traj;              // [P][16] export of the selected trajectory (NULL = skip)
  // candidate split over ranks (jssp_set_split; 0 / NULL = off): this launch holds seeds [split_seed_lo, +n_seeds) of
  // split_ns_global; the record of the selection carries the GLOBAL candidate index w * split_ns_global + split_seed_lo + k
  long long split_seed_lo, split_ns_global;
  SplitRecord* split_rec;         // NCCL transport: the rank's record, all-gathered by jssp_allreduce_argmin
  const PeerView* peers;          // peer-mailbox transport: exchange inside this launch (epoch = split_epoch)
  unsigned long long split_epoch;
  // batched replay (NULL / unused for a single plan cycle)
#ifdef JSSP_PROFILE_PHASES
  unsigned long long* phase_ts;   // [blocks][8] globaltimer stamps (tools/phase_probe.py; profiling builds only)
#endif
  struct BatchSlot* batch;        // per-scenario replay state, updated by the hand-off at the end of every cycle
  EvalParams* self;               // this scenario's parameter record in device memory (the hand-off writes the next cycle's inputs)
};

// Replay state of one scenario of a batch (device memory).  The hand-off at the end of a plan cycle - the last block of
// the evaluation kernel - moves the ego to sample 1 of the selected trajectory (code/onsite_trans/env.py:24-30), makes
// that sample the next cycle's Frenet start (intersection_planner.py:312), advances the clock (controller.py:303) and
// evaluates the end-of-scenario status (controller.py:338-380), so that a whole replay runs without host round trips.
struct BatchSlot {
  int live;                       // 1 while the scenario is running
  int cycle;                      // plan cycles done
  int final_step;                 // int(max_t / dt)
  int gt_O, gt_Tt;                // ground-truth background vehicles and steps of their table
  int n_cycles_cap;               // the per-cycle tables hold n_cycles_cap + 1 entries: the replay pauses after that many cycles
  int seed_overflow;              // a cycle enumerated more seeds than max_seeds (the surplus was dropped): reported, never silent
  int last_n_pts;                 // Cartesian points of the trajectory in best_traj (0 = no trajectory selected yet)
  int is_fop;                     // the scenario's parameter record is a FopParams (FOP baseline in the batch)
  int kin;                        // ego_update_mode "kinematics": kinematic ego + pure pursuit + projected start state
  int n_poly, pad3;
  const double* poly;             // [n_poly][6] re-discretised reference line: x, y, yaw, arc, cos yaw, sin yaw
  double pp_kv, pp_ld0, pp_wb;    // pure pursuit: ld = kv * v + ld0, wheel base
  double end;                     // end status: -1 running, 1 max_t, 1.5 goal, 2 collision, -2 no solution, -3 short trajectory
  double max_t, horizon;
  double goal_x, goal_y, goal_cos, goal_sin;
  double ego_l, ego_w;
  double ego_x, ego_y, ego_yaw, ego_v;   // the ego as the last Env.step left it (initially the scenario's start pose)
  double dt;
  double bx_min, bx_max, by_min, by_max; // map bounds (+-inf without a map)
  double targets[JSSP_MAX_WIDTHS];
  const double* cycle_time;       // [C+1] clock at the start of cycle c (host-accumulated float additions)
  const int* cycle_now;           // [C+1] int(t / dt)
  const int* cycle_end_step;      // [C+1] table step whose key equals str(around(t, 3)), -1 = none
  const double* gt_state;         // [gt_O][gt_Tt][3] x, y, yaw of the background vehicles
  const uint8_t* gt_valid;        // [gt_O][gt_Tt]
  const double* gt_size;          // [gt_O][2] length, width
  jssp_cycle_record* records;     // [C]
};

__device__ __forceinline__ SeedState seed_state_of(const EvalParams* slots, int sc) { return SeedState{slots[sc].s0, slots[sc].v0, slots[sc].a0}; }
__device__ __forceinline__ bool slot_live(const EvalParams* slots, int sc) { return slots[sc].batch->live != 0; }

struct BlockTables {
  const float4* broad;   // window rows [win_rows][OMpad] (shared) or the global table [Tt][OMpad]
  int broad_row0;        // 0 for the shared window, time_step_now for the global table
  int broad_stride;      // rows advance by check_res in the global table, 1 in the window
  SplineView sp;
  const double* time;
  const double* lat;     // [W][P][4]
  // mode 2: per-row lists of the obstacle states this block's candidates can reach: (x', y', |o'|^2 - R^2, om bits)
  const float4* list;
  const int* row_off;
  const int* row_cnt;
  bool use_lists;
};

// Walks one candidate through the horizon and hands every derived quantity to `sink`, in the order the reference
// derives them (get_one_frenet_trajectory -> calc_global_paths).  Used by the evaluation, export and dump kernels so
// that all three see the same positions.
//   sink.sample(i, t, s, sd, sdd, sddd)                      every i in [0, P)
//   sink.point(i, x, y)                                      i < n_pts
//   sink.segment(i, x_i, y_i, dx, dy, ds_i, cos, sin)        i < n_pts - 1: forward difference p_{i+1} - p_i, its length
//                                                            and unit vector ((1, 0) when the length is 0: atan2(0,0) = 0)
//   sink.finish(n_pts, x_L, y_L)                             once; L = n_pts - 1
// `lon.eval(t, s, sd, sdd, sddd)` is the longitudinal model (the 5-segment jerk profile of the JSSP seeds, the quartic of
// the FOP baseline), `lat` the candidate's lateral samples [P][4], P its number of samples.
template <class Sink, class Lon>
__device__ __forceinline__ void walk_core(const BlockTables& tb, Lon& lon, const double* lat, int P, int seg, Sink& sink) {
  int n_pts = P;
  double xp = 0, yp = 0;
  int i = 0;
  // live part: ends at the first sample whose s is off the reference line (calc_position returned None -> break,
  // jerk_space_sampling_planner.py:138-139)
  for (; i < P; ++i) {
    double s, sd, sdd, sddd;
    lon.eval(tb.time[i], s, sd, sdd, sddd);
    sink.sample(i, tb.time[i], s, sd, sdd, sddd);
    if (!spline_in_range(tb.sp, s)) { n_pts = i; ++i; break; }
    seg = spline_walk(tb.sp, s, seg);
    double x, y;
    frenet_point(tb.sp, seg, s, lat[i * 4], x, y);
    sink.point(i, x, y);
    if (i >= 1) {
      const double dx = x - xp, dy = y - yp;
      const double q = dx * dx + dy * dy;
      double ds, c = 1.0, sn = 0.0;
      if (q > 1e-200 && q < 1e200) { const double inv = rsqrt_fast(q); ds = q * inv; c = dx * inv; sn = dy * inv; }
      else { ds = hypot(dx, dy); if (ds > 0.0) { c = dx / ds; sn = dy / ds; } }
      sink.segment(i - 1, xp, yp, dx, dy, ds, c, sn);
      if (sink.vetoed()) return;          // select-only evaluation: this candidate can no longer be selected
    }
    xp = x; yp = y;
  }
  // Frenet-only tail (the cost still sums over every sample)
  for (; i < P; ++i) {
    double s, sd, sdd, sddd;
    lon.eval(tb.time[i], s, sd, sdd, sddd);
    sink.sample(i, tb.time[i], s, sd, sdd, sddd);
  }
  sink.finish(n_pts, xp, yp);
}

struct JerkLon {
  LonWalk& w;
  __device__ __forceinline__ void eval(double t, double& s, double& sd, double& sdd, double& sddd) { lon_walk_eval(w, t, s, sd, sdd, sddd); }
};

template <class Sink>
__device__ __forceinline__ void walk_candidate(const EvalParams& p, const BlockTables& tb, LonWalk& lp, int w,
                                               int seg, Sink& sink) {
  JerkLon lon{lp};
  walk_core(tb, lon, tb.lat + (size_t)w * p.P * 4, p.P, seg, sink);
}

__device__ __forceinline__ double heading_of(double dx, double dy) { return atan2(dy, dx); }   // np.arctan2(y_d, x_d)

// c_{i-1} = (yaw_i - yaw_{i-1}) / ds_{i-1} with yaw = arctan2 of the forward differences and NO unwrapping
// (jerk_space_sampling_planner.py:155-163); the candidate fails when |c| > kmax (:175).  (p*) = segment i-1, the others
// segment i, each as difference vector, length and unit vector.  Fast path: sin / cos of the wrapped angle from the unit
// vectors, compared against the series of sin(kmax * ds); the +-2*pi jump of the un-unwrapped difference is detected from
// the exact signs; inside a 1e-6 relative guard band, for zero-length segments and for jumps the float64 atan2 decides.
__device__ __forceinline__ bool curvature_fails(double kmax, double pdx, double pdy, double pds, double pc, double psn, double dx,
                                                double dy, double ds, double c, double sn) {
  const double phi = kmax * pds;
  bool exact = !(pds > 0.0) || !(ds > 0.0) || !(phi < 0.5);
  if (!exact) {
    const double cr = pc * sn - psn * c, dt = pc * c + psn * sn;      // sin / cos of the wrapped angle (unit vectors)
    const bool nb = signbit(dy);
    if ((signbit(pdy) != nb) && (nb ? (cr > 0.0) : (cr < 0.0))) exact = true;   // the difference jumps by 2*pi: rare
    else if (!(dt > 0.0)) return true;                                           // |angle| >= pi/2 > phi
    else {
      // |angle| < pi/2:  |angle| > phi  <=>  |sin angle| > sin phi;  sin phi by its series (phi < 0.5: 1e-11 relative)
      const double p2 = phi * phi;
      const double sp = phi * fma(p2, fma(p2, fma(p2, fma(p2, 1.0 / 362880.0, -1.0 / 5040.0), 1.0 / 120.0), -1.0 / 6.0), 1.0);
      const double a = fabs(cr);
      if (fabs(a - sp) > 1e-6 * fmax(a, sp)) return a > sp;
      exact = true;
    }
  }
  const double dyaw = heading_of(dx, dy) - heading_of(pdx, pdy);
  return fabs(dyaw) > phi;      // pds == 0: inf fails (dyaw != 0), nan passes (dyaw == 0)
}

// separating-axis test of the ego rectangle at (x, y, heading (c, sn)) against obstacle mode om at abs_step
__device__ __forceinline__ bool narrow_hit(const EvalParams& p, int abs_step, int om, float dxf, float dyf, double x, double y, double c,
                                           double sn) {
  const size_t gi = (size_t)abs_step * p.OM + om;
  const float4 nb = __ldg(p.narrow + gi);
  const int cls = sat_classify_f32(dxf, dyf, (float)c, (float)sn, (float)p.ea, (float)p.eb, nb.x, nb.y, nb.z, nb.w, SAT_EPS);
  if (cls != 0) return cls < 0;
  const double* q = p.obs64 + gi * 4;
  return sat_intersects_f64(q[0] - x, q[1] - y, c, sn, p.ea, p.eb, q[2], q[3], p.ext64[2 * om], p.ext64[2 * om + 1]);
}

// Collision row r (sample i of the candidate, pose (x, y, heading (c, sn))) against the obstacle states of that step: a
// hit by a mode with prob >= p_min vetoes the candidate, a hit by a weaker mode adds its probability to the risk term.
__device__ __forceinline__ void collide_row(const EvalParams& p, const BlockTables& tb, int r, int i, double x, double y, double c, double sn,
                                            bool& collided, double& risk) {
  if (collided || p.n_dict <= 0 || i >= p.tcap) return;
  const int abs_step = p.now + i;
  if (abs_step >= p.Tt) return;
  const float4* row = tb.broad + (size_t)(tb.broad_row0 + r * tb.broad_stride) * p.OMpad;
  const float fx = (float)(x - p.ox), fy = (float)(y - p.oy);
  const float m2x = -2.0f * fx, m2y = -2.0f * fy, ne2 = -fmaf(fx, fx, fy * fy);
  if (tb.use_lists) {      // block-level culled rows: a handful of entries instead of OMpad
    const int cnt = tb.row_cnt[r];
    const float4* e = tb.list + tb.row_off[r];
    for (int k = 0; k < cnt; ++k) {
      const float4 b = e[k];
      if (fmaf(b.x, m2x, fmaf(b.y, m2y, b.z)) <= ne2) {
        const int om = __float_as_int(b.w);
        if (narrow_hit(p, abs_step, om, b.x - fx, b.y - fy, x, y, c, sn)) {
          const float pr = __ldg(p.probf + om);
          if ((double)pr >= p.p_min) { collided = true; return; }
          risk = risk + (double)pr;
        }
      }
    }
    return;
  }
  for (int base = 0; base < p.OM; base += 32) {
    unsigned int mask = 0u;
    if (p.OM - base >= 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float4 b = row[base + j];
        const float t = fmaf(b.x, m2x, fmaf(b.y, m2y, b.z));
        mask |= (t <= ne2) ? (1u << j) : 0u;
      }
    } else {                                  // last, partial chunk: the padding up to OMpad is never touched
      const int nj = p.OM - base;
#pragma unroll 4
      for (int j = 0; j < nj; ++j) {
        const float4 b = row[base + j];
        const float t = fmaf(b.x, m2x, fmaf(b.y, m2y, b.z));
        mask |= (t <= ne2) ? (1u << j) : 0u;
      }
    }
    while (mask) {      // rare: bounding circles overlap -> separating-axis test
      const int om = base + __ffs(mask) - 1;
      mask &= mask - 1u;
      const float4 b = row[om];
      if (narrow_hit(p, abs_step, om, b.x - fx, b.y - fy, x, y, c, sn)) {
        if ((double)b.w >= p.p_min) { collided = true; return; }
        risk = risk + (double)b.w;
      }
    }
  }
}

// -- the evaluation sink ---------------------------------------------------------------------------------------------
struct EvalSink {
  const EvalParams& p;
  const BlockTables& tb;
  double sum_a = 0, sum_j = 0, sum_v = 0, s_first = 0, s_last = 0, risk = 0;
  double pdx = 0, pdy = 0, pds = 0, pc = 1, psn = 0;   // previous segment (difference, length, unit vector)
  int next_check = 0, row = 0;                         // next sample index to collision-check and its window row
  bool curv_fail = false, collided = false;
  __device__ __forceinline__ EvalSink(const EvalParams& p_, const BlockTables& tb_) : p(p_), tb(tb_) {}

  __device__ __forceinline__ void sample(int i, double, double s, double sd, double sdd, double sddd) {
    if (i == 0) s_first = s;
    s_last = s;
    const double ea = fmax(fabs(sdd) - p.thr_lon_a, 0.0);
    const double ej = fmax(fabs(sddd) - p.thr_lon_j, 0.0);
    const double ev = (p.vmax - sd) * p.inv_vmax;
    sum_a = sum_a + ea * ea;
    sum_j = sum_j + ej * ej;
    sum_v = sum_v + ev * ev;
  }
  // the same accumulation from per-sample cost terms computed elsewhere (group kernel: by the lane that owns the sample)
  __device__ __forceinline__ void sample_terms(double ea2, double ej2, double ev2) {
    sum_a = sum_a + ea2;
    sum_j = sum_j + ej2;
    sum_v = sum_v + ev2;
  }
  __device__ __forceinline__ void point(int, double, double) {}
  __device__ __forceinline__ bool vetoed() const { return p.select_only && (curv_fail || collided); }

  __device__ __forceinline__ void curvature_step(double dx, double dy, double ds, double c, double sn) {
    if (curvature_fails(p.kmax, pdx, pdy, pds, pc, psn, dx, dy, ds, c, sn)) curv_fail = true;
  }

  __device__ __forceinline__ void collide(int i, double x, double y, double c, double sn) {
    if (i != next_check) return;               // i % check_res == 0 without the division
    next_check += p.check_res;
    collide_row(p, tb, row++, i, x, y, c, sn, collided, risk);
  }
  __device__ __forceinline__ void segment(int i, double x, double y, double dx, double dy, double ds, double c, double sn) {
    if (i >= 1) curvature_step(dx, dy, ds, c, sn);
    collide(i, x, y, c, sn);
    pdx = dx; pdy = dy; pds = ds; pc = c; psn = sn;
  }
  __device__ __forceinline__ void finish(int n_pts, double x, double y) {
    if (n_pts >= 2) collide(n_pts - 1, x, y, pc, psn);                      // last point reuses the previous heading (:158-160)
    else if (n_pts == 1 && p.n_dict > 0 && p.tcap >= 1) collided = true;    // bare except -> collision (:250-254)
  }
};

// ------------------------------------------------------------------------------------------------------------------
// Group form of the walk: the W candidates that share a longitudinal seed (one per lateral target) sit in W consecutive
// lanes.  Everything that depends on the seed only - the unrolled s, s_d, s_dd, s_ddd, the cost terms, the spline segment,
// the reference point and its unit normal - is computed ONCE per sample by the lane that owns the sample (lane j owns
// samples j, j + W, ...) and handed to the group by warp shuffles; every lane then adds its own lateral offset and runs the
// per-candidate part (difference, length, curvature test, collision rows).  Same expressions, same order of the cost
// sums: results are bit-identical to walk_candidate.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_f64(unsigned int mask, double v, int src) { return __shfl_sync(mask, v, src); }

__device__ __forceinline__ void walk_group(const EvalParams& p, const BlockTables& tb, LonWalk& lp, int w, int W, unsigned int gmask,
                                           int glane0, EvalSink& sink) {
  const int P = p.P;
  const double* lat = tb.lat + (size_t)w * P * 4;
  int n_pts = P;
  bool live = true;
  double xp = 0, yp = 0;
  for (int base = 0; base < P; base += W) {
    // the sample this lane owns in this round
    const int io = base + w;
    double o_s = 0, o_ea = 0, o_ej = 0, o_ev = 0, o_ix = 0, o_iy = 0, o_u = 0, o_v = 0;
    int o_in = 0;
    if (io < P) {
      double sd, sdd, sddd;
      lon_walk_eval(lp, tb.time[io], o_s, sd, sdd, sddd);
      const double ea = fmax(fabs(sdd) - p.thr_lon_a, 0.0);
      const double ej = fmax(fabs(sddd) - p.thr_lon_j, 0.0);
      const double ev = (p.vmax - sd) * p.inv_vmax;
      o_ea = ea * ea; o_ej = ej * ej; o_ev = ev * ev;
      if (live && spline_in_range(tb.sp, o_s)) {
        o_in = 1;
        frenet_base(tb.sp, spline_walk(tb.sp, o_s, 0), o_s, o_ix, o_iy, o_u, o_v);
      }
    }
    const int cnt = min(W, P - base);
    for (int jj = 0; jj < cnt; ++jj) {
      const int i = base + jj, src = glane0 + jj;
      sink.sample_terms(shfl_f64(gmask, o_ea, src), shfl_f64(gmask, o_ej, src), shfl_f64(gmask, o_ev, src));
      if (i == 0) sink.s_first = shfl_f64(gmask, o_s, src);
      if (i == P - 1) sink.s_last = shfl_f64(gmask, o_s, src);
      if (!live) continue;
      if (!__shfl_sync(gmask, o_in, src)) { n_pts = i; live = false; continue; }   // first sample off the reference line (:138-139)
      const double ix = shfl_f64(gmask, o_ix, src), iy = shfl_f64(gmask, o_iy, src);
      const double u = shfl_f64(gmask, o_u, src), v = shfl_f64(gmask, o_v, src);
      const double d = lat[i * 4];
      const double x = ix - d * u, y = iy + d * v;
      if (i >= 1) {
        const double dx = x - xp, dy = y - yp;
        const double q = dx * dx + dy * dy;
        double ds, c = 1.0, sn = 0.0;
        if (q > 1e-200 && q < 1e200) { const double inv = rsqrt_fast(q); ds = q * inv; c = dx * inv; sn = dy * inv; }
        else { ds = hypot(dx, dy); if (ds > 0.0) { c = dx / ds; sn = dy / ds; } }
        sink.segment(i - 1, xp, yp, dx, dy, ds, c, sn);
      }
      xp = x; yp = y;
    }
  }
  sink.finish(n_pts, xp, yp);
}

__device__ __forceinline__ double candidate_cost(const EvalParams& p, double sum_a, double sum_j, double sum_v, double s_first,
                                                 double s_last, double risk, const double* latcost) {
  const double P = (double)p.P;
  double cost = p.cw[1] * (1.0 - (s_last - s_first) / p.max_distance);
  cost = cost + p.cw[2] * (latcost[0] / P / (p.half_w * p.half_w));
  cost = cost + p.cw[3] * (latcost[1] / P);
  cost = cost + p.cw[4] * (latcost[2] / P);
  cost = cost + p.cw[5] * (sum_a / P);
  cost = cost + p.cw[6] * (sum_j / P);
  cost = cost + p.cw[7] * (sum_v / P);
  cost = cost + p.w_risk * risk;
  return cost;
}
__device__ __forceinline__ double candidate_cost(const EvalParams& p, const EvalSink& k, const double* latcost) {
  return candidate_cost(p, k.sum_a, k.sum_j, k.sum_v, k.s_first, k.s_last, k.risk, latcost);
}

// dynamic shared memory carve-up (the host uses the same function to size the launch)
struct SmemLayout {
  size_t off_broad, off_rows, off_knot, off_coef, off_time, off_lat, off_latcost, total;
};
// mode: 0 = obstacle table read from global memory, 1 = the cycle's window staged by TMA, 2 = per-block culled lists
__host__ __device__ inline SmemLayout eval_smem_layout(int win_rows, int OMpad, int mode, int K, int kpad, int P, int W, int list_cap = 0) {
  SmemLayout L;
  size_t o = 16;   // mbarrier / list cursor
  L.off_broad = o; o += mode == 1 ? (size_t)win_rows * OMpad * 16 : (mode == 2 ? (size_t)list_cap * 16 : 0);
  L.off_rows = o; o += mode == 2 ? (size_t)((4 * win_rows + 3) & ~3) * 4 : 0;   // row_off, row_cnt, row_smin, row_smax
  L.off_knot = o; o += (size_t)((K + 1) & ~1) * 8;
  L.off_coef = o; o += (size_t)8 * kpad * 8;
  L.off_time = o; o += (size_t)((P + 1) & ~1) * 8;
  L.off_lat = o; o += (size_t)W * P * 4 * 8;
  L.off_latcost = o; o += (size_t)W * 4 * 8;
  L.total = o;
  return L;
}

// Lateral quintic samples [W][P][4] = d, d_d, d_dd, d_ddd (jerk_space_sampling_planner.py:118-122) and the per-width lateral
// cost sums: one warp per width, lane l adds samples l, l + 32, ... in time order, then a fixed shuffle tree (deterministic;
// every kernel stages its tables through this function, so all of them see the same sums).  Ends with a block barrier.
__device__ __forceinline__ void stage_lateral(const EvalParams& p, const double* sm_time, double* sm_lat, double* sm_latcost) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < p.W * p.P; e += nt) {
    const int w = e / p.P, i = e - w * p.P;
    double d, d1, d2, d3;
    quintic_eval(p.quintic[w], sm_time[i], d, d1, d2, d3);
    double* o = sm_lat + (size_t)e * 4;
    o[0] = d; o[1] = d1; o[2] = d2; o[3] = d3;
  }
  __syncthreads();
  for (int w = tid >> 5; w < p.W; w += nt >> 5) {
    double s0 = 0, s1 = 0, s2 = 0;
    for (int i = tid & 31; i < p.P; i += 32) {
      const double* o = sm_lat + ((size_t)w * p.P + i) * 4;
      const double e1 = fmax(fabs(o[2]) - p.thr_lat_a, 0.0), e2 = fmax(fabs(o[3]) - p.thr_lat_j, 0.0);
      s0 = s0 + o[0] * o[0]; s1 = s1 + e1 * e1; s2 = s2 + e2 * e2;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s0 = s0 + __shfl_down_sync(0xffffffffu, s0, off);
      s1 = s1 + __shfl_down_sync(0xffffffffu, s1, off);
      s2 = s2 + __shfl_down_sync(0xffffffffu, s2, off);
    }
    if ((tid & 31) == 0) { sm_latcost[w * 4] = s0; sm_latcost[w * 4 + 1] = s1; sm_latcost[w * 4 + 2] = s2; }
  }
  __syncthreads();
}

// Stage the per-block tables: obstacle window by bulk copy (TMA engine; one row per issued copy, all completing on one
// mbarrier), spline, time grid and the lateral quintic samples (jerk_space_sampling_planner.py:118-122) by ordinary loads.
template <int kMode>
__device__ __forceinline__ void stage_tables(const EvalParams& p, unsigned char* smem, BlockTables& tb, const double*& latcost) {
  constexpr bool kSmem = kMode == 1;
  const SmemLayout L = eval_smem_layout(p.win_rows, p.OMpad, kMode, p.K, p.kpad, p.P, p.W, p.list_cap);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  float4* sm_broad = reinterpret_cast<float4*>(smem + L.off_broad);
  double* sm_knot = reinterpret_cast<double*>(smem + L.off_knot);
  double* sm_coef = reinterpret_cast<double*>(smem + L.off_coef);
  double* sm_time = reinterpret_cast<double*>(smem + L.off_time);
  double* sm_lat = reinterpret_cast<double*>(smem + L.off_lat);
  double* sm_latcost = reinterpret_cast<double*>(smem + L.off_latcost);
  const int tid = threadIdx.x, nt = blockDim.x;
  const uint32_t row_bytes = (uint32_t)p.OMpad * 16u;
  // Everything that is copied verbatim goes through the TMA engine (1-D bulk copies completing on one mbarrier): the
  // reference-line knots and coefficients, the time grid and - mode 1 - the rows of the obstacle window.
  int rows_avail = 0;
  if (kSmem) for (int r = 0; r < p.win_rows; ++r) rows_avail += (p.now + r * p.check_res < p.Tt) ? 1 : 0;   // rows that exist in the scenario table
  const uint32_t knot_bytes = (uint32_t)p.kpad * 8u, coef_bytes = (uint32_t)p.kpad * 64u, time_bytes = (uint32_t)((p.P + 1) & ~1) * 8u;
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  if (tid < 32) {
    if (tid == 0) {
      mbar_arrive_expect_tx(bar, knot_bytes + coef_bytes + time_bytes + row_bytes * (uint32_t)rows_avail);
      bulk_copy_g2s(sm_knot, p.knot, knot_bytes, bar);
      bulk_copy_g2s(sm_coef, p.coef, coef_bytes, bar);
      bulk_copy_g2s(sm_time, p.time, time_bytes, bar);
    }
    __syncwarp();
    for (int r = tid; r < rows_avail; r += 32)
      bulk_copy_g2s(sm_broad + (size_t)r * p.OMpad, p.broad + (size_t)(p.now + r * p.check_res) * p.OMpad, row_bytes, bar);
  }
  if (kSmem) for (int e = rows_avail * p.OMpad + tid; e < p.win_rows * p.OMpad; e += nt) sm_broad[e] = make_float4(0.f, 0.f, FAR_AWAY, 0.f);
  mbar_wait(bar, 0);
  stage_lateral(p, sm_time, sm_lat, sm_latcost);
  if (kSmem) {
    tb.broad = sm_broad; tb.broad_row0 = 0; tb.broad_stride = 1;
  } else {
    tb.broad = p.broad; tb.broad_row0 = p.now; tb.broad_stride = p.check_res;
  }
  tb.sp.knot = sm_knot; tb.sp.coef = sm_coef; tb.sp.K = p.K; tb.sp.kpad = p.kpad;
  tb.sp.inv_h = (double)(p.K - 1) / sm_knot[p.K - 1];
  tb.time = sm_time; tb.lat = sm_lat;
  tb.list = sm_broad; tb.row_off = reinterpret_cast<const int*>(smem + L.off_rows); tb.row_cnt = tb.row_off + p.win_rows;
  tb.use_lists = false;
  latcost = sm_latcost;
}

// Mode 2, after stage_tables: cull the obstacle states of this cycle against the arc-length range the block's candidates
// cover at every collision row, and compact the survivors into shared memory (order within a row = obstacle-mode order).
//   phase A  every thread unrolls its longitudinal profile at the collision steps; warp REDUX + shared atomics give
//            [s_min, s_max] per row (fp32, rounded outwards, clamped to the reference line);
//   phase B  one warp per row: ballot-compaction of the states whose reachable interval (pack_obstacles_kernel)
//            overlaps the row's range.
// Falls back to the global table (tb.use_lists stays false) when the lists would not fit.
// Phase A lets every thread scan the rows row_first, row_first + row_stride, ... of ITS seed: (0, 1) when every thread has a
// seed of its own, (lane-in-group, W) when the W lanes of a group share one.
// `s_at(t)` returns the arc length of the thread's seed at time t (called with non-decreasing t).
struct WalkSAt {
  LonWalk lp;
  __device__ __forceinline__ double operator()(double t) { double s, sd, sdd, sddd; lon_walk_eval(lp, t, s, sd, sdd, sddd); return s; }
};
// The list builder in three steps (build_lists runs them for the kernels whose threads each own one seed):
//   lists_begin    clear the per-row ranges / counters (ends with a block barrier)
//   phase A        every seed of the block widens [row_min, row_max] of every collision row with its arc length there
//   lists_compact  one warp per row: ballot-compaction of the states whose reachable interval overlaps the row's range
struct ListsView {
  float4* list; int* row_off; int* row_cnt; unsigned int* row_min; unsigned int* row_max; int* cursor;
};
__device__ __forceinline__ ListsView lists_view(const EvalParams& p, unsigned char* smem) {
  const SmemLayout L = eval_smem_layout(p.win_rows, p.OMpad, 2, p.K, p.kpad, p.P, p.W, p.list_cap);
  ListsView v;
  v.list = reinterpret_cast<float4*>(smem + L.off_broad);
  v.row_off = reinterpret_cast<int*>(smem + L.off_rows);
  v.row_cnt = v.row_off + p.win_rows;
  v.row_min = reinterpret_cast<unsigned int*>(v.row_cnt + p.win_rows);
  v.row_max = v.row_min + p.win_rows;
  v.cursor = reinterpret_cast<int*>(smem);        // [0] next free list entry, [1] overflow flag
  return v;
}
__device__ __forceinline__ void lists_begin(const EvalParams& p, const ListsView& v) {
  for (int r = threadIdx.x; r < p.win_rows; r += blockDim.x) { v.row_min[r] = 0x7f800000u; v.row_max[r] = 0u; v.row_cnt[r] = 0; v.row_off[r] = 0; }
  if (threadIdx.x == 0) { v.cursor[0] = 0; v.cursor[1] = 0; }
  __syncthreads();
}
__device__ __forceinline__ void lists_compact(const EvalParams& p, const ListsView& v, BlockTables& tb) {
  float4* list = v.list;
  int* row_off = v.row_off; int* row_cnt = v.row_cnt; unsigned int* row_min = v.row_min; unsigned int* row_max = v.row_max; int* cursor = v.cursor;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  (void)tid;
  __syncthreads();
  constexpr int CH = 8;                               // chunks of 32 obstacle states held in registers per pass
  for (int r = wid; r < p.win_rows; r += nw) {
    const int t = p.now + r * p.check_res;
    if (t >= p.Tt) continue;                          // no obstacle state exists beyond the scenario
    const float smin = __uint_as_float(row_min[r]) - 1e-3f, smax = __uint_as_float(row_max[r]) + 1e-3f;
    const float2* iv = p.sint + (size_t)t * p.OMpad;
    const float4* src = p.broad + (size_t)t * p.OMpad;
    if (p.OMpad <= 32 * CH) {
      // single pass: every load of the row is in flight at once
      float2 v[CH];
      float4 bb[CH];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int e = c * 32 + lane;
        v[c] = e < p.OMpad ? __ldg(iv + e) : make_float2(FAR_AWAY, -FAR_AWAY);
        bb[c] = e < p.OMpad ? __ldg(src + e) : make_float4(0.f, 0.f, FAR_AWAY, 0.f);
      }
      unsigned int m[CH];
      int cnt = 0;
#pragma unroll
      for (int c = 0; c < CH; ++c) { m[c] = __ballot_sync(0xffffffffu, v[c].y >= smin && v[c].x <= smax); cnt += __popc(m[c]); }
      int off = 0;
      if (lane == 0) off = atomicAdd(&cursor[0], cnt);
      off = __shfl_sync(0xffffffffu, off, 0);
      if (off + cnt > p.list_cap) { if (lane == 0) cursor[1] = 1; continue; }
      if (lane == 0) { row_off[r] = off; row_cnt[r] = cnt; }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if ((m[c] >> lane) & 1u) {
          float4 b = bb[c];
          b.w = __int_as_float(c * 32 + lane);
          list[off + __popc(m[c] & ((1u << lane) - 1u))] = b;
        }
        off += __popc(m[c]);
      }
      continue;
    }
    int cnt = 0;
    for (int base = 0; base < p.OMpad; base += 32) {
      const float2 v = iv[base + lane];
      cnt += __popc(__ballot_sync(0xffffffffu, v.y >= smin && v.x <= smax));
    }
    int off = 0;
    if (lane == 0) off = atomicAdd(&cursor[0], cnt);
    off = __shfl_sync(0xffffffffu, off, 0);
    if (off + cnt > p.list_cap) { if (lane == 0) cursor[1] = 1; continue; }
    if (lane == 0) { row_off[r] = off; row_cnt[r] = cnt; }
    for (int base = 0; base < p.OMpad; base += 32) {
      const float2 v = iv[base + lane];
      const bool keep = v.y >= smin && v.x <= smax;
      const unsigned int m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        float4 b = src[base + lane];
        b.w = __int_as_float(base + lane);
        list[off + __popc(m & ((1u << lane) - 1u))] = b;
      }
      off += __popc(m);
    }
  }
  __syncthreads();
#ifdef JSSP_PROFILE_PHASES
  // profiling builds: slot 15 of the block's stamps = entries of the block's lists (bit 32 = overflow: the block falls back to full rows)
  if (p.phase_ts && threadIdx.x == 0)
    p.phase_ts[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + 15] = (unsigned long long)cursor[0] | ((unsigned long long)cursor[1] << 32);
#endif
  tb.use_lists = cursor[1] == 0;
  tb.broad = p.broad; tb.broad_row0 = p.now; tb.broad_stride = p.check_res;   // fallback path
}


template <class SAt>
__device__ __forceinline__ void build_lists(const EvalParams& p, unsigned char* smem, BlockTables& tb, SAt& s_at, bool active,
                                            int row_first = 0, int row_stride = 1) {
  const ListsView v = lists_view(p, smem);
  unsigned int* row_min = v.row_min; unsigned int* row_max = v.row_max;
  const int lane = threadIdx.x & 31;
  lists_begin(p, v);
  const double s_end = tb.sp.knot[p.K - 1];
  if (row_stride == 1) {
    for (int r = 0; r < p.win_rows; ++r) {
      unsigned int lo = 0x7f800000u, hi = 0u;
      if (active) {
        double s = s_at(tb.time[r * p.check_res]);
        s = fmin(fmax(s, 0.0), s_end);                  // off-line samples are never collision-checked
        lo = __float_as_uint(__double2float_rd(s));     // s >= 0: the bit patterns order like the values
        hi = __float_as_uint(__double2float_ru(s));
      }
      lo = __reduce_min_sync(0xffffffffu, lo);
      hi = __reduce_max_sync(0xffffffffu, hi);
      if (lane == 0) { atomicMin(&row_min[r], lo); atomicMax(&row_max[r], hi); }
    }
  } else {
    // the lanes of a group share the seed: lane j of the group evaluates the rows r = j (mod W); in one round every lane
    // with the same j holds the same row, so a stride-W shuffle tree leaves the warp's range in the lanes of group 0
    const int W = row_stride;
    for (int r0 = 0; r0 < p.win_rows; r0 += W) {
      const int r = r0 + row_first;
      unsigned int lo = 0x7f800000u, hi = 0u;
      if (active && r < p.win_rows) {
        double s = s_at(tb.time[r * p.check_res]);
        s = fmin(fmax(s, 0.0), s_end);
        lo = __float_as_uint(__double2float_rd(s));
        hi = __float_as_uint(__double2float_ru(s));
      }
      for (int off = W; off < 32; off <<= 1) {
        const unsigned int lo2 = __shfl_down_sync(0xffffffffu, lo, off), hi2 = __shfl_down_sync(0xffffffffu, hi, off);
        if (lane + off < 32) { lo = min(lo, lo2); hi = max(hi, hi2); }
      }
      if (lane < W && r < p.win_rows) { atomicMin(&row_min[r], lo); atomicMax(&row_max[r], hi); }
    }
  }
  lists_compact(p, v, tb);
}

// Block index within the plan cycle and blocks per cycle.  The batch kernels run a TRANSPOSED grid (x = scenario, y = block):
// blocks are dispatched x-fastest, so block 0 of every scenario starts before any scenario's block 1, and the trailing
// blocks that find no candidates (the grid is sized for max_seeds) come last instead of sitting between the working ones.
__device__ __forceinline__ int blk_index(const EvalParams& p) { return p.batch ? (int)blockIdx.y : (int)blockIdx.x; }
__device__ __forceinline__ int blk_count(const EvalParams& p) { return p.batch ? (int)gridDim.y : (int)gridDim.x; }

__device__ __forceinline__ int resolve_n_seeds(const EvalParams& p) {
  return p.n_seeds_ptr ? min(p.n_seeds_ptr[2], p.seed_cap) : p.n_seeds_fixed;
}

// one seed row (j1, t1, ..., j5, t5) through L2-coherent loads: inside the persistent replay kernel the seeds are rewritten
// every plan cycle of the same launch, so they must not come through the read-only (non-coherent) path
__device__ __forceinline__ void load_seed_row(double (&r)[10], const double* src) {
  const double2* q = reinterpret_cast<const double2*>(src);       // rows are 80 B: 16-byte aligned
#pragma unroll
  for (int c = 0; c < 5; ++c) { const double2 v = __ldcg(q + c); r[2 * c] = v.x; r[2 * c + 1] = v.y; }
}

__device__ __forceinline__ bool lex_less(double ca, int ia, double cb, int ib) {
  // (cost, index) lexicographic; index -1 = empty
  if (ib < 0) return ia >= 0;
  if (ia < 0) return false;
  return ca < cb || (ca == cb && ia < ib);
}

// Export of the selected trajectory by one thread block: thread i owns sample i.  Positions come from the same
// lon_eval / frenet_point arithmetic as the evaluation walk (bisect and walk select the same spline segment), headings are
// the float64 arctan2 of the forward differences exactly as jerk_space_sampling_planner.py:155-165 derives them.
// `out` (global, [P][16]) doubles as the exchange buffer between the phases.
__device__ __forceinline__ void export_block_global(const EvalParams& p, const BlockTables& tb, int best, double* __restrict__ out,
                                                    int* sm_npts);

// Second half of the export, common to every candidate model: thread i holds row[0..8] of sample i (t, s.., d..) and
// *sm_npts the index of the first sample off the reference line (P if none).  Positions, forward-difference headings with
// the last one repeated, ds, c, c_d, c_dd exactly as jerk_space_sampling_planner.py:131-165; neighbours through `ex`.
constexpr int EXPORT_MAX_P = 128;
__device__ __forceinline__ void export_tail(const BlockTables& tb, double (&row)[JSSP_TRAJ_COLS], int P, double d_i, double dt,
                                            double* __restrict__ out, int* sm_npts, BestRecord* best_rec, double (*ex)[EXPORT_MAX_P]) {
  const int i = threadIdx.x;
  __syncthreads();
  const int np = *sm_npts;
  if (i < np) {
    frenet_point(tb.sp, spline_bisect(tb.sp, row[1]), row[1], d_i, row[9], row[10]);
    ex[0][i] = row[9]; ex[1][i] = row[10];
  }
  __syncthreads();
  if (i + 1 < np) {
    const double dx = ex[0][i + 1] - row[9], dy = ex[1][i + 1] - row[10];
    const double q = dx * dx + dy * dy;
    row[11] = heading_of(dx, dy);
    row[12] = (q > 1e-200 && q < 1e200) ? q * rsqrt_fast(q) : hypot(dx, dy);
    ex[2][i] = row[11];
  }
  __syncthreads();
  if (np >= 2 && i == np - 1) { row[11] = ex[2][np - 2]; ex[2][i] = row[11]; }     // yaw[-1] repeated (:158-160)
  __syncthreads();
  if (i + 1 < np) { row[13] = (ex[2][i + 1] - row[11]) / row[12]; ex[4][i] = row[13]; }
  __syncthreads();
  if (i + 2 < np) { row[14] = (ex[4][i + 1] - row[13]) / dt; ex[5][i] = row[14]; }
  __syncthreads();
  if (i + 3 < np) row[15] = (ex[5][i + 1] - row[14]) / dt;
  if (i < P) {
    double* r = out + (size_t)i * JSSP_TRAJ_COLS;
#pragma unroll
    for (int c = 0; c < JSSP_TRAJ_COLS; ++c) r[c] = row[c];
  }
  if (i == 0) best_rec->n_pts = np;
}

// Fast form for P <= EXPORT_MAX_P samples and blockDim >= P: thread i owns sample i, neighbours are exchanged through
// shared memory and every row is written to global memory once.  Same expressions as export_block_global.
__device__ __forceinline__ void export_block(const EvalParams& p, const BlockTables& tb, int best, double* __restrict__ out,
                                             int* sm_npts) {
  const int P = p.P, i = threadIdx.x;
  if (P > EXPORT_MAX_P || (int)blockDim.x < P) { export_block_global(p, tb, best, out, sm_npts); return; }
  __shared__ double ex[6][EXPORT_MAX_P];          // x, y, yaw, ds, c, c_d
  const int Ns = resolve_n_seeds(p);
  const int w = best / Ns, k = best - w * Ns;
  const double* lat = tb.lat + (size_t)w * P * 4;
  if (i == 0) *sm_npts = P;
  __syncthreads();
  const double nanv = nan("");
  double row[JSSP_TRAJ_COLS];
  if (i < P) {
    LonProfile lp;
    double seed[10];
    load_seed_row(seed, p.seeds + (size_t)k * 10);
    lon_profile_init(lp, seed, p.s0, p.v0, p.a0);
    lon_eval(lp, p.s0, p.v0, p.a0, tb.time[i], row[1], row[2], row[3], row[4]);
    row[0] = tb.time[i];
    row[5] = lat[i * 4]; row[6] = lat[i * 4 + 1]; row[7] = lat[i * 4 + 2]; row[8] = lat[i * 4 + 3];
#pragma unroll
    for (int c = 9; c < JSSP_TRAJ_COLS; ++c) row[c] = nanv;
    if (!spline_in_range(tb.sp, row[1])) atomicMin(sm_npts, i);   // first sample off the reference line ends the Cartesian path (:138-139)
  }
  export_tail(tb, row, P, i < P ? lat[i * 4] : 0.0, tb.time[1] - tb.time[0], out, sm_npts, p.best, ex);
}

__device__ __forceinline__ void export_block_global(const EvalParams& p, const BlockTables& tb, int best, double* __restrict__ out,
                                                    int* sm_npts) {
  const int P = p.P, i = threadIdx.x;
  const int Ns = resolve_n_seeds(p);
  const int w = best / Ns, k = best - w * Ns;
  const double* lat = tb.lat + (size_t)w * P * 4;
  if (i == 0) *sm_npts = P;
  __syncthreads();
  double s = 0.0;
  bool inr = false;
  for (int j = i; j < P; j += blockDim.x) {      // blockDim >= P in practice; the loop keeps small blocks correct
    LonProfile lp;
    double seed[10];
    load_seed_row(seed, p.seeds + (size_t)k * 10);
    lon_profile_init(lp, seed, p.s0, p.v0, p.a0);
    double sd, sdd, sddd;
    lon_eval(lp, p.s0, p.v0, p.a0, tb.time[j], s, sd, sdd, sddd);
    double* r = out + (size_t)j * JSSP_TRAJ_COLS;
    r[0] = tb.time[j]; r[1] = s; r[2] = sd; r[3] = sdd; r[4] = sddd;
    r[5] = lat[j * 4]; r[6] = lat[j * 4 + 1]; r[7] = lat[j * 4 + 2]; r[8] = lat[j * 4 + 3];
    for (int c = 9; c < JSSP_TRAJ_COLS; ++c) r[c] = nan("");
    inr = spline_in_range(tb.sp, s);
    if (!inr) atomicMin(sm_npts, j);             // first sample off the reference line ends the Cartesian path (:138-139)
  }
  __syncthreads();
  const int np = *sm_npts;
  for (int j = i; j < np; j += blockDim.x) {
    const double sj = out[(size_t)j * JSSP_TRAJ_COLS + 1];
    double x, y;
    frenet_point(tb.sp, spline_bisect(tb.sp, sj), sj, lat[j * 4], x, y);
    out[(size_t)j * JSSP_TRAJ_COLS + 9] = x; out[(size_t)j * JSSP_TRAJ_COLS + 10] = y;
  }
  __syncthreads();
  for (int j = i; j + 1 < np; j += blockDim.x) {
    const double dx = out[(size_t)(j + 1) * JSSP_TRAJ_COLS + 9] - out[(size_t)j * JSSP_TRAJ_COLS + 9];
    const double dy = out[(size_t)(j + 1) * JSSP_TRAJ_COLS + 10] - out[(size_t)j * JSSP_TRAJ_COLS + 10];
    const double q = dx * dx + dy * dy;
    out[(size_t)j * JSSP_TRAJ_COLS + 11] = heading_of(dx, dy);
    out[(size_t)j * JSSP_TRAJ_COLS + 12] = (q > 1e-200 && q < 1e200) ? q * rsqrt_fast(q) : hypot(dx, dy);
  }
  __syncthreads();
  if (i == 0 && np >= 2) out[(size_t)(np - 1) * JSSP_TRAJ_COLS + 11] = out[(size_t)(np - 2) * JSSP_TRAJ_COLS + 11];   // yaw[-1] repeated
  __syncthreads();
  for (int j = i; j + 1 < np; j += blockDim.x)
    out[(size_t)j * JSSP_TRAJ_COLS + 13] = (out[(size_t)(j + 1) * JSSP_TRAJ_COLS + 11] - out[(size_t)j * JSSP_TRAJ_COLS + 11]) / out[(size_t)j * JSSP_TRAJ_COLS + 12];
  __syncthreads();
  const double dt = tb.time[1] - tb.time[0];
  for (int j = i; j + 2 < np; j += blockDim.x)
    out[(size_t)j * JSSP_TRAJ_COLS + 14] = (out[(size_t)(j + 1) * JSSP_TRAJ_COLS + 13] - out[(size_t)j * JSSP_TRAJ_COLS + 13]) / dt;
  __syncthreads();
  for (int j = i; j + 3 < np; j += blockDim.x)
    out[(size_t)j * JSSP_TRAJ_COLS + 15] = (out[(size_t)(j + 1) * JSSP_TRAJ_COLS + 14] - out[(size_t)j * JSSP_TRAJ_COLS + 14]) / dt;
  if (i == 0) p.best->n_pts = np;
}

// Closed-set overlap of two centred rectangles, float64 (ReplayController._collision_detect, controller.py:382-415).
__device__ __forceinline__ bool rect_overlap_f64(double x1, double y1, double yaw1, double l1, double w1, double x2, double y2,
                                                 double yaw2, double l2, double w2) {
  double s1, c1, s2, c2;
  sincos(yaw1, &s1, &c1);
  sincos(yaw2, &s2, &c2);
  return sat_intersects_f64(x2 - x1, y2 - y1, c1, s1, l1 / 2.0, w1 / 2.0, c2, s2, l2 / 2.0, w2 / 2.0);
}

__device__ __forceinline__ void fop_store_lateral_start(EvalParams* self, double d, double d_d, double d_dd);   // jssp_fop.cuh

// a % m with Python's sign convention (result in [0, m) for m > 0): controller.py:312 wraps the yaw with it
__device__ __forceinline__ double py_mod(double a, double m) {
  double r = fmod(a, m);
  if (r != 0.0 && (r < 0.0) != (m < 0.0)) r += m;
  return r;
}

// Front-wheel angle of the pure-pursuit tracker on the selected path (intersection_planner.py:365-389, 416-443, tracing
// "plan_path"): nearest way point to the ego, walk along the path until the accumulated chord length reaches
// ld = kv * v + ld0, steer towards that point.  `traj` = exported rows [n_pts][JSSP_TRAJ_COLS] (x, y in columns 9, 10).
__device__ inline double pure_pursuit_angle(const double* traj, int n_pts, double x, double y, double yaw, double v, double kv,
                                            double ld0, double wheel_base) {
  const double ld = kv * v + ld0;
  int idx = 0;
  double bd = INFINITY;
  for (int i = 0; i < n_pts; ++i) {                  // np.argmin(np.linalg.norm(ref_pos - pos, axis=1)): first minimum
    const double dx = traj[(size_t)i * JSSP_TRAJ_COLS + 9] - x, dy = traj[(size_t)i * JSSP_TRAJ_COLS + 10] - y;
    const double d = sqrt(dx * dx + dy * dy);
    if (d < bd) { bd = d; idx = i; }
  }
  double l_steps = 0.0;
  while (l_steps < ld && idx + 1 < n_pts) {
    const double dx = traj[(size_t)(idx + 1) * JSSP_TRAJ_COLS + 9] - traj[(size_t)idx * JSSP_TRAJ_COLS + 9];
    const double dy = traj[(size_t)(idx + 1) * JSSP_TRAJ_COLS + 10] - traj[(size_t)idx * JSSP_TRAJ_COLS + 10];
    l_steps = l_steps + sqrt(dx * dx + dy * dy);
    ++idx;
  }
  const double px = traj[(size_t)idx * JSSP_TRAJ_COLS + 9], py = traj[(size_t)idx * JSSP_TRAJ_COLS + 10];
  const double alpha = atan2(py - y, px - x) - yaw;
  return atan2(2.0 * wheel_base * sin(alpha), ld);
}

// Batched replay: end-of-cycle hand-off of one scenario, run by the block that exported the selected trajectory.
__device__ __forceinline__ void batch_handoff(const EvalParams& p, int best, int N, int* sm_flag) {
  __shared__ double hk[5];                           // the ego after Env.step's kinematic update: x, y, yaw, v, a
  __shared__ double pr_d[32];
  __shared__ int pr_i[32];
  BatchSlot* b = p.batch;
  __syncthreads();                                   // the exported rows (global) are visible to the whole block
  if (threadIdx.x == 0) *sm_flag = 0;
  __syncthreads();
  const int c = b->cycle;
  // "No solution" keeps the previous best_traj (jerk_space_sampling_planner.py:359-362) and the driver re-applies ITS sample 1:
  // the export buffer still holds that trajectory.  Without a previous one (first cycle) the replay ends (-2).
  const bool fresh = best >= 0;
  const int np = fresh ? p.best->n_pts : b->last_n_pts;
  const bool ok = np >= 2;                           // env.py:26 needs best_traj.x[1]
  const double* row1 = p.best_traj + JSSP_TRAJ_COLS;
  double tn = 0.0, x = 0.0, y = 0.0, yaw = 0.0;
  // Env.step in planner_expected_value mode (env.py:21-33 -> controller.py:256-263): the ego is first advanced by the
  // kinematic model with the action (0, 0) (controller.py:301-319) and the END STATUS is evaluated on that pose; only then is
  // the ego overwritten with sample 1 of the selected trajectory.
  // In kinematics mode the action is (planned acceleration s_dd[1], pure-pursuit angle) (code/__main__.py:129-138; after "No
  // solution" plan_traj hands the previous trajectory back, so both come from it again) and that pose IS the new ego.
  const bool kin = b->kin != 0;
  if (threadIdx.x == 0) {
    double a = 0.0, rot = 0.0;
    if (kin && ok) {
      a = row1[3];
      rot = pure_pursuit_angle(p.best_traj, np, b->ego_x, b->ego_y, b->ego_yaw, b->ego_v, b->pp_kv, b->pp_ld0, b->pp_wb);
    }
    a = fmin(fmax(a, -15.0), 15.0); rot = fmin(fmax(rot, -1.0), 1.0);                          // _action_cheaker (controller.py:265-268)
    const double PI = 3.14159265358979323846;
    hk[0] = b->ego_x + b->ego_v * cos(b->ego_yaw) * b->dt;
    hk[1] = b->ego_y + b->ego_v * sin(b->ego_yaw) * b->dt;
    hk[2] = py_mod(b->ego_yaw + b->ego_v / b->ego_l * 1.7 * tan(rot) * b->dt + PI, 2.0 * PI) - PI;
    const double vn = b->ego_v + a * b->dt;
    hk[3] = vn < 0.0 ? 0.0 : vn;
    hk[4] = a;
  }
  __syncthreads();
  const double kx = hk[0], ky = hk[1], kyaw = hk[2];
  if (ok) {
    tn = b->cycle_time[c + 1];                       // t + dt (controller.py:303)
    x = row1[9]; y = row1[10]; yaw = row1[11];
    const int k = b->cycle_end_step[c + 1];
    if (tn > 0.5 && k >= 0 && k < b->gt_Tt) {        // collision with a background vehicle (controller.py:343-349, 382-398)
      for (int o = threadIdx.x; o < b->gt_O; o += blockDim.x) {
        if (!b->gt_valid[(size_t)o * b->gt_Tt + k]) continue;
        const double* st = b->gt_state + ((size_t)o * b->gt_Tt + k) * 3;
        if (rect_overlap_f64(kx, ky, kyaw, b->ego_l, b->ego_w, st[0], st[1], st[2], b->gt_size[2 * o], b->gt_size[2 * o + 1])) atomicOr(sm_flag, 1);
      }
    }
  }
  // kinematics mode: the next cycle starts from FrenetState.from_state(new ego, polyline) - nearest polyline point (first
  // minimum of the squared distance, like numpy.argmin), found by the whole block
  int near = 0;
  if (kin && ok) {
    double bd = INFINITY;
    int bi = -1;
    for (int m = threadIdx.x; m < b->n_poly; m += blockDim.x) {
      const double ex = b->poly[(size_t)m * 6] - kx, ey = b->poly[(size_t)m * 6 + 1] - ky;
      const double d2 = ex * ex + ey * ey;
      if (d2 < bd) { bd = d2; bi = m; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_down_sync(0xffffffffu, bd, o);
      const int oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || od < bd || (od == bd && oi < bi))) { bd = od; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { pr_d[threadIdx.x >> 5] = bd; pr_i[threadIdx.x >> 5] = bi; }
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  if (kin && ok) {
    double bd = pr_d[0];
    int bi = pr_i[0];
    for (int q = 1; q < (int)((blockDim.x + 31) >> 5); ++q)
      if (pr_i[q] >= 0 && (bi < 0 || pr_d[q] < bd || (pr_d[q] == bd && pr_i[q] < bi))) { bd = pr_d[q]; bi = pr_i[q]; }
    near = bi < 0 ? 0 : bi;
  }
  jssp_cycle_record* rec = b->records + c;
  if (p.n_seeds_ptr && p.n_seeds_ptr[3]) b->seed_overflow = 1;
  rec->best_idx = best; rec->n_candidates = N; rec->best_cost = best >= 0 ? p.best->cost : nan("");
  if (!ok) {
    rec->t = rec->x = rec->y = rec->v = rec->a = rec->yaw = nan("");
    b->end = (!fresh && b->last_n_pts == 0) ? -2.0 : -3.0;
    b->live = 0;
  } else {
    if (kin) { x = kx; y = ky; yaw = kyaw; }
    const double vnew = kin ? hk[3] : row1[2], anew = kin ? hk[4] : row1[3];
    rec->t = tn; rec->x = x; rec->y = y; rec->v = vnew; rec->a = anew; rec->yaw = yaw;
    double status = -1.0;
    if (*sm_flag) status = 2.0;
    if (tn >= b->max_t) status = fmax(status, 1.0);                                          // controller.py:352-353
    if (kx > b->bx_max || kx < b->bx_min || ky > b->by_max || ky < b->by_min) status = fmax(status, 1.5);   // out of the map (:362-369)
    const double dx = kx - b->goal_x, dy = ky - b->goal_y;
    if (b->goal_sin * dy + b->goal_cos * dx >= 0.0) status = fmax(status, 1.5);               // goal line passed (:372-377)
    b->ego_x = x; b->ego_y = y; b->ego_yaw = yaw; b->ego_v = vnew;                            // env.py:26-30 (kinematics: the model's pose)
    if (fresh) b->last_n_pts = np;
    b->end = status;
    if (status != -1.0) b->live = 0;
    // next cycle starts from sample 1 of the selected trajectory (intersection_planner.py:312) or, in kinematics mode, from
    // the projection of the new ego on the re-discretised reference line (:308-310; dlii_jssp_b200.frenet.FrenetState.from_state)
    double fs = row1[1], fsd = row1[2], fsdd = row1[3], fd = row1[5], fdd = row1[6], fddd = row1[7];
    if (kin) {
      const double* pl = b->poly + (size_t)near * 6;
      const double ex = x - pl[0], ey = y - pl[1], tx = pl[4], ty = pl[5];
      const double dyaw = yaw - pl[2];
      fs = pl[3] + (ex * tx + ey * ty);
      fd = -ex * ty + ey * tx;
      fsd = vnew * cos(dyaw); fdd = vnew * sin(dyaw);
      fsdd = anew * cos(dyaw); fddd = anew * sin(dyaw);
    }
    EvalParams* q = p.self;
    q->s0 = fs; q->v0 = fsd; q->a0 = fsdd;
    if (b->is_fop) fop_store_lateral_start(q, fd, fdd, fddd);
    else for (int w = 0; w < p.W; ++w) quintic_coeffs(fd, fdd, fddd, b->targets[w], 0.0, 0.0, b->horizon, q->quintic[w]);
    const int now = b->cycle_now[c + 1];
    q->now = now; q->tcap = b->final_step - now;
    int rows = 0;
    if (q->tcap > 0) rows = (min(p.P, q->tcap) + p.check_res - 1) / p.check_res;
    q->win_rows = p.OM > 0 ? rows : 0;
  }
  b->cycle = c + 1;
  if (c + 1 >= b->n_cycles_cap) b->live = 0;         // table end: stays "running" (end = -1) like a max_cycles stop
}

// Selection shared by both evaluation kernels: block argmin of (cost, index) by warp shuffles, one record per block,
// and the last block to finish (ticket counter) reduces the per-block winners - deterministic, lexicographic, no
// floating-point atomics - and exports the winning trajectory in the same launch.
template <int kMode>
__device__ __forceinline__ void select_and_export(const EvalParams& p, unsigned char* smem, BlockTables& tb, const double*& latcost,
                                                  bool staged, double my_cost, int my_idx, int N, int Ns, double* red_cost,
                                                  int* red_idx, bool* is_last_p, int* sm_npts_p) {
  const int nwarps = blockDim.x >> 5;
  // block argmin (cost, index) - warp shuffles, then one value per warp through shared memory
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oc = __shfl_down_sync(0xffffffffu, my_cost, o);
    const int oi = __shfl_down_sync(0xffffffffu, my_idx, o);
    if (lex_less(oc, oi, my_cost, my_idx)) { my_cost = oc; my_idx = oi; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { red_cost[wid] = my_cost; red_idx[wid] = my_idx; }
  PHASE(p, 7);
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < nwarps; ++q)
      if (lex_less(red_cost[q], red_idx[q], my_cost, my_idx)) { my_cost = red_cost[q]; my_idx = red_idx[q]; }
    p.blk_cost[blk_index(p)] = my_cost;
    p.blk_idx[blk_index(p)] = my_idx;
    __threadfence();
    const unsigned int t = atomicAdd(p.ticket, 1u);
    *is_last_p = (t == (unsigned int)blk_count(p) - 1u);
  }
  __syncthreads();
  if (!*is_last_p) return;
  // the last block to finish reduces the per-block winners: deterministic, no floating-point atomics
  __threadfence();
  my_cost = INFINITY; my_idx = -1;
  for (int b = threadIdx.x; b < blk_count(p); b += blockDim.x) {
    const double c = __ldcg(p.blk_cost + b);
    const int i = __ldcg(p.blk_idx + b);
    if (lex_less(c, i, my_cost, my_idx)) { my_cost = c; my_idx = i; }
  }
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
    *p.ticket = 0u;   // re-arm for the next launch (stream-ordered)
    red_idx[0] = my_idx;
    red_cost[0] = my_cost;
  }
  __syncthreads();
  const int best = red_idx[0];
  // candidate split: the rank's record with the GLOBAL candidate index - pushed to the peers' mailboxes right away (their
  // wait overlaps this block's export) and / or left in device memory for the NCCL all-gather
  const bool split = p.split_ns_global > 0;
  unsigned long long skey = 0xffffffffffffffffull, sidx = 0xffffffffffffffffull;
  if (split && threadIdx.x < 32) {
    if (best >= 0) {
      const long long w = best / Ns, k = best - w * Ns;
      skey = orderable_bits(red_cost[0]);
      sidx = (unsigned long long)(w * p.split_ns_global + p.split_seed_lo + k);
    }
    if (threadIdx.x == 0 && p.split_rec) { p.split_rec->key = skey; p.split_rec->idx = sidx; }
    if (p.peers) split_push_peers(*p.peers, p.split_epoch, skey, sidx);
  }
  // the same block exports the winner: no second launch, no host round trip in between
  if (p.best_traj != nullptr) {
    PHASE(p, 5);
    if (best >= 0) {
      if (!staged) stage_tables<kMode>(p, smem, tb, latcost);
      export_block(p, tb, best, p.best_traj, sm_npts_p);
    }
    PHASE(p, 6);
  }
  if (split && p.peers && threadIdx.x < 32) split_wait_peers(*p.peers, p.split_epoch);
  if (p.best_traj == nullptr) return;
  if (p.batch) batch_handoff(p, best, N, sm_npts_p);
}


template <int kMode>
__device__ __forceinline__ void evaluate_body(const EvalParams& p, unsigned char* smem, double* red_cost, int* red_idx, bool* is_last,
                                              int* sm_npts) {
  const int Ns = resolve_n_seeds(p);
  const int N = Ns * p.W;
  const int n = blk_index(p) * blockDim.x + threadIdx.x;
  double my_cost = INFINITY;
  int my_idx = -1;
  BlockTables tb;
  const double* latcost;
  const bool staged = blk_index(p) * (int)blockDim.x < N;   // block-uniform
  PHASE(p, 0);
  if (staged) {
    stage_tables<kMode>(p, smem, tb, latcost);
    PHASE(p, 1);
    const bool active = n < N;
    const int w = active ? n / Ns : 0, k = active ? n - w * Ns : 0;
    LonWalk lp;
    lon_walk_init(lp, p.seeds + (size_t)k * 10, p.s0, p.v0, p.a0);
    if (kMode == 2) { WalkSAt s_at{lp}; build_lists(p, smem, tb, s_at, active); }
    PHASE(p, 2);
    if (active) {
      EvalSink sink(p, tb);
      const int seg0 = spline_bisect(tb.sp, fmin(fmax(p.s0, tb.sp.knot[0]), tb.sp.knot[p.K - 1]));
      walk_candidate(p, tb, lp, w, seg0, sink);
      const double cost = candidate_cost(p, sink, latcost + w * 4);
      if (p.feasible) p.feasible[n] = sink.curv_fail ? 0 : 1;
      if (p.collision) p.collision[n] = sink.collided ? 1 : 0;
      if (p.cost) p.cost[n] = (float)cost;
      if (!sink.curv_fail && !sink.collided && isfinite(cost)) { my_cost = cost; my_idx = n; }
    }
  }
#ifdef JSSP_PROFILE_PHASES
  __syncthreads();
#endif
  PHASE(p, 3);
  select_and_export<kMode>(p, smem, tb, latcost, staged, my_cost, my_idx, N, Ns, red_cost, red_idx, is_last, sm_npts);
  PHASE(p, 4);
}

template <int kMode>
__global__ void __launch_bounds__(EVAL_THREADS, EVAL_MIN_BLOCKS) evaluate_kernel(const __grid_constant__ EvalParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  evaluate_body<kMode>(p, smem, red_cost, red_idx, &is_last, &sm_npts);
}

// K2 (default for large cycles): the group form - W consecutive lanes evaluate the W lateral targets of one seed and share
// the seed-only work (walk_group).  Seeds per block = warps * floor(32 / W); lanes beyond floor(32 / W) * W idle.
__host__ __device__ inline int group_seeds_per_block(int W, int threads = EVAL_THREADS) { return (threads / 32) * (32 / W); }

template <int kMode>
__device__ __forceinline__ void evaluate_group_body(const EvalParams& p, unsigned char* smem, double* red_cost, int* red_idx,
                                                    bool* is_last, int* sm_npts) {
  const int Ns = resolve_n_seeds(p);
  const int W = p.W;
  const int N = Ns * W;
  const int gpw = 32 / W;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int g = lane / W, j = lane - g * W;
  const int spb = (blockDim.x >> 5) * gpw;     // the host picks the block size (<= EVAL_THREADS)
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
    LonWalk lp;
    lon_walk_init(lp, p.seeds + (size_t)(active ? k : 0) * 10, p.s0, p.v0, p.a0);
    if (kMode == 2) { WalkSAt s_at{lp}; build_lists(p, smem, tb, s_at, active, j, W); }
    PHASE(p, 2);
    if (active) {
      const unsigned int gmask = (W >= 32 ? 0xffffffffu : ((1u << W) - 1u)) << (g * W);
      EvalSink sink(p, tb);
      walk_group(p, tb, lp, j, W, gmask, g * W, sink);
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
__global__ void __launch_bounds__(EVAL_THREADS, EVAL_MIN_BLOCKS) evaluate_group_kernel(const __grid_constant__ EvalParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  evaluate_group_body<kMode>(p, smem, red_cost, red_idx, &is_last, &sm_npts);
}

// Batched replay: blockIdx.x = scenario, blockIdx.y = block of the scenario (see blk_index); the scenario's parameter record is copied into shared memory once per block.
__device__ __forceinline__ void load_slot(EvalParams* dst, const EvalParams* src) {
  static_assert(sizeof(EvalParams) % 8 == 0, "EvalParams is copied as 8-byte words");
  const unsigned long long* s = reinterpret_cast<const unsigned long long*>(src);
  unsigned long long* d = reinterpret_cast<unsigned long long*>(dst);
  for (int i = threadIdx.x; i < (int)(sizeof(EvalParams) / 8); i += blockDim.x) d[i] = s[i];
  __syncthreads();
}

template <int kMode>
__global__ void __launch_bounds__(BATCH_EVAL_THREADS, BATCH_MIN_BLOCKS) evaluate_batch_kernel(const EvalParams* __restrict__ slots) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(16) EvalParams sp;
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  pdl_launch_dependents(); pdl_wait();                 // seeds and slot records come from the kernels ahead in the stream
  if (!slot_live(slots, blockIdx.x)) return;           // block-uniform; stable until the scenario's last block hands off
  load_slot(&sp, slots + blockIdx.x);
  evaluate_body<kMode>(sp, smem, red_cost, red_idx, &is_last, &sm_npts);
}

template <int kMode>
__global__ void __launch_bounds__(BATCH_EVAL_THREADS, BATCH_MIN_BLOCKS) evaluate_group_batch_kernel(const EvalParams* __restrict__ slots) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(16) EvalParams sp;
  __shared__ double red_cost[EVAL_THREADS / 32];
  __shared__ int red_idx[EVAL_THREADS / 32];
  __shared__ bool is_last;
  __shared__ int sm_npts;
  pdl_launch_dependents(); pdl_wait();
  if (!slot_live(slots, blockIdx.x)) return;
  load_slot(&sp, slots + blockIdx.x);
  evaluate_group_body<kMode>(sp, smem, red_cost, red_idx, &is_last, &sm_npts);
}

}  // namespace jssp
