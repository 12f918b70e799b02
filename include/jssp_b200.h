/*
 * jssp_b200.h - C ABI of libjssp_b200.so: the B200 (sm_100a) implementation of the DLII-JSSP
 * jerk-space sampling planner's per-cycle candidate evaluation.
 *
 * Reference = byChenZhifa/DLII-JSSP; every entry point names the reference interface it replaces
 * (paths relative to the reference root, `jssp.py` = code/intersection_local_planner/jerk_space_sampling_planner.py,
 * `polyvt.py` = code/intersection_local_planner/polyvt_sampling.py).  The reference is pure Python, so there is no
 * existing FFI; INTEGRATION.md shows the ctypes stub a maintainer adds.
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns an int32 status (0 = JSSP_OK, < 0 = error, see
 *     jssp_strerror) and never throws.
 *   - "dev" pointers are CUDA device pointers owned by the CALLER (e.g. torch tensors, pass tensor.data_ptr());
 *     "host" pointers are ordinary host memory.  The library allocates only its private workspace in jssp_create
 *     (sized from the capacities in jssp_config) and frees it in jssp_destroy; nothing is allocated per cycle.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  Calls are stream-ordered; only the calls
 *     documented as synchronising block the host.
 *   - a handle is bound to one device and is not thread-safe.
 */
#ifndef JSSP_B200_H_
#define JSSP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JSSP_ABI_VERSION 1
#define JSSP_MAX_WIDTHS 16     /* lateral targets per cycle */
#define JSSP_TRAJ_COLS 16      /* columns of an exported trajectory row, see jssp_cycle_out.best_traj */
#define JSSP_STATE_COLS 8      /* columns of the optional per-candidate state dump */

enum jssp_status {
  JSSP_OK = 0,
  JSSP_ERR_INVALID_ARG = -1,
  JSSP_ERR_CUDA = -2,
  JSSP_ERR_CAPACITY = -3,      /* more seeds / knots / obstacles / steps than the handle was created for */
  JSSP_ERR_NOT_READY = -4,     /* reference line, obstacles or seeds missing */
  JSSP_ERR_SEED_OVERFLOW = -5, /* seed enumeration produced more than max_seeds seeds */
  JSSP_ERR_NO_DEVICE = -6
};

/* Planner constants.  Defaults (jssp_default_config) are the reference's values:
 *   JerkSpaceSamplingPlannerSettings  jssp.py:35-64 with configuration_parameters.py:52-61
 *   parameters["vehicle_para"], ["para_threshold"]  configuration_parameters.py:7-22
 *   PlannerSettings.max_road_width  intersection_planner.py:60 */
typedef struct jssp_config {
  double delta_t;            /* 0.1 */
  double plan_horizon;       /* 5.0  -> n_steps = int(plan_horizon / 0.1), polyvt.py:44 */
  double max_road_width;     /* 3.5 */
  int32_t num_width;         /* 3 */
  int32_t num_jerk;          /* 25 (PolyVTSampling num_samples) */
  double highest_jerk;       /* 10.0 */
  double speed_max;          /* 18.0 */
  double lon_accel_max;      /* 8.0 */
  double ego_length;         /* ego footprint, centred rectangle */
  double ego_width;
  double max_curvature;      /* Vehicle.max_curvature, builder-specified 0.2 1/m */
  double obstacle_inflation; /* 0.1, jssp.py:263-265 */
  int32_t check_res;         /* 2, jssp.py:280 */
  int32_t reserved0;
  double cost_weights[8];    /* w_time, w_distance, w_laneoffset, w_lat_a, w_lat_j, w_lon_a, w_lon_j, w_lon_v */
  double max_distance;       /* 140.0 */
  double thr_lon_accel, thr_lon_jerk, thr_lat_accel, thr_lat_jerk; /* 3.0 6.0 0.5 1.0 */
  double p_min;              /* multi-modal extension: modes with prob >= p_min veto (0 = every mode) */
  double w_risk;             /* weight of the summed probability of the non-vetoing modes hit */
  /* capacities of the private workspace */
  int32_t max_seeds;         /* longitudinal seeds per cycle (enumerated or external) */
  int32_t max_knots;         /* reference-line knots */
  int32_t max_obstacle_modes;/* O * M */
  int32_t max_total_steps;   /* scenario length in steps (Tt) */
} jssp_config;

typedef struct jssp_handle jssp_handle;

/* One plan cycle's inputs = the arguments of JerkSpaceSamplingPlanner.plan_traj (jssp.py:296-300):
 * frenet_state (s, s_d, s_dd, d, d_d, d_dd), time_step_now = int(t / dt), final_time_step = int(max_t / dt)
 * (jssp.py:244). */
typedef struct jssp_cycle_in {
  double s0, v0, a0;           /* frenet_state.s, s_d, s_dd */
  double d0, d_d0, d_dd0;      /* frenet_state.d, d_d, d_dd */
  int32_t time_step_now;
  int32_t final_time_step;
  int32_t n_widths;            /* 0 -> cfg.num_width targets from jssp.py:100-104; else use targets[] */
  int32_t flags;               /* JSSP_FLAG_* */
  double targets[JSSP_MAX_WIDTHS];
  const double* seeds_dev;     /* NULL -> the seeds of the last jssp_enumerate_seeds; else dev [n_seeds,10] */
  int32_t n_seeds;             /* rows of seeds_dev (ignored when seeds_dev == NULL) */
  int32_t reserved0;
} jssp_cycle_in;

#define JSSP_FLAG_NO_SYNC 1      /* do not wait / do not fill the host fields of jssp_cycle_out */
#define JSSP_FLAG_NO_EXPORT 2    /* skip the export of the selected trajectory */
#define JSSP_FLAG_ENUMERATE 8    /* run jssp_enumerate_seeds(s0, v0, a0) first: one call = one whole plan cycle */
#define JSSP_FLAG_FORCE_WARP 16  /* use the warp-per-candidate kernel (default for <= 16384 candidates) */
#define JSSP_FLAG_FORCE_THREAD 32 /* use the thread-per-candidate kernel; same results */
#define JSSP_FLAG_FORCE_GROUP 64 /* use the group kernel: the lateral targets of a seed sit in consecutive lanes and share the
                                    seed-only work by warp shuffles; same results */
#define JSSP_FLAG_FORCE_TILED 128 /* use the tiled two-phase kernel (default for larger cycles): per tile of the horizon the seed-only
                                    work runs in parallel into shared memory, then one lane per candidate walks the tile; same masks
                                    and selection, costs equal to rounding */
#define JSSP_FLAG_LOCKSTEP 256    /* jssp_batch_run only: the lock-step pipeline (two launches per plan cycle for the whole batch; the default) */
#define JSSP_FLAG_PERSISTENT 512  /* jssp_batch_run only: the persistent replay kernel - one CTA per scenario loops seed enumeration -> evaluation ->
                                    selection -> export -> hand-off over the scenario's own plan cycles, ONE launch per call, no lock-step between
                                    scenarios; same results (measured slower than the lock-step pipeline on B200) */
#define JSSP_FLAG_CHUNKED 1024    /* jssp_batch_run only: the chunked tiled kernel - two lane groups per seed split the horizon, the halves are
                                    combined in time order with the serial walk's semantics (same selection and rows); chosen automatically
                                    when the batch leaves the GPU half empty (JSSP_FLAG_FORCE_TILED keeps the whole-horizon walk) */
#define JSSP_FLAG_BEST_FIRST 2048 /* jssp_batch_run only: the persistent replay kernel with BEST-FIRST selection - a cycle's candidates are ordered by a
                                    lower bound of their cost (everything but the risk term, which needs no walk) and checked in that order, one
                                    warp per candidate with the lanes over the samples, until no unchecked candidate can cost less than the best
                                    survivor: the selection, its cost and the exported trajectory are those of the exhaustive evaluation */
#define JSSP_FLAG_SMALL_CTAS 4096 /* with JSSP_FLAG_BEST_FIRST: 256-thread CTAs (two per SM) also for a launch of few scenarios - for callers that run
                                    several sub-batches concurrently on separate streams.  Default: 512-thread CTAs (one per SM, 16 warps checking
                                    candidates) up to 3 scenarios per SM, 256-thread CTAs above (measured: 200 scenarios 12.8 vs 18.1 ms, 400
                                    18.9 vs 20.5 ms, 800 29.7 vs 28.4 ms) */
#define JSSP_FLAG_NO_CULL 4      /* test every obstacle state instead of the per-block culled lists (same results) */

/* Outputs.  Candidate n = w * n_seeds + k (width-major, jssp.py:106,124-125).  All dev pointers may be NULL.  With
 * feasible_dev, collision_dev and cost_dev all NULL the evaluation is select-only: a candidate's walk stops at its first
 * veto - the reference filters the same way (check_constraints_curvature -> check_collisions on the survivors -> cost on the
 * survivors, has_collision returning at the first hit; jssp.py:311-319) - and the selection is unchanged. */
typedef struct jssp_cycle_out {
  uint8_t* feasible_dev;       /* [N] 1 = passes check_constraints_curvature (jssp.py:171-180) */
  uint8_t* collision_dev;      /* [N] 1 = has_collision (jssp.py:233-272), evaluated for every candidate */
  float* cost_dev;             /* [N] cost_total (jssp.py:288-294) rounded to fp32; +inf is never written */
  float* states_dev;           /* [N, n_steps+1, 8] = s, s_d, s_dd, d, x, y, yaw, c (NaN beyond n_pts); optional */
  double* best_traj_host;      /* host [n_steps+1, 16] = t, s, s_d, s_dd, s_ddd, d, d_d, d_dd, d_ddd, x, y, yaw,
                                  ds, c, c_d, c_dd of the selected candidate; filled unless NO_SYNC/NO_EXPORT */
  /* host results (valid after a synchronising call) */
  int32_t n_candidates;
  int32_t best_idx;            /* -1 = no candidate survived ("No solution", jssp.py:359-360) */
  int32_t best_n_pts;          /* Cartesian points of the best trajectory (< n_steps+1 when s leaves the spline) */
  int32_t n_seeds;
  double best_cost;
} jssp_cycle_out;

/* -- lifecycle ------------------------------------------------------------------------------------------ */
int32_t jssp_abi_version(void);
/* 16 hex digits: hash of the sources and compiler flags the library was built from (dlii_jssp_b200/build.py).  The Python
 * loader rebuilds / refuses the library when it differs from the sources on disk; profiles/ records it next to every ncu
 * capture so that a counter can be tied to the code that produced it. */
const char* jssp_build_id(void);
/* sizeof of the public structs as this library was compiled, for a binding to check its own declarations against:
 * which = 0 jssp_config, 1 jssp_cycle_in, 2 jssp_cycle_out, 3 jssp_fop_in, 4 jssp_imm_params, 5 jssp_cycle_record,
 * 6 jssp_scenario_result, 7 jssp_scenario_in; -1 for anything else */
int32_t jssp_struct_size(int32_t which);
const char* jssp_strerror(int32_t status);
/* last CUDA error string seen by this handle (for JSSP_ERR_CUDA) */
const char* jssp_last_cuda_error(const jssp_handle* h);
void jssp_default_config(jssp_config* cfg);
/* replaces JerkSpaceSamplingPlanner.__init__ (jssp.py:67-78) */
int32_t jssp_create(const jssp_config* cfg, int32_t device, jssp_handle** out);
int32_t jssp_destroy(jssp_handle* h);

/* -- host-only helpers (no GPU needed) --------------------------------------------------------------------- */
/* The sampling grids of PolyVTSampling.__init__ (polyvt.py:33-58) and of the two samplers (:181, :274-279),
 * bit-identical to numpy's linspace / arange.  Each out array must hold `cap` doubles; n_* receive the lengths in
 * the order time_array, j1, t1, t2, j3, t3, t4, js, ts. */
int32_t jssp_make_grids(double total_time, double jerk, int32_t num_samples, int32_t cap, double* time_array, double* j1,
                        double* t1, double* t2, double* j3, double* t3, double* t4, double* js, double* ts,
                        int32_t n_out[9]);
/* Lateral quintic coefficients a0..a5 (QuinticPolynomial, jssp.py:109-117), closed form. */
void jssp_quintic_coeffs(double xs, double vxs, double axs, double xe, double vxe, double axe, double T, double out[6]);

/* -- per-scenario state ---------------------------------------------------------------------------------- */
/* replaces the assignment `planner.fplanner.cubic_spline = CubicSpline2D(...)` (code/__main__.py:100, jssp.py:365):
 * host knot_s[K] (cumulative chord length, knot_s[0] = 0) and host coef[K-1][8] = ax,bx,cx,dx,ay,by,cy,dy per segment. */
int32_t jssp_set_refline(jssp_handle* h, const double* knot_s, const double* coef, int32_t K, void* stream);
/* replaces the `obstacles` dict argument of plan_traj / has_collision (jssp.py:233-272; format controller.py:92-107):
 * host state[O][M][Tt][3] = x, y, yaw; host valid[O][M][Tt]; host size[O][2] = length, width (not inflated);
 * host prob[O][M]; n_dict = len(obstacles) for the `len(obstacles) <= 0` early-out (jssp.py:241).  M = 1, prob = 1
 * is the reference's ground-truth replay case. */
/* The upload and the table packing run on a stream private to the handle (after the last evaluation has finished reading the
 * tables); the next jssp_evaluate / jssp_fop_evaluate makes its stream wait for them.  `stream` is therefore not blocked: the
 * caller can keep it busy with the other input of the cycle (e.g. the H2D copy of external seeds). */
int32_t jssp_set_obstacles(jssp_handle* h, const double* state, const uint8_t* valid, const double* size,
                           const double* prob, int32_t O, int32_t M, int32_t Tt, int32_t n_dict, void* stream);
/* same, from caller-owned DEVICE buffers (e.g. the output of jssp_imm_update / jssp_predict_modes) */
int32_t jssp_set_obstacles_dev(jssp_handle* h, const double* state_dev, const uint8_t* valid_dev, const double* size_dev,
                               const double* prob_dev, int32_t O, int32_t M, int32_t Tt, int32_t n_dict, void* stream);

/* -- per-cycle ------------------------------------------------------------------------------------------- */
/* replaces PolyVTSampling.sampling + sampling_stop_seeds_supplement (polyvt.py:167-305; jssp.py:95-97).
 * Seeds stay on the device.  n_main / n_stop may be NULL (then the call does not synchronise). */
int32_t jssp_enumerate_seeds(jssp_handle* h, double s0, double v0, double a0, void* stream, int32_t* n_main,
                             int32_t* n_stop);
/* copy the enumerated seeds to host [cap,10]; synchronises; returns the count in *n */
int32_t jssp_get_seeds(jssp_handle* h, double* seeds_host, int32_t cap, int32_t* n, void* stream);
/* replaces the body of JerkSpaceSamplingPlanner.plan_traj after sampling (jssp.py:303-327):
 * get_all_frenet_trajectory (polyvt.py:307-385), calc_global_paths, check_constraints_curvature, check_collisions,
 * calc_all_trajs_cost and the PriorityQueue argmin, fused. */
int32_t jssp_evaluate(jssp_handle* h, const jssp_cycle_in* in, void* stream, jssp_cycle_out* out);
/* fetch the results of a JSSP_FLAG_NO_SYNC evaluate (synchronises the stream) */
int32_t jssp_fetch_best(jssp_handle* h, void* stream, jssp_cycle_out* out);
/* write the selection record of the last evaluate into caller-owned device memory: record_dev[0] = orderable bits of
 * the float64 cost (unsigned order == numeric order), record_dev[1] = candidate index + index_offset (UINT64_MAX when
 * no candidate survived) - what a candidate-split run feeds to its NCCL reduction. */
int32_t jssp_best_record_dev(jssp_handle* h, int64_t index_offset, void* stream, uint64_t* record_dev);
/* -- candidate split over ranks (BASELINE configs[4]) ------------------------------------------
 * The reference selects over ALL candidates of a cycle (PriorityQueue over calc_all_trajs_cost, jssp.py:288-294, 326-327).
 * When one huge cycle is split over G ranks - contiguous seed ranges, every width on every rank - the only exchange is the
 * per-rank selection record (orderable bits of the float64 cost, GLOBAL candidate index w * n_seeds_global + seed_lo + k):
 * 16 bytes, lexicographic minimum = lowest cost, then lowest global index = the 1-GPU selection.
 *
 * jssp_set_split: this handle's external seeds are rows [seed_lo, seed_lo + n_seeds) of a cycle with n_seeds_global seeds
 *   (0, 0 switches the split off).  Every following jssp_evaluate leaves the rank's record on the device.
 * Transport A - NCCL (the "single NCCL min-reduce" of the north star, exact for a float64 cost + 64-bit index):
 *   jssp_allreduce_argmin enqueues on `stream` one ncclAllGather of the 16-byte record (2 x ncclUint64) on `nccl_comm`
 *   (an ncclComm_t as void*: the caller's own communicator, or one made by jssp_nccl_comm_create from a
 *   jssp_nccl_unique_id broadcast by the caller) and a one-warp kernel that takes the lexicographic minimum.  world = 1
 *   skips the collective.  NCCL is resolved at run time from the libnccl.so.2 already in the process (dlopen; override with
 *   JSSP_NCCL_LIB); JSSP_ERR_NOT_READY when it cannot be found.
 * Transport B - peer mailboxes over NVLink, fused into the evaluation launch: jssp_peer_export allocates this rank's
 *   mailbox and returns its 64-byte CUDA IPC handle; after the caller has all-gathered the handles, jssp_peer_connect maps
 *   every peer's mailbox.  From then on the block that finishes a rank's evaluation stores the record into every peer's
 *   mailbox and one of its warps waits for the world's records and reduces them - no collective kernel, no host step.
 * jssp_fetch_global_best synchronises `stream` and returns the global selection (best_idx = -1: no survivor anywhere) and
 *   the rank that holds it (that rank's exported trajectory is the global best_traj).  Identical on every rank. */
int32_t jssp_set_split(jssp_handle* h, int64_t seed_lo, int64_t n_seeds_global);
int32_t jssp_nccl_unique_id(uint8_t id_out[128]);
int32_t jssp_nccl_comm_create(const uint8_t id[128], int32_t world, int32_t rank, int32_t device, void** comm_out);
int32_t jssp_nccl_comm_destroy(void* comm);
int32_t jssp_allreduce_argmin(jssp_handle* h, void* nccl_comm, int32_t world, void* stream);
int32_t jssp_peer_export(jssp_handle* h, int32_t world, uint8_t handle_out[64]);
int32_t jssp_peer_connect(jssp_handle* h, const uint8_t* handles /* [world][64] */, int32_t world, int32_t rank);
int32_t jssp_peer_disconnect(jssp_handle* h);
int32_t jssp_fetch_global_best(jssp_handle* h, void* stream, int64_t* best_idx, double* best_cost, int32_t* winner_rank);
/* number of kernels the library has launched through this handle so far */
int64_t jssp_launch_count(const jssp_handle* h);
/* bench utility: measured dense FMA throughput of the CUDA cores in TFLOP/s (bits = 32 or 64): the roofline
 * denominator bench.py reports the evaluation kernel against (MEASURED_PEAKS.json has no CUDA-core figure). */
int32_t jssp_measure_fma_peak(int32_t device, int32_t bits, double* tflops);

/* -- secondary kernel: batched IMM lane-intention update (no reference code, README.md:19,23,98-110) ------- */
typedef struct jssp_imm_params {
  double dt;
  double q_pos, q_vel;         /* process noise of the per-mode lateral KF */
  double r_pos;                /* measurement noise of the lateral offset */
  double k_lat;                /* lane-keeping gain of a mode towards its own lane centre */
  double p_stay;               /* diagonal of the Markov transition matrix; off-diagonals share 1 - p_stay */
  double mu_floor;             /* lower clamp on the mode probabilities before normalisation */
} jssp_imm_params;
/* All pointers are device pointers.  Per obstacle o and mode m: mean[o][m][2], cov[o][m][3] (p00,p01,p11),
 * mu[o][m]; z[o][m] = measured lateral offset of the obstacle from mode m's lane centre.  Updates in place. */
int32_t jssp_imm_update(jssp_handle* h, const jssp_imm_params* p, double* mean_dev, double* cov_dev, double* mu_dev,
                        const double* z_dev, int32_t O, int32_t M, void* stream);
/* Roll out one constant-speed trajectory per (obstacle, mode) along the mode's lane polyline:
 * lanes_dev[L][P][2] polylines, lane_of[o][m] lane index, s_start[o][m] arc length at t0, speed[o], ->
 * state_dev[O][M][Tt][3] (x, y, yaw), valid_dev[O][M][Tt]. */
int32_t jssp_predict_modes(jssp_handle* h, const double* lanes_dev, int32_t L, int32_t P, const int32_t* lane_of_dev,
                           const double* s_start_dev, const double* speed_dev, double dt, int32_t O, int32_t M,
                           int32_t Tt, double* state_dev, uint8_t* valid_dev, void* stream);

/* -- FOP baseline planner through the same stages ------------------------------------
 * replaces the body of FrenetOptimalPlanner.plan_traj (code/intersection_local_planner/frenet_optimal_planner.py:301-333):
 * sampling_frenet_trajectories (:97-144: lateral quintic x horizon x longitudinal quartic to a target speed, samples
 * arange(0, T, dt)), calc_global_paths (:146-183), check_constraints (:185-225: |s_dd| <= max_accel, |s_ddd| <= max_jerk),
 * check_collisions (:249-299: check_res 2, rectangles NOT inflated - create the handle with obstacle_inflation = 0), cost and
 * the `min_cost >= cost` selection (:328-333: the LAST minimum wins).  Candidate n = (w * num_t + it) * num_speed + iv.
 * The handle supplies reference line, obstacles, delta_t, ego footprint, cost weights (configuration_parameters.py:48:
 * {1.0, 0.0, 99.4, 0.2, 0.1, 0.8, 0.4, 10.0}), thresholds, speed_max = target speed (highest_speed) and must have
 * plan_horizon >= max_t.  Outputs as jssp_evaluate (masks / costs [N], selected trajectory rows; rows beyond the
 * candidate's own length are NaN).  QuarticPolynomial and GoalSamplingCostFunction are absent from the reference tree:
 * builder-specified, see oracle/fop_oracle.py. */
typedef struct jssp_fop_in {
  double s0, v0, a0;           /* frenet_state.s, s_d, s_dd */
  double d0, d_d0, d_dd0;      /* frenet_state.d, d_d, d_dd */
  int32_t time_step_now, final_time_step;
  int32_t num_width, num_speed, num_t;   /* FrenetOptimalPlannerSettings (:56-83); paras_planner["planner_frenet"]: 3, 250, 10 */
  int32_t flags;               /* JSSP_FLAG_NO_SYNC / JSSP_FLAG_NO_EXPORT */
  double min_t, max_t;         /* 5.0, 5.5 */
  double lowest_speed, highest_speed;    /* 0.0, speed_max */
  double max_accel, max_jerk;  /* Vehicle.max_accel, max_jerk = lon_accel_max, lon_jerk_max (8.0, 10.0) */
} jssp_fop_in;
int32_t jssp_fop_evaluate(jssp_handle* h, const jssp_fop_in* in, void* stream, jssp_cycle_out* out);

/* -- batched Cartesian -> Frenet projection ------------------------------------------
 * replaces FrenetState().from_state(ego_state, ref_ego_lane_pts) (intersection_planner.py:140-142, 308-310) for n states
 * at once on the handle's reference line re-discretised every `step` metres (0.1 in jssp.py:366): state_dev [n][5] =
 * x, y, yaw, v, a -> frenet_dev [n][6] = s, s_d, s_dd, d, d_d, d_dd (both device pointers).  The class is absent from the
 * reference tree; dlii_jssp_b200/frenet.py documents the formulation. */
int32_t jssp_project_states(jssp_handle* h, const double* state_dev, int32_t n, double step, double* frenet_dev, void* stream);

/* -- batched, device-resident replay ------------------------------------------------
 * Replaces the per-scenario driver loop of code/__main__.py:124-167 for MANY independent scenarios at once: in every
 * step all running scenarios execute one plan cycle (seed enumeration -> evaluation -> selection -> export) and the
 * device itself performs the hand-off the reference does on the host - ego := sample 1 of best_traj
 * (code/onsite_trans/env.py:24-30), next Frenet start := best_traj.frenet_state_at_time_step(1)
 * (intersection_planner.py:312), t += dt (controller.py:303), end status (controller.py:338-380: collision with a
 * background vehicle for t > 0.5, t >= max_t, map bounds, goal line passed - evaluated, as in the reference, on the ego
 * advanced by the kinematic model with the (0, 0) action) - so that a whole replay of S
 * scenarios costs 2 kernel launches per step and one result copy at the end.  Scenarios share one jssp_config
 * (sampling grids, limits, cost weights; the ego footprint may differ per scenario); capacities per scenario come from cfg.max_seeds / max_knots / max_obstacle_modes / max_total_steps. */
typedef struct jssp_batch jssp_batch;

typedef struct jssp_cycle_record {  /* one plan cycle of one scenario = one row of the reference's recorder */
  int32_t best_idx;                 /* selected candidate, -1 = "No solution" */
  int32_t n_candidates;
  double best_cost;
  double t, x, y, v, a, yaw;        /* ego after the cycle (NaN when the cycle ended the replay without a move) */
} jssp_cycle_record;

typedef struct jssp_scenario_result {
  double end;                       /* -1 still running, 1 max_t, 1.5 goal, 2 collision, -2 no solution, -3 short trajectory */
  int32_t cycles;                   /* plan cycles executed */
  int32_t seed_overflow;            /* 1: some cycle enumerated more than cfg.max_seeds seeds (jssp_batch_fetch then returns
                                       JSSP_ERR_SEED_OVERFLOW; results are still filled in) */
} jssp_scenario_result;

typedef struct jssp_scenario_in {
  /* reference line, as jssp_set_refline */
  const double* knot_s; const double* coef; int32_t K; int32_t reserved0;
  /* obstacles the planner sees, as jssp_set_obstacles (host pointers) */
  const double* state; const uint8_t* valid; const double* size; const double* prob;
  int32_t O, M, Tt, n_dict;
  /* ground-truth background vehicles for the end-status collision test: state [gt_O][Tt][3], valid [gt_O][Tt],
   * size [gt_O][2]; NULL -> mode 0 of the planner's table (the replay case M = 1) */
  const double* gt_state; const uint8_t* gt_valid; const double* gt_size; int32_t gt_O; int32_t reserved1;
  /* initial Frenet state (s, s_d, s_dd, d, d_d, d_dd) of cycle 0 (FrenetState.from_state, intersection_planner.py:309-310) */
  double frenet0[6];
  double max_t;                     /* test_setting["max_t"] */
  int32_t final_time_step;          /* int(max_t / dt), jssp.py:244 */
  int32_t n_cycles;                 /* entries - 1 of the three per-cycle tables below */
  /* host-computed clock tables, index c = plan cycles done: cycle_time[c] = t after c float additions of dt,
   * cycle_now[c] = int(t / dt) (jssp.py:300), cycle_end_step[c] = table step whose key equals str(around(t, 3))
   * (controller.py:345) or -1 */
  const double* cycle_time; const int32_t* cycle_now; const int32_t* cycle_end_step;
  double goal_x, goal_y, goal_yaw;  /* goal pose: passed when sin(yaw)*dy + cos(yaw)*dx >= 0 */
  double ego_length, ego_width;     /* this scenario's ego footprint (observation["vehicle_info"]["ego"]); 0 -> jssp_config's */
  double ego0[4];                   /* initial ego x, y, yaw, v: the end status of a step is evaluated on the ego advanced by the
                                       kinematic model with the (0, 0) action (controller.py:256-263, 301-319), before the ego is
                                       overwritten with sample 1 of the selected trajectory (env.py:23-30) */
  double bounds[4];                 /* x_min, x_max, y_min, y_max of the map's lane vertices (controller.py:279-288) */
  int32_t has_bounds;               /* 0: no map, the bounds test is skipped */
  /* parameters["sim_config"]["ego_update_mode"] (configuration_parameters.py:4): 0 = "planner_expected_value" (default: the ego
   * jumps to sample 1 of the selected trajectory and the next cycle starts from its Frenet state); 1 = "kinematics"
   * (code/__main__.py:136-138): the ego follows the kinematic model (controller.py:301-319) driven by the planned acceleration
   * s_dd[1] and the pure-pursuit front-wheel angle on the selected path (intersection_planner.py:365-443), and every cycle
   * starts from the projection of that ego on the re-discretised reference line (FrenetState.from_state, :308-310) */
  int32_t ego_update_mode;
  /* kinematics mode only: the re-discretised reference line [n_polyline][6] = x, y, yaw, arc length of the polyline up to the
   * point (sum of the chord lengths before it), cos(yaw), sin(yaw)  (generate_frenet_frame, jssp.py:366-371) */
  const double* polyline; int32_t n_polyline; int32_t reserved2;
  double pp_kv, pp_ld0;             /* pure pursuit look-ahead ld = kv * v + ld0  (paras_control["pure_pursuit"]["kv_ld0"]) */
  double pp_wheel_base;             /* ego length / coff_wheel_base */
} jssp_scenario_in;

int32_t jssp_batch_create(const jssp_config* cfg, int32_t device, int32_t max_scenarios, int32_t max_cycles, jssp_batch** out);
int32_t jssp_batch_destroy(jssp_batch* b);
const char* jssp_batch_last_cuda_error(const jssp_batch* b);
/* switch the batch to the FOP baseline planner (jssp_fop_evaluate's candidate model) for every scenario loaded afterwards:
 * `grid` supplies num_width (= cfg.num_width), num_speed, num_t, min_t, max_t, lowest/highest speed, max_accel, max_jerk; its
 * start-state and time-step fields are ignored.  The jssp_config of the batch should be the FOP one (obstacle_inflation 0,
 * plan_horizon >= max_t, speed_max = highest_speed, planner_frenet cost weights).  Two launches per lock-step cycle. */
int32_t jssp_batch_set_fop(jssp_batch* b, const jssp_fop_in* grid);
/* upload one scenario into `slot` and reset its replay state.  The host arrays are copied into a pinned staging ring before
 * the call returns (they may be freed); the transfers themselves are asynchronous on `stream` */
int32_t jssp_batch_set_scenario(jssp_batch* b, int32_t slot, const jssp_scenario_in* in, void* stream);
/* put slots [0, n_scenarios) back to the state jssp_batch_set_scenario left them in (device-to-device; asynchronous):
 * a sweep can be repeated without uploading the scenarios again */
int32_t jssp_batch_reset(jssp_batch* b, int32_t n_scenarios, void* stream);
/* enqueue up to `n_cycles` plan cycles of every slot in [0, n_scenarios) (finished scenarios idle).  Default: the lock-step
 * pipeline - every step runs one plan cycle of all running scenarios with two launches (seed enumeration, evaluation +
 * selection + export + hand-off); JSSP_FLAG_FORCE_* choose its evaluation kernel.  JSSP_FLAG_PERSISTENT: ONE launch of the
 * persistent replay kernel, in which every scenario's CTA loops over its own plan cycles.  Same results.  Asynchronous. */
int32_t jssp_batch_run(jssp_batch* b, int32_t n_scenarios, int32_t n_cycles, int32_t flags, void* stream);
/* copy results back (synchronises): results_host[n_scenarios], records_host[n_scenarios][max_cycles] (may be NULL) */
int32_t jssp_batch_fetch(jssp_batch* b, int32_t n_scenarios, jssp_scenario_result* results_host,
                         jssp_cycle_record* records_host, void* stream);
/* copy the trajectory selected in the last cycle of `slot` ([n_steps+1][16] float64, columns as jssp_cycle_out.best_traj)
 * into `out`, a host or device buffer (unified addressing decides); synchronises */
int32_t jssp_batch_best_traj(jssp_batch* b, int32_t slot, double* out, void* stream);
int64_t jssp_batch_launch_count(const jssp_batch* b);

#ifdef __cplusplus
}
#endif
#endif /* JSSP_B200_H_ */
