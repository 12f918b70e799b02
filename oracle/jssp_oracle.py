"""CPU oracle (float64, NumPy, structure-of-arrays) for the JSSP per-cycle candidate evaluation.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu-baseline / ``--impl reference`` legs may import it.  The product path
(``dlii_jssp_b200``) never routes through this module and has no CPU fallback.

What it restates (reference = byChenZhifa/DLII-JSSP, paths relative to the reference root):

  stage                              reference file:line                                        parity
  ---------------------------------  ---------------------------------------------------------  ---------------------------
  grids / PolyVTSampling.__init__    code/intersection_local_planner/polyvt_sampling.py:33-58   PINNED (golden from the real file)
  main seeds      sampling()         polyvt_sampling.py:167-263 (+check_state :88-113)          PINNED (tests/golden/polyvt_golden.npz)
  stop seeds      ..._supplement()   polyvt_sampling.py:265-305                                 PINNED
  unroll  get_one_frenet_trajectory  polyvt_sampling.py:311-385                                 PINNED
  lateral grid + quintic             jerk_space_sampling_planner.py:99-122                      grid pinned by reading; quintic UNPINNED
  Frenet->Cartesian, yaw, ds, c      jerk_space_sampling_planner.py:131-169                     logic pinned by reading; spline UNPINNED
  curvature feasibility              jerk_space_sampling_planner.py:171-180                     logic pinned; max_curvature UNPINNED
  rectangle collision, check_res=2   jerk_space_sampling_planner.py:211-286                     logic pinned; shapely -> SAT, ego footprint UNPINNED
  cost + argmin                      jerk_space_sampling_planner.py:288-294,326-327             cost function and tie rule UNPINNED

Second pin (tests/golden/jssp_cycle_golden.npz, oracle/make_golden_cycle.py): the reference's own
JerkSpaceSamplingPlanner.plan_traj is executed on four plan cycles of inputs/8_3_4_565 with stubs for the absent LEAF
classes only, so the stage logic above the leaves (sampling order, truncation, forward-difference headings, un-unwrapped
curvature, collision rows / time keys / inflation / window, survivor filtering, heap selection) is pinned by executable
reference code: tests/test_oracle.py::test_oracle_matches_the_reference_plan_cycle.

"parity unpinned" pieces: the reference's ``code/common/**`` package (CubicSpline2D, QuinticPolynomial,
JerkSpaceSamplingCostFunction, Vehicle, the trajectory ordering) is absent from the published tree and
the upstream it derives from (SS47816/fiss_plus_planner, cited README.md:199, which itself follows the
PythonRobotics cubic-spline / quintic formulation) is neither vendored nor version-pinned; shapely is
an unpinned third-party dependency.  For those the oracle restates the published algorithm /
the builder specification of SURVEY.md section 8c, and says so next to each function.  The multi-modal
(M > 1, probability-weighted) collision semantics are a builder-specified extension with no reference
code; for M = 1, prob = 1 they reduce exactly to the reference's single-trajectory boolean check.

Floating-point contract: everything is IEEE float64 with the reference's operand order (NumPy element
wise ops never fuse), so seeds and the unrolled s, s_d, s_dd, s_ddd are bit-identical to the reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

__all__ = [
    "OracleConfig", "make_grids", "sample_main_seeds", "sample_stop_seeds", "sample_seeds", "unroll_seeds",
    "QuinticPolynomial", "lateral_targets", "lateral_profiles", "CubicSpline1D", "CubicSpline2D",
    "frenet_to_cartesian", "heading_curvature", "curvature_fail", "pack_obstacles", "collision_mask",
    "cost_total", "select_best", "plan_cycle", "obb_intersects",
]


# ----------------------------------------------------------------------------------------------
# configuration (defaults = reference values; citations in the comments)
# ----------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    # JerkSpaceSamplingPlannerSettings (jerk_space_sampling_planner.py:35-64) with the intended values of
    # configuration_parameters.py:52-61 (the published wiring never instantiates them)
    delta_t: float = 0.1
    plan_horizon: float = 5.0
    max_road_width: float = 3.5          # intersection_planner.py:60
    num_width: int = 3
    num_jerk: int = 25
    highest_jerk: float = 10.0           # = vehicle_para.lon_jerk_max (decision recorded in SURVEY section 0.4)
    # parameters["vehicle_para"] (configuration_parameters.py:7-14)
    speed_max: float = 18.0
    lon_accel_max: float = 8.0
    # ego footprint (xosc <Dimensions> of the first ScenarioObject; controller.py:170-175)
    ego_length: float = 5.4246098261480151
    ego_width: float = 1.9938293892797307
    # Vehicle.max_curvature: class absent -> builder-specified constant, 1/R_min, R_min = 5 m
    max_curvature: float = 0.2
    # collision check (jerk_space_sampling_planner.py:263-265, :280)
    obstacle_inflation: float = 0.1
    check_res: int = 2
    # cost (configuration_parameters.py:15-22 thresholds, :56-59 weights)
    cost_weights: tuple = (0.0, 3.0, 5.4, 0.2, 0.1, 0.8, 0.4, 4.0)
    max_distance: float = 140.0
    thr_lon_accel: float = 3.0
    thr_lon_jerk: float = 6.0
    thr_lat_accel: float = 0.5
    thr_lat_jerk: float = 1.0
    # multi-modal extension (builder-specified): modes with prob >= p_min veto, the others add w_risk * prob
    p_min: float = 0.0
    w_risk: float = 0.0

    @property
    def n_steps(self) -> int:  # PolyVTSampling.__init__ :43-44 -> len(time_array) - 1
        return int(self.plan_horizon / self.delta_t)


# ----------------------------------------------------------------------------------------------
# closed forms (polyvt_sampling.py:60-86); operand order is the reference's
# ----------------------------------------------------------------------------------------------
def _st(s0, v0, a0, j, t):
    return s0 + v0 * t + 0.5 * a0 * t * t + (1 / 6.0) * j * t * t * t


def _vt(v0, a0, j, t):
    return v0 + a0 * t + 0.5 * j * t * t


def _at(a0, j, t):
    return a0 + j * t


def make_grids(total_time: float, jerk: float, num_samples: int) -> dict:
    """Grids of PolyVTSampling.__init__ (polyvt_sampling.py:33-58) and of the two samplers (:181, :274-279)."""
    g = {}
    g["total_time"] = float(total_time)
    g["time_array"] = np.linspace(0.0, total_time, int(total_time / 0.1) + 1)
    g["j1"] = np.linspace(-jerk, jerk, num_samples)
    g["t1"] = np.arange(0.0, total_time + 1e-9, 0.5)
    g["t2"] = np.concatenate((np.array([5.0]), np.arange(0.0, total_time + 1e-9, 0.8)))
    g["t3"] = np.arange(0.0, total_time + 1e-9, 1.0)
    g["t4"] = np.arange(0.0, total_time + 1e-9, 1.5)
    g["j3"] = np.linspace(-jerk * 0.6, jerk * 0.6, int(num_samples / 4))
    g["jmin5"], g["jmax5"] = -jerk * 0.3, jerk * 0.3
    g["js"] = np.linspace(-jerk, jerk, num_samples * 2)       # stop-seed jerk grid (:274, :286)
    g["ts"] = np.arange(0.0, total_time + 1e-9, 0.1)           # stop-seed time grid (:278, :290)
    return g


def _check_state(st, vt, at, s_prev, a_max, v_max):
    """check_state (polyvt_sampling.py:88-113), vectorised; returns the PASS mask."""
    acc_max = a_max - 1e-9
    vel_max = v_max - 1e-9
    bad = (at > acc_max) | (at < -acc_max) | (vt > vel_max) | (vt < 0.0 - 1e-9) | (st < s_prev - 1e-9)
    return ~bad


def sample_main_seeds(s0, v0, a0, g, a_max=8.0, v_max=18.0) -> np.ndarray:
    """PolyVTSampling.sampling (polyvt_sampling.py:167-263) as one flat predicate over the
    (j1, t1, t2, j3, t3, t4) grid; survivors are emitted in the reference's loop (= C) order."""
    T = g["total_time"]
    _err = np.seterr(all="ignore")      # inf * 0 etc. inside rejected grid points, as in the reference's scalar code
    J1, T1, T2, J3, T3, T4 = np.meshgrid(g["j1"], g["t1"], g["t2"], g["j3"], g["t3"], g["t4"], indexing="ij")
    ok = ~(T1 < 0.55)                                                    # :183-184
    s1, v1, a1 = _st(s0, v0, a0, J1, T1), _vt(v0, a0, J1, T1), _at(a0, J1, T1)
    ok &= _check_state(s1, v1, a1, s0, a_max, v_max)                      # :190
    ok &= ~(T1 + T2 >= T - 0.8 * 3)                                       # :195
    s2, v2, a2 = _st(s1, v1, a1, 0.0, T2), _vt(v1, a1, 0.0, T2), _at(a1, 0.0, T2)
    ok &= _check_state(s2, v2, a2, s1, a_max, v_max)                      # :202
    ok &= ~(T1 + T2 > T + 1e-9)                                           # :205
    ok &= ~(J3 * J1 > 0)                                                  # :209
    ok &= ~(T1 + T2 + T3 > T - 1.0 * 2)                                   # :213
    s3, v3, a3 = _st(s2, v2, a2, J3, T3), _vt(v2, a2, J3, T3), _at(a2, J3, T3)
    ok &= _check_state(s3, v3, a3, s2, a_max, v_max)                      # :220
    ok &= ~(a3 * a1 > 0)                                                  # :223
    with np.errstate(divide="ignore", invalid="ignore"):
        t25 = (0.0 - a2) / J3
        v25 = _vt(v2, a2, J3, t25)
    ok &= ~((a3 * a2 < 0) & ((v25 > v_max) | (v25 < 0.0 + 1e-9)))         # :226-230
    ok &= ~(T1 + T2 + T3 + T4 > T - 1.5)                                  # :234
    s4, v4, a4 = _st(s3, v3, a3, 0.0, T4), _vt(v3, a3, 0.0, T4), _at(a3, 0.0, T4)
    ok &= _check_state(s4, v4, a4, s3, a_max, v_max)                      # :241
    T5 = T - T1 - T2 - T3 - T4
    with np.errstate(divide="ignore", invalid="ignore"):
        J5 = (0.0 - a4) / T5
    ok &= ~((J5 > g["jmax5"]) | (J5 < g["jmin5"]))                        # :246-247
    s5, v5, a5 = _st(s4, v4, a4, J5, T5), _vt(v4, a4, J5, T5), _at(a4, J5, T5)
    ok &= _check_state(s5, v5, a5, s4, a_max, v_max)                      # :252
    np.seterr(**_err)
    idx = np.nonzero(ok.ravel())[0]                                       # C order == loop order
    z = np.zeros(len(idx))
    cols = [J1.ravel()[idx], T1.ravel()[idx], z, T2.ravel()[idx], J3.ravel()[idx], T3.ravel()[idx],
            z, T4.ravel()[idx], J5.ravel()[idx], T5.ravel()[idx]]
    return np.stack(cols, 1).reshape(-1, 10)


def sample_stop_seeds(s0, v0, a0, g, a_max=8.0, v_max=18.0) -> np.ndarray:
    """PolyVTSampling.sampling_stop_seeds_supplement (polyvt_sampling.py:265-305), flat predicate."""
    T = g["total_time"]
    js = g["js"]
    J1, T1, J2, T2 = np.meshgrid(js[~(js > 0)], g["ts"], js[~(js < 0)], g["ts"], indexing="ij")  # :275, :287
    ok = np.ones(J1.shape, dtype=bool)
    s1, v1, a1 = _st(s0, v0, a0, J1, T1), _vt(v0, a0, J1, T1), _at(a0, J1, T1)
    ok &= _check_state(s1, v1, a1, s0, a_max, v_max)                      # :283
    ok &= ~(T1 + T2 >= T)                                                 # :291
    a2 = _at(a1, J2, T2)
    ok &= ~((a2 > 0.0) | (a2 < -0.3))                                     # :293-294
    v2 = _vt(v1, a1, J2, T2)
    ok &= ~((v2 > 0.0) | (v2 < -0.3))                                     # :296-297
    s2 = _st(s1, v1, a1, J2, T2)
    ok &= ~(s2 < s1 - 0.1)                                                # :299
    idx = np.nonzero(ok.ravel())[0]
    z = np.zeros(len(idx))
    T3 = T - T1 - T2                                                      # :302
    cols = [J1.ravel()[idx], T1.ravel()[idx], J2.ravel()[idx], T2.ravel()[idx], z, T3.ravel()[idx], z, z, z, z]
    return np.stack(cols, 1).reshape(-1, 10)


def sample_seeds(s0, v0, a0, cfg: OracleConfig) -> np.ndarray:
    """seeds_temp + seeds_stop (jerk_space_sampling_planner.py:95-97)."""
    g = make_grids(cfg.plan_horizon, cfg.highest_jerk, cfg.num_jerk)
    return np.concatenate([sample_main_seeds(s0, v0, a0, g, cfg.lon_accel_max, cfg.speed_max),
                           sample_stop_seeds(s0, v0, a0, g, cfg.lon_accel_max, cfg.speed_max)], 0)


def unroll_seeds(seeds: np.ndarray, s0, v0, a0, time_array: np.ndarray):
    """get_one_frenet_trajectory (polyvt_sampling.py:311-385) for all seeds -> s, s_d, s_dd, s_ddd [Ns, T+1]."""
    seeds = np.asarray(seeds, dtype=np.float64).reshape(-1, 10)
    j1, t1, j2, t2, j3, t3, j4, t4, j5, t5 = [seeds[:, k:k + 1] for k in range(10)]
    s1, v1, a1 = _st(s0, v0, a0, j1, t1), _vt(v0, a0, j1, t1), _at(a0, j1, t1)
    s2, v2, a2 = _st(s1, v1, a1, j2, t2), _vt(v1, a1, j2, t2), _at(a1, j2, t2)
    s3, v3, a3 = _st(s2, v2, a2, j3, t3), _vt(v2, a2, j3, t3), _at(a2, j3, t3)
    s4, v4, a4 = _st(s3, v3, a3, j4, t4), _vt(v3, a3, j4, t4), _at(a3, j4, t4)
    t = np.asarray(time_array, dtype=np.float64)[None, :]
    # segment choice with the reference's asymmetric epsilons (:334, :343, :352, :363)
    in1 = t < t1 - 1e-9
    in2 = ~in1 & (t < t1 + t2 + 1e-9)
    in3 = ~in1 & ~in2 & (t < t1 + t2 + t3 + 1e-9)
    in4 = ~in1 & ~in2 & ~in3 & (t < t1 + t2 + t3 + t4 + 1e-9)
    segs = [(in1, s0, v0, a0, j1, t), (in2, s1, v1, a1, j2, t - t1), (in3, s2, v2, a2, j3, t - t1 - t2),
            (in4, s3, v3, a3, j4, t - t1 - t2 - t3), (None, s4, v4, a4, j5, t - t1 - t2 - t3 - t4)]
    shape = np.broadcast(t, t1).shape
    s = np.empty(shape); sd = np.empty(shape); sdd = np.empty(shape); sddd = np.empty(shape)
    done = np.zeros(shape, dtype=bool)
    for m, sb, vb, ab, jb, dt in segs:
        m = ~done if m is None else (m & ~done)
        S, V, A = _st(sb, vb, ab, jb, dt), _vt(vb, ab, jb, dt), _at(ab, jb, dt)
        Jb = np.broadcast_to(jb, shape)
        s[m], sd[m], sdd[m], sddd[m] = np.broadcast_to(S, shape)[m], np.broadcast_to(V, shape)[m], np.broadcast_to(A, shape)[m], Jb[m]
        done |= m
    return s, sd, sdd, sddd


# ----------------------------------------------------------------------------------------------
# lateral quintic (class absent from the tree -> PythonRobotics formulation; parity unpinned)
# ----------------------------------------------------------------------------------------------
class QuinticPolynomial:
    """QuinticPolynomial(xs, vxs, axs, xe, vxe, axe, time) as called at jerk_space_sampling_planner.py:109-117.
    Published algorithm (PythonRobotics / fiss_plus_planner): a0..a2 from the start state, a3..a5 from the
    3x3 linear system at ``time``."""

    def __init__(self, xs, vxs, axs, xe, vxe, axe, time):
        self.a0, self.a1, self.a2 = xs, vxs, axs / 2.0
        A = np.array([[time ** 3, time ** 4, time ** 5],
                      [3 * time ** 2, 4 * time ** 3, 5 * time ** 4],
                      [6 * time, 12 * time ** 2, 20 * time ** 3]])
        b = np.array([xe - self.a0 - self.a1 * time - self.a2 * time ** 2, vxe - self.a1 - 2 * self.a2 * time, axe - 2 * self.a2])
        self.a3, self.a4, self.a5 = np.linalg.solve(A, b)

    def calc_point(self, t):
        return self.a0 + self.a1 * t + self.a2 * t ** 2 + self.a3 * t ** 3 + self.a4 * t ** 4 + self.a5 * t ** 5

    def calc_first_derivative(self, t):
        return self.a1 + 2 * self.a2 * t + 3 * self.a3 * t ** 2 + 4 * self.a4 * t ** 3 + 5 * self.a5 * t ** 4

    def calc_second_derivative(self, t):
        return 2 * self.a2 + 6 * self.a3 * t + 12 * self.a4 * t ** 2 + 20 * self.a5 * t ** 3

    def calc_third_derivative(self, t):
        return 6 * self.a3 + 24 * self.a4 * t + 60 * self.a5 * t ** 2


def lateral_targets(cfg: OracleConfig) -> np.ndarray:
    """di_list (jerk_space_sampling_planner.py:100-104)."""
    if cfg.num_width <= 1:
        return np.array([0.0])
    w = cfg.max_road_width - cfg.ego_width
    return np.linspace(-w / 2, w / 2, cfg.num_width)


def lateral_profiles(d0, dd0, ddd0, targets, horizon, n_steps):
    """d, d_d, d_dd, d_ddd [W, T+1] (jerk_space_sampling_planner.py:106-122)."""
    t = np.linspace(0.0, horizon, n_steps + 1)
    out = np.zeros((len(targets), 4, n_steps + 1))
    for w, di in enumerate(targets):
        q = QuinticPolynomial(d0, dd0, ddd0, di, 0.0, 0.0, horizon)
        out[w, 0], out[w, 1], out[w, 2], out[w, 3] = (q.calc_point(t), q.calc_first_derivative(t),
                                                      q.calc_second_derivative(t), q.calc_third_derivative(t))
    return out[:, 0], out[:, 1], out[:, 2], out[:, 3]


# ----------------------------------------------------------------------------------------------
# reference line: natural cubic spline over cumulative chord length (class absent -> PythonRobotics
# formulation as used by fiss_plus_planner; parity unpinned)
# ----------------------------------------------------------------------------------------------
class CubicSpline1D:
    def __init__(self, x, y):
        x = np.asarray(x, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
        h = np.diff(x)
        if np.any(h <= 0):
            raise ValueError("knots must be strictly increasing")
        n = len(x)
        A = np.zeros((n, n)); B = np.zeros(n)
        A[0, 0] = 1.0; A[n - 1, n - 1] = 1.0          # natural boundary: c0 = c_{n-1} = 0
        for i in range(1, n - 1):
            A[i, i - 1] = h[i - 1]; A[i, i] = 2.0 * (h[i - 1] + h[i]); A[i, i + 1] = h[i]
            B[i] = 3.0 * (y[i + 1] - y[i]) / h[i] - 3.0 * (y[i] - y[i - 1]) / h[i - 1]
        c = np.linalg.solve(A, B)
        self.x, self.a, self.c = x, y.copy(), c
        self.d = (c[1:] - c[:-1]) / (3.0 * h)
        self.b = (y[1:] - y[:-1]) / h - h / 3.0 * (2.0 * c[:-1] + c[1:])

    def _seg(self, x):
        """bisect_right - 1, clamped so that x == x[-1] evaluates on the last segment (builder-specified:
        the PythonRobotics original indexes one past the end there)."""
        i = np.searchsorted(self.x, x, side="right") - 1
        return np.clip(i, 0, len(self.x) - 2)

    def position(self, x):
        i = self._seg(x); dx = x - self.x[i]
        return ((self.a[i] + self.b[i] * dx) + self.c[i] * (dx * dx)) + self.d[i] * ((dx * dx) * dx)

    def first(self, x):
        i = self._seg(x); dx = x - self.x[i]
        return (self.b[i] + (2.0 * self.c[i]) * dx) + (3.0 * self.d[i]) * (dx * dx)

    def second(self, x):
        i = self._seg(x); dx = x - self.x[i]
        return 2.0 * self.c[i] + (6.0 * self.d[i]) * dx


class CubicSpline2D:
    """CubicSpline2D(x, y) with ``s_list``, ``calc_position``, ``calc_yaw``, ``calc_curvature`` as used at
    jerk_space_sampling_planner.py:137-141,365-369.  Scalar calls return (None, None) outside [0, s_list[-1]]."""

    def __init__(self, x, y):
        x = np.asarray(x, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
        self.s_list = np.concatenate(([0.0], np.cumsum(np.hypot(np.diff(x), np.diff(y)))))
        self.sx, self.sy = CubicSpline1D(self.s_list, x), CubicSpline1D(self.s_list, y)

    def in_range(self, s):
        return ~((s < self.s_list[0]) | (s > self.s_list[-1]))

    def calc_position(self, s):
        if np.ndim(s) == 0:
            if not self.in_range(s):
                return None, None
            return float(self.sx.position(s)), float(self.sy.position(s))
        return self.sx.position(s), self.sy.position(s)

    def calc_yaw(self, s):
        return np.arctan2(self.sy.first(s), self.sx.first(s))

    def calc_curvature(self, s):
        dx, ddx, dy, ddy = self.sx.first(s), self.sx.second(s), self.sy.first(s), self.sy.second(s)
        return (ddy * dx - ddx * dy) / ((dx ** 2 + dy ** 2) ** (3 / 2))

    def coefficient_table(self):
        """[K-1, 8] = (ax, bx, cx, dx, ay, by, cy, dy) per segment - the layout jssp_set_refline takes."""
        return np.stack([self.sx.a[:-1], self.sx.b, self.sx.c[:-1], self.sx.d, self.sy.a[:-1], self.sy.b, self.sy.c[:-1], self.sy.d], 1)


# ----------------------------------------------------------------------------------------------
# Frenet -> Cartesian, heading, curvature (jerk_space_sampling_planner.py:131-169)
# ----------------------------------------------------------------------------------------------
def frenet_to_cartesian(spline: CubicSpline2D, s: np.ndarray, d: np.ndarray):
    """x, y [N, T+1] and n_pts [N] = number of leading samples whose s lies on the spline (:134-146:
    the loop ``break``s at the first s for which calc_position returns None)."""
    inr = spline.in_range(s)
    n_pts = np.where(inr.all(1), s.shape[1], np.argmin(inr, 1))
    sc = np.clip(s, spline.s_list[0], spline.s_list[-1])
    ix, iy = spline.calc_position(sc)
    yaw = spline.calc_yaw(sc)
    x = ix + d * np.cos(yaw + math.pi / 2.0)
    y = iy + d * np.sin(yaw + math.pi / 2.0)
    live = np.arange(s.shape[1])[None, :] < n_pts[:, None]
    return np.where(live, x, np.nan), np.where(live, y, np.nan), n_pts


def heading_curvature(x, y, n_pts, dt):
    """yaw [N,T+1], ds [N,T], c [N,T] with NaN padding beyond each candidate's own length (:148-165).
    yaw = forward difference with the last value repeated, c = diff(yaw)/ds with NO angle unwrapping and
    IEEE inf/nan when ds == 0 - all as in the reference.  Only defined when n_pts >= 2."""
    N, P = x.shape
    idx = np.arange(P)[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        xd, yd = np.diff(x, axis=1), np.diff(y, axis=1)
        seg_live = idx[:, :-1] < (n_pts[:, None] - 1)
        yaw_f = np.where(seg_live, np.arctan2(yd, xd), np.nan)
        ds = np.where(seg_live, np.hypot(xd, yd), np.nan)
        yaw = np.full((N, P), np.nan)
        yaw[:, :-1] = yaw_f
        has = n_pts >= 2
        rows = np.nonzero(has)[0]
        yaw[rows, n_pts[rows] - 1] = yaw[rows, n_pts[rows] - 2]
        c = np.where(seg_live, np.diff(yaw, axis=1) / ds, np.nan)
        c_d = np.diff(c, axis=1) / dt
        c_dd = np.diff(c_d, axis=1) / dt
    return yaw, ds, c, c_d, c_dd


def curvature_fail(c, max_curvature):
    """any(abs(c) > max_curvature) (:175); NaN compares False, inf True, empty passes."""
    with np.errstate(invalid="ignore"):
        return np.any(np.abs(c) > max_curvature, axis=1)


# ----------------------------------------------------------------------------------------------
# obstacles: OnSite replay dict -> dense tensors (format: controller.py:92-107,186-213)
# ----------------------------------------------------------------------------------------------
def pack_obstacles(obstacles: dict, dt: float, n_total: int):
    """{id: {"shape": {length,width}, "<t>": {x,y,v,a,yaw}}} -> state [O,1,Tt,3] (x,y,yaw), valid [O,1,Tt],
    size [O,2] (length,width), prob [O,1].  The time key is the reference's ``str(round(step*dt, 2))``
    (jerk_space_sampling_planner.py:257-260)."""
    ids = list(obstacles.keys())
    O = len(ids)
    state = np.zeros((O, 1, n_total, 3)); valid = np.zeros((O, 1, n_total), dtype=bool); size = np.zeros((O, 2))
    for o, k in enumerate(ids):
        ob = obstacles[k]
        size[o] = (ob["shape"]["length"], ob["shape"]["width"])
        for step in range(n_total):
            key = str(round(step * dt, 2))
            if key in ob:
                st = ob[key]
                state[o, 0, step] = (st["x"], st["y"], st["yaw"])
                valid[o, 0, step] = True
    return state, valid, size, np.ones((O, 1))


def obb_intersects(c1x, c1y, cos1, sin1, a1, b1, c2x, c2y, cos2, sin2, a2, b2):
    """Closed-set overlap of two oriented rectangles (half extents a = length/2 along the heading, b = width/2)
    by the separating-axis theorem.  Stands in for shapely ``Polygon.intersects`` (jerk_space_sampling_planner.py:268):
    touching counts as intersecting, so an axis separates only on strict '>'."""
    dx, dy = c2x - c1x, c2y - c1y
    cd = np.abs(cos1 * cos2 + sin1 * sin2)
    sd = np.abs(cos1 * sin2 - sin1 * cos2)
    sep = np.abs(dx * cos1 + dy * sin1) > a1 + a2 * cd + b2 * sd
    sep |= np.abs(-dx * sin1 + dy * cos1) > b1 + a2 * sd + b2 * cd
    sep |= np.abs(dx * cos2 + dy * sin2) > a2 + a1 * cd + b1 * sd
    sep |= np.abs(-dx * sin2 + dy * cos2) > b2 + a1 * sd + b1 * cd
    return ~sep


def collision_mask(x, y, yaw, n_pts, state, valid, size, prob, time_step_now, final_time_step, cfg: OracleConfig,
                   n_obstacles_in_dict=None):
    """has_collision / check_collisions (jerk_space_sampling_planner.py:233-286), vectorised over candidates.

    Returns (collision [N] bool, risk [N] float).  Step i is tested iff i % check_res == 0 and
    i < min(n_pts, final_time_step - time_step_now) (:245-247); an obstacle mode takes part at step i iff it has a
    state at absolute step i + time_step_now (:256-260).  Obstacle rectangles are inflated by +0.1 m in length and
    width (:263-265); the ego rectangle is ego_length x ego_width centred on (x_i, y_i) with heading yaw_i (builder-
    specified footprint).  A 1-point trajectory raises inside the reference's ``try`` and is reported as
    a collision (:250-254); a 0-point trajectory passes vacuously.
    Multi-modal extension: a hit on mode (o, m) vetoes iff prob[o, m] >= p_min, else adds prob[o, m] to ``risk``."""
    N, P = x.shape
    O, M, Tt, _ = state.shape
    n_dict = O if n_obstacles_in_dict is None else n_obstacles_in_dict
    coll = np.zeros(N, dtype=bool); risk = np.zeros(N)
    if n_dict <= 0:                                                      # :241-242
        return coll, risk
    t_max = np.minimum(n_pts, final_time_step - time_step_now)
    coll |= (n_pts == 1) & (t_max >= 1)                                  # bare except -> "collision" (:250-254)
    ea, eb = cfg.ego_length / 2.0, cfg.ego_width / 2.0
    oa = (size[:, 0] + cfg.obstacle_inflation) / 2.0
    ob = (size[:, 1] + cfg.obstacle_inflation) / 2.0
    for i in range(0, P, cfg.check_res):
        step = i + time_step_now
        if step >= Tt or step < 0:
            continue
        live = (i < t_max) & (n_pts >= 2)
        if not live.any():
            continue
        ex, ey, eyaw = x[live, i][:, None, None], y[live, i][:, None, None], yaw[live, i][:, None, None]
        ox, oy, oyaw = state[None, :, :, step, 0], state[None, :, :, step, 1], state[None, :, :, step, 2]
        hit = obb_intersects(ex, ey, np.cos(eyaw), np.sin(eyaw), ea, eb,
                             ox, oy, np.cos(oyaw), np.sin(oyaw), oa[None, :, None], ob[None, :, None])
        hit &= valid[None, :, :, step]
        hard = hit & (prob[None] >= cfg.p_min)
        coll[live] |= hard.any(axis=(1, 2))
        risk[live] += (np.where(hit & ~hard, prob[None], 0.0)).sum(axis=(1, 2))
    return coll, risk


# ----------------------------------------------------------------------------------------------
# cost + selection (cost class absent -> builder specification of SURVEY section 8c; parity unpinned)
# ----------------------------------------------------------------------------------------------
def _seq_mean(a):
    """mean over the time axis with a strictly sequential float64 sum (the order the CUDA kernel uses)."""
    acc = np.zeros(a.shape[0])
    for k in range(a.shape[1]):
        acc = acc + a[:, k]
    return acc / a.shape[1]


def cost_total(s, s_d, s_dd, s_ddd, d, d_d, d_dd, d_ddd, cfg: OracleConfig, risk=None):
    """JerkSpaceSamplingCostFunction("WX1").cost_total(traj, target_speed=speed_max) as specified in SURVEY section 8c:
    weights = configuration_parameters.py:56, thresholds = :15-22, target speed = :8 (jerk_space_sampling_planner.py:291)."""
    w_time, w_dist, w_off, w_lat_a, w_lat_j, w_lon_a, w_lon_j, w_v = cfg.cost_weights
    relu = lambda z: np.maximum(z, 0.0)
    half = cfg.max_road_width / 2.0
    cost = w_dist * (1.0 - (s[:, -1] - s[:, 0]) / cfg.max_distance)
    cost = cost + w_off * (_seq_mean(d * d) / (half * half))
    cost = cost + w_lat_a * _seq_mean(relu(np.abs(d_dd) - cfg.thr_lat_accel) ** 2)
    cost = cost + w_lat_j * _seq_mean(relu(np.abs(d_ddd) - cfg.thr_lat_jerk) ** 2)
    cost = cost + w_lon_a * _seq_mean(relu(np.abs(s_dd) - cfg.thr_lon_accel) ** 2)
    cost = cost + w_lon_j * _seq_mean(relu(np.abs(s_ddd) - cfg.thr_lon_jerk) ** 2)
    cost = cost + w_v * _seq_mean(((cfg.speed_max - s_d) / cfg.speed_max) ** 2)
    if risk is not None:
        cost = cost + cfg.w_risk * risk
    return cost


def select_best(cost, feasible, collision):
    """PriorityQueue argmin over the survivors (jerk_space_sampling_planner.py:288-294,326-327); ties -> lowest
    candidate index (builder-specified: the trajectory class that would break ties is absent).  -1 = no solution."""
    ok = feasible & ~collision & np.isfinite(cost)
    if not ok.any():
        return -1
    c = np.where(ok, cost, np.inf)
    return int(np.argmin(c))  # np.argmin returns the first minimum


# ----------------------------------------------------------------------------------------------
# one plan cycle (JerkSpaceSamplingPlanner.plan_traj, jerk_space_sampling_planner.py:296-327)
# ----------------------------------------------------------------------------------------------
def plan_cycle(cfg: OracleConfig, spline: CubicSpline2D, frenet0, obstacles, time_step_now: int, final_time_step: int,
               seeds=None, targets=None, n_obstacles_in_dict=None) -> dict:
    """frenet0 = (s, s_d, s_dd, d, d_d, d_dd); obstacles = (state[O,M,Tt,3], valid, size[O,2], prob[O,M]).
    Candidate order is width-major: n = w * Ns + k (jerk_space_sampling_planner.py:106,124-125)."""
    s0, v0, a0, d0, dd0, ddd0 = [float(v) for v in frenet0]
    if seeds is None:
        seeds = sample_seeds(s0, v0, a0, cfg)
    seeds = np.asarray(seeds, dtype=np.float64).reshape(-1, 10)
    if targets is None:
        targets = lateral_targets(cfg)
    P = cfg.n_steps + 1
    time_array = np.linspace(0.0, cfg.plan_horizon, P)
    Ns, W = len(seeds), len(targets)
    sl, sdl, sddl, sdddl = unroll_seeds(seeds, s0, v0, a0, time_array)
    dl, ddl, dddl, ddddl = lateral_profiles(d0, dd0, ddd0, targets, cfg.plan_horizon, cfg.n_steps)
    tile = lambda a: np.tile(a, (W, 1))
    rep = lambda a: np.repeat(a, Ns, axis=0)
    s, s_d, s_dd, s_ddd = tile(sl), tile(sdl), tile(sddl), tile(sdddl)
    d, d_d, d_dd, d_ddd = rep(dl), rep(ddl), rep(dddl), rep(ddddl)
    N = Ns * W
    if N == 0:
        z = np.zeros((0, P))
        return dict(seeds=seeds, targets=targets, n=0, best=-1, feasible=np.zeros(0, bool), collision=np.zeros(0, bool),
                    cost=np.zeros(0), s=z, s_d=z, s_dd=z, s_ddd=z, d=z, d_d=z, d_dd=z, d_ddd=z, x=z, y=z, yaw=z,
                    n_pts=np.zeros(0, int), time=time_array)
    x, y, n_pts = frenet_to_cartesian(spline, s, d)
    yaw, ds, c, c_d, c_dd = heading_curvature(x, y, n_pts, cfg.delta_t)
    feasible = ~curvature_fail(c, cfg.max_curvature)
    state, valid, size, prob = obstacles
    collision, risk = collision_mask(x, y, yaw, n_pts, state, valid, size, prob, time_step_now, final_time_step, cfg,
                                     n_obstacles_in_dict)
    cost = cost_total(s, s_d, s_dd, s_ddd, d, d_d, d_dd, d_ddd, cfg, risk)
    best = select_best(cost, feasible, collision)
    return dict(seeds=seeds, targets=targets, n=N, best=best, feasible=feasible, collision=collision, cost=cost, risk=risk,
                s=s, s_d=s_d, s_dd=s_dd, s_ddd=s_ddd, d=d, d_d=d_d, d_dd=d_dd, d_ddd=d_ddd, x=x, y=y, yaw=yaw,
                ds=ds, c=c, c_d=c_d, c_dd=c_dd, n_pts=n_pts, time=time_array)
