"""CPU oracle (float64, NumPy) of the FOP baseline planner's plan cycle - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates code/intersection_local_planner/frenet_optimal_planner.py of byChenZhifa/DLII-JSSP:

  stage                                   reference file:line                      parity
  --------------------------------------  ---------------------------------------  ---------------------------------------------
  sampling grid  (di, Ti, tv) loops       frenet_optimal_planner.py:97-144         pinned by reading (loop order, linspace/arange)
  lateral quintic / longitudinal quartic  :113, :125 (common.geometry.polynomial)  UNPINNED: class absent -> PythonRobotics form
  Frenet->Cartesian, yaw, ds, c           :146-183                                 same code as the JSSP planner (jssp_oracle)
  constraints: |s_dd|, |s_ddd| limits     :185-225 (curvature check commented out) pinned by reading
  collision: L x W rectangles, check_res 2  :249-299 (no +0.1 inflation, :280)     logic pinned; shapely -> SAT
  cost  GoalSamplingCostFunction("WX1")   :128-133 (common.cost.cost_function)     UNPINNED: class absent -> builder specification
  best = LAST minimum (min_cost >= cost)  :328-333                                 pinned by reading

Pin (tests/golden/fop_cycle_golden.npz, oracle/make_golden_fop.py): the reference's own FrenetOptimalPlanner methods
(sampling_frenet_trajectories, calc_global_paths, check_constraints, check_collisions and the `>=` loop) are executed on
plan cycles of inputs/8_3_4_565 with stubs for the absent leaf classes only; this oracle and the CUDA path reproduce the
masks, costs, selection and trajectory (tests/test_oracle.py, tests/test_gpu_fop.py).

Builder-specified cost (the class is absent; weights and layout are configuration_parameters.py:42-51, the same 8-entry
vector the JSSP cost uses): the JSSP terms of jssp_oracle.cost_total over the candidate's own samples, with
target speed = highest_speed, plus w_time * (1 - (Ti - t_min) / (t_max - t_min)) (0 when t_max == t_min).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import jssp_oracle as O


@dataclass
class FopConfig:
    # FrenetOptimalPlannerSettings as wired at intersection_planner.py:101-109 with configuration_parameters.py:42-51
    num_width: int = 3
    num_speed: int = 250
    num_t: int = 10
    delta_t: float = 0.1
    max_road_width: float = 3.5
    highest_speed: float = 18.0          # = PlannerSettings.speed_max
    lowest_speed: float = 0.0
    min_t: float = 5.0
    max_t: float = 5.5
    max_accel: float = 8.0               # Vehicle.max_accel = lon_accel_max (intersection_planner.py:117-126)
    max_jerk: float = 10.0               # Vehicle.max_jerk = lon_jerk_max
    ego_length: float = 4.924
    ego_width: float = 1.872
    check_res: int = 2
    cost_weights: tuple = (1.0, 0.0, 99.4, 0.2, 0.1, 0.8, 0.4, 10.0)
    max_distance: float = 140.0
    thr_lon_accel: float = 3.0
    thr_lon_jerk: float = 6.0
    thr_lat_accel: float = 0.5
    thr_lat_jerk: float = 1.0

    def jssp_view(self) -> O.OracleConfig:
        """The fields the shared stages (collision, cost terms) read, as an OracleConfig: no obstacle inflation (:280),
        target speed = highest_speed."""
        return O.OracleConfig(delta_t=self.delta_t, max_road_width=self.max_road_width, num_width=self.num_width, speed_max=self.highest_speed,
                              ego_length=self.ego_length, ego_width=self.ego_width, obstacle_inflation=0.0, check_res=self.check_res,
                              cost_weights=self.cost_weights, max_distance=self.max_distance, thr_lon_accel=self.thr_lon_accel,
                              thr_lon_jerk=self.thr_lon_jerk, thr_lat_accel=self.thr_lat_accel, thr_lat_jerk=self.thr_lat_jerk)


class QuarticPolynomial:
    """QuarticPolynomial(xs, vxs, axs, vxe, axe, time) (frenet_optimal_planner.py:125): a0..a2 from the start state,
    a3, a4 from the 2x2 system at ``time`` (PythonRobotics formulation), solved in closed form."""

    def __init__(self, xs, vxs, axs, vxe, axe, time):
        T = float(time)
        self.a0, self.a1, self.a2 = xs, vxs, axs / 2.0
        b0 = vxe - self.a1 - 2.0 * self.a2 * T
        b1 = axe - 2.0 * self.a2
        self.a3 = b0 / (T * T) - b1 / (3.0 * T)
        self.a4 = b1 / (4.0 * T * T) - b0 / (2.0 * T * T * T)

    def calc_point(self, t):
        return self.a0 + self.a1 * t + self.a2 * t ** 2 + self.a3 * t ** 3 + self.a4 * t ** 4

    def calc_first_derivative(self, t):
        return self.a1 + 2 * self.a2 * t + 3 * self.a3 * t ** 2 + 4 * self.a4 * t ** 3

    def calc_second_derivative(self, t):
        return 2 * self.a2 + 6 * self.a3 * t + 12 * self.a4 * t ** 2

    def calc_third_derivative(self, t):
        return 6 * self.a3 + 24 * self.a4 * t


def grids(cfg: FopConfig):
    sw = cfg.max_road_width - cfg.ego_width
    widths = np.linspace(-sw / 2, sw / 2, cfg.num_width)                               # :101
    horizons = np.linspace(cfg.min_t, cfg.max_t, cfg.num_t)                            # :106
    speeds = np.linspace(cfg.lowest_speed, cfg.highest_speed, cfg.num_speed)           # :122
    n_samples = np.array([len(np.arange(0.0, T, cfg.delta_t)) for T in horizons])      # :114
    return widths, horizons, speeds, n_samples


def plan_cycle(cfg: FopConfig, spline: O.CubicSpline2D, frenet0, obstacles, time_step_now: int, final_time_step: int,
               n_obstacles_in_dict=None) -> dict:
    """One FrenetOptimalPlanner.plan_traj (:301-360).  Candidate n = (w * num_t + it) * num_speed + iv (loop order :101-122).
    Per-candidate arrays are padded with NaN to the longest horizon."""
    s0, v0, a0, d0, dd0, ddd0 = [float(v) for v in frenet0]
    widths, horizons, speeds, n_samples = grids(cfg)
    W, NT, NV = len(widths), len(horizons), len(speeds)
    N, Pmax = W * NT * NV, int(n_samples.max())
    jc = cfg.jssp_view()
    keys = ("s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd", "x", "y", "yaw")
    out = {k: np.full((N, Pmax), np.nan) for k in keys}
    feasible = np.zeros(N, dtype=bool); collision = np.zeros(N, dtype=bool); cost = np.zeros(N); n_pts = np.zeros(N, dtype=int)
    state, valid, size, prob = obstacles
    w_time = cfg.cost_weights[0]
    for it, (T, P) in enumerate(zip(horizons, n_samples)):
        t = np.arange(0.0, T, cfg.delta_t)                                                               # :114
        idx = np.array([(w * NT + it) * NV + iv for w in range(W) for iv in range(NV)])
        lat = np.zeros((W, 4, P))
        for w, di in enumerate(widths):
            q = O.QuinticPolynomial(d0, dd0, ddd0, di, 0.0, 0.0, T)                                      # :113
            lat[w] = (q.calc_point(t), q.calc_first_derivative(t), q.calc_second_derivative(t), q.calc_third_derivative(t))
        lon = np.zeros((NV, 4, P))
        for iv, tv in enumerate(speeds):
            q = QuarticPolynomial(s0, v0, a0, tv, 0.0, T)                                                # :125
            lon[iv] = (q.calc_point(t), q.calc_first_derivative(t), q.calc_second_derivative(t), q.calc_third_derivative(t))
        s, s_d, s_dd, s_ddd = [np.tile(lon[:, k], (W, 1)) for k in range(4)]
        d, d_d, d_dd, d_ddd = [np.repeat(lat[:, k], NV, axis=0) for k in range(4)]
        x, y, npt = O.frenet_to_cartesian(spline, s, d)                                                  # :146-160
        yaw, ds, c, c_d, c_dd = O.heading_curvature(x, y, npt, cfg.delta_t)                              # :162-181
        feas = ~(np.any(np.abs(s_dd) > cfg.max_accel, axis=1) | np.any(np.abs(s_ddd) > cfg.max_jerk, axis=1))   # :211-216
        coll, _ = O.collision_mask(x, y, yaw, npt, state, valid, size, prob, time_step_now, final_time_step, jc, n_obstacles_in_dict)
        cst = O.cost_total(s, s_d, s_dd, s_ddd, d, d_d, d_dd, d_ddd, jc)
        if cfg.max_t > cfg.min_t:
            cst = cst + w_time * (1.0 - (T - cfg.min_t) / (cfg.max_t - cfg.min_t))
        for k, a in zip(keys, (s, s_d, s_dd, s_ddd, d, d_d, d_dd, d_ddd, x, y, yaw)):
            out[k][idx, :P] = a
        feasible[idx], collision[idx], cost[idx], n_pts[idx] = feas, coll, cst, npt
    ok = feasible & ~collision & np.isfinite(cost)
    best = -1
    if ok.any():                                                                                         # :328-333: `>=` keeps the LAST minimum
        c = np.where(ok, cost, np.inf)
        best = int(N - 1 - np.argmin(c[::-1]))
    return dict(n=N, best=best, feasible=feasible, collision=collision, cost=cost, n_pts=n_pts, n_samples=n_samples, horizons=horizons,
                widths=widths, speeds=speeds, **out)
