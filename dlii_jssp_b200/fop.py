"""Drop-in mirror of the reference's FOP baseline planner (code/intersection_local_planner/frenet_optimal_planner.py:56-360),
backed by the same CUDA library: ``FrenetOptimalPlannerSettings`` / ``FrenetOptimalPlanner`` with the reference's
constructor arguments, ``plan_traj`` contract, ``cubic_spline`` attribute and ``best_traj`` - the baseline of the paper's
ablation, evaluated through the kernels of the JSSP path."""
from __future__ import annotations

import time
from typing import Optional

import numpy as np

from .engine import JsspConfig, JsspEngine
from .frenet import FrenetState, JerkSpaceSamplingFrenetTrajectory, State, Vehicle
from .geometry import CubicSpline2D, spline_tables
from .scenario import ObstacleTensors, pack_obstacle_dict

# paras_planner["planner_frenet"] (configuration_parameters.py:42-51)
PLANNER_FRENET = {"num_width": 3, "num_speed": 250, "num_t": 10, "min_t": 5.0, "max_t": 5.5,
                  "cost_wights": [1.0, 0.0, 99.4, 0.2, 0.1, 0.8, 0.4, 10.0], "max_distance": 140.0}
PARA_THRESHOLD = {"lon_accel_max": 3.0, "lon_jerk_max": 6.0, "lateral_accel_max": 0.5, "lateral_jerk_max": 1.0}


class GoalSamplingFrenetTrajectory(JerkSpaceSamplingFrenetTrajectory):
    """Same fields as the JSSP trajectory container (frenet_optimal_planner.py:26); arrays end at the candidate's own length."""


class FrenetOptimalPlannerSettings(object):
    """Same signature and attributes as the reference class (frenet_optimal_planner.py:56-83)."""

    def __init__(self, num_width: int = 5, num_speed: int = 5, num_t: int = 5, delta_t: float = 0.1, max_road_width: float = 3.5,
                 highest_speed: float = 13.4112, lowest_speed: float = 0.0, min_planning_t: float = 5.0, max_planning_t: float = 8.0):
        self.delta_t = delta_t
        self.num_width, self.num_t, self.num_speed = num_width, num_t, num_speed
        self.max_road_width = max_road_width
        self.highest_speed, self.lowest_speed = highest_speed, lowest_speed
        self.min_t, self.max_t = min_planning_t, max_planning_t

    @classmethod
    def intended(cls, speed_max: float = 18.0) -> "FrenetOptimalPlannerSettings":
        """The settings IntersectionPlanner.init_local_planner builds for method "frenet" (intersection_planner.py:101-109)."""
        f = PLANNER_FRENET
        return cls(num_width=f["num_width"], num_speed=f["num_speed"], num_t=f["num_t"], highest_speed=speed_max, lowest_speed=0.0,
                   min_planning_t=f["min_t"], max_planning_t=f["max_t"])


class FrenetOptimalPlanner(object):
    def __init__(self, planner_settings: FrenetOptimalPlannerSettings, ego_vehicle: Vehicle, device: int = 0, max_knots: int = 2048,
                 max_obstacle_modes: int = 256, max_total_steps: int = 1024):
        self.settings = planner_settings
        self.ego_vehicle = ego_vehicle
        self.best_traj: Optional[GoalSamplingFrenetTrajectory] = None
        self.all_trajs = []
        s = planner_settings
        n = s.num_width * s.num_t * s.num_speed
        cfg = JsspConfig(delta_t=s.delta_t, plan_horizon=s.max_t, max_road_width=s.max_road_width, num_width=s.num_width,
                         speed_max=s.highest_speed,                      # cost_total(tft, self.settings.highest_speed, ...) (:128-133)
                         ego_length=ego_vehicle.l, ego_width=ego_vehicle.w, obstacle_inflation=0.0,   # :280: plain L x W rectangles
                         cost_weights=tuple(PLANNER_FRENET["cost_wights"]), max_distance=PLANNER_FRENET["max_distance"],
                         thr_lon_accel=PARA_THRESHOLD["lon_accel_max"], thr_lon_jerk=PARA_THRESHOLD["lon_jerk_max"],
                         thr_lat_accel=PARA_THRESHOLD["lateral_accel_max"], thr_lat_jerk=PARA_THRESHOLD["lateral_jerk_max"],
                         max_seeds=max(64, -(-n // max(s.num_width, 1))), max_knots=max_knots, max_obstacle_modes=max_obstacle_modes,
                         max_total_steps=max_total_steps)
        self.engine = JsspEngine(cfg, device=device, max_candidates=n)
        self._cubic_spline = None
        self._obstacle_key = None
        self.last_result = None

    @property
    def cubic_spline(self):
        return self._cubic_spline

    @cubic_spline.setter
    def cubic_spline(self, spline):
        self._cubic_spline = spline
        if spline is not None:
            self.engine.set_refline(*spline_tables(spline))

    def generate_frenet_frame(self, centerline_pts: np.ndarray):
        self.cubic_spline = CubicSpline2D(centerline_pts[:, 0], centerline_pts[:, 1])
        s = np.arange(0, self._cubic_spline.s_list[-1], 0.1)
        return self._cubic_spline, self._cubic_spline.sample(s)

    def _set_obstacles(self, obstacles, dt: float, final_time_step: int):
        # cache by object identity (strong reference + ``is``) and length / version, like JerkSpaceSamplingPlanner.set_obstacles
        if isinstance(obstacles, ObstacleTensors):
            key, t = ("tensors", getattr(obstacles, "version", 0), 0), obstacles
        else:
            n_total = final_time_step + int(np.ceil(self.settings.max_t / self.settings.delta_t)) + 2
            key = ("dict", len(obstacles) if obstacles is not None else 0, n_total)
            t = None
        if getattr(self, "_obstacle_ref", None) is not obstacles or key != self._obstacle_key or obstacles is None:
            t = t if t is not None else pack_obstacle_dict(obstacles, dt, key[2])
            self.engine.set_obstacles(t.state, t.valid, t.size, t.prob, t.n_dict)
            self._obstacle_key, self._obstacle_ref = key, obstacles

    def invalidate_obstacles(self):
        self._obstacle_key = self._obstacle_ref = None

    def reset(self):
        """Forget the previous scenario (reference line, obstacles, best trajectory); the device workspace is kept."""
        self.best_traj, self.all_trajs, self._cubic_spline, self._obstacle_key, self.last_result = None, [], None, None, None
        self._obstacle_ref = None

    def plan_traj(self, frenet_state: FrenetState, ego_state: State, obstacles, observation) -> Optional[GoalSamplingFrenetTrajectory]:
        """Same contract as frenet_optimal_planner.py:301-360 (without the plotting side effects)."""
        if self._cubic_spline is None:
            raise RuntimeError("cubic_spline has not been set")
        t0 = time.perf_counter()
        ts = observation["test_setting"]
        now, final = int(ts["t"] / ts["dt"]), int(ts["max_t"] / ts["dt"])
        self._set_obstacles(obstacles, ts["dt"], final)
        s, v = self.settings, self.ego_vehicle
        res = self.engine.evaluate_fop((frenet_state.s, frenet_state.s_d, frenet_state.s_dd, frenet_state.d, frenet_state.d_d, frenet_state.d_dd),
                                       now, final, s.num_width, s.num_speed, s.num_t, s.min_t, s.max_t, s.lowest_speed, s.highest_speed,
                                       v.max_accel, v.max_jerk)
        self.last_result = res
        if res.best_idx >= 0:
            rows = res.best_traj
            n_own = int(np.count_nonzero(~np.isnan(rows[:, 1])))          # the candidate's own number of samples
            tr = GoalSamplingFrenetTrajectory.from_rows(rows[:n_own], res.best_n_pts, res.best_idx,
                                                        res.best_idx // (s.num_t * s.num_speed), res.best_cost)
            self.best_traj = tr
        self.all_trajs.append((now, res.n_candidates, res.best_idx, res.best_cost, time.perf_counter() - t0))
        return self.best_traj
