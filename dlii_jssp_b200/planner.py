"""Drop-in mirror of the reference's local planner interface (code/intersection_local_planner/jerk_space_sampling_planner.py
and the JSSP branch of intersection_planner.py), backed by the CUDA library through ``JsspEngine``.

Same class names, constructor arguments, method names and attribute names as the reference, so that its replay driver
(code/__main__.py:95-167) can use this planner unchanged:

    planner.fplanner.cubic_spline = ...            (__main__.py:100)   -> uploads the spline coefficient table
    planner.fplanner.plan_traj(frenet_state, ego_state, obstacles=traj, observation=observation)   (intersection_planner.py:330)
    planner.fplanner.best_traj                     (__main__.py:164, env.py:26-30)
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Optional

import numpy as np

from .engine import JsspConfig, JsspEngine, CycleResult
from .frenet import FrenetState, JerkSpaceSamplingFrenetTrajectory, State, Vehicle
from .geometry import CubicSpline2D, spline_tables
from .scenario import ObstacleTensors, pack_obstacle_dict

# parameters["vehicle_para"] / ["para_threshold"] / paras_planner["planner_JSSP"] (configuration_parameters.py:7-22,52-61)
VEHICLE_PARA = {"speed_max": 18.0, "lon_accel_max": 8.0, "lon_jerk_max": 10.0, "lateral_accel_max": 1.2, "lateral_jerk_max": 2.0,
                "centripetal_accel_max": 2.0}
PARA_THRESHOLD = {"speed_max": 18.0, "lon_accel_max": 3.0, "lon_jerk_max": 6.0, "lateral_accel_max": 0.5, "lateral_jerk_max": 1.0,
                  "centripetal_accel_max": 1.0}
# paras_control["pure_pursuit"] (configuration_parameters.py:27-33)
PURE_PURSUIT = {"kv_ld0": [0.35, 4.8], "coff_wheel_base": 1.7}
PLANNER_JSSP = {"num_width": 3, "num_jerk": 25, "plan_horizon": 5.0, "cost_wights": [0.0, 3.0, 5.4, 0.2, 0.1, 0.8, 0.4, 4.0],
                "max_distance": 140.0, "traj_num_max": 50}


class JerkSpaceSamplingPlannerSettings(object):
    """Same signature and attributes as the reference class (jerk_space_sampling_planner.py:35-64)."""

    def __init__(self, num_width: int = 5, num_jerk: int = 5, delta_t: float = 0.1, max_road_width: float = 3.5, highest_jerk: float = 13.4112,
                 lowest_jerk: float = 0.0, plan_horizon: float = 5.0, traj_num_max: int = 5):
        self.delta_t = delta_t
        self.plan_horizon = plan_horizon
        self.max_road_width = max_road_width
        self.num_width = num_width
        self.highest_jerk = highest_jerk
        self.lowest_jerk = lowest_jerk
        self.num_jerk = num_jerk
        self.jerk_values_sampling = np.linspace(-self.highest_jerk, self.highest_jerk, self.num_jerk)
        self.traj_num_max = traj_num_max

    @classmethod
    def intended(cls) -> "JerkSpaceSamplingPlannerSettings":
        """The settings the published configuration intends (paras_planner["planner_JSSP"], highest_jerk = lon_jerk_max;
        SURVEY.md section 0.4 - the published wiring never builds them)."""
        return cls(num_width=PLANNER_JSSP["num_width"], num_jerk=PLANNER_JSSP["num_jerk"], delta_t=0.1, max_road_width=3.5,
                   highest_jerk=VEHICLE_PARA["lon_jerk_max"], plan_horizon=PLANNER_JSSP["plan_horizon"], traj_num_max=PLANNER_JSSP["traj_num_max"])


@dataclass
class CycleStats:
    time_step_now: int
    n_sampled: int
    n_passed: int
    best_idx: int
    best_cost: float
    seconds: float


class JerkSpaceSamplingPlanner(object):
    def __init__(self, planner_settings: JerkSpaceSamplingPlannerSettings, ego_vehicle: Vehicle, device: int = 0, verbose: bool = False,
                 max_seeds: int = 4096, max_knots: int = 2048, max_obstacle_modes: int = 256, max_total_steps: int = 1024,
                 p_min: float = 0.0, w_risk: float = 0.0):
        self.settings = planner_settings
        self.ego_vehicle = ego_vehicle
        self.ego_frenet_state = None
        self.ego_state = None
        self.best_traj: Optional[JerkSpaceSamplingFrenetTrajectory] = None
        self.all_trajs = []                 # per cycle: CycleStats (the reference keeps the surviving trajectory objects)
        self.all_path_sampling_trajs = []   # per cycle: number of sampled candidates
        self.time_step_max = int(round(self.settings.plan_horizon / self.settings.delta_t))
        self.verbose = verbose
        cfg = JsspConfig(
            delta_t=planner_settings.delta_t, plan_horizon=planner_settings.plan_horizon, max_road_width=planner_settings.max_road_width,
            num_width=planner_settings.num_width, num_jerk=planner_settings.num_jerk, highest_jerk=planner_settings.highest_jerk,
            speed_max=VEHICLE_PARA["speed_max"], lon_accel_max=VEHICLE_PARA["lon_accel_max"],
            ego_length=ego_vehicle.l, ego_width=ego_vehicle.w, max_curvature=ego_vehicle.max_curvature,
            cost_weights=tuple(PLANNER_JSSP["cost_wights"]), max_distance=PLANNER_JSSP["max_distance"],
            thr_lon_accel=PARA_THRESHOLD["lon_accel_max"], thr_lon_jerk=PARA_THRESHOLD["lon_jerk_max"],
            thr_lat_accel=PARA_THRESHOLD["lateral_accel_max"], thr_lat_jerk=PARA_THRESHOLD["lateral_jerk_max"],
            p_min=p_min, w_risk=w_risk, max_seeds=max_seeds, max_knots=max_knots, max_obstacle_modes=max_obstacle_modes,
            max_total_steps=max_total_steps)
        self.engine = JsspEngine(cfg, device=device)
        self._cubic_spline = None
        self._obstacle_key = self._obstacle_ref = None
        self.last_result = None            # jssp_cycle_out of the last cycle (n_candidates, n_seeds, best_idx, best_cost, best_n_pts)

    # -- reference line -------------------------------------------------------------------------------------------
    @property
    def cubic_spline(self):
        return self._cubic_spline

    @cubic_spline.setter
    def cubic_spline(self, spline):
        """The driver assigns the spline from outside (__main__.py:100); any object with s_list, sx.{a,b,c,d}, sy.{a,b,c,d}
        works (this package's CubicSpline2D or the reference's own)."""
        self._cubic_spline = spline
        if spline is not None:
            knot, coef = spline_tables(spline)
            self.engine.set_refline(knot, coef)

    def generate_frenet_frame(self, centerline_pts: np.ndarray):
        """Same contract as jerk_space_sampling_planner.py:364-371: returns (spline, [x, y, yaw, curvature] every 0.1 m)."""
        self.cubic_spline = CubicSpline2D(centerline_pts[:, 0], centerline_pts[:, 1])
        s = np.arange(0, self._cubic_spline.s_list[-1], 0.1)
        return self._cubic_spline, self._cubic_spline.sample(s)

    # -- obstacles ------------------------------------------------------------------------------------------------
    def set_obstacles(self, obstacles, dt: float, final_time_step: int):
        """``obstacles`` is the OnSite dictionary (packed once per scenario and cached) or ObstacleTensors (multi-modal
        predictions).  The cache holds a strong reference to the object it packed and compares with ``is`` (an ``id()`` alone
        can be recycled by a fresh temporary), plus the dictionary's length / the tensors' ``version``.  An object mutated IN
        PLACE between cycles (per-cycle IMM updates) must either bump ``ObstacleTensors.version`` or be followed by
        ``invalidate_obstacles()``; otherwise the device keeps the tables of the previous contents."""
        if isinstance(obstacles, ObstacleTensors):
            key = ("tensors", getattr(obstacles, "version", 0))
            if self._obstacle_ref is not obstacles or key != self._obstacle_key:
                self.engine.set_obstacles(obstacles.state, obstacles.valid, obstacles.size, obstacles.prob, obstacles.n_dict)
        else:
            n_total = final_time_step + self.time_step_max + 2
            key = ("dict", len(obstacles) if obstacles is not None else 0, n_total)
            if self._obstacle_ref is not obstacles or key != self._obstacle_key or obstacles is None:
                t = pack_obstacle_dict(obstacles, dt, n_total)
                self.engine.set_obstacles(t.state, t.valid, t.size, t.prob, t.n_dict)
        self._obstacle_key, self._obstacle_ref = key, obstacles

    def invalidate_obstacles(self):
        """Forget the cached obstacle tables: the next plan_traj packs and uploads ``obstacles`` again."""
        self._obstacle_key = self._obstacle_ref = None

    # -- one plan cycle ---------------------------------------------------------------------------------------------
    def sampling_frenet_trajectories(self, frenet_state: FrenetState, obstacles=None, time_step_now=None, observation=None) -> int:
        """Longitudinal seed enumeration on the device (PolyVTSampling.sampling + sampling_stop_seeds_supplement,
        jerk_space_sampling_planner.py:85-97).  Returns the number of candidates (num_width x seeds); the candidates
        themselves are never materialised on the host."""
        n_main, n_stop = self.engine.enumerate_seeds(frenet_state.s, frenet_state.s_d, frenet_state.s_dd, sync=True)
        return (n_main + n_stop) * max(self.settings.num_width, 1)

    def plan_traj(self, frenet_state: FrenetState, ego_state: State, obstacles, observation) -> Optional[JerkSpaceSamplingFrenetTrajectory]:
        """Same contract as jerk_space_sampling_planner.py:296-362: returns the minimum-cost trajectory that passes the
        curvature and collision checks, or the previous ``best_traj`` when no candidate survives ("No solution")."""
        if self._cubic_spline is None:
            raise RuntimeError("cubic_spline has not been set (the driver assigns it, __main__.py:100)")
        t0 = time.perf_counter()
        self.ego_frenet_state = frenet_state
        self.ego_state = ego_state
        ts = observation["test_setting"]
        time_step_now = int(ts["t"] / ts["dt"])                       # :300 (float truncation is part of the contract)
        final_time_step = int(ts["max_t"] / ts["dt"])                 # :244
        self.set_obstacles(obstacles, ts["dt"], final_time_step)
        res = self.engine.plan_cycle((frenet_state.s, frenet_state.s_d, frenet_state.s_dd, frenet_state.d, frenet_state.d_d, frenet_state.d_dd),
                                     time_step_now, final_time_step)
        self.last_result = res
        self.all_path_sampling_trajs.append(res.n_candidates)
        if res.best_idx >= 0:
            self.best_traj = JerkSpaceSamplingFrenetTrajectory.from_rows(self.engine._best_host, res.best_n_pts, res.best_idx,
                                                                       res.best_idx // max(res.n_seeds, 1), res.best_cost)
        elif self.verbose:
            print("##log## No solution.")
        dt_s = time.perf_counter() - t0
        self.all_trajs.append(CycleStats(time_step_now, res.n_candidates, -1, res.best_idx, res.best_cost, dt_s))
        if self.verbose:
            print("##log##", res.n_candidates, "trajectories are sampled.")
            print(f"###log### plan_traj time:{dt_s}\n")
        return self.best_traj

    def reset(self):
        """Forget the previous scenario (reference line, obstacles, best trajectory); the device workspace is kept."""
        self.best_traj = None
        self.all_trajs, self.all_path_sampling_trajs = [], []
        self._cubic_spline = None
        self._obstacle_key = self._obstacle_ref = None
        self.last_result = None

    def count_passed(self) -> int:
        """Number of candidates that passed both checks in the last cycle (the reference prints it, :324)."""
        r = self.last_result
        if r is None:
            return 0
        n = r.n_candidates
        return int(((self.engine.feasible[:n] == 1) & (self.engine.collision[:n] == 0)).sum().item())


class PlannerSettings(object):
    """intersection_planner.py:44-62."""

    def __init__(self):
        self.dt = 0.1
        self.speed_max = VEHICLE_PARA["speed_max"]
        self.accel_max = VEHICLE_PARA["lon_accel_max"]
        self.jerk_max = VEHICLE_PARA["lon_jerk_max"]
        self.lateral_accel_max = VEHICLE_PARA["lateral_accel_max"]
        self.lateral_jerk_max = VEHICLE_PARA["lateral_jerk_max"]
        self.centripetal_accel_max = VEHICLE_PARA["centripetal_accel_max"]
        self.max_road_width = 3.5
        self.speeds_ref = None


class IntersectionPlanner:
    """The JSSP branch of the reference façade (intersection_planner.py:65-163, 304-343): ego-state extraction, Frenet
    frame generation and the per-cycle state hand-off.  Global route search and pure pursuit are out of scope - the
    route centre line is given to ``generate_frenet_frame`` directly."""

    def __init__(self, planner_settings: PlannerSettings, observation: dict, method: str = "JSSP", device: int = 0, **planner_kwargs):
        if method not in ("JSSP", "frenet"):
            print("##log## ERROR: Planning method entered is not recognized!")
            raise ValueError(method)
        self.fplanner_settings = planner_settings
        self.dt = observation["test_setting"]["dt"]
        self.method = method
        self.ego_states, self.ego_frenet_states, self.time_list = [], [], []
        self.ego_frenet_state = FrenetState()
        self.cubic_spline = None
        self.ref_ego_lane_pts = None
        self.ego_vehicle = self._get_ego_vehicle_parameters(observation)
        self.ego_state = self._get_ego_states(observation)
        if method == "frenet":                      # the baseline branch of init_local_planner (intersection_planner.py:99-110)
            from .fop import FrenetOptimalPlanner, FrenetOptimalPlannerSettings
            self.fplanner = FrenetOptimalPlanner(FrenetOptimalPlannerSettings.intended(planner_settings.speed_max), self.ego_vehicle, device=device)
        else:
            self.fplanner = JerkSpaceSamplingPlanner(JerkSpaceSamplingPlannerSettings.intended(), self.ego_vehicle, device=device, **planner_kwargs)

    def _get_ego_vehicle_parameters(self, observation) -> Vehicle:
        s = self.fplanner_settings
        e = observation["vehicle_info"]["ego"]
        return Vehicle(e["length"], e["width"], longitudinal_v_max=s.speed_max, longitudinal_a_max=s.accel_max, longitudinal_jerk_max=s.jerk_max,
                       lateral_accel_max=s.lateral_accel_max, lateral_jerk_max=s.lateral_jerk_max, centripetal_accel_max=s.centripetal_accel_max)

    def _get_ego_states(self, observation) -> State:
        e = observation["vehicle_info"]["ego"]
        self.ego_state = State(t=observation["test_setting"]["t"], x=e["x"], y=e["y"], yaw=e["yaw"], v=e["v"], a=e["a"])
        return self.ego_state

    def reset(self, observation):
        """Start a new scenario with the same ego footprint."""
        self.dt = observation["test_setting"]["dt"]
        self.ego_states, self.ego_frenet_states, self.time_list = [], [], []
        self.ego_frenet_state = FrenetState()
        self.ego_state = self._get_ego_states(observation)
        self.fplanner.reset()

    def generate_frenet_frame(self, centerline_pts: np.ndarray):
        self.cubic_spline, self.ref_ego_lane_pts = self.fplanner.generate_frenet_frame(centerline_pts)
        return self.cubic_spline, self.ref_ego_lane_pts

    # -- path tracking of the "kinematics" ego-update mode (intersection_planner.py:363-443) ---------------------------------
    @staticmethod
    def find_lookahead_point_by_curve_distance(pos: np.ndarray, ref_pos: np.ndarray, ld: float = 5.0):
        """Same contract as intersection_planner.py:365-389: from the way point nearest to ``pos`` walk along the path until
        the accumulated arc length reaches ``ld``; returns (point, index)."""
        dist = np.linalg.norm(ref_pos - pos, axis=1)
        idx = int(np.argmin(dist))
        l_steps = 0
        while l_steps < ld and idx + 1 < len(ref_pos):
            l_steps += np.linalg.norm(ref_pos[idx + 1] - ref_pos[idx])
            idx += 1
        return ref_pos[idx], idx

    def pure_pursuit_control(self, kv: float = PURE_PURSUIT["kv_ld0"][0], ld0: float = PURE_PURSUIT["kv_ld0"][1],
                             tracing: str = "plan_path") -> float:
        """Front-wheel angle of the pure-pursuit tracker (intersection_planner.py:416-443) on the selected trajectory
        (``plan_path``, the published configuration) or on the re-discretised reference line (``ref_path``)."""
        ref_pos = self.ref_ego_lane_pts[:, 0:2] if tracing == "ref_path" else np.column_stack((self.fplanner.best_traj.x, self.fplanner.best_traj.y))
        ld = kv * self.ego_state.v + ld0
        pos = np.array([self.ego_state.x, self.ego_state.y])
        point, idx = self.find_lookahead_point_by_curve_distance(pos, ref_pos, ld)
        alpha = np.arctan2(point[1] - pos[1], point[0] - pos[0]) - self.ego_state.yaw
        wheel_base = self.ego_vehicle.l / PURE_PURSUIT["coff_wheel_base"]
        return np.arctan2(2 * wheel_base * np.sin(alpha), ld)

    def plan(self, observation, traj_future=None, ego_update_mode="planner_expected_value"):
        self.ego_state = self._get_ego_states(observation)
        self.ego_states.append(self.ego_state)
        if ego_update_mode == "planner_expected_value" and observation["test_setting"]["t"] > 0.05 and self.fplanner.best_traj is not None:
            self.ego_frenet_state = self.fplanner.best_traj.frenet_state_at_time_step(t=1)      # :312
        else:
            self.ego_frenet_state = FrenetState().from_state(self.ego_state, self.ref_ego_lane_pts)   # :309-310
        self.ego_frenet_states.append(self.ego_frenet_state)
        best = self.fplanner.plan_traj(self.ego_frenet_state, self.ego_state, obstacles=traj_future, observation=observation)
        if best is None:
            print("###log### No solution available for problem:")
            raise ValueError("no feasible trajectory")
        return best.s_dd[1]                                                                        # :342
