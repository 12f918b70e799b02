"""OnSite replay loop around the planner: the shape of the reference driver's per-scenario loop (code/__main__.py:95-167)
with the default ``ego_update_mode = "planner_expected_value"`` (configuration_parameters.py:4), i.e. after every plan cycle
the ego is moved to sample 1 of the selected trajectory (code/onsite_trans/env.py:24-30) and the next cycle starts from
``best_traj.frenet_state_at_time_step(1)`` (intersection_planner.py:312).

Only what the hot path needs is here: the state hand-off, the clock, and the end-of-scenario status of
ReplayController._update_end_status (controller.py:338-380: collision with a background vehicle after t > 0.5, t >= max_t,
map bounds, goal line passed) - evaluated, as in the reference, on the ego advanced by the kinematic model with the (0, 0)
action, before Env.step overwrites it with sample 1 of the selected trajectory; pinned by tests/golden/replay_golden_*.npz,
which oracle/make_golden_replay.py records from the reference's own Env.step / ReplayController.step.  Map parsing, route search, visualisation, CSV recording and the kinematic-bicycle mode are out of scope
(SURVEY.md section 2); the route centre line comes with the scenario."""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .planner import IntersectionPlanner, PlannerSettings
from .scenario import ReplayScenario

END_RUNNING, END_MAX_T, END_GOAL, END_COLLISION, END_NO_SOLUTION, END_SHORT_TRAJ = -1, 1, 1.5, 2, -2, -3


@dataclass
class ReplayResult:
    name: str
    end: float
    cycles: int
    best_idx: List[int] = field(default_factory=list)
    n_candidates: List[int] = field(default_factory=list)
    ego_rows: List[tuple] = field(default_factory=list)      # per step after the cycle: (t, x, y, v, a, yaw) - the Recorder's ego columns
    plan_seconds: List[float] = field(default_factory=list)
    seed_overflow: bool = False      # batched replay: some cycle of THIS scenario enumerated more than max_seeds seeds (surplus dropped)

    @property
    def p50_ms(self) -> float:
        return 1e3 * float(np.median(self.plan_seconds)) if self.plan_seconds else float("nan")


def rect_overlap(x1, y1, yaw1, l1, w1, x2, y2, yaw2, l2, w2) -> bool:
    """Closed-set overlap of two centred rectangles (ReplayController._collision_detect builds them with _get_poly and
    asks shapely ``intersects``, controller.py:382-415); separating-axis form."""
    c1, s1, c2, s2 = math.cos(yaw1), math.sin(yaw1), math.cos(yaw2), math.sin(yaw2)
    dx, dy = x2 - x1, y2 - y1
    a1, b1, a2, b2 = l1 / 2.0, w1 / 2.0, l2 / 2.0, w2 / 2.0
    cd, sd = abs(c1 * c2 + s1 * s2), abs(c1 * s2 - s1 * c2)
    return not (abs(dx * c1 + dy * s1) > a1 + a2 * cd + b2 * sd or abs(-dx * s1 + dy * c1) > b1 + a2 * sd + b2 * cd
                or abs(dx * c2 + dy * s2) > a2 + a1 * cd + b1 * sd or abs(-dx * s2 + dy * c2) > b2 + a1 * sd + b1 * cd)


def goal_pose_of(scenario: ReplayScenario) -> tuple:
    """(x_avg, y_avg, yaw of the last centre-line segment) (intersection_planner.py:181-182, 265-268, 283)."""
    gx, gy = sum(scenario.goal["x"]) / 2, sum(scenario.goal["y"]) / 2
    r = scenario.route_xy
    return gx, gy, math.atan2(r[-1, 1] - r[-2, 1], r[-1, 0] - r[-2, 0])


def kinematic_ego(ego: dict, dt: float, action=(0.0, 0.0)) -> dict:
    """ReplayController._update_ego_and_t (controller.py:301-319) after _action_cheaker's clipping (:265-268).  With the action
    (a, rot) = (0.0, 0.0) the driver passes in ``planner_expected_value`` mode (code/__main__.py:139-140) this is the pose on
    which the end status of the new step is evaluated, before Env.step overwrites the ego with sample 1 of the selected
    trajectory (env.py:23-30); in ``kinematics`` mode (acceleration, front-wheel angle) it IS the new ego."""
    a, rot = np.clip(action[0], -15, 15), np.clip(action[1], -1, 1)
    x, y, v, yaw = float(ego["x"]), float(ego["y"]), float(ego["v"]), float(ego["yaw"])
    out = dict(ego)
    out["x"] = x + v * np.cos(yaw) * dt
    out["y"] = y + v * np.sin(yaw) * dt
    yaw_new = yaw + v / float(ego["length"]) * 1.7 * np.tan(rot) * dt
    out["yaw"] = (yaw_new + np.pi) % (2 * np.pi) - np.pi
    out["v"] = v + a * dt
    if out["v"] < 0:
        out["v"] = 0
    out["a"] = a
    return out


def end_status(observation: dict, scenario: ReplayScenario, goal_pose: Optional[tuple]) -> float:
    """ReplayController._update_end_status (controller.py:338-380): collision of the ego with a background vehicle for
    t > 0.5 (:382-398), t >= max_t, the bounds of an "intersection" map (``scenario.bounds``; scenarios without a map skip
    it), goal line passed (``goal_pose`` None at initialisation, :297)."""
    status = [END_RUNNING]
    t = observation["test_setting"]["t"]
    ego = observation["vehicle_info"]["ego"]
    if t > 0.5:
        key = str(np.around(t, 3))
        for tr in scenario.vehicle_traj.values():
            st = tr.get(key)
            if st and rect_overlap(ego["x"], ego["y"], ego["yaw"], ego["length"], ego["width"], st["x"], st["y"], st["yaw"],
                                   tr["shape"]["length"], tr["shape"]["width"]):
                status.append(END_COLLISION)
                break
    if t >= scenario.max_t:
        status.append(END_MAX_T)
    if scenario.bounds is not None:
        x_min, x_max, y_min, y_max = scenario.bounds
        if ego["x"] > x_max or ego["x"] < x_min or ego["y"] > y_max or ego["y"] < y_min:
            status.append(END_GOAL)                            # 1.5: out of the map (:362-369)
    if goal_pose is not None:
        dx, dy = ego["x"] - goal_pose[0], ego["y"] - goal_pose[1]
        if math.sin(goal_pose[2]) * dy + math.cos(goal_pose[2]) * dx >= 0.0:
            status.append(END_GOAL)
    return max(status)


def advance(observation: dict, scenario: ReplayScenario, goal_pose: tuple, best, ego_update_mode: str = "planner_expected_value",
            action=(0.0, 0.0)) -> None:
    """One Env.step (env.py:21-33 -> controller.py:256-263): clock, kinematic ego with ``action``, end status ON THAT POSE;
    in ``planner_expected_value`` mode (action = (0, 0)) the ego is then overwritten with sample 1 of the selected trajectory."""
    ts = observation["test_setting"]
    ego = observation["vehicle_info"]["ego"]
    ts["t"] = float(ts["t"] + ts["dt"])                        # controller.py:303
    observation["vehicle_info"]["ego"] = kinematic_ego(ego, ts["dt"], action)
    ts["end"] = end_status(observation, scenario, goal_pose)   # controller.py:260
    if ego_update_mode == "planner_expected_value":
        ego["x"], ego["y"], ego["yaw"] = float(best.x[1]), float(best.y[1]), float(best.yaw[1])     # env.py:26-30
        ego["v"], ego["a"] = float(best.s_d[1]), float(best.s_dd[1])
        observation["vehicle_info"]["ego"] = ego


def run_replay(scenario: ReplayScenario, device: int = 0, max_cycles: Optional[int] = None, planner: Optional[IntersectionPlanner] = None,
               method: str = "JSSP", ego_update_mode: str = "planner_expected_value", **planner_kwargs) -> ReplayResult:
    """Replay one scenario with the CUDA planner (``method`` = "JSSP" or the "frenet" baseline, paras_planner["planner_type"];
    ``ego_update_mode`` = parameters["sim_config"]["ego_update_mode"]: the default "planner_expected_value", or "kinematics" -
    every cycle starts from the projected ego state and the ego follows the kinematic model driven by the planned acceleration
    and the pure-pursuit steering angle, code/__main__.py:136-138).  Returns per-cycle selections and the ego rows the reference
    would record."""
    obs = scenario.observation()
    if planner is None:
        planner = IntersectionPlanner(PlannerSettings(), obs, method, device=device, **planner_kwargs)
    route = scenario.route_xy
    planner.generate_frenet_frame(route)                       # __main__.py:100
    goal_pose = goal_pose_of(scenario)
    res = ReplayResult(scenario.name, END_RUNNING, 0)
    obs["test_setting"]["end"] = end_status(obs, scenario, None)       # _get_initial_observation: no goal pose yet (controller.py:297)
    while obs["test_setting"]["end"] == END_RUNNING and (max_cycles is None or res.cycles < max_cycles):
        t0 = time.perf_counter()
        plan_acc = 0.0
        try:
            plan_acc = planner.plan(obs, scenario.vehicle_traj, ego_update_mode=ego_update_mode)    # __main__.py:129
        except ValueError:                                         # "No solution available" on the very first cycle
            pass
        res.plan_seconds.append(time.perf_counter() - t0)
        r = planner.fplanner.last_result
        res.best_idx.append(r.best_idx); res.n_candidates.append(r.n_candidates); res.cycles += 1
        # "No solution" keeps the previous best_traj (jerk_space_sampling_planner.py:359-362): the driver then re-applies ITS
        # sample 1 - the ego stays where it is, the clock advances, the next cycle starts from the same Frenet state.  Without a
        # previous trajectory (first cycle) the reference's fallback raises (intersection_planner.py:332-337) and the scenario ends.
        best = planner.fplanner.best_traj
        if best is None:
            obs["test_setting"]["end"] = END_NO_SOLUTION
            break
        if len(best.x) < 2:
            obs["test_setting"]["end"] = END_SHORT_TRAJ        # env.py:26 would raise IndexError on best_traj.x[1]
            break
        action = (plan_acc, planner.pure_pursuit_control()) if ego_update_mode == "kinematics" else (0.0, 0.0)     # __main__.py:136-140
        advance(obs, scenario, goal_pose, best, ego_update_mode, action)
        ts, ego = obs["test_setting"], obs["vehicle_info"]["ego"]
        res.ego_rows.append((ts["t"], ego["x"], ego["y"], ego["v"], ego["a"], ego["yaw"]))
    res.end = obs["test_setting"]["end"]
    return res


def run_sweep(scenarios, device: int = 0, max_cycles: Optional[int] = None, planner: Optional[IntersectionPlanner] = None) -> List[ReplayResult]:
    """Replay a shard of independent scenarios on one GPU (BASELINE config 4: each rank calls this on its shard).  The
    planner (device workspace) is reused while the ego footprint stays the same."""
    out, dims = [], None
    if planner is not None:
        dims = (planner.ego_vehicle.l, planner.ego_vehicle.w)
    for sc in scenarios:
        d = (sc.ego["length"], sc.ego["width"])
        if planner is None or d != dims:
            planner, dims = IntersectionPlanner(PlannerSettings(), sc.observation(), "JSSP", device=device), d
        else:
            planner.reset(sc.observation())
        out.append(run_replay(sc, device=device, max_cycles=max_cycles, planner=planner))
    return out
