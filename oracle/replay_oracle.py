"""CPU replay of one scenario with the oracle planner (TEST INFRASTRUCTURE): the same loop shape as the reference driver
(code/__main__.py:124-167, env.py:21-33) in ``planner_expected_value`` mode, used to check the CUDA planner's replay
cycle by cycle (selected index, ego rows)."""
from __future__ import annotations

import math

import numpy as np

from . import jssp_oracle as O


def replay(scenario, frenet0, pack_obstacles, end_status, goal_pose, max_cycles=None, cfg_overrides=None, fop_cfg=None):
    """``frenet0`` = initial Frenet state (s, s_d, s_dd, d, d_d, d_dd); ``pack_obstacles`` / ``end_status`` are the
    host-logic callables of the product (dict -> tensors, end-of-scenario test), passed in so that this file only
    supplies the planner arithmetic."""
    """With ``fop_cfg`` (oracle.fop_oracle.FopConfig) the FOP baseline planner replaces the JSSP planner in the loop."""
    cfg = O.OracleConfig(ego_length=scenario.ego["length"], ego_width=scenario.ego["width"], **(cfg_overrides or {}))
    horizon_steps = cfg.n_steps if fop_cfg is None else int(math.ceil(fop_cfg.max_t / fop_cfg.delta_t))
    spline = O.CubicSpline2D(scenario.route_xy[:, 0], scenario.route_xy[:, 1])
    final = int(scenario.max_t / scenario.dt)
    t = pack_obstacles(scenario.vehicle_traj, scenario.dt, final + horizon_steps + 2)
    obs_t = (t.state, t.valid.astype(bool), t.size, t.prob)
    obs = scenario.observation()
    out = dict(best_idx=[], n=[], rows=[], end=-1)
    fr = tuple(float(v) for v in frenet0)
    prev = None
    obs["test_setting"]["end"] = end_status(obs, scenario, None)
    while obs["test_setting"]["end"] == -1 and (max_cycles is None or len(out["best_idx"]) < max_cycles):
        ts = obs["test_setting"]
        now = int(ts["t"] / ts["dt"])
        with np.errstate(all="ignore"):
            if fop_cfg is None:
                r = O.plan_cycle(cfg, spline, fr, obs_t, now, final, n_obstacles_in_dict=t.n_dict)
            else:
                from . import fop_oracle
                r = fop_oracle.plan_cycle(fop_cfg, spline, fr, obs_t, now, final, n_obstacles_in_dict=t.n_dict)
        b = r["best"]
        out["best_idx"].append(b); out["n"].append(r["n"])
        if b < 0:
            # "No solution": plan_traj returns the PREVIOUS best_traj (jerk_space_sampling_planner.py:359-362), whose sample 1 the
            # driver re-applies; on the first cycle there is none and the reference's fallback raises (intersection_planner.py:332-337)
            if prev is None:
                ts["end"] = -2
                break
            r, b = prev
        prev = (r, b)
        if r["n_pts"][b] < 2:
            ts["end"] = -3
            break
        # Env.step in planner_expected_value mode (env.py:21-33, controller.py:256-263, 301-319): clock, kinematic ego with the
        # (0, 0) action, end status on THAT pose, then the ego is overwritten with sample 1 of the selected trajectory
        ts["t"] = float(ts["t"] + ts["dt"])
        ego = obs["vehicle_info"]["ego"]
        x0, y0, v0, yaw0 = float(ego["x"]), float(ego["y"]), float(ego["v"]), float(ego["yaw"])
        kin = dict(ego)
        kin["x"], kin["y"] = x0 + v0 * np.cos(yaw0) * ts["dt"], y0 + v0 * np.sin(yaw0) * ts["dt"]
        kin["yaw"] = (yaw0 + v0 / float(ego["length"]) * 1.7 * np.tan(0.0) * ts["dt"] + np.pi) % (2 * np.pi) - np.pi
        obs["vehicle_info"]["ego"] = kin
        ts["end"] = end_status(obs, scenario, goal_pose)
        ego["x"], ego["y"], ego["yaw"] = float(r["x"][b, 1]), float(r["y"][b, 1]), float(r["yaw"][b, 1])
        ego["v"], ego["a"] = float(r["s_d"][b, 1]), float(r["s_dd"][b, 1])
        obs["vehicle_info"]["ego"] = ego
        out["rows"].append((ts["t"], ego["x"], ego["y"], ego["v"], ego["a"], ego["yaw"]))
        fr = (r["s"][b, 1], r["s_d"][b, 1], r["s_dd"][b, 1], r["d"][b, 1], r["d_d"][b, 1], r["d_dd"][b, 1])
    out["end"] = obs["test_setting"]["end"]
    return out
