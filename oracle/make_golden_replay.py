#!/usr/bin/env python
"""Generate tests/golden/replay_golden_8_3_4_565.npz by driving the reference's own replay environment.

TEST INFRASTRUCTURE - build container only (needs /root/reference).  The scenario inputs/8_3_4_565 is parsed by the
reference's own ``ReplayParser`` (xosc + xodr, lxml -> ElementTree shim), initialised by its own ``ReplayController.init``
and advanced by its own ``Env.step`` -> ``Controller.step`` -> ``ReplayController.step`` (code/onsite_trans/env.py:21-33,
controller.py:256-263,301-415) in the default ``planner_expected_value`` mode, i.e. with action = (0.0, 0.0)
(code/__main__.py:139-140).  shapely is replaced by the convex-polygon stand-in of make_golden_cycle.py.  The planner in the
loop is the oracle (the reference's IntersectionPlanner cannot be imported), handing ``Env.step`` a trajectory object with the
fields it reads (x, y, yaw, s_d, s_dd).

What this pins - by executable reference code - is the driver-side semantics of one replay step:
  * t <- float(t + dt); the ego is first advanced by the kinematic model with the (0, 0) action,
  * the background vehicles are looked up with the key str(np.around(t, 3)),
  * the END STATUS is evaluated on that kinematically advanced ego (collision for t > 0.5, t >= max_t, map bounds of an
    "intersection" map, goal line) - BEFORE Env.step overwrites the ego with sample 1 of the selected trajectory,
  * the overwritten ego is what the next step starts from.

With ``--full`` the planner in the loop is the reference's own too (IntersectionPlanner.plan -> JerkSpaceSamplingPlanner.plan_traj,
leaf classes stubbed; ~7 minutes of Python); the result is stored under ``full_*`` and must equal the oracle-planner loop.

``--kinematics`` records the same complete loop in the reference's other ego-update mode (parameters["sim_config"]
["ego_update_mode"] = "kinematics": every cycle starts from FrenetState.from_state of the ego, the action is the planned
acceleration and the reference's own pure_pursuit_control, the ego follows the kinematic model) under ``kin_*``.

Usage:  python oracle/make_golden_replay.py  [--full] [--kinematics]
"""
import contextlib
import io
import os
import sys
import types
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
NAME = "8_3_4_565"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def full_reference_loop(parser0, sc, goal_pose, mode="planner_expected_value", prefix="full_"):
    """The WHOLE reference loop (code/__main__.py:124-167): its own IntersectionPlanner.plan (intersection_planner.py:304-343)
    calling its own JerkSpaceSamplingPlanner.plan_traj, and its own Env.step - with the leaf stubs only.  Two pieces of wiring
    are supplied here because the published tree lacks them: init_local_planner has no "JSSP" branch (the planner object is
    built with the intended settings) and plan_global_route needs the absent common.tool (the route
    centre line of the fixture is handed to the reference's own generate_frenet_frame)."""
    import types
    from onsite_trans.controller import Controller, ReplayController, ReplayParser
    from onsite_trans.env import Env
    sys.path.insert(0, os.path.join(REF, "code", "intersection_local_planner"))
    import intersection_local_planner.intersection_planner as ip   # the unmodified reference façade
    import intersection_local_planner.jerk_space_sampling_planner as jp
    base = os.path.join(REF, "inputs", NAME)
    parser = ReplayParser()
    parser._parse_openscenario(os.path.join(base, NAME + "_exam.xosc"))
    parser._parse_opendrive(os.path.join(base, NAME + ".xodr"))
    info = parser.replay_info
    ctl = Controller()
    ctl.mode, ctl.parser, ctl.control_info, ctl.controller = "replay", parser, info, ReplayController()
    ctl.observation = ctl.controller.init(info)
    ctl.traj = info.vehicle_traj
    env = object.__new__(Env)
    env.controller = ctl
    env.recorder = types.SimpleNamespace(record=lambda obs: None)
    obs = ctl.observation.format()
    with contextlib.redirect_stdout(io.StringIO()):
        planner = ip.IntersectionPlanner(ip.PlannerSettings(), obs, "frenet")          # the only branch the published wiring has
        planner.method = "JSSP"
        planner.fplanner = jp.JerkSpaceSamplingPlanner(
            jp.JerkSpaceSamplingPlannerSettings(num_width=3, num_jerk=25, delta_t=0.1, max_road_width=3.5, highest_jerk=10.0, plan_horizon=5.0,
                                                traj_num_max=50), planner.ego_vehicle)
        planner.fplanner.cubic_spline, _ = planner.generate_frenet_frame(sc.route_xy)   # __main__.py:100
    rows, ends, best_idx, actions = [], [float(obs["test_setting"]["end"])], [], []
    while obs["test_setting"]["end"] == -1:
        with contextlib.redirect_stdout(io.StringIO()):
            before = planner.fplanner.best_traj
            plan_acc = planner.plan(obs, info.vehicle_traj, ego_update_mode=mode)                    # __main__.py:129
            best = planner.fplanner.best_traj
            sampled = planner.fplanner.all_path_sampling_trajs[-1]
            pos = [n for n, tr in enumerate(sampled) if tr is best]
            best_idx.append(pos[0] if pos and best is not before else -1)
            action = (plan_acc, planner.pure_pursuit_control()) if mode == "kinematics" else (0.0, 0.0)   # __main__.py:136-140
            obs = env.step(action=action, observation_last=100, traj=info.vehicle_traj, best_traj_ego=best,
                           ego_update_mode=mode, goal_pose=goal_pose)                                 # __main__.py:158-165
        e = obs["vehicle_info"]["ego"]
        rows.append((obs["test_setting"]["t"], e["x"], e["y"], e["v"], e["a"], e["yaw"]))
        ends.append(float(obs["test_setting"]["end"]))
        actions.append((float(action[0]), float(action[1])))
        print(f"  {prefix}loop step {len(rows)} t={rows[-1][0]:.1f} best={best_idx[-1]} end={ends[-1]}", flush=True)
    return {prefix + "rows": np.array(rows), prefix + "ends": np.array(ends), prefix + "best_idx": np.array(best_idx),
            prefix + "actions": np.array(actions)}


def main(out):
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not mounted")
    import make_golden_cycle as G
    lx = types.ModuleType("lxml"); lx.etree = ET
    sys.modules["lxml"] = lx; sys.modules["lxml.etree"] = ET
    G.install_stubs()
    G.Polygon.convex_hull = property(lambda self: self)             # controller._get_poly asks for .convex_hull of a rectangle
    sys.path.insert(0, os.path.join(REF, "code"))
    from onsite_trans.controller import Controller, ReplayController, ReplayParser      # unmodified reference classes
    from onsite_trans.env import Env
    from oracle import jssp_oracle as O
    import dlii_jssp_b200 as J
    from dlii_jssp_b200 import replay as R, batch as B

    base = os.path.join(REF, "inputs", NAME)
    parser = ReplayParser()
    parser._parse_openscenario(os.path.join(base, NAME + "_exam.xosc"))
    parser._parse_opendrive(os.path.join(base, NAME + ".xodr"))
    info = parser.replay_info
    ctl = Controller()
    ctl.mode, ctl.parser, ctl.control_info, ctl.controller = "replay", parser, info, ReplayController()
    ctl.observation = ctl.controller.init(info)                     # ReplayController._get_initial_observation (:265-299)
    ctl.traj = info.vehicle_traj
    env = object.__new__(Env)                                       # Env.__init__ would build a Recorder that writes files
    env.controller = ctl
    env.recorder = types.SimpleNamespace(record=lambda obs: None)
    obs = ctl.observation.format()
    ts = obs["test_setting"]
    bounds = np.array([ts["x_min"], ts["x_max"], ts["y_min"], ts["y_max"]], dtype=np.float64)
    print("map_type", info.test_setting["map_type"], "bounds", bounds, "initial end", ts["end"])

    sc = J.scenario.scenario_from_fixture(os.path.join(ROOT, "tests", "golden", "scenario_8_3_4_565.npz"))
    assert abs(sc.ego["x"] - obs["vehicle_info"]["ego"]["x"]) < 1e-12
    goal_pose = R.goal_pose_of(sc)
    cfg = O.OracleConfig(ego_length=sc.ego["length"], ego_width=sc.ego["width"])
    spline = O.CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1])
    final = int(ts["max_t"] / ts["dt"])
    state, valid, size = O.pack_obstacles(info.vehicle_traj, ts["dt"], final + cfg.n_steps + 2)[:3]
    prob = np.ones(valid.shape[:2])
    fr = tuple(B.pack_scenario(sc).frenet0)
    rows, ends, best_idx, prev = [], [float(ts["end"])], [], None
    while obs["test_setting"]["end"] == -1:
        t = obs["test_setting"]["t"]
        now = int(t / obs["test_setting"]["dt"])
        with np.errstate(all="ignore"):
            r = O.plan_cycle(cfg, spline, fr, (state, valid.astype(bool), size, prob), now, final)
        b = r["best"]
        best_idx.append(b)
        if b < 0:                                                   # "No solution": plan_traj returns the previous best_traj (jssp.py:359-362)
            if prev is None:
                break                                               # first cycle: the reference's fallback raises (intersection_planner.py:332-337)
            r, b = prev
        prev = (r, b)
        if r["n_pts"][b] < 2:
            break
        best = types.SimpleNamespace(x=r["x"][b], y=r["y"][b], yaw=r["yaw"][b], s_d=r["s_d"][b], s_dd=r["s_dd"][b])
        with contextlib.redirect_stdout(io.StringIO()):
            obs = env.step(action=(0.0, 0.0), observation_last=100, traj=info.vehicle_traj, best_traj_ego=best,
                           ego_update_mode="planner_expected_value", goal_pose=goal_pose)          # __main__.py:158-165
        e = obs["vehicle_info"]["ego"]
        rows.append((obs["test_setting"]["t"], e["x"], e["y"], e["v"], e["a"], e["yaw"]))
        ends.append(float(obs["test_setting"]["end"]))
        fr = (r["s"][b, 1], r["s_d"][b, 1], r["s_dd"][b, 1], r["d"][b, 1], r["d_d"][b, 1], r["d_dd"][b, 1])
    extra = {}
    if "--full" in sys.argv:
        extra = full_reference_loop(parser, sc, goal_pose)
        assert extra["full_best_idx"].tolist() == best_idx and np.array_equal(extra["full_ends"], np.array(ends))
        np.testing.assert_allclose(extra["full_rows"], np.array(rows), rtol=1e-12, atol=1e-12)
        print("full reference loop == oracle-planner loop:", len(extra["full_rows"]), "steps")
    elif os.path.exists(out):                                        # keep previously recorded full loops
        old = np.load(out)
        extra = {k: old[k] for k in old.files if k.startswith("full_") or k.startswith("kin_")}
    if "--kinematics" in sys.argv:                                   # ego_update_mode = "kinematics": projected start state every cycle,
        extra.update(full_reference_loop(parser, sc, goal_pose, mode="kinematics", prefix="kin_"))   # pure-pursuit steering, kinematic ego
    np.savez_compressed(out, rows=np.array(rows), ends=np.array(ends), best_idx=np.array(best_idx), bounds=bounds,
                        goal_pose=np.array(goal_pose), map_type=np.array(str(info.test_setting["map_type"])), **extra)
    print(f"{len(rows)} steps, final end {ends[-1]}, last t {rows[-1][0] if rows else None}")
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main(os.path.join(ROOT, "tests", "golden", f"replay_golden_{NAME}.npz"))
