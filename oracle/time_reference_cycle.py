"""Timing of the reference's OWN plan cycle next to the two CPU forms bench.py reports (TEST / MEASUREMENT INFRASTRUCTURE).

Runs, on the four plan cycles of inputs/8_3_4_565 that oracle/make_golden_cycle.py records, with time.perf_counter() around the
call exactly where the reference times it (code/__main__.py:127-130):
  * the unmodified reference JerkSpaceSamplingPlanner.plan_traj (jerk_space_sampling_planner.py:296-362; absent leaf classes
    stubbed as in make_golden_cycle.py),
  * the loop-form float64 port (oracle/jssp_loop.py) - bench.py's ``cpu_baseline`` / ``--impl reference`` arm,
  * the vectorised NumPy oracle (oracle/jssp_oracle.py).
Needs /root/reference, so it only runs in the build container; its output is quoted in BASELINE.md section 2b.

    python oracle/time_reference_cycle.py
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_golden_cycle as G      # noqa: E402


def main():
    if not os.path.isdir(G.REF):
        raise SystemExit("reference tree not mounted")
    State, FrenetState, Vehicle = G.install_stubs()
    sys.path.insert(0, os.path.join(G.REF, "intersection_local_planner"))
    sys.path.insert(0, G.REF)
    import jerk_space_sampling_planner as ref
    from oracle import jssp_oracle as O, jssp_loop as L
    import dlii_jssp_b200 as J
    sc = J.scenario.scenario_from_fixture(os.path.join(ROOT, "tests", "golden", "scenario_8_3_4_565.npz"))
    settings = ref.JerkSpaceSamplingPlannerSettings(num_width=3, num_jerk=25, delta_t=0.1, max_road_width=3.5, highest_jerk=10.0,
                                                    plan_horizon=5.0, traj_num_max=50)
    cfg = O.OracleConfig(ego_length=sc.ego["length"], ego_width=sc.ego["width"])
    spline = O.CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1])
    final = int(sc.max_t / sc.dt)
    t_obs = J.scenario.pack_obstacle_dict(sc.vehicle_traj, sc.dt, final + cfg.n_steps + 2)
    obs_t = (t_obs.state, t_obs.valid.astype(bool), t_obs.size, t_obs.prob)
    print(f"{'cycle':>10} {'N':>6} {'reference plan_traj':>20} {'loop port':>10} {'vectorised':>11}   [s, 1 core]")
    for label, fr, t in G.CYCLES:
        planner = ref.JerkSpaceSamplingPlanner(settings, Vehicle(sc.ego["length"], sc.ego["width"]))
        planner.cubic_spline = O.CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1])
        obs = {"test_setting": {"t": t, "dt": sc.dt, "max_t": sc.max_t, "scenario_name": sc.name}}
        fs = FrenetState(t=t, s=fr[0], s_d=fr[1], s_dd=fr[2], d=fr[3], d_d=fr[4], d_dd=fr[5])
        with contextlib.redirect_stdout(io.StringIO()):
            t0 = time.perf_counter()
            planner.plan_traj(fs, State(t=t), sc.vehicle_traj, obs)
            t_ref = time.perf_counter() - t0
        n = len(planner.all_path_sampling_trajs[-1])
        now = int(t / sc.dt)
        with np.errstate(all="ignore"):
            t0 = time.perf_counter()
            seeds = O.sample_seeds(fr[0], fr[1], fr[2], cfg)
            L.plan_cycle_loop(cfg, spline, tuple(fr), obs_t, now, final, seeds=seeds)
            t_loop = time.perf_counter() - t0
            t0 = time.perf_counter()
            O.plan_cycle(cfg, spline, tuple(fr), obs_t, now, final)
            t_vec = time.perf_counter() - t0
        print(f"{label:>10} {n:>6} {t_ref:>20.3f} {t_loop:>10.3f} {t_vec:>11.3f}")


if __name__ == "__main__":
    main()
