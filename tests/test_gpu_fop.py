"""FOP baseline planner through the CUDA path (jssp_fop_evaluate) against its CPU oracle (oracle/fop_oracle.py):
masks and the selected index bit-exact, costs and trajectory rows to tolerance."""
import numpy as np
import pytest

import dlii_jssp_b200 as J
from dlii_jssp_b200 import scenario as S
from oracle import fop_oracle as F, jssp_oracle as O

pytestmark = pytest.mark.gpu


def _run(sc, cfg: F.FopConfig, frenet0, now):
    spline = O.CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1])
    final = int(sc.max_t / sc.dt)
    obs = S.pack_obstacle_dict(sc.vehicle_traj, sc.dt, final + int(np.ceil(cfg.max_t / cfg.delta_t)) + 2)
    with np.errstate(all="ignore"):
        ref = F.plan_cycle(cfg, spline, frenet0, (obs.state, obs.valid.astype(bool), obs.size, obs.prob), now, final)
    veh = J.Vehicle(cfg.ego_length, cfg.ego_width, longitudinal_a_max=cfg.max_accel, longitudinal_jerk_max=cfg.max_jerk)
    st = J.FrenetOptimalPlannerSettings(num_width=cfg.num_width, num_speed=cfg.num_speed, num_t=cfg.num_t, highest_speed=cfg.highest_speed,
                                        lowest_speed=cfg.lowest_speed, min_planning_t=cfg.min_t, max_planning_t=cfg.max_t)
    pl = J.FrenetOptimalPlanner(st, veh)
    pl.cubic_spline = J.CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1])
    observation = sc.observation()
    observation["test_setting"]["t"] = now * sc.dt
    best = pl.plan_traj(J.FrenetState(s=frenet0[0], s_d=frenet0[1], s_dd=frenet0[2], d=frenet0[3], d_d=frenet0[4], d_dd=frenet0[5]),
                        J.State(), sc.vehicle_traj, observation)
    return ref, pl, best


def _compare(ref, pl, best):
    res = pl.last_result
    n = ref["n"]
    assert res.n_candidates == n
    assert np.array_equal(res.feasible.cpu().numpy().astype(bool), ref["feasible"])
    assert np.array_equal(res.collision.cpu().numpy().astype(bool), ref["collision"])
    np.testing.assert_allclose(res.cost.cpu().numpy(), ref["cost"], rtol=1e-4)
    assert res.best_idx == ref["best"], (res.best_idx, ref["best"])
    if ref["best"] >= 0:
        b = ref["best"]
        own = int(np.count_nonzero(~np.isnan(ref["s"][b])))
        assert len(best.s) == own and len(best.x) == int(ref["n_pts"][b])
        for key in ("s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd"):
            np.testing.assert_allclose(getattr(best, key), ref[key][b][:own], rtol=1e-9, atol=1e-9, err_msg=key)
        npts = int(ref["n_pts"][b])
        for key in ("x", "y", "yaw"):
            np.testing.assert_allclose(getattr(best, key), ref[key][b][:npts], rtol=1e-9, atol=1e-9, err_msg=key)
        assert abs(best.cost_total - ref["cost"][b]) <= 1e-9 * max(1.0, abs(ref["cost"][b]))


def test_fop_default_grid_on_bundled_scenario(scenario_fixture):
    """paras_planner["planner_frenet"]: 3 widths x 10 horizons (5.0 .. 5.5 s) x 250 target speeds = 7500 candidates."""
    sc = scenario_fixture
    cfg = F.FopConfig(ego_length=sc.ego["length"], ego_width=sc.ego["width"])
    for fr, now in (((2.93, 5.67, 0.0, 0.6, -0.4, 0.0), 0), ((14.0, 5.2, -0.3, 0.1, 0.0, 0.0), 20), ((27.0, 4.0, 0.2, 0.0, 0.0, 0.0), 50)):
        ref, pl, best = _run(sc, cfg, fr, now)
        _compare(ref, pl, best)
        assert ref["collision"].any() or now == 50
        pl.engine.close()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fop_reference_class_defaults_on_synthetic_intersections(seed):
    """FrenetOptimalPlannerSettings' own defaults (5 x 5 x 5, horizons 5 .. 8 s => per-candidate lengths 50 .. 80), trajectories that
    leave the reference line, end-of-scenario windows."""
    sc = S.synthetic_intersection(seed)
    cfg = F.FopConfig(num_width=5, num_speed=5, num_t=5, highest_speed=13.4112, min_t=5.0, max_t=8.0, max_accel=2.5, max_jerk=2.0)
    s_end = O.CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1]).s_list[-1]
    some_infeasible = False
    for fr, now in (((2.0, 6.0, 0.0, 0.3, 0.0, 0.0), 3), ((s_end - 25.0, 9.0, 0.5, -0.2, 0.1, 0.0), 40), ((s_end - 0.3, 2.0, 0.0, 0.0, 0.0, 0.0), 110)):
        ref, pl, best = _run(sc, cfg, fr, now)
        _compare(ref, pl, best)
        some_infeasible |= not ref["feasible"].all()
        pl.engine.close()
    assert some_infeasible                                    # tight acceleration / jerk limits reject the fast targets


def test_fop_replay_matches_oracle_replay(scenario_fixture):
    """The baseline planner in the replay loop (paras_planner["planner_type"] = "frenet"): cycle-by-cycle selections and ego
    rows against the oracle's FOP replay, reduced grid (3 x 4 x 40) to keep the CPU side short."""
    from dlii_jssp_b200 import replay as R, batch as B
    from oracle import replay_oracle
    sc = scenario_fixture
    cfg = F.FopConfig(num_width=3, num_speed=40, num_t=4, ego_length=sc.ego["length"], ego_width=sc.ego["width"])
    planner = J.IntersectionPlanner(J.PlannerSettings(), sc.observation(), "frenet")
    planner.fplanner.engine.close()
    planner.fplanner = J.FrenetOptimalPlanner(J.FrenetOptimalPlannerSettings(num_width=3, num_speed=40, num_t=4, highest_speed=18.0, min_planning_t=5.0,
                                                                            max_planning_t=5.5), planner.ego_vehicle)
    gpu = R.run_replay(sc, max_cycles=15, planner=planner, method="frenet")
    fr0 = tuple(B.pack_scenario(sc).frenet0)
    cpu = replay_oracle.replay(sc, fr0, S.pack_obstacle_dict, R.end_status, R.goal_pose_of(sc), max_cycles=15, fop_cfg=cfg)
    assert gpu.best_idx == cpu["best_idx"] and gpu.n_candidates == cpu["n"] and gpu.end == cpu["end"]
    np.testing.assert_allclose(np.array(gpu.ego_rows), np.array(cpu["rows"]), rtol=1e-9, atol=1e-9)


def test_fop_cuda_path_matches_the_reference_fop_methods(golden_dir, scenario_fixture):
    """jssp_fop_evaluate directly against tests/golden/fop_cycle_golden.npz (the reference's own FOP methods run with stubbed
    leaf classes, oracle/make_golden_fop.py): masks, costs, the last-minimum selection and the selected trajectory."""
    import os
    sc = scenario_fixture
    g = np.load(os.path.join(golden_dir, "fop_cycle_golden.npz"))
    W, NV, NT = [int(v) for v in g["grid"]]
    veh = J.Vehicle(sc.ego["length"], sc.ego["width"], longitudinal_a_max=float(g["limits"][0]), longitudinal_jerk_max=float(g["limits"][1]))
    st = J.FrenetOptimalPlannerSettings(num_width=W, num_speed=NV, num_t=NT, highest_speed=18.0, lowest_speed=0.0, min_planning_t=5.0, max_planning_t=5.5)
    pl = J.FrenetOptimalPlanner(st, veh)
    pl.cubic_spline = J.CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1])
    for k, label in enumerate(g["labels"]):
        label, fr, t = str(label), tuple(g["frenet0"][k]), float(g["t"][k])
        observation = sc.observation()
        observation["test_setting"]["t"] = t
        best = pl.plan_traj(J.FrenetState(s=fr[0], s_d=fr[1], s_dd=fr[2], d=fr[3], d_d=fr[4], d_dd=fr[5]), J.State(), sc.vehicle_traj, observation)
        res = pl.last_result
        feas = g[f"{label}_feasible"]
        assert res.n_candidates == int(g[f"{label}_n"])
        assert np.array_equal(res.feasible.cpu().numpy().astype(bool), feas)
        assert np.array_equal(res.collision.cpu().numpy().astype(bool)[feas], ~g[f"{label}_collision_passed"][feas])
        np.testing.assert_allclose(res.cost.cpu().numpy(), g[f"{label}_cost"], rtol=1e-4)
        assert res.best_idx == int(g[f"{label}_best"])
        rows = g[f"{label}_best_rows"]
        assert len(best.s) == len(rows) and len(best.x) == int(g[f"{label}_n_pts"][res.best_idx])
        for col, key in ((1, "s"), (2, "s_d"), (3, "s_dd"), (4, "s_ddd"), (5, "d"), (6, "d_d"), (7, "d_dd"), (8, "d_ddd")):
            np.testing.assert_allclose(getattr(best, key), rows[:, col], rtol=1e-9, atol=1e-9, err_msg=key)
        n = len(best.x)
        for col, key in ((9, "x"), (10, "y"), (11, "yaw")):
            np.testing.assert_allclose(getattr(best, key), rows[:n, col], rtol=1e-9, atol=1e-9, err_msg=key)
    pl.engine.close()


def test_fop_batched_replay_matches_single_replays_and_oracle(scenario_fixture):
    """The FOP baseline in the batched device-resident replay (jssp_batch_set_fop): every scenario of the batch reproduces the
    one-at-a-time FOP replay exactly and the oracle's FOP replay to tolerance (reduced grid 3 x 4 x 40)."""
    from dlii_jssp_b200 import replay as R, batch as B
    from oracle import replay_oracle
    st = J.FrenetOptimalPlannerSettings(num_width=3, num_speed=40, num_t=4, highest_speed=18.0, lowest_speed=0.0, min_planning_t=5.0, max_planning_t=5.5)
    scs = [scenario_fixture, S.perturbed_scenario(scenario_fixture, 2)]
    got = B.run_batch(scs, max_cycles=15, fop=st)
    for sc, g in zip(scs, got):
        planner = J.IntersectionPlanner(J.PlannerSettings(), sc.observation(), "frenet")
        planner.fplanner.engine.close()
        planner.fplanner = J.FrenetOptimalPlanner(st, planner.ego_vehicle)
        solo = R.run_replay(sc, max_cycles=15, planner=planner, method="frenet")
        assert g.best_idx == solo.best_idx and g.n_candidates == solo.n_candidates and g.end == solo.end and g.ego_rows == solo.ego_rows
        planner.fplanner.engine.close()
    sc = scs[0]
    cfg = F.FopConfig(num_width=3, num_speed=40, num_t=4, ego_length=sc.ego["length"], ego_width=sc.ego["width"])
    cpu = replay_oracle.replay(sc, tuple(B.pack_scenario(sc).frenet0), S.pack_obstacle_dict, R.end_status, R.goal_pose_of(sc), max_cycles=15, fop_cfg=cfg)
    assert got[0].best_idx == cpu["best_idx"] and got[0].end == cpu["end"]
    np.testing.assert_allclose(np.array(got[0].ego_rows), np.array(cpu["rows"]), rtol=1e-9, atol=1e-9)


def test_fop_batched_replay_in_kinematics_mode(scenario_fixture):
    """FOP baseline + "kinematics" ego-update mode together in the batched replay against the host-driven loop."""
    from dlii_jssp_b200 import replay as R, batch as B
    st = J.FrenetOptimalPlannerSettings(num_width=3, num_speed=40, num_t=4, highest_speed=18.0, lowest_speed=0.0, min_planning_t=5.0, max_planning_t=5.5)
    sc = scenario_fixture
    got = B.run_batch([sc], max_cycles=20, fop=st, ego_update_mode="kinematics")[0]
    planner = J.IntersectionPlanner(J.PlannerSettings(), sc.observation(), "frenet")
    planner.fplanner.engine.close()
    planner.fplanner = J.FrenetOptimalPlanner(st, planner.ego_vehicle)
    solo = R.run_replay(sc, max_cycles=20, planner=planner, method="frenet", ego_update_mode="kinematics")
    assert got.best_idx == solo.best_idx and got.end == solo.end and got.cycles == solo.cycles == 20
    np.testing.assert_allclose(np.array(got.ego_rows), np.array(solo.ego_rows), rtol=1e-9, atol=1e-9)
    planner.fplanner.engine.close()
