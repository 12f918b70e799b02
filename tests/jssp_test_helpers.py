"""Shared test plumbing: build oracle inputs and engine inputs from the same scenario / synthetic cycle."""
import numpy as np

from oracle import jssp_oracle as O


def oracle_config(ego_length, ego_width, **kw):
    return O.OracleConfig(ego_length=ego_length, ego_width=ego_width, **kw)


def oracle_obstacles(t):
    """ObstacleTensors -> the tuple oracle.plan_cycle takes."""
    return (t.state, t.valid.astype(bool), t.size, t.prob)


def engine_for(J, cfg: O.OracleConfig, **caps):
    jc = J.JsspConfig(delta_t=cfg.delta_t, plan_horizon=cfg.plan_horizon, max_road_width=cfg.max_road_width, num_width=cfg.num_width,
                      num_jerk=cfg.num_jerk, highest_jerk=cfg.highest_jerk, speed_max=cfg.speed_max, lon_accel_max=cfg.lon_accel_max,
                      ego_length=cfg.ego_length, ego_width=cfg.ego_width, max_curvature=cfg.max_curvature,
                      obstacle_inflation=cfg.obstacle_inflation, check_res=cfg.check_res, cost_weights=tuple(cfg.cost_weights),
                      max_distance=cfg.max_distance, thr_lon_accel=cfg.thr_lon_accel, thr_lon_jerk=cfg.thr_lon_jerk,
                      thr_lat_accel=cfg.thr_lat_accel, thr_lat_jerk=cfg.thr_lat_jerk, p_min=cfg.p_min, w_risk=cfg.w_risk, **caps)
    return J.JsspEngine(jc)


def compare_cycle(res, ref, rtol=1e-4, check_traj=True):
    """res = CycleResult from the CUDA path, ref = oracle.plan_cycle dict.  Masks and index bit-exact, floats to rtol."""
    n = ref["n"]
    assert res.n_candidates == n, (res.n_candidates, n)
    feas = res.feasible.cpu().numpy().astype(bool)
    coll = res.collision.cpu().numpy().astype(bool)
    cost = res.cost.cpu().numpy().astype(np.float64)
    bad_f = np.nonzero(feas != ref["feasible"])[0]
    bad_c = np.nonzero(coll != ref["collision"])[0]
    assert bad_f.size == 0, f"feasible mask differs at {bad_f[:10]} ({bad_f.size} of {n})"
    assert bad_c.size == 0, f"collision mask differs at {bad_c[:10]} ({bad_c.size} of {n})"
    # the cost is defined for every candidate, except that the soft-risk term of a vetoed candidate stops accumulating at
    # the veto (the reference only ever costs survivors, jerk_space_sampling_planner.py:313-320)
    sel = ~ref["collision"] if np.any(ref.get("risk", 0) != 0) else np.ones(n, dtype=bool)
    np.testing.assert_allclose(cost[sel], ref["cost"][sel], rtol=rtol, atol=0)
    assert res.best_idx == ref["best"], (res.best_idx, ref["best"], res.best_cost, ref["cost"][ref["best"]] if ref["best"] >= 0 else None)
    if ref["best"] >= 0 and check_traj and res.best_traj is not None:
        b = ref["best"]
        tr = res.best_traj
        assert abs(res.best_cost - ref["cost"][b]) <= 1e-9 * max(1.0, abs(ref["cost"][b]))
        npts = int(ref["n_pts"][b])
        assert res.best_n_pts == npts
        for col, key in ((1, "s"), (2, "s_d"), (3, "s_dd"), (4, "s_ddd")):
            assert np.array_equal(tr[:, col], ref[key][b]), key          # unroll is bit-identical to the reference
        for col, key in ((5, "d"), (6, "d_d"), (7, "d_dd"), (8, "d_ddd")):
            np.testing.assert_allclose(tr[:, col], ref[key][b], rtol=1e-9, atol=1e-12, err_msg=key)
        for col, key in ((9, "x"), (10, "y"), (11, "yaw")):
            np.testing.assert_allclose(tr[:npts, col], ref[key][b][:npts], rtol=1e-9, atol=1e-9, err_msg=key)
        if npts >= 2:
            with np.errstate(invalid="ignore"):
                np.testing.assert_allclose(tr[:npts - 1, 12], ref["ds"][b][:npts - 1], rtol=1e-9, atol=1e-12, err_msg="ds")
                np.testing.assert_allclose(tr[:npts - 1, 13], ref["c"][b][:npts - 1], rtol=1e-6, atol=1e-7, err_msg="c")
