"""GPU parity tests: the CUDA path (through the C ABI) against the float64 oracle on the same inputs.
Bar: feasible mask, collision mask and selected index bit-exact; costs within 1e-4 relative (fp32 output); the exported
best trajectory's s, s_d, s_dd, s_ddd bit-identical, the rest within 1e-9."""
import os

import numpy as np
import pytest
import torch

import dlii_jssp_b200 as J
from oracle import jssp_oracle as O
import jssp_test_helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c1(scenario_fixture):
    sc = scenario_fixture
    cfg = H.oracle_config(sc.ego["length"], sc.ego["width"])
    spline = O.CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1])
    final = int(sc.max_t / sc.dt)
    obs = J.scenario.pack_obstacle_dict(sc.vehicle_traj, sc.dt, final + cfg.n_steps + 2)
    eng = H.engine_for(J, cfg)
    eng.set_refline(spline.s_list, spline.coefficient_table())
    eng.set_obstacles(obs.state, obs.valid, obs.size, obs.prob, obs.n_dict)
    return sc, cfg, spline, obs, final, eng


def test_seed_enumeration_matches_reference_golden(polyvt_golden):
    g = polyvt_golden
    for k, (s0, v0, a0, Jk, T, ns) in enumerate(g["meta"]):
        eng = J.JsspEngine(J.JsspConfig(plan_horizon=float(T), highest_jerk=float(Jk), num_jerk=int(ns)))
        n_main, n_stop = eng.enumerate_seeds(float(s0), float(v0), float(a0))
        seeds = eng.get_seeds()
        want = np.concatenate([g[f"c{k}_main"], g[f"c{k}_stop"]], 0)
        assert (n_main, n_stop) == (len(g[f"c{k}_main"]), len(g[f"c{k}_stop"])), k
        assert np.array_equal(seeds, want), f"case {k}: seeds differ from the reference's"
        eng.close()


@pytest.mark.parametrize("now", [0, 7, 30, 60, 100, 119])
def test_c1_cycle_matches_oracle(c1, now):
    sc, cfg, spline, obs, final, eng = c1
    # a plausible Frenet start state along the route at this time
    fr = (0.4 + 0.35 * now, max(5.686 - 0.04 * now, 0.0), -0.2 if now else 0.0, 0.1 * np.sin(now), 0.02, -0.01)
    ref = O.plan_cycle(cfg, spline, fr, H.oracle_obstacles(obs), now, final)
    n_main, n_stop = eng.enumerate_seeds(*fr[:3])
    assert n_main + n_stop == len(ref["seeds"])
    assert np.array_equal(eng.get_seeds(), ref["seeds"])
    res = eng.evaluate(fr, now, final)
    H.compare_cycle(res, ref)


def test_c1_state_dump(c1):
    sc, cfg, spline, obs, final, eng = c1
    fr = (2.0, 5.0, 0.3, 0.2, 0.0, 0.0)
    ref = O.plan_cycle(cfg, spline, fr, H.oracle_obstacles(obs), 3, final)
    eng.enumerate_seeds(*fr[:3])
    res = eng.evaluate(fr, 3, final, want_states=True)
    st = res.states.cpu().numpy().astype(np.float64)
    for col, key in enumerate(("s", "s_d", "s_dd", "d", "x", "y", "yaw")):
        a, b = st[:, :, col], ref[key]
        assert np.array_equal(np.isnan(a), np.isnan(b)), key
        np.testing.assert_allclose(np.nan_to_num(a), np.nan_to_num(b), rtol=1e-4, atol=1e-4, err_msg=key)
    c = st[:, :-1, 7]
    with np.errstate(invalid="ignore"):
        fin = np.isfinite(ref["c"]) & (np.abs(ref["c"]) < 1e3)
        np.testing.assert_allclose(c[fin], ref["c"][fin], rtol=1e-4, atol=1e-4)


def _dense(n_seeds, n_widths, O_, M, horizon, seed, **kw):
    cyc = J.scenario.dense_cycle(n_seeds=n_seeds, n_widths=n_widths, O=O_, M=M, plan_horizon=horizon, seed=seed)
    cfg = H.oracle_config(cyc.ego_length, cyc.ego_width, plan_horizon=horizon, **kw)
    spline = O.CubicSpline2D(cyc.route_xy[:, 0], cyc.route_xy[:, 1])
    eng = H.engine_for(J, cfg, max_seeds=n_seeds, max_obstacle_modes=max(O_ * M, 1), max_total_steps=cyc.obstacles.state.shape[2])
    eng.set_refline(spline.s_list, spline.coefficient_table())
    o = cyc.obstacles
    eng.set_obstacles(o.state, o.valid, o.size, o.prob, o.n_dict)
    return cyc, cfg, spline, eng


@pytest.mark.parametrize("n_seeds,n_widths,O_,M,horizon,seed", [
    (512, 4, 32, 3, 5.0, 1), (777, 3, 6, 1, 5.0, 2), (300, 5, 0, 1, 5.0, 3), (640, 8, 64, 3, 8.0, 4), (33, 1, 2, 2, 5.0, 5)])
def test_synthetic_cycle_matches_oracle(n_seeds, n_widths, O_, M, horizon, seed):
    cyc, cfg, spline, eng = _dense(n_seeds, n_widths, O_, M, horizon, seed)
    ref = O.plan_cycle(cfg, spline, cyc.frenet0, H.oracle_obstacles(cyc.obstacles), cyc.time_step_now, cyc.final_time_step,
                       seeds=cyc.seeds, targets=cyc.targets)
    seeds = torch.from_numpy(cyc.seeds).cuda()
    for kern in (None, "group", "thread", "warp", "tiled"):
        res = eng.evaluate(cyc.frenet0, cyc.time_step_now, cyc.final_time_step, seeds=seeds, targets=cyc.targets, kernel=kern)
        H.compare_cycle(res, ref)
    assert ref["collision"].any() or O_ == 0
    eng.close()


def test_multimodal_pmin_and_risk():
    """Modes below p_min do not veto; they add w_risk * prob to the cost (builder-specified extension)."""
    cyc, cfg, spline, eng = _dense(512, 4, 32, 3, 5.0, 11, p_min=0.3, w_risk=2.5)
    ref = O.plan_cycle(cfg, spline, cyc.frenet0, H.oracle_obstacles(cyc.obstacles), 0, cyc.final_time_step, seeds=cyc.seeds, targets=cyc.targets)
    res = eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=torch.from_numpy(cyc.seeds).cuda(), targets=cyc.targets)
    H.compare_cycle(res, ref)
    assert (ref["risk"] > 0).any()
    eng.close()


def test_single_mode_equals_reference_boolean_check():
    """M = 1, prob = 1 reduces to the reference's single-trajectory check: same masks as taking mode 0 alone."""
    cyc, cfg, spline, eng = _dense(400, 3, 16, 3, 5.0, 21)
    o = cyc.obstacles
    seeds = torch.from_numpy(cyc.seeds).cuda()
    any_mode = eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=seeds, targets=cyc.targets).collision.cpu().numpy().astype(bool)
    per_mode = []
    for m in range(3):
        eng.set_obstacles(o.state[:, m:m + 1], o.valid[:, m:m + 1], o.size, np.ones((o.state.shape[0], 1)), o.n_dict)
        per_mode.append(eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=seeds, targets=cyc.targets).collision.cpu().numpy().astype(bool))
    assert np.array_equal(any_mode, per_mode[0] | per_mode[1] | per_mode[2])
    eng.close()


def test_truncated_trajectories_and_end_of_scenario():
    """s leaving the spline truncates the Cartesian path (jerk_space_sampling_planner.py:138-139); the collision window is
    cut at final_time_step - time_step_now (:245)."""
    cyc, cfg, spline, eng = _dense(600, 3, 24, 2, 5.0, 31)
    s_end = spline.s_list[-1]
    seeds = torch.from_numpy(cyc.seeds).cuda()
    for fr, now, final in [((s_end - 12.0, 9.0, 0.0, 0.0, 0.0, 0.0), 0, 100), ((s_end - 0.05, 3.0, 0.0, 0.1, 0.0, 0.0), 0, 100),
                           ((s_end + 1.0, 3.0, 0.0, 0.0, 0.0, 0.0), 0, 100), ((5.0, 6.0, 0.0, 0.0, 0.0, 0.0), 40, 47),
                           ((5.0, 6.0, 0.0, 0.0, 0.0, 0.0), 49, 50), ((5.0, 6.0, 0.0, 0.0, 0.0, 0.0), 50, 50)]:
        ref = O.plan_cycle(cfg, spline, fr, H.oracle_obstacles(cyc.obstacles), now, final, seeds=cyc.seeds, targets=cyc.targets)
        for kern in (None, "group", "tiled"):
            H.compare_cycle(eng.evaluate(fr, now, final, seeds=seeds, targets=cyc.targets, kernel=kern), ref)
    eng.close()


def _fullsize_golden(name, feas, coll, cost, res):
    """The CUDA path at FULL size against the oracle's results for every candidate (tests/golden/fullsize_golden.npz, written by
    oracle/make_golden_fullsize.py from the same seeded generator): both masks bit-exact, the selected index equal to the
    oracle's argmin (lowest index among equal costs), its cost, and the costs of every 64th candidate."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize_golden.npz"))
    n = int(g[f"{name}_n"][0])
    assert len(feas) == n
    gf = np.unpackbits(g[f"{name}_feasible_bits"])[:n].astype(bool)
    gc = np.unpackbits(g[f"{name}_collision_bits"])[:n].astype(bool)
    assert np.array_equal(feas, gf), f"feasible mask differs at {np.nonzero(feas != gf)[0][:10]}"
    assert np.array_equal(coll, gc), f"collision mask differs at {np.nonzero(coll != gc)[0][:10]}"
    assert res.best_idx == int(g[f"{name}_best"][0]), (res.best_idx, int(g[f"{name}_best"][0]))
    assert abs(res.best_cost - float(g[f"{name}_best_cost"][0])) <= 1e-9
    np.testing.assert_allclose(cost[::64], g[f"{name}_cost_stride64"], rtol=1e-4)


def test_c3_full_size_properties_and_subsample_parity():
    """BASELINE config 3 at full size: the oracle's full-size golden (every mask bit, the argmin), size-independent properties,
    and a live oracle run on a candidate subsample."""
    cyc, cfg, spline, eng = _dense(16384, 4, 32, 3, 5.0, 20251019)
    seeds = torch.from_numpy(cyc.seeds).cuda()
    res = eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=seeds, targets=cyc.targets)
    assert res.n_candidates == 65536
    feas = res.feasible.cpu().numpy().astype(bool); coll = res.collision.cpu().numpy().astype(bool); cost = res.cost.cpu().numpy()
    _fullsize_golden("C3", feas, coll, cost, res)
    ok = feas & ~coll
    assert ok.any() and res.best_idx >= 0 and ok[res.best_idx]
    # argmin property: nothing that survived is cheaper (fp32 costs: the float64 argmin rounds to the fp32 minimum)
    assert cost[res.best_idx] <= cost[ok].min()
    # subsample parity: every 8th seed, all widths
    sub = cyc.seeds[::8]
    ref = O.plan_cycle(cfg, spline, cyc.frenet0, H.oracle_obstacles(cyc.obstacles), 0, cyc.final_time_step, seeds=sub, targets=cyc.targets)
    Ns, ns = len(cyc.seeds), len(sub)
    pick = np.concatenate([w * Ns + np.arange(0, Ns, 8) for w in range(4)])
    assert np.array_equal(feas[pick], ref["feasible"]) and np.array_equal(coll[pick], ref["collision"])
    np.testing.assert_allclose(cost[pick], ref["cost"], rtol=1e-4)
    # removing obstacles can only clear collisions (monotonicity)
    o = cyc.obstacles
    eng.set_obstacles(o.state[:16], o.valid[:16], o.size[:16], o.prob[:16], 16)
    res2 = eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=seeds, targets=cyc.targets)
    coll2 = res2.collision.cpu().numpy().astype(bool)
    assert not (coll2 & ~coll).any()
    # permutation invariance of the selected candidate (as a seed) under reversal of the seed order
    rev = torch.from_numpy(cyc.seeds[::-1].copy()).cuda()
    eng.set_obstacles(o.state, o.valid, o.size, o.prob, o.n_dict)
    res3 = eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=rev, targets=cyc.targets)
    w3, k3 = divmod(res3.best_idx, Ns); w1, k1 = divmod(res.best_idx, Ns)
    assert abs(res3.best_cost - res.best_cost) <= 1e-12 and w3 == w1
    assert np.array_equal(cyc.seeds[::-1][k3], cyc.seeds[k1])       # the same seed wins (the golden's runner-up margin is 2e-3)
    eng.close()


@pytest.mark.parametrize("n_seeds,n_widths,O_,M,horizon,seed", [(4096, 4, 32, 3, 5.0, 51), (3000, 3, 6, 1, 5.0, 52), (2048, 8, 64, 3, 8.0, 53)])
def test_select_only_evaluation_selects_the_same_trajectory(n_seeds, n_widths, O_, M, horizon, seed):
    """No per-candidate outputs requested (select-only: walks stop at the first veto) -> same selected index, cost and exported
    trajectory as the full evaluation, for every kernel."""
    cyc, cfg, spline, eng = _dense(n_seeds, n_widths, O_, M, horizon, seed, p_min=0.2, w_risk=0.5)
    seeds = torch.from_numpy(cyc.seeds).cuda()
    for kern in ("group", "thread", "warp", "tiled"):
        full = eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=seeds, targets=cyc.targets, kernel=kern)
        assert full.best_idx >= 0 and bool(full.collision.any())
        sel = eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=seeds, targets=cyc.targets, kernel=kern, select_only=True)
        assert (sel.best_idx, sel.best_cost, sel.best_n_pts) == (full.best_idx, full.best_cost, full.best_n_pts)
        assert np.array_equal(np.nan_to_num(sel.best_traj), np.nan_to_num(full.best_traj))
    eng.close()


def test_c5_full_size_properties_and_subsample_parity():
    """BASELINE config 5 at full size on one GPU (1,048,576 candidates x 81 samples x 64 obstacles x 3 modes): argmin property,
    exact mask parity on a candidate subsample, and the candidate split of the multi-GPU path (8 contiguous seed ranges
    evaluated one after the other, reduced by (cost, index)) selects the same candidate."""
    cyc, cfg, spline, eng = _dense(131072, 8, 64, 3, 8.0, 20251020)
    seeds = torch.from_numpy(cyc.seeds).cuda()
    res = eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=seeds, targets=cyc.targets)
    Ns = len(cyc.seeds)
    assert res.n_candidates == 8 * Ns == 1048576
    feas = res.feasible.cpu().numpy().astype(bool); coll = res.collision.cpu().numpy().astype(bool); cost = res.cost.cpu().numpy()
    _fullsize_golden("C5", feas, coll, cost, res)
    ok = feas & ~coll
    assert ok.any() and ok[res.best_idx] and cost[res.best_idx] <= cost[ok].min()
    step = 256
    ref = O.plan_cycle(cfg, spline, cyc.frenet0, H.oracle_obstacles(cyc.obstacles), 0, cyc.final_time_step, seeds=cyc.seeds[::step], targets=cyc.targets)
    pick = np.concatenate([w * Ns + np.arange(0, Ns, step) for w in range(8)])
    assert np.array_equal(feas[pick], ref["feasible"]) and np.array_equal(coll[pick], ref["collision"])
    np.testing.assert_allclose(cost[pick], ref["cost"], rtol=1e-4)
    best = (np.inf, -1)
    for r in range(8):                                       # distributed.candidate_range: contiguous seed ranges per rank
        lo, hi = r * Ns // 8, (r + 1) * Ns // 8
        part = eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=seeds[lo:hi].contiguous(), targets=cyc.targets)
        if part.best_idx >= 0:
            w, k = divmod(part.best_idx, hi - lo)
            best = min(best, (part.best_cost, w * Ns + lo + k))
    assert best[1] == res.best_idx and abs(best[0] - res.best_cost) <= 1e-12
    eng.close()


@pytest.mark.parametrize("n_seeds,n_widths,O_,M,horizon,seed", [(1024, 4, 32, 3, 5.0, 41), (700, 3, 6, 1, 5.0, 42), (512, 8, 64, 3, 8.0, 43)])
def test_block_level_culling_changes_nothing(n_seeds, n_widths, O_, M, horizon, seed):
    """The per-block culled obstacle lists (default) and the exhaustive table walk (JSSP_FLAG_NO_CULL) give identical
    masks, costs and selection - and both match the oracle."""
    cyc, cfg, spline, eng = _dense(n_seeds, n_widths, O_, M, horizon, seed, p_min=0.15, w_risk=0.7)
    seeds = torch.from_numpy(cyc.seeds).cuda()
    out = []
    for kern in ("group", "thread", "tiled"):  # the kernels that build the per-block lists (small cycles default to the warp kernel)
        for cull in (True, False):
            r = eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=seeds, targets=cyc.targets, cull=cull, kernel=kern)
            out.append((r.feasible.cpu().numpy().copy(), r.collision.cpu().numpy().copy(), r.cost.cpu().numpy().copy(), r.best_idx, r.best_cost))
    ok = out[0][1] == 0          # a vetoed candidate stops accumulating risk at the veto, wherever the walk is then
    for other in out[1:]:
        assert np.array_equal(out[0][0], other[0]) and np.array_equal(out[0][1], other[1]) and np.array_equal(out[0][2][ok], other[2][ok])
        assert out[0][3:] == other[3:]
    ref = O.plan_cycle(cfg, spline, cyc.frenet0, H.oracle_obstacles(cyc.obstacles), 0, cyc.final_time_step, seeds=cyc.seeds, targets=cyc.targets)
    res = eng.evaluate(cyc.frenet0, 0, cyc.final_time_step, seeds=seeds, targets=cyc.targets, kernel="group")
    H.compare_cycle(res, ref)
    # a start offset beyond the lateral cap switches culling off instead of risking a miss
    wide = (cyc.frenet0[0], cyc.frenet0[1], 0.0, 4.2, 0.0, 0.0)
    ref = O.plan_cycle(cfg, spline, wide, H.oracle_obstacles(cyc.obstacles), 0, cyc.final_time_step, seeds=cyc.seeds, targets=cyc.targets)
    for kern in ("group", "thread", "tiled"):
        H.compare_cycle(eng.evaluate(wide, 0, cyc.final_time_step, seeds=seeds, targets=cyc.targets, kernel=kern), ref)
    eng.close()


@pytest.mark.parametrize("n_seeds,n_widths,O_,M,horizon,seed", [(300, 3, 6, 1, 5.0, 51), (257, 4, 40, 3, 5.0, 52), (90, 5, 70, 2, 8.0, 53), (64, 2, 0, 1, 5.0, 54)])
def test_warp_and_thread_kernels_agree(n_seeds, n_widths, O_, M, horizon, seed):
    """The warp-per-candidate kernel (small cycles), the group kernel (large cycles) and the thread-per-candidate kernel give bit-identical masks, fp32
    costs, selection and exported trajectory - and match the oracle, including truncated trajectories."""
    cyc, cfg, spline, eng = _dense(n_seeds, n_widths, O_, M, horizon, seed, p_min=0.2, w_risk=0.9)
    seeds = torch.from_numpy(cyc.seeds).cuda()
    s_end = spline.s_list[-1]
    for fr, now, final in [(cyc.frenet0, 0, cyc.final_time_step), ((s_end - 9.0, 8.0, 0.5, -0.3, 0.1, 0.0), 0, 100),
                           ((s_end - 0.04, 3.0, 0.0, 0.1, 0.0, 0.0), 0, 100), ((4.0, 5.0, 0.0, 0.0, 0.0, 0.0), 43, 50)]:
        out = {}
        for kern in ("warp", "thread", "group", "tiled"):
            r = eng.evaluate(fr, now, final, seeds=seeds, targets=cyc.targets, kernel=kern)
            out[kern] = (r.feasible.cpu().numpy().copy(), r.collision.cpu().numpy().copy(), r.cost.cpu().numpy().copy(), r.best_idx, r.best_cost,
                         r.best_n_pts, None if r.best_traj is None else r.best_traj.copy())
        b = out["thread"]
        for a in (out["warp"], out["group"], out["tiled"]):      # the tiled kernel adds its cost terms in time order too: bit-identical costs
            ok = a[1] == 0
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2][ok], b[2][ok])
            assert a[3:6] == b[3:6]
            if a[6] is not None:
                assert np.array_equal(a[6], b[6], equal_nan=True)
        ref = O.plan_cycle(cfg, spline, fr, H.oracle_obstacles(cyc.obstacles), now, final, seeds=cyc.seeds, targets=cyc.targets)
        H.compare_cycle(eng.evaluate(fr, now, final, seeds=seeds, targets=cyc.targets, kernel="warp"), ref)
    eng.close()


def test_cuda_path_matches_the_reference_plan_cycle(golden_dir, scenario_fixture):
    """The CUDA plan cycle (device seed enumeration + evaluation + selection + export) directly against
    tests/golden/jssp_cycle_golden.npz = the reference's own JerkSpaceSamplingPlanner.plan_traj run by
    oracle/make_golden_cycle.py: candidate count, curvature mask, collision mask (where the reference evaluates it), survivor
    costs, selected index and the selected trajectory - through every evaluation kernel."""
    import os
    sc = scenario_fixture
    g = np.load(os.path.join(golden_dir, "jssp_cycle_golden.npz"))
    final = int(sc.max_t / sc.dt)
    obs = J.scenario.pack_obstacle_dict(sc.vehicle_traj, sc.dt, final + 52)
    eng = J.JsspEngine(J.JsspConfig(ego_length=sc.ego["length"], ego_width=sc.ego["width"]))
    sp = J.CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1])
    eng.set_refline(sp.s_list, sp.coefficient_table())
    eng.set_obstacles(obs.state, obs.valid, obs.size, obs.prob, obs.n_dict)
    for k, label in enumerate(g["labels"]):
        label, fr, now = str(label), tuple(g["frenet0"][k]), int(float(g["t"][k]) / sc.dt)
        eng.enumerate_seeds(*fr[:3])
        for kern in (None, "thread", "group", "tiled"):
            res = eng.evaluate(fr, now, final, kernel=kern)
            feas_ref = g[f"{label}_feasible"]
            assert res.n_candidates == int(g[f"{label}_n"])
            feas = res.feasible.cpu().numpy().astype(bool); coll = res.collision.cpu().numpy().astype(bool)
            assert np.array_equal(feas, feas_ref), (label, kern)
            assert np.array_equal(coll[feas_ref], ~g[f"{label}_collision_passed"][feas_ref]), (label, kern)
            surv = ~np.isnan(g[f"{label}_cost"])
            np.testing.assert_allclose(res.cost.cpu().numpy()[surv], g[f"{label}_cost"][surv], rtol=1e-4)
            assert res.best_idx == int(g[f"{label}_best"]), (label, kern, res.best_idx)
            if res.best_idx >= 0:
                rows, tr = g[f"{label}_best_rows"], res.best_traj
                npts = int(g[f"{label}_n_pts"][res.best_idx])
                assert res.best_n_pts == npts
                assert np.array_equal(tr[:, 1:5], rows[:, 1:5])                                  # s, s_d, s_dd, s_ddd bit-identical
                np.testing.assert_allclose(tr[:, 5:9], rows[:, 5:9], rtol=1e-9, atol=1e-12)     # lateral quintic (closed form vs solve)
                np.testing.assert_allclose(tr[:npts, 9:12], rows[:npts, 9:12], rtol=1e-9, atol=1e-9)
                assert abs(res.best_cost - g[f"{label}_cost"][res.best_idx]) <= 1e-9 * abs(res.best_cost)
    eng.close()


@pytest.mark.parametrize("yaw_o", [0.0, 0.3])
def test_guard_band_decisions_near_touching_rectangles(yaw_o):
    """The fp32 separating-axis classification with its 2 mm guard band must hand every near-tie to the float64 test: obstacles
    placed so that their rectangle is within -3 mm .. +3 mm (down to 1e-12 m, and exactly touching) of the ego footprint's side
    give the oracle's collision mask bit for bit - shapely's closed-set ``intersects`` counts touching as a hit."""
    route = np.column_stack([np.linspace(0.0, 120.0, 241), np.zeros(241)])
    rng = np.random.default_rng(5)
    seeds = J.scenario.random_seeds(96, rng)
    cfg = H.oracle_config(4.924, 1.872)
    spline = O.CubicSpline2D(route[:, 0], route[:, 1])
    eng = H.engine_for(J, cfg, max_seeds=96, max_obstacle_modes=1, max_total_steps=60)
    eng.set_refline(spline.s_list, spline.coefficient_table())
    seeds_d = torch.from_numpy(seeds).cuda()
    fr = (0.0, 6.0, 0.0, 0.0, 0.0, 0.0)                     # d = 0 towards target 0: the ego rectangle stays axis-aligned
    L, Wd = 8.0, 2.0
    a, b = (L + 0.1) / 2.0, (Wd + 0.1) / 2.0                 # inflated half sizes (jssp.py:263-265)
    # lateral extent of the rotated obstacle rectangle; gap = y_o - extent - ego half width
    ext = a * abs(np.sin(yaw_o)) + b * abs(np.cos(yaw_o))
    seen = set()
    for delta in (0.0, 1e-12, -1e-12, 1e-9, -1e-9, 1e-6, -1e-6, 1e-4, -1e-4, 1.9e-3, -1.9e-3, 2.1e-3, -2.1e-3, 3e-3, -3e-3):
        y_o = 1.872 / 2.0 + ext + delta
        state = np.zeros((1, 1, 60, 3)); state[0, 0, :, 0] = 14.0; state[0, 0, :, 1] = y_o; state[0, 0, :, 2] = yaw_o
        valid = np.ones((1, 1, 60), dtype=np.uint8); size = np.array([[L, Wd]]); prob = np.ones((1, 1))
        eng.set_obstacles(state, valid, size, prob, 1)
        ref = O.plan_cycle(cfg, spline, fr, (state, valid.astype(bool), size, prob), 0, 55, seeds=seeds, targets=[0.0])
        for kern in ("warp", "thread", "group", "tiled"):
            res = eng.evaluate(fr, 0, 55, seeds=seeds_d, targets=[0.0], kernel=kern)
            assert np.array_equal(res.collision.cpu().numpy().astype(bool), ref["collision"]), (delta, kern)
            assert res.best_idx == ref["best"]
        seen.add(bool(ref["collision"].any()))
    assert seen == {True, False}                              # both sides of the tie were exercised
    eng.close()


def test_guard_band_decisions_near_the_curvature_limit():
    """The square-root/atan-free curvature fast path must hand every near-tie to the float64 atan2 form: with max_curvature set
    within 1e-10 .. 1e-3 (relative) above and below a candidate's own largest |c|, the feasible mask equals the oracle's."""
    cyc = J.scenario.dense_cycle(n_seeds=48, n_widths=1, O=0, M=1, seed=9)
    spline = O.CubicSpline2D(cyc.route_xy[:, 0], cyc.route_xy[:, 1])
    fr = (22.0, 7.0, 0.0, 0.3, 0.0, 0.0)                     # starts just before the 15 m-radius arc
    base = H.oracle_config(cyc.ego_length, cyc.ego_width)
    empty = H.oracle_obstacles(cyc.obstacles)
    with np.errstate(all="ignore"):
        r0 = O.plan_cycle(base, spline, fr, empty, 0, 100, seeds=cyc.seeds, targets=[0.0])
        cmax = np.nanmax(np.abs(r0["c"]), axis=1)
    pivot = float(np.sort(cmax[np.isfinite(cmax) & (cmax < 1.0)])[len(cmax) // 2])      # a candidate in the middle of the pack
    seeds_d = torch.from_numpy(cyc.seeds).cuda()
    flips = set()
    for eps in (1e-10, -1e-10, 1e-8, -1e-8, 9e-7, -9e-7, 1.1e-6, -1.1e-6, 1e-5, -1e-5, 1e-3, -1e-3):
        cfg = H.oracle_config(cyc.ego_length, cyc.ego_width, max_curvature=pivot * (1.0 + eps))
        with np.errstate(all="ignore"):
            ref = O.plan_cycle(cfg, spline, fr, empty, 0, 100, seeds=cyc.seeds, targets=[0.0])
        eng = H.engine_for(J, cfg, max_seeds=48, max_obstacle_modes=1, max_total_steps=cyc.obstacles.state.shape[2])
        eng.set_refline(spline.s_list, spline.coefficient_table())
        o = cyc.obstacles
        eng.set_obstacles(o.state, o.valid, o.size, o.prob, o.n_dict)
        for kern in ("warp", "thread", "group", "tiled"):
            res = eng.evaluate(fr, 0, 100, seeds=seeds_d, targets=[0.0], kernel=kern)
            assert np.array_equal(res.feasible.cpu().numpy().astype(bool), ref["feasible"]), (eps, kern)
        flips.add(int(ref["feasible"].sum()))
        eng.close()
    assert len(flips) >= 2                                    # the pivot candidate changes side across the sweep
