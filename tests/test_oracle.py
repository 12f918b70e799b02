"""CPU tests of the oracle itself: pinned against golden vectors produced by RUNNING the reference's own
polyvt_sampling.py (oracle/make_golden_polyvt.py), plus known-answer and property tests for the builder-specified pieces."""
import math
import warnings

import numpy as np
import pytest

import os

from oracle import jssp_oracle as O
from oracle import jssp_loop as L

warnings.filterwarnings("ignore", category=RuntimeWarning)

# seed counts measured on the reference file: (v0, a0, J) -> (main, stop)
REFERENCE_COUNTS = {(5.686, 0.0, 10.0): (175, 142), (0.0, 0.0, 10.0): (103, 625), (10.0, 1.0, 10.0): (112, 60), (17.0, 0.0, 10.0): (115, 13),
                    (3.0, -2.0, 10.0): (42, 307), (0.0, 0.0, 13.4112): (97, 625), (10.0, 1.0, 13.4112): (95, 41), (3.0, -2.0, 13.4112): (37, 138)}


def test_seeds_and_unroll_bit_identical_to_reference(polyvt_golden):
    g = polyvt_golden
    for k, (s0, v0, a0, J, T, ns) in enumerate(g["meta"]):
        gr = O.make_grids(T, J, int(ns))
        assert np.array_equal(gr["time_array"], g[f"c{k}_time"])
        main, stop = O.sample_main_seeds(s0, v0, a0, gr), O.sample_stop_seeds(s0, v0, a0, gr)
        assert np.array_equal(main, g[f"c{k}_main"]), k
        assert np.array_equal(stop, g[f"c{k}_stop"]), k
        seeds = np.concatenate([main, stop], 0)
        s, sd, sdd, sddd = O.unroll_seeds(seeds[g[f"c{k}_pick"]], s0, v0, a0, gr["time_array"])
        for a, b in zip((s, sd, sdd, sddd), g[f"c{k}_unroll"].transpose(1, 0, 2)):
            assert np.array_equal(a, b), k
        key = (round(float(v0), 3), float(a0), float(J))
        if T == 5.0 and key in REFERENCE_COUNTS:
            assert (len(main), len(stop)) == REFERENCE_COUNTS[key]


def test_obb_known_answers():
    hit = lambda *a: bool(O.obb_intersects(*[np.float64(v) for v in a]))
    # unit squares (half extents 0.5): overlapping, exactly touching edge (closed sets -> intersect), separated
    assert hit(0, 0, 1, 0, .5, .5, 0.9, 0, 1, 0, .5, .5)
    assert hit(0, 0, 1, 0, .5, .5, 1.0, 0, 1, 0, .5, .5)
    assert not hit(0, 0, 1, 0, .5, .5, 1.0 + 1e-12, 0, 1, 0, .5, .5)
    # touching corner to corner
    assert hit(0, 0, 1, 0, .5, .5, 1.0, 1.0, 1, 0, .5, .5)
    # 45 degree square: corner reaches sqrt(0.5) along x
    c = math.sqrt(0.5)
    assert hit(0, 0, 1, 0, .5, .5, 0.5 + c - 1e-9, 0, c, c, .5, .5)
    assert not hit(0, 0, 1, 0, .5, .5, 0.5 + c + 1e-9, 0, c, c, .5, .5)
    # a long box crossing another without any corner inside (needs the second box's axes)
    assert hit(0, 0, 1, 0, 5, .2, 0, 0, 0, 1, 5, .2)
    # parallel offset boxes separated only along the ego's lateral axis
    assert not hit(0, 0, 1, 0, 2.5, 1.0, 0, 2.2, 1, 0, 2.5, 1.0)


def test_spline_and_quintic_known_answers():
    th = np.linspace(0, np.pi / 2, 60)
    sp = O.CubicSpline2D(20 * np.cos(th), 20 * np.sin(th))            # circle arc, R = 20
    s = np.linspace(2.0, sp.s_list[-1] - 2.0, 50)
    assert np.allclose(sp.calc_curvature(s), 1 / 20, rtol=2e-3)
    x, y = sp.calc_position(sp.s_list)
    assert np.allclose(x, 20 * np.cos(th), atol=1e-12) and np.allclose(y, 20 * np.sin(th), atol=1e-12)   # interpolates the knots
    assert sp.calc_position(-1e-9) == (None, None) and sp.calc_position(sp.s_list[-1] + 1e-9) == (None, None)
    assert sp.calc_position(sp.s_list[-1])[0] is not None                                           # end point is on the line
    q = O.QuinticPolynomial(0.2, -0.1, 0.05, 0.75, 0.0, 0.0, 5.0)
    assert np.isclose(q.calc_point(0), 0.2) and np.isclose(q.calc_first_derivative(0), -0.1) and np.isclose(q.calc_second_derivative(0), 0.05)
    assert np.isclose(q.calc_point(5.0), 0.75) and abs(q.calc_first_derivative(5.0)) < 1e-12 and abs(q.calc_second_derivative(5.0)) < 1e-12


def _cycle(n_seeds=96, n_widths=3, O_=8, M=2, seed=5, **cfg_kw):
    from dlii_jssp_b200 import scenario as S
    cyc = S.dense_cycle(n_seeds=n_seeds, n_widths=n_widths, O=O_, M=M, seed=seed)
    cfg = O.OracleConfig(ego_length=cyc.ego_length, ego_width=cyc.ego_width, **cfg_kw)
    spline = O.CubicSpline2D(cyc.route_xy[:, 0], cyc.route_xy[:, 1])
    o = cyc.obstacles
    return cyc, cfg, spline, (o.state, o.valid.astype(bool), o.size, o.prob)


def test_loop_form_agrees_with_vectorised_oracle():
    cyc, cfg, spline, obs = _cycle(p_min=0.25, w_risk=1.5)
    a = O.plan_cycle(cfg, spline, cyc.frenet0, obs, 0, cyc.final_time_step, seeds=cyc.seeds, targets=cyc.targets)
    b = L.plan_cycle_loop(cfg, spline, cyc.frenet0, obs, 0, cyc.final_time_step, seeds=cyc.seeds, targets=cyc.targets)
    assert np.array_equal(a["feasible"], b["feasible"]) and np.array_equal(a["collision"], b["collision"])
    np.testing.assert_allclose(a["cost"], b["cost"], rtol=1e-12)
    assert a["best"] == b["best"]


@pytest.mark.parametrize("kw", [dict(), dict(p_min=0.25, w_risk=1.5), dict(O_=24, M=3, seed=9, p_min=0.4, w_risk=0.7)])
def test_best_first_order_selects_the_exhaustive_argmin(kw):
    """The rule behind the replay kernel's best-first selection (csrc/jssp_bestfirst.cuh), stated on the oracle: the cost without the
    risk term needs no Frenet->Cartesian walk and never exceeds the cost (w_risk >= 0, risk >= 0); checking candidates in the order
    of that bound and stopping when the next bound exceeds the best surviving (cost, index) returns exactly the argmin over the
    survivors of the filter chain (jerk_space_sampling_planner.py:311-327) - without looking at the other candidates."""
    cfg_kw = {k: v for k, v in kw.items() if k in ("p_min", "w_risk")}
    cyc, cfg, spline, obs = _cycle(**{k: v for k, v in kw.items() if k not in cfg_kw}, **cfg_kw)
    with np.errstate(all="ignore"):
        r = O.plan_cycle(cfg, spline, cyc.frenet0, obs, 0, cyc.final_time_step, seeds=cyc.seeds, targets=cyc.targets)
        bound = O.cost_total(r["s"], r["s_d"], r["s_dd"], r["s_ddd"], r["d"], r["d_d"], r["d_dd"], r["d_ddd"], cfg, risk=None)
    fin = np.isfinite(r["cost"])
    assert (bound[fin] <= r["cost"][fin]).all()
    if kw.get("O_", 0) >= 24:
        assert (r["risk"] > 0).any()                       # soft hits are in play: some bounds are strictly below their costs
    ok = r["feasible"] & ~r["collision"] & fin
    order = np.lexsort((np.arange(r["n"]), bound))         # (bound, index)
    best, best_cost, checked = -1, np.inf, 0
    for n in order:
        if best >= 0 and bound[n] > best_cost:
            break                                          # no unchecked candidate can cost less
        checked += 1
        if ok[n] and (r["cost"][n] < best_cost or (r["cost"][n] == best_cost and n < best)):
            best, best_cost = int(n), float(r["cost"][n])
    assert best == r["best"]
    assert r["best"] < 0 or checked < r["n"]               # and it got there without checking everything


def test_reference_edge_semantics():
    """Truncation, 1-point / 0-point trajectories, stopped car (ds == 0 -> nan passes), reversing car (|c| = pi/ds fails)."""
    cyc, cfg, spline, obs = _cycle()
    s_end = spline.s_list[-1]
    stop = np.array([[0.0, 0.0, 0.0, 0.0, 0.0, 5.0, 0.0, 0.0, 0.0, 0.0]])       # all-zero jerk seed
    # stopped on the line with d(t) = const: every forward difference is exactly zero -> c = 0/0 = nan -> feasible
    r = O.plan_cycle(cfg, spline, (10.0, 0.0, 0.0, 0.0, 0.0, 0.0), obs, 0, 100, seeds=stop, targets=np.array([0.0]))
    assert r["n_pts"][0] == 51 and np.all(r["ds"][0] == 0) and np.all(np.isnan(r["c"][0])) and r["feasible"][0]
    assert np.all(r["yaw"][0] == 0.0)                                              # arctan2(0, 0) = 0
    # slowly reversing: heading flips by pi between the first forward and the first backward step
    rev = np.array([[-1.0, 1.0, 0.0, 0.0, 0.0, 4.0, 0.0, 0.0, 0.0, 0.0]])
    r = O.plan_cycle(cfg, spline, (10.0, 0.2, 0.0, 0.0, 0.0, 0.0), obs, 0, 100, seeds=rev, targets=np.array([0.0]))
    assert not r["feasible"][0] and np.nanmax(np.abs(r["c"][0])) > 10
    # start beyond the end of the line: 0 Cartesian points -> passes both checks vacuously
    r = O.plan_cycle(cfg, spline, (s_end + 1.0, 3.0, 0.0, 0.0, 0.0, 0.0), obs, 0, 100, seeds=stop, targets=np.array([0.0]))
    assert r["n_pts"][0] == 0 and r["feasible"][0] and not r["collision"][0]
    # exactly one point on the line -> the reference's bare except reports a collision
    one = np.array([[0.0, 1.0, 0.0, 1.0, 0.0, 3.0, 0.0, 0.0, 0.0, 0.0]])
    r = O.plan_cycle(cfg, spline, (s_end - 1e-3, 5.0, 0.0, 0.0, 0.0, 0.0), obs, 0, 100, seeds=one, targets=np.array([0.0]))
    assert r["n_pts"][0] == 1 and r["collision"][0]
    # ... unless there are no obstacles at all (len(obstacles) <= 0 early-out)
    empty = (np.zeros((0, 1, 60, 3)), np.zeros((0, 1, 60), bool), np.zeros((0, 2)), np.zeros((0, 1)))
    r = O.plan_cycle(cfg, spline, (s_end - 1e-3, 5.0, 0.0, 0.0, 0.0, 0.0), empty, 0, 100, seeds=one, targets=np.array([0.0]))
    assert not r["collision"][0]


def test_collision_properties():
    cyc, cfg, spline, obs = _cycle(n_seeds=128, O_=12, M=3, seed=9)
    state, valid, size, prob = obs
    base = O.plan_cycle(cfg, spline, cyc.frenet0, obs, 0, cyc.final_time_step, seeds=cyc.seeds, targets=cyc.targets)
    # monotone under obstacle removal
    fewer = O.plan_cycle(cfg, spline, cyc.frenet0, (state[:6], valid[:6], size[:6], prob[:6]), 0, cyc.final_time_step, seeds=cyc.seeds, targets=cyc.targets)
    assert not (fewer["collision"] & ~base["collision"]).any()
    # any-mode semantics: union of the per-mode single-trajectory checks
    per = [O.plan_cycle(cfg, spline, cyc.frenet0, (state[:, m:m + 1], valid[:, m:m + 1], size, np.ones((len(size), 1))), 0, cyc.final_time_step,
                        seeds=cyc.seeds, targets=cyc.targets)["collision"] for m in range(3)]
    assert np.array_equal(base["collision"], per[0] | per[1] | per[2])
    # argmin invariant under a permutation of the seeds (as a seed/width pair)
    perm = np.random.default_rng(0).permutation(len(cyc.seeds))
    shuf = O.plan_cycle(cfg, spline, cyc.frenet0, obs, 0, cyc.final_time_step, seeds=cyc.seeds[perm], targets=cyc.targets)
    if base["best"] >= 0:
        w0, k0 = divmod(base["best"], len(cyc.seeds)); w1, k1 = divmod(shuf["best"], len(cyc.seeds))
        assert w0 == w1 and np.array_equal(cyc.seeds[k0], cyc.seeds[perm][k1])
    # the end of the scenario cuts the collision window
    cut = O.plan_cycle(cfg, spline, cyc.frenet0, obs, 0, 4, seeds=cyc.seeds, targets=cyc.targets)
    assert not (cut["collision"] & ~base["collision"]).any()


def test_time_key_lookup_matches_reference_rounding(scenario_fixture):
    from dlii_jssp_b200 import scenario as S
    sc = scenario_fixture
    n_total = int(sc.max_t / sc.dt) + 60
    a = S.pack_obstacle_dict(sc.vehicle_traj, sc.dt, n_total)
    st, va, sz, pr = O.pack_obstacles(sc.vehicle_traj, sc.dt, n_total)
    assert np.array_equal(a.state, st) and np.array_equal(a.valid.astype(bool), va) and np.array_equal(a.size, sz)
    assert a.valid.sum() == sum(len([k for k in tr if k != "shape"]) for tr in sc.vehicle_traj.values()) == 271
    assert int(sc.max_t / sc.dt) == 121                       # float truncation of 12.2 / 0.1 (jerk_space_sampling_planner.py:244)


def test_fop_oracle_known_answers():
    """FOP baseline oracle: quartic boundary conditions, per-horizon sample counts, and the `>=` rule (last minimum wins)."""
    from oracle import fop_oracle as F
    q = F.QuarticPolynomial(3.0, 5.0, -1.0, 9.0, 0.0, 5.5)
    assert np.isclose(q.calc_point(0.0), 3.0) and np.isclose(q.calc_first_derivative(0.0), 5.0) and np.isclose(q.calc_second_derivative(0.0), -1.0)
    assert np.isclose(q.calc_first_derivative(5.5), 9.0) and abs(q.calc_second_derivative(5.5)) < 1e-12
    cfg = F.FopConfig(num_width=3, num_speed=4, num_t=3, min_t=5.0, max_t=5.5, ego_length=4.9, ego_width=1.9)
    widths, horizons, speeds, n = F.grids(cfg)
    assert list(n) == [len(np.arange(0.0, T, 0.1)) for T in horizons] and n[0] == 50 and n[-1] == 55
    route = np.column_stack([np.linspace(0, 200, 401), np.zeros(401)])
    spline = O.CubicSpline2D(route[:, 0], route[:, 1])
    empty = (np.zeros((0, 1, 10, 3)), np.zeros((0, 1, 10), bool), np.zeros((0, 2)), np.zeros((0, 1)))
    r = F.plan_cycle(cfg, spline, (5.0, 6.0, 0.0, 0.0, 0.0, 0.0), empty, 0, 100)
    assert r["n"] == 36 and not r["collision"].any()
    # d0 = 0 on a straight line: the outer lateral targets are mirror images -> equal costs; the LAST one must be selected
    ok = r["feasible"]
    best_cost = r["cost"][ok].min()
    ties = np.nonzero(ok & (r["cost"] == best_cost))[0]
    assert r["best"] == ties[-1]


def _golden_cycles(golden_dir):
    g = np.load(os.path.join(golden_dir, "jssp_cycle_golden.npz"))
    for k, label in enumerate(g["labels"]):
        yield str(label), tuple(g["frenet0"][k]), float(g["t"][k]), g


def _check_against_reference_cycle(res, label, g, rtol_cost):
    """res = dict with n, feasible, collision, cost, best, n_pts, and per-candidate arrays (oracle layout)."""
    n = int(g[f"{label}_n"])
    assert res["n"] == n
    feas = g[f"{label}_feasible"]
    assert np.array_equal(np.asarray(res["feasible"], bool), feas), label
    # the reference only collision-checks the candidates that passed the curvature check (jssp.py:310-311)
    coll_ref = ~g[f"{label}_collision_passed"][feas]
    assert np.array_equal(np.asarray(res["collision"], bool)[feas], coll_ref), label
    assert np.array_equal(np.asarray(res["n_pts"]), g[f"{label}_n_pts"]), label
    surv = ~np.isnan(g[f"{label}_cost"])
    assert np.array_equal(surv, feas & ~np.asarray(res["collision"], bool))
    np.testing.assert_allclose(np.asarray(res["cost"], np.float64)[surv], g[f"{label}_cost"][surv], rtol=rtol_cost)
    assert res["best"] == int(g[f"{label}_best"]), (label, res["best"], int(g[f"{label}_best"]))


def test_oracle_matches_the_reference_plan_cycle(golden_dir, scenario_fixture):
    """oracle.plan_cycle against tests/golden/jssp_cycle_golden.npz, produced by oracle/make_golden_cycle.py from the
    reference's own JerkSpaceSamplingPlanner.plan_traj (leaf classes stubbed): sampled count, curvature mask, collision mask,
    survivor costs, selected index, the selected trajectory and the Cartesian paths of a candidate subsample."""
    sc = scenario_fixture
    cfg = O.OracleConfig(ego_length=sc.ego["length"], ego_width=sc.ego["width"])
    spline = O.CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1])
    final = int(sc.max_t / sc.dt)
    state, valid, size = O.pack_obstacles(sc.vehicle_traj, sc.dt, final + cfg.n_steps + 2)[:3]
    prob = np.ones(valid.shape[:2])
    seen_no_solution = False
    for label, fr, t, g in _golden_cycles(golden_dir):
        with np.errstate(all="ignore"):
            r = O.plan_cycle(cfg, spline, fr, (state, valid.astype(bool), size, prob), int(t / sc.dt), final)
        _check_against_reference_cycle(r, label, g, rtol_cost=1e-12)
        b = r["best"]
        seen_no_solution |= b < 0
        if b >= 0:
            rows = g[f"{label}_best_rows"]
            for col, key in ((1, "s"), (2, "s_d"), (3, "s_dd"), (4, "s_ddd")):
                assert np.array_equal(r[key][b], rows[:, col]), key                      # the unroll is the reference's, bit for bit
            for col, key in ((5, "d"), (6, "d_d"), (7, "d_dd"), (8, "d_ddd")):
                np.testing.assert_allclose(r[key][b], rows[:, col], rtol=1e-12, atol=1e-14, err_msg=key)
            npts = int(r["n_pts"][b])
            for col, key in ((9, "x"), (10, "y"), (11, "yaw")):
                np.testing.assert_allclose(r[key][b][:npts], rows[:npts, col], rtol=1e-12, atol=1e-12, err_msg=key)
            np.testing.assert_allclose(r["c"][b][:npts - 1], rows[:npts - 1, 13], rtol=1e-9, atol=1e-12)
        pick = g[f"{label}_pick"]
        xy = g[f"{label}_pick_xyyawc"]
        for q, n in enumerate(pick):
            k = int(r["n_pts"][n])
            np.testing.assert_allclose(r["x"][n][:k], xy[q, :k, 0], rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(r["y"][n][:k], xy[q, :k, 1], rtol=1e-12, atol=1e-12)
            if k >= 2:
                np.testing.assert_allclose(r["yaw"][n][:k], xy[q, :k, 2], rtol=1e-12, atol=1e-12)
                with np.errstate(invalid="ignore"):
                    fin = np.isfinite(xy[q, :k - 1, 3])
                    np.testing.assert_allclose(r["c"][n][:k - 1][fin], xy[q, :k - 1, 3][fin], rtol=1e-9, atol=1e-12)
    assert seen_no_solution                      # the t = 2 s cycle has no feasible candidate ("No solution", jssp.py:359-360)


def _fop_golden_cfg(g, sc):
    from oracle import fop_oracle as F
    W, NV, NT = [int(v) for v in g["grid"]]
    return F.FopConfig(num_width=W, num_speed=NV, num_t=NT, min_t=5.0, max_t=5.5, highest_speed=18.0, max_accel=float(g["limits"][0]),
                       max_jerk=float(g["limits"][1]), ego_length=sc.ego["length"], ego_width=sc.ego["width"])


def test_fop_oracle_matches_the_reference_fop_methods(golden_dir, scenario_fixture):
    """oracle.fop_oracle.plan_cycle against tests/golden/fop_cycle_golden.npz = the reference's own FrenetOptimalPlanner methods
    (sampling_frenet_trajectories, calc_global_paths, check_constraints, check_collisions, `>=` selection) run by
    oracle/make_golden_fop.py with the leaf classes stubbed."""
    from oracle import fop_oracle as F
    sc = scenario_fixture
    g = np.load(os.path.join(golden_dir, "fop_cycle_golden.npz"))
    cfg = _fop_golden_cfg(g, sc)
    spline = O.CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1])
    final = int(sc.max_t / sc.dt)
    state, valid, size = O.pack_obstacles(sc.vehicle_traj, sc.dt, final + 57)[:3]
    prob = np.ones(valid.shape[:2])
    for k, label in enumerate(g["labels"]):
        label, fr, now = str(label), tuple(g["frenet0"][k]), int(float(g["t"][k]) / sc.dt)
        with np.errstate(all="ignore"):
            r = F.plan_cycle(cfg, spline, fr, (state, valid.astype(bool), size, prob), now, final)
        feas = g[f"{label}_feasible"]
        assert r["n"] == int(g[f"{label}_n"]) and np.array_equal(r["feasible"], feas)
        assert np.array_equal(r["collision"][feas], ~g[f"{label}_collision_passed"][feas])
        assert np.array_equal(r["n_pts"], g[f"{label}_n_pts"])
        assert np.array_equal(np.count_nonzero(~np.isnan(r["s"]), axis=1), g[f"{label}_n_own"])       # per-candidate sample counts
        np.testing.assert_allclose(r["cost"], g[f"{label}_cost"], rtol=1e-12)
        assert r["best"] == int(g[f"{label}_best"])
        b, rows = r["best"], g[f"{label}_best_rows"]
        own = len(rows)
        for col, key in ((1, "s"), (2, "s_d"), (3, "s_dd"), (4, "s_ddd"), (5, "d"), (6, "d_d"), (7, "d_dd"), (8, "d_ddd")):
            np.testing.assert_allclose(r[key][b][:own], rows[:, col], rtol=1e-12, atol=1e-13, err_msg=key)
        npts = int(r["n_pts"][b])
        for col, key in ((9, "x"), (10, "y"), (11, "yaw")):
            np.testing.assert_allclose(r[key][b][:npts], rows[:npts, col], rtol=1e-12, atol=1e-12, err_msg=key)
