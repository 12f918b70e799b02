"""GPU tests of the batched, device-resident replay (jssp_batch_*): every scenario of a batch must reproduce, cycle by cycle,
what the one-scenario-at-a-time replay through the reference-shaped planner interface produces (and hence the oracle's
replay), with the hand-off, the clock and the end-of-scenario status evaluated on the device."""
import numpy as np
import pytest

import dlii_jssp_b200 as J
from dlii_jssp_b200 import batch as B, replay as R, scenario as S
from oracle import replay_oracle

pytestmark = pytest.mark.gpu


def _same(a, b, exact=True):
    assert a.best_idx == b.best_idx, (a.name, a.best_idx[:12], b.best_idx[:12])
    assert a.n_candidates == b.n_candidates
    assert a.end == b.end and a.cycles == b.cycles, (a.name, a.end, b.end, a.cycles, b.cycles)
    if exact:
        assert a.ego_rows == b.ego_rows
    elif b.ego_rows:
        np.testing.assert_allclose(np.array(a.ego_rows), np.array(b.ego_rows), rtol=1e-9, atol=1e-9)


def test_batch_matches_single_scenario_replays(scenario_fixture):
    """Bundled scenario + synthetic left / straight / right intersections in ONE batch, full length."""
    scs = [scenario_fixture] + [S.synthetic_intersection(i) for i in (0, 1, 2, 7)]
    got = B.run_batch(scs)
    for sc, g in zip(scs, got):
        _same(g, R.run_replay(sc))
    assert any(g.end in (1, 1.5, 2) for g in got)           # at least one scenario ran to a real end status


@pytest.mark.parametrize("kernel", [None, "persistent", "bestfirst", "warp", "thread", "group", "tiled", "chunked"])
def test_batch_kernels_agree_with_oracle_replay(kernel):
    scs = [S.synthetic_intersection(i) for i in (3, 4)]
    got = B.run_batch(scs, max_cycles=20, kernel=kernel)
    for sc, g in zip(scs, got):
        p = B.pack_scenario(sc)
        cpu = replay_oracle.replay(sc, tuple(p.frenet0), S.pack_obstacle_dict, R.end_status, R.goal_pose_of(sc), max_cycles=20)
        assert g.best_idx == cpu["best_idx"] and g.n_candidates == cpu["n"] and g.end == cpu["end"]
        np.testing.assert_allclose(np.array(g.ego_rows), np.array(cpu["rows"]), rtol=1e-9, atol=1e-9)


def test_batch_larger_than_capacity_is_chunked_and_perturbed_scenes(scenario_fixture):
    """BASELINE config 2 shape: the bundled scene + seeded variants, replayed through a 3-slot batch."""
    scs = [S.perturbed_scenario(scenario_fixture, s) for s in range(5)]
    cfg = J.JsspConfig(ego_length=scenario_fixture.ego["length"], ego_width=scenario_fixture.ego["width"], max_seeds=1024, max_knots=256,
                       max_obstacle_modes=8, max_total_steps=256)
    bt = B.BatchReplay(cfg, max_scenarios=3, max_cycles=16)
    got = B.run_batch(scs, max_cycles=16, batch=bt)
    for sc, g in zip(scs, got):
        _same(g, R.run_replay(sc, max_cycles=16))
    assert bt.launch_count >= 2 * 16 * 2
    tr = bt.best_traj(0)
    assert tr.shape == (51, 16) and np.isfinite(tr[:, :9]).all()
    bt.close()


def test_batch_resume_continues_where_it_stopped():
    """run(k) twice == run(2k): the replay state lives on the device between calls."""
    sc = S.synthetic_intersection(5)
    p = B.pack_scenario(sc, max_cycles=12)
    cfg = J.JsspConfig(max_seeds=1024, max_knots=512, max_obstacle_modes=16, max_total_steps=256)
    bt = B.BatchReplay(cfg, max_scenarios=1, max_cycles=12)
    bt.load([p]); bt.run(12)
    whole = bt.results()[0]
    bt.load([p]); bt.run(5); bt.run(7)
    _same(bt.results()[0], whole)
    # reset = back to the freshly loaded state without an upload; array-valued results are views of the same records
    bt.reset(); bt.run(12)
    arr = bt.results(lists=False)[0]
    assert arr.best_idx.tolist() == whole.best_idx and arr.n_candidates.tolist() == whole.n_candidates
    assert [tuple(r) for r in arr.ego_rows.tolist()] == whole.ego_rows and arr.end == whole.end
    bt.close()


def test_batch_scenario_without_obstacles_and_finished_scenario():
    sc = S.synthetic_intersection(6)
    empty = S.ReplayScenario("empty", dict(sc.ego), {}, sc.goal, sc.dt, sc.max_t, sc.route_xy)
    done = S.ReplayScenario("done", dict(sc.ego), sc.vehicle_traj, sc.goal, sc.dt, 0.0, sc.route_xy)     # t >= max_t before the first cycle
    got = B.run_batch([empty, done, sc], max_cycles=10)
    _same(got[0], R.run_replay(empty, max_cycles=10))
    assert got[1].cycles == 0 and got[1].end == 1
    _same(got[2], R.run_replay(sc, max_cycles=10))


def test_c2_multimodal_predictions_batch_matches_oracle(scenario_fixture):
    """BASELINE config 2 shape: scenes with M = 3 predicted intention modes per vehicle (the planner sees the predictions,
    the end-of-scenario test the ground truth); p_min / w_risk semantics as in the single-cycle tests."""
    scs = [S.perturbed_scenario(scenario_fixture, s) for s in (0, 2, 3)]
    preds = [S.multimodal_predictions(sc, M=3, seed=i) for i, sc in enumerate(scs)]
    packed = [B.pack_scenario(sc, max_cycles=18, predictions=pr) for sc, pr in zip(scs, preds)]
    cfg = J.JsspConfig(ego_length=scenario_fixture.ego["length"], ego_width=scenario_fixture.ego["width"], p_min=0.3, w_risk=1.0,
                       max_seeds=1024, max_knots=256, max_obstacle_modes=32, max_total_steps=256)
    bt = B.BatchReplay(cfg, max_scenarios=4, max_cycles=18)
    got = B.run_batch(packed, max_cycles=18, batch=bt)
    differs = False
    for sc, pr, p, g in zip(scs, preds, packed, got):
        cpu = replay_oracle.replay(sc, tuple(p.frenet0), lambda *_a, _pr=pr: _pr, R.end_status, R.goal_pose_of(sc), max_cycles=18,
                                   cfg_overrides=dict(p_min=0.3, w_risk=1.0))
        assert g.best_idx == cpu["best_idx"] and g.n_candidates == cpu["n"] and g.end == cpu["end"], (g.best_idx, cpu["best_idx"])
        np.testing.assert_allclose(np.array(g.ego_rows), np.array(cpu["rows"]), rtol=1e-9, atol=1e-9)
        differs |= g.best_idx != R.run_replay(sc, max_cycles=18).best_idx
    assert differs                   # the predicted modes change at least one selection w.r.t. the ground-truth-only replay
    # best-first selection with soft hits in play (modes below p_min add w_risk * prob to a survivor's cost: the bound leaves the
    # risk term out, the exact cost adds it in the serial walk's order): the same replay
    for a, b in zip(got, B.run_batch(packed, max_cycles=18, batch=bt, kernel="bestfirst")):
        _same(a, b)
    bt.close()


def test_batch_kinematics_mode_matches_host_loop_and_reference_golden(scenario_fixture, golden_dir):
    """ego_update_mode "kinematics" in the batched device-resident replay (kinematic ego, pure pursuit on the selected path and
    the Frenet projection of the new ego all in the hand-off): same selections as the host-driven loop and as the reference's own
    complete loop in that mode (kin_* of replay_golden_8_3_4_565.npz), ego rows to 1e-8 (device vs libm trigonometry)."""
    import os
    g = np.load(os.path.join(golden_dir, "replay_golden_8_3_4_565.npz"))
    scs = [scenario_fixture, S.perturbed_scenario(scenario_fixture, 3)]
    got = B.run_batch(scs, ego_update_mode="kinematics")
    assert got[0].best_idx == g["kin_best_idx"].tolist()
    assert got[0].end == g["kin_ends"][-1] and got[0].cycles == len(g["kin_rows"])
    np.testing.assert_allclose(np.array(got[0].ego_rows), g["kin_rows"], rtol=1e-8, atol=1e-8)
    for sc, r in zip(scs, got):
        solo = R.run_replay(sc, ego_update_mode="kinematics")
        assert r.best_idx == solo.best_idx and r.n_candidates == solo.n_candidates and r.end == solo.end
        np.testing.assert_allclose(np.array(r.ego_rows), np.array(solo.ego_rows), rtol=1e-9, atol=1e-9)
    # the default mode is untouched by the packed polyline
    a = B.run_batch([B.pack_scenario(scs[0], kinematics=True)], max_cycles=12)[0]
    b = B.run_batch([scs[0]], max_cycles=12)[0]
    assert a.best_idx == b.best_idx and a.ego_rows == b.ego_rows


def test_sub_batches_on_separate_streams_replay_identically(scenario_fixture):
    """MultiStreamReplay (k sub-batches on k CUDA streams, scenarios dealt round-robin) returns, in the caller's order, exactly
    what one BatchReplay returns - also when there are fewer scenarios than sub-batches."""
    scs = [scenario_fixture] + [S.synthetic_intersection(i) for i in (2, 5, 9, 11)]
    packed = [B.pack_scenario(sc, max_cycles=20) for sc in scs]
    cfg = J.JsspConfig(ego_length=float(packed[0].ego[5]), ego_width=float(packed[0].ego[6]), max_seeds=1024,
                       max_knots=max(64, max(len(p.knot_s) for p in packed)), max_obstacle_modes=max(p.state.shape[0] for p in packed),
                       max_total_steps=max(p.state.shape[2] for p in packed))
    one = B.BatchReplay(cfg, max_scenarios=len(packed), max_cycles=20)
    one.load(packed); one.run(20)
    want = one.results()
    for k in (2, 3):
        ms = B.MultiStreamReplay(cfg, max_scenarios=len(packed), max_cycles=20, k=k)
        ms.load(packed); ms.run(20)
        for a, b in zip(want, ms.results()):
            assert a.name == b.name
            _same(b, a)
        ms.reset(); ms.run(20)
        for a, b in zip(want, ms.results(lists=False)):
            assert b.best_idx.tolist() == a.best_idx
        ms.load(packed[:1]); ms.run(20)                      # fewer scenarios than sub-batches
        _same(ms.results()[0], want[0])
        # upload and replay sub-batch by sub-batch (the replay of one overlaps the upload of the next), every pipeline
        for kernel in (None, "bestfirst", "auto"):
            ms.load_and_run(packed, 20, kernel=kernel)
            for a, b in zip(want, ms.results()):
                assert a.name == b.name
                _same(b, a)
        ms.close()
    one.close()


@pytest.mark.parametrize("kernel", ["persistent", "bestfirst"])
def test_persistent_and_lockstep_replays_are_identical(scenario_fixture, kernel):
    """("bestfirst" = the persistent kernel with best-first selection: candidates checked in the order of a lower bound of their cost.)
    The persistent replay kernel (one CTA per scenario loops over its plan cycles in one launch) and the lock-step pipeline
    (two launches per cycle for the whole batch; the default) produce the same records, cycle by cycle, for a mixed batch: bundled
    scene, synthetic intersections with 3..12 vehicles, an obstacle-free scenario - with 256 and (few scenarios) 512 threads."""
    scs = [scenario_fixture] + [S.synthetic_intersection(i) for i in range(20, 31)]
    empty = S.synthetic_intersection(9)
    empty.vehicle_traj = {}
    scs.append(empty)
    a = B.run_batch(scs, max_cycles=60, kernel=kernel)
    b = B.run_batch(scs, max_cycles=60)
    for x, y in zip(a, b):
        _same(x, y)
    many = [S.synthetic_intersection(i) for i in range(100, 260)]         # > 148 scenarios: 256-thread CTAs, several per SM
    for x, y in zip(B.run_batch(many, max_cycles=25, kernel=kernel), B.run_batch(many, max_cycles=25)):
        _same(x, y)
