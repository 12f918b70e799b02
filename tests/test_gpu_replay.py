"""GPU replay tests through the reference-shaped planner interface (IntersectionPlanner.plan -> JerkSpaceSamplingPlanner.plan_traj):
the CUDA planner's replay against the oracle's, cycle by cycle."""
import numpy as np
import pytest

import dlii_jssp_b200 as J
from dlii_jssp_b200 import replay as R, scenario as S
from oracle import replay_oracle

pytestmark = pytest.mark.gpu


def _both(sc, max_cycles=None):
    gpu = R.run_replay(sc, max_cycles=max_cycles)
    # initial Frenet state exactly as the product façade computes it (host logic, FrenetState.from_state)
    obs = sc.observation()
    sp = J.CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1])
    s = np.arange(0, sp.s_list[-1], 0.1)
    ref = np.column_stack(([sp.calc_position(v) for v in s], [sp.calc_yaw(v) for v in s], [sp.calc_curvature(v) for v in s]))
    e = obs["vehicle_info"]["ego"]
    fr0 = J.FrenetState().from_state(J.State(0.0, e["x"], e["y"], e["yaw"], e["v"], e["a"]), ref).as_tuple()
    cpu = replay_oracle.replay(sc, fr0, S.pack_obstacle_dict, R.end_status, R.goal_pose_of(sc), max_cycles=max_cycles)
    return gpu, cpu


def _assert_same(gpu, cpu):
    assert gpu.best_idx == cpu["best_idx"], (gpu.best_idx[:20], cpu["best_idx"][:20])
    assert gpu.n_candidates == cpu["n"]
    assert gpu.end == cpu["end"]
    if cpu["rows"]:
        np.testing.assert_allclose(np.array(gpu.ego_rows), np.array(cpu["rows"]), rtol=1e-9, atol=1e-9)


def test_c1_bundled_scenario_replay_matches_oracle(scenario_fixture):
    """BASELINE config 1: inputs/8_3_4_565, default sampling grid, full replay."""
    gpu, cpu = _both(scenario_fixture)
    assert gpu.cycles >= 1
    _assert_same(gpu, cpu)


@pytest.mark.parametrize("seed", [1, 4])
def test_c2_perturbed_scene_replay_matches_oracle(scenario_fixture, seed):
    """BASELINE config 2 (the reference's 8 demo scenes are 0-byte placeholders): seeded variants of 8_3_4_565."""
    gpu, cpu = _both(S.perturbed_scenario(scenario_fixture, seed), max_cycles=25)
    _assert_same(gpu, cpu)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_c4_synthetic_intersection_replay_matches_oracle(seed):
    """BASELINE config 4 (scenes_all(800) is absent): seeded synthetic intersections - left turn, straight, right turn."""
    gpu, cpu = _both(S.synthetic_intersection(seed), max_cycles=30)
    assert gpu.cycles >= 5
    _assert_same(gpu, cpu)


def test_sweep_reuses_planner_and_matches_individual_replays():
    scs = [S.synthetic_intersection(s) for s in (3, 4, 5)]
    sweep = R.run_sweep(scs, max_cycles=8)
    solo = [R.run_replay(sc, max_cycles=8) for sc in scs]
    for a, b in zip(sweep, solo):
        assert a.best_idx == b.best_idx and a.ego_rows == b.ego_rows and a.end == b.end


def test_c1_replay_matches_the_reference_environment_golden(scenario_fixture, golden_dir):
    """Full replay of inputs/8_3_4_565 - one scenario at a time and through the batched device-resident replay - against
    tests/golden/replay_golden_8_3_4_565.npz, recorded from the reference's own Env.step / ReplayController.step driven by the
    oracle planner (oracle/make_golden_replay.py): selected index per cycle, ego rows, end status."""
    import os
    from dlii_jssp_b200 import batch as B
    g = np.load(os.path.join(golden_dir, "replay_golden_8_3_4_565.npz"))
    for res in (R.run_replay(scenario_fixture), B.run_batch([scenario_fixture])[0]):
        assert res.best_idx == g["best_idx"].tolist() == g["full_best_idx"].tolist()          # full_*: the reference's own planner in the loop
        assert res.end == g["ends"][-1] == g["full_ends"][-1] and res.cycles == len(g["rows"])
        np.testing.assert_allclose(np.array(res.ego_rows), g["full_rows"], rtol=1e-9, atol=1e-9)


def test_c1_kinematics_mode_replay_matches_the_reference_loop(scenario_fixture, golden_dir):
    """parameters["sim_config"]["ego_update_mode"] = "kinematics": every cycle starts from the projected ego state, the ego follows
    the kinematic model driven by the planned acceleration and the pure-pursuit steering angle.  Against kin_* of
    replay_golden_8_3_4_565.npz = the reference's own complete loop in that mode (make_golden_replay.py --kinematics)."""
    import os
    g = np.load(os.path.join(golden_dir, "replay_golden_8_3_4_565.npz"))
    res = R.run_replay(scenario_fixture, ego_update_mode="kinematics")
    assert res.best_idx == g["kin_best_idx"].tolist()
    assert res.end == g["kin_ends"][-1] and res.cycles == len(g["kin_rows"])
    np.testing.assert_allclose(np.array(res.ego_rows), g["kin_rows"], rtol=1e-8, atol=1e-8)


def _sweep_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sweep_golden.npz"))


def test_sweep_sample_matches_oracle_replay_golden():
    """A seeded sample of 40 of the 800 scenarios bench.py sweeps (BASELINE configs[3]), <= 100 cycles each, replayed by the
    batched device-resident replay against the oracle planner in the reference driver's loop (tests/golden/sweep_golden.npz,
    oracle/make_golden_sweep.py): end status, cycle count, selected index and candidate count of EVERY cycle identical, ego rows
    to 1e-9."""
    from dlii_jssp_b200 import batch as B
    g = _sweep_golden()
    ids, C = g["ids"].tolist(), int(g["max_cycles"][0])
    scs = [S.synthetic_intersection(i) for i in ids]
    got = B.run_batch(scs, max_cycles=C)
    assert len(got) >= 32
    for kernel in ("persistent", "bestfirst"):       # the persistent replay kernel, exhaustive and with best-first selection: the same rows
        for a, b in zip(got, B.run_batch(scs, max_cycles=C, kernel=kernel)):
            assert (a.end, a.cycles, list(a.best_idx), list(a.n_candidates), a.ego_rows) == (b.end, b.cycles, list(b.best_idx), list(b.n_candidates), b.ego_rows), kernel
    for k, r in enumerate(got):
        c = int(g["cycles"][k])
        assert (r.end, r.cycles) == (float(g["end"][k]), c), (ids[k], r.end, r.cycles, g["end"][k], c)
        assert list(r.best_idx) == g["best_idx"][k, :c].tolist(), ids[k]
        assert list(r.n_candidates) == g["n_candidates"][k, :c].tolist(), ids[k]
        nr = int(g["n_rows"][k])
        assert len(r.ego_rows) == nr
        if nr:
            np.testing.assert_allclose(np.array(r.ego_rows), g["rows"][k, :nr], rtol=1e-9, atol=1e-9)
    assert not any(r.seed_overflow for r in got)


def test_sharded_sweep_returns_the_same_rows_as_one_gpu():
    """The scenario sweep shards over ranks with no collective: dealing the sample to 8 ranks'
    shards and replaying shard by shard gives, per scenario, exactly the rows of the single-batch run - and the digest bench.py
    prints for every --gpus N is therefore the same."""
    from dlii_jssp_b200 import batch as B, distributed as D
    ids = _sweep_golden()["ids"].tolist()
    packed = [B.pack_scenario(S.synthetic_intersection(i), max_cycles=100) for i in ids]
    whole = {r.name: r for r in B.run_batch(packed, max_cycles=100)}
    parts = {}
    for rank in range(8):
        for r in B.run_batch(D.shard_scenarios(packed, 8, rank), max_cycles=100):
            parts[r.name] = r
    assert parts.keys() == whole.keys()
    for name, r in whole.items():
        q = parts[name]
        assert (r.end, r.cycles, list(r.best_idx), list(r.n_candidates), r.ego_rows) == (q.end, q.cycles, list(q.best_idx), list(q.n_candidates), q.ego_rows)
    assert D.sweep_digest([(r.name, r.end, r.cycles, list(r.best_idx)) for r in whole.values()]) == \
        D.sweep_digest([(r.name, r.end, r.cycles, list(r.best_idx)) for r in parts.values()])
