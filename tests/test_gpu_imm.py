"""Secondary kernels (builder-specified, parity unpinned - the reference has no predictor code): batched IMM lane-intention
update and per-mode roll-out against their NumPy specification, and the predicted modes feeding the collision check."""
import numpy as np
import pytest
import torch

import dlii_jssp_b200 as J
from oracle import imm_oracle, jssp_oracle as O
import jssp_test_helpers as H

pytestmark = pytest.mark.gpu


def test_imm_update_matches_spec_and_converges():
    rng = np.random.default_rng(0)
    Ob, M = 37, 3
    mean = rng.normal(0, 0.5, (Ob, M, 2)); cov = np.tile(np.array([0.5, 0.0, 0.5]), (Ob, M, 1)); mu = np.full((Ob, M), 1.0 / M)
    eng = J.JsspEngine(J.JsspConfig())
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    mean_d, cov_d, mu_d = d(mean), d(cov), d(mu)
    lane_centres = np.array([-3.5, 0.0, 3.5])
    y = rng.uniform(-4, 4, Ob)
    target = rng.integers(0, M, Ob)
    for step in range(40):
        y = y + 0.15 * (lane_centres[target] - y) + rng.normal(0, 0.02, Ob)        # the vehicle converges to its target lane
        z = y[:, None] - lane_centres[None, :]
        mean, cov, mu = imm_oracle.imm_update(mean, cov, mu, z)
        eng.imm_update(mean_d, cov_d, mu_d, d(z))
        np.testing.assert_allclose(mu_d.cpu().numpy(), mu, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(mean_d.cpu().numpy(), mean, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(cov_d.cpu().numpy(), cov, rtol=1e-9, atol=1e-12)
    assert np.allclose(mu.sum(1), 1.0)
    assert (np.argmax(mu, 1) == target).mean() > 0.9          # the intended lane wins
    eng.close()


def test_predicted_modes_feed_the_collision_check():
    """IMM probabilities + per-mode roll-outs (device tensors) -> jssp_set_obstacles_dev -> same masks as the oracle."""
    cyc = J.scenario.dense_cycle(n_seeds=256, n_widths=3, O=4, M=1, seed=3)
    rng = np.random.default_rng(5)
    Ob, M, Tt = 6, 3, 60
    lanes = np.stack([np.stack([np.linspace(-40, 40, 81), np.full(81, yv)], 1) for yv in (-16.5, -13.0, -9.5)] +
                     [np.stack([np.full(81, xv), np.linspace(-40, 40, 81)], 1) for xv in (11.5, 18.5)])
    lane_of = np.stack([rng.permutation(5)[:M] for _ in range(Ob)]).astype(np.int32)
    s_start = rng.uniform(0, 40, (Ob, M)); speed = rng.uniform(1, 9, Ob)
    size = np.stack([rng.uniform(3.6, 5.4, Ob), rng.uniform(1.7, 2.0, Ob)], 1)
    prob = rng.dirichlet(np.ones(M), Ob)
    cfg = H.oracle_config(cyc.ego_length, cyc.ego_width, p_min=0.2, w_risk=1.0)
    eng = H.engine_for(J, cfg, max_seeds=256, max_obstacle_modes=Ob * M, max_total_steps=Tt)
    spline = O.CubicSpline2D(cyc.route_xy[:, 0], cyc.route_xy[:, 1])
    eng.set_refline(spline.s_list, spline.coefficient_table())
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    state_d, valid_d = eng.predict_modes(d(lanes), d(lane_of), d(s_start), d(speed), 0.1, Tt)
    state, valid = imm_oracle.predict_modes(lanes, lane_of, s_start, speed, 0.1, Tt)
    assert np.array_equal(valid_d.cpu().numpy().astype(bool), valid)
    np.testing.assert_allclose(state_d.cpu().numpy()[valid], state[valid], rtol=1e-9, atol=1e-9)
    eng.set_obstacles_dev(state_d, valid_d, d(size), d(prob))
    res = eng.evaluate(cyc.frenet0, 0, 55, seeds=d(cyc.seeds), targets=cyc.targets)
    ref = O.plan_cycle(cfg, spline, cyc.frenet0, (state_d.cpu().numpy(), valid, size, prob), 0, 55, seeds=cyc.seeds, targets=cyc.targets)
    H.compare_cycle(res, ref)
    assert ref["collision"].any() and not ref["collision"].all() and (ref["risk"] > 0).any()
    eng.close()


def test_batched_frenet_projection_matches_oracle(scenario_fixture):
    """jssp_project_states against the oracle's projection (oracle/frenet_oracle.py: scalar loops over the oracle's own spline
    sampling) on the bundled route and on a synthetic turn: states scattered around the reference line, all headings.  The
    product's host formulation (FrenetState.from_state, used for cycle 0 of a replay) is held to the same oracle."""
    from dlii_jssp_b200 import scenario as S
    from oracle import frenet_oracle as FO
    rng = np.random.default_rng(4)
    for route in (scenario_fixture.route_xy, S.synthetic_intersection(0).route_xy):
        sp = J.CubicSpline2D(route[:, 0], route[:, 1])
        pts = FO.rediscretise(O.CubicSpline2D(route[:, 0], route[:, 1]))
        ref = sp.sample(np.arange(0, sp.s_list[-1], 0.1))
        eng = J.JsspEngine(J.JsspConfig(max_knots=512))
        eng.set_refline(sp.s_list, sp.coefficient_table())
        k = rng.integers(0, len(ref), 64)
        st = np.column_stack([ref[k, 0] + rng.normal(0, 1.5, 64), ref[k, 1] + rng.normal(0, 1.5, 64), rng.uniform(-np.pi, np.pi, 64),
                              rng.uniform(0, 15, 64), rng.normal(0, 2, 64)])
        st[0] = (ref[0, 0], ref[0, 1], ref[0, 2], 5.0, 0.0)                       # exactly on the first sample
        got = eng.project_states(torch.from_numpy(st).cuda()).cpu().numpy()
        want = np.array([FO.project_state(pts, *row) for row in st])
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)
        host = np.array([J.FrenetState().from_state(J.State(0.0, *row), ref).as_tuple() for row in st])
        np.testing.assert_allclose(host, want, rtol=1e-9, atol=1e-9)
        eng.close()
