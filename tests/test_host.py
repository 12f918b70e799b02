"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header declares, host-only entry
points agree with NumPy / the oracle, the reference-shaped containers behave, and nothing falls back to the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import dlii_jssp_b200 as J
from dlii_jssp_b200 import _lib, scenario as S
from oracle import jssp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "jssp_b200.h")).read()
    declared = set(re.findall(r"\b(jssp_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found in the header"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/jssp_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.jssp_abi_version() == 1
    assert b"no CPU fallback" in lib.jssp_strerror(-6)


def test_flag_constants_match_header():
    """Every JSSP_FLAG_* of include/jssp_b200.h has the same value in the ctypes binding, the flags are distinct bits, and the replay
    pipelines a caller can name map onto them."""
    header = open(os.path.join(ROOT, "include", "jssp_b200.h")).read()
    flags = {m[0]: int(m[1]) for m in re.findall(r"#define\s+JSSP_FLAG_([A-Z_]+)\s+(\d+)", header)}
    assert len(flags) >= 13
    for name, value in flags.items():
        assert getattr(_lib, "FLAG_" + name) == value, name
        assert value & (value - 1) == 0, name                    # one bit each
    assert len(set(flags.values())) == len(flags)
    import inspect
    from dlii_jssp_b200 import batch
    src = inspect.getsource(batch.BatchReplay.run)
    for kernel in ("lockstep", "persistent", "bestfirst", "tiled", "chunked", "auto"):
        assert f'"{kernel}"' in src
    assert batch.auto_pipeline(800) == "bestfirst" and batch.auto_pipeline(100) is None and batch.auto_pipeline(800, fop=True) is None


def test_struct_layouts_match_header():
    lib = _lib.load()
    cfg = _lib.JsspConfigC()
    lib.jssp_default_config(C.byref(cfg))
    d = J.JsspConfig()
    for f, _ in _lib.JsspConfigC._fields_:
        if f in ("reserved0",):
            continue
        v = getattr(cfg, f)
        if f == "cost_weights":
            assert list(v) == list(d.cost_weights)
        else:
            assert v == getattr(d, f), f          # ctypes layout == C layout, field by field
    for which, st in enumerate(_lib.STRUCTS):        # every ctypes declaration against the compiled header
        assert lib.jssp_struct_size(which) == C.sizeof(st), st.__name__
    assert lib.jssp_struct_size(len(_lib.STRUCTS)) == -1
    assert C.sizeof(_lib.JsspCycleIn) == 6 * 8 + 4 * 4 + 16 * 8 + 8 + 8
    assert C.sizeof(_lib.JsspCycleOut) == 5 * 8 + 4 * 4 + 8


@pytest.mark.parametrize("T,jerk,ns", [(5.0, 10.0, 25), (5.0, 13.4112, 25), (8.0, 10.0, 13), (3.0, 7.5, 9)])
def test_host_grids_bit_identical_to_numpy(T, jerk, ns):
    lib = _lib.load()
    n = (C.c_int32 * 9)()
    arr = [np.zeros(512) for _ in range(9)]
    assert lib.jssp_make_grids(T, jerk, ns, 512, *[a.ctypes.data for a in arr], C.byref(n)) == 0
    g = O.make_grids(T, jerk, ns)
    for k, name in enumerate(["time_array", "j1", "t1", "t2", "j3", "t3", "t4", "js", "ts"]):
        assert np.array_equal(arr[k][: n[k]], g[name]), name


def test_quintic_closed_form_matches_linear_solve():
    lib = _lib.load()
    rng = np.random.default_rng(3)
    for _ in range(50):
        xs, vs, a_s, xe = rng.normal(0, 1, 4)
        T = rng.uniform(2, 9)
        out = (C.c_double * 6)()
        lib.jssp_quintic_coeffs(xs, vs, a_s, xe, 0.0, 0.0, T, C.byref(out))
        q = O.QuinticPolynomial(xs, vs, a_s, xe, 0.0, 0.0, T)
        np.testing.assert_allclose(list(out), [q.a0, q.a1, q.a2, q.a3, q.a4, q.a5], rtol=1e-9, atol=1e-12)
        p = J.QuinticPolynomial(xs, vs, a_s, xe, 0.0, 0.0, T)
        assert np.allclose([p.a3, p.a4, p.a5], list(out)[3:], rtol=1e-13, atol=0)


def test_product_spline_matches_oracle_spline(scenario_fixture):
    xy = scenario_fixture.route_xy
    a, b = J.CubicSpline2D(xy[:, 0], xy[:, 1]), O.CubicSpline2D(xy[:, 0], xy[:, 1])
    assert np.array_equal(a.s_list, b.s_list)
    np.testing.assert_allclose(a.coefficient_table(), b.coefficient_table(), rtol=1e-9, atol=1e-11)   # Thomas vs dense solve
    knot, coef = J.spline_tables(b)                     # any PythonRobotics-shaped object can be uploaded
    assert np.array_equal(coef, b.coefficient_table()) and np.array_equal(knot, b.s_list)
    for s in (0.0, 3.3, a.s_list[-1]):
        pa, pb = a.calc_position(s), b.calc_position(s)
        assert np.allclose(pa, pb, atol=1e-9) and np.isclose(a.calc_yaw(s), float(b.calc_yaw(s)), atol=1e-9)
    assert a.calc_position(a.s_list[-1] + 1e-6) == (None, None)


def test_openscenario_parser_matches_reference_parser(scenario_fixture):
    """Only where the reference tree is mounted (the build container): our parser against the fixture that the
    reference's own ReplayParser produced."""
    path = "/root/reference/inputs/8_3_4_565/8_3_4_565_exam.xosc"
    if not os.path.exists(path):
        pytest.skip("reference data not mounted")
    sc, fx = S.parse_openscenario(path, "8_3_4_565"), scenario_fixture
    assert sc.dt == fx.dt and sc.max_t == fx.max_t and sc.goal == fx.goal
    for k in ("x", "y", "v", "yaw", "length", "width"):
        assert sc.ego[k] == fx.ego[k], k
    assert set(sc.vehicle_traj) == set(fx.vehicle_traj)
    for vid, tr in fx.vehicle_traj.items():
        assert tr == sc.vehicle_traj[vid], vid


def test_trajectory_container_contract():
    rows = np.full((51, 16), np.nan)
    rows[:, :9] = np.arange(51 * 9).reshape(51, 9)
    rows[:40, 9:12] = 1.0; rows[:39, 12:14] = 2.0; rows[:38, 14] = 3.0; rows[:37, 15] = 4.0
    tr = J.JerkSpaceSamplingFrenetTrajectory.from_rows(rows, 40, idx=7, path_id=1, cost=3.5)
    assert len(tr.s) == 51 and len(tr.x) == 40 and len(tr.yaw) == 40 and len(tr.c) == 39 and len(tr.c_d) == 38 and len(tr.c_dd) == 37
    fs = tr.frenet_state_at_time_step(t=1)
    assert (fs.s, fs.s_d, fs.s_dd, fs.d, fs.d_d, fs.d_dd) == tuple(rows[1, [1, 2, 3, 5, 6, 7]])
    other = J.JerkSpaceSamplingFrenetTrajectory.from_rows(rows, 40, idx=3, path_id=0, cost=3.5)
    assert other < tr                                    # equal cost -> lowest candidate index first (PriorityQueue tie)
    s = J.JerkSpaceSamplingPlannerSettings()
    assert (s.num_width, s.num_jerk, s.highest_jerk, s.plan_horizon) == (5, 5, 13.4112, 5.0)     # reference class defaults
    i = J.JerkSpaceSamplingPlannerSettings.intended()
    assert (i.num_width, i.num_jerk, i.highest_jerk) == (3, 25, 10.0)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        J.JsspEngine()
    h = C.c_void_p()
    cfg = J.JsspConfig().to_c()
    assert _lib.load().jssp_create(C.byref(cfg), 0, C.byref(h)) == -6      # JSSP_ERR_NO_DEVICE
    # nothing under the product package imports the oracle
    pkg = os.path.join(ROOT, "dlii_jssp_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert not re.search(r"^\s*(from|import)\s+oracle", open(os.path.join(pkg, fn)).read(), re.M), fn


def test_synthetic_generators_are_seeded_and_valid():
    a, b = S.dense_cycle(n_seeds=64, n_widths=2, O=4, M=2, seed=7), S.dense_cycle(n_seeds=64, n_widths=2, O=4, M=2, seed=7)
    assert np.array_equal(a.seeds, b.seeds) and np.array_equal(a.obstacles.state, b.obstacles.state)
    s, sd, sdd, _ = O.unroll_seeds(a.seeds, 0.0, 6.0, 0.0, np.linspace(0, 5, 51))
    assert (sd > -1e-6).all() and (sd < 18.0).all() and (np.abs(sdd) < 8.0 + 1e-9).all()      # the grammar's limits hold
    sc = S.synthetic_intersection(5)
    assert 3 <= len(sc.vehicle_traj) <= 12 and sc.route_xy.shape[1] == 2
    assert np.all(np.hypot(*np.diff(sc.route_xy, axis=0).T) > 1e-6)
    t = S.pack_obstacle_dict(sc.vehicle_traj, sc.dt, 130)
    assert t.valid.any()
    d = S.tensors_to_dict(t, sc.dt)
    t2 = S.pack_obstacle_dict(d, sc.dt, 130)
    assert np.array_equal(t.valid, t2.valid) and np.allclose(t.state, t2.state)


def test_clock_tables_follow_the_reference_clock(scenario_fixture):
    """The batched replay's per-cycle tables against the reference's own expressions evaluated step by step."""
    from dlii_jssp_b200 import batch as B
    dt, n_total = 0.1, 180
    times, now, end = B.clock_tables(dt, 120, n_total)
    t = 0.0
    for c in range(121):
        assert times[c] == t and now[c] == int(t / dt)
        key = str(np.around(t, 3))
        assert (end[c] == -1 and key not in {S.time_key(k, dt) for k in range(n_total)}) or S.time_key(int(end[c]), dt) == key
        t = float(t + dt)
    assert now[8] == 7 and end[8] == 8          # 0.7999999999999999 / 0.1 truncates to 7, but its 3-dp key is "0.8"


def test_packed_scenario_cache_round_trip(tmp_path, scenario_fixture):
    from dlii_jssp_b200 import batch as B
    packed = [B.pack_scenario(scenario_fixture), B.pack_scenario(S.synthetic_intersection(2), max_cycles=30),
              B.pack_scenario(S.synthetic_intersection(3), max_cycles=10, kinematics=True)]      # with the kinematics-mode polyline table
    B.save_packed(str(tmp_path / "shard.npz"), packed)
    back = B.load_packed(str(tmp_path / "shard.npz"))
    assert [p.name for p in back] == [p.name for p in packed]
    for a, b in zip(packed, back):
        for f in B._ARRAY_FIELDS:
            va, vb = getattr(a, f), getattr(b, f)
            assert (va is None and vb is None) or (va.dtype == vb.dtype and np.array_equal(va, vb)), f
        assert (a.n_dict, a.dt, a.max_t, a.final_time_step, a.initial_end) == (b.n_dict, b.dt, b.max_t, b.final_time_step, b.initial_end)
    assert back[2].polyline is not None and back[0].polyline is None and back[2].c_view(True).n_polyline == len(packed[2].polyline)
    # the packed obstacle table holds exactly the dictionary's states
    p = packed[0]
    tr = scenario_fixture.vehicle_traj
    for o, vid in enumerate(tr):
        keys = [k for k in tr[vid] if k != "shape"]
        assert int(p.valid[o, 0].sum()) == len([k for k in keys if round(float(k) / p.dt) < p.state.shape[2]])


def test_polyline_table_reproduces_from_state(scenario_fixture):
    """The table the kinematics mode uploads (pack_scenario(..., kinematics=True)) carries exactly the numbers
    FrenetState.from_state derives from the re-discretised reference line: the device formula (nearest row, arc + tangential
    offset, left-normal offset, heading-error split) evaluated on the table equals from_state bit for bit; the packed C view is
    built once and marks the mode."""
    import math
    from dlii_jssp_b200 import batch as B
    from dlii_jssp_b200.frenet import FrenetState, State
    from dlii_jssp_b200.geometry import CubicSpline2D
    sc = scenario_fixture
    p = B.pack_scenario(sc, kinematics=True, max_cycles=5)
    spline = CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1])
    ref = spline.sample(np.arange(0, spline.s_list[-1], 0.1))
    assert p.polyline.shape == (len(ref), 6)
    rng = np.random.default_rng(5)
    for _ in range(50):
        i = int(rng.integers(0, len(ref)))
        st = State(0.0, ref[i, 0] + rng.normal(0, 0.7), ref[i, 1] + rng.normal(0, 0.7), ref[i, 2] + rng.normal(0, 0.2), rng.uniform(0, 12), rng.normal(0, 1))
        want = FrenetState().from_state(st, ref).as_tuple()
        t = p.polyline
        k = int(np.argmin((t[:, 0] - st.x) ** 2 + (t[:, 1] - st.y) ** 2))
        ex, ey, dyaw = st.x - t[k, 0], st.y - t[k, 1], st.yaw - t[k, 2]
        got = (t[k, 3] + (ex * t[k, 4] + ey * t[k, 5]), st.v * math.cos(dyaw), st.a * math.cos(dyaw),
               -ex * t[k, 5] + ey * t[k, 4], st.v * math.sin(dyaw), st.a * math.sin(dyaw))
        assert got == want
    v0, v1 = p.c_view(False), p.c_view(True)
    assert v0 is p.c_view(False) and v1 is p.c_view(True) and v0 is not v1
    assert (v0.ego_update_mode, v1.ego_update_mode, v1.n_polyline) == (0, 1, len(ref)) and v1.pp_wheel_base > 0
    with pytest.raises(ValueError):
        B.pack_scenario(sc, max_cycles=5).c_view(True)


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Argument checks of the batched-replay / FOP / projection entry points run before any device work; without a GPU the
    constructors report JSSP_ERR_NO_DEVICE instead of falling back."""
    import torch
    lib = _lib.load()
    cfg = J.JsspConfig().to_c()
    h = C.c_void_p()
    assert lib.jssp_batch_create(C.byref(cfg), 0, 0, 10, C.byref(h)) == -1            # max_scenarios < 1
    assert lib.jssp_batch_create(C.byref(cfg), 0, 4, 0, C.byref(h)) == -1             # max_cycles < 1
    assert lib.jssp_batch_run(None, 1, 1, 0, None) == -1 and lib.jssp_batch_reset(None, 1, None) == -1
    assert lib.jssp_fop_evaluate(None, None, None, None) == -1
    assert lib.jssp_project_states(None, None, 0, 0.1, None, None) == -1
    assert lib.jssp_batch_destroy(None) == 0 and lib.jssp_batch_launch_count(None) == 0
    if not torch.cuda.is_available():
        assert lib.jssp_batch_create(C.byref(cfg), 0, 4, 10, C.byref(h)) == -6        # JSSP_ERR_NO_DEVICE
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            J.BatchReplay()
    assert C.sizeof(_lib.JsspCycleRecord) == 64 and C.sizeof(_lib.JsspScenarioResult) == 16 and C.sizeof(_lib.JsspFopIn) == 6 * 8 + 6 * 4 + 6 * 8


def test_replay_step_semantics_match_the_reference_environment(golden_dir, scenario_fixture):
    """tests/golden/replay_golden_8_3_4_565.npz was recorded from the reference's own Env.step / ReplayController.step
    (oracle/make_golden_replay.py).  Replaying its ego rows through the product's host logic must give the same end status
    at every step: clock, kinematic ego with the (0, 0) action, status on THAT pose (collision, max_t, map bounds, goal line)."""
    from dlii_jssp_b200 import replay as R
    g = np.load(os.path.join(golden_dir, "replay_golden_8_3_4_565.npz"))
    sc = scenario_fixture
    assert sc.bounds is not None and np.allclose(sc.bounds, g["bounds"])
    gp = R.goal_pose_of(sc)
    assert np.allclose(gp, g["goal_pose"])
    obs = sc.observation()
    assert R.end_status(obs, sc, None) == g["ends"][0]
    rows, ends = g["rows"], g["ends"]
    for k in range(len(rows)):
        class Best:                                    # the fields Env.step reads from best_traj (env.py:26-30)
            x = [None, rows[k, 1]]; y = [None, rows[k, 2]]; yaw = [None, rows[k, 5]]; s_d = [None, rows[k, 3]]; s_dd = [None, rows[k, 4]]
        R.advance(obs, sc, gp, Best)
        assert obs["test_setting"]["t"] == rows[k, 0]
        assert obs["test_setting"]["end"] == ends[k + 1], (k, obs["test_setting"]["end"], ends[k + 1])
        e = obs["vehicle_info"]["ego"]
        assert (e["x"], e["y"], e["v"], e["a"], e["yaw"]) == tuple(rows[k, 1:6])
    assert ends[-1] == 1.5 and (ends[:-1] == -1).all()
    # the same file holds the run of the WHOLE reference loop (its own IntersectionPlanner.plan -> JerkSpaceSamplingPlanner.plan_traj
    # -> Env.step, leaf classes stubbed; make_golden_replay.py --full): identical, cycle by cycle
    assert g["full_best_idx"].tolist() == g["best_idx"].tolist() and np.array_equal(g["full_ends"], ends)
    np.testing.assert_allclose(g["full_rows"], rows, rtol=1e-12, atol=1e-12)
    # the status is taken on the kinematic pose, not on the planner's: a goal line just behind the planner pose of the last
    # step but ahead of the kinematic pose must not end the replay
    k = len(rows) - 1
    prev = dict(sc.ego); prev.update(x=rows[k - 1, 1], y=rows[k - 1, 2], v=rows[k - 1, 3], a=rows[k - 1, 4], yaw=rows[k - 1, 5])
    kin = R.kinematic_ego(prev, sc.dt)
    assert np.hypot(kin["x"] - rows[k, 1], kin["y"] - rows[k, 2]) > 1e-6          # the two poses differ


def test_pure_pursuit_and_kinematic_update_match_the_reference_loop(golden_dir, scenario_fixture):
    """kin_* of the replay golden (the reference's own loop in "kinematics" mode): applying the recorded actions through the
    product's kinematic update reproduces the recorded ego rows and end statuses; the pure-pursuit look-ahead walk is the
    reference's (intersection_planner.py:365-389)."""
    from dlii_jssp_b200 import replay as R
    g = np.load(os.path.join(golden_dir, "replay_golden_8_3_4_565.npz"))
    sc = scenario_fixture
    gp = R.goal_pose_of(sc)
    obs = sc.observation()
    for k, (acc, steer) in enumerate(g["kin_actions"]):
        R.advance(obs, sc, gp, None, "kinematics", (acc, steer))
        e = obs["vehicle_info"]["ego"]
        np.testing.assert_allclose((obs["test_setting"]["t"], e["x"], e["y"], e["v"], e["a"], e["yaw"]), g["kin_rows"][k], rtol=0, atol=1e-12)
        assert obs["test_setting"]["end"] == g["kin_ends"][k + 1]
    path = np.column_stack([np.linspace(0, 10, 21), np.zeros(21)])
    pt, idx = J.IntersectionPlanner.find_lookahead_point_by_curve_distance(np.array([1.1, 0.3]), path, 2.2)
    assert idx == 7 and np.allclose(pt, [3.5, 0.0])                    # nearest way point 2 (x = 1.0), then 5 steps of 0.5 m >= 2.2 m
