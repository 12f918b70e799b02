#!/usr/bin/env python
"""Generate tests/golden/fop_cycle_golden.npz by RUNNING the reference's own FOP baseline planner.

TEST INFRASTRUCTURE - build container only (needs /root/reference).  Imports the *unmodified*

    /root/reference/code/intersection_local_planner/frenet_optimal_planner.py

with the same leaf stubs as oracle/make_golden_cycle.py (plus QuarticPolynomial = oracle.fop_oracle.QuarticPolynomial and a
GoalSamplingCostFunction holding the builder-specified cost of oracle/fop_oracle.py) and calls, in the order of its
``plan_traj`` (:301-333), the reference's own ``sampling_frenet_trajectories``, ``calc_global_paths``, ``check_constraints``,
``check_collisions`` and the `min_cost >= cost` loop.  ``plan_traj`` itself cannot be called: its unconditional plotting
block (:335-358) indexes ``self.all_trajs[time_step_now]`` and creates directories inside the (read-only) reference tree.
Reduced grid (3 widths x 4 horizons x 40 speeds) to keep the Python run short.

Usage:  python oracle/make_golden_fop.py  [out.npz]
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/code"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

CYCLES = [("t0", (2.93, 5.67, 0.0, 0.6, -0.4, 0.0), 0.0), ("t5", (27.0, 4.0, 0.2, 0.0, 0.0, 0.0), 5.0), ("t10", (40.0, 6.0, 0.0, -0.2, 0.1, 0.0), 10.0)]
GRID = dict(num_width=3, num_speed=40, num_t=4, min_t=5.0, max_t=5.5, highest_speed=18.0)
MAX_ACCEL, MAX_JERK = 2.5, 2.0          # tight limits so that check_constraints rejects part of the grid


def main(out):
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not mounted")
    import make_golden_cycle as G
    State, FrenetState, Vehicle = G.install_stubs()
    import types
    from oracle import fop_oracle as F, jssp_oracle as O
    sys.modules["common.geometry.polynomial"].QuarticPolynomial = F.QuarticPolynomial
    fr_mod = sys.modules["common.scenario.frenet"]
    fr_mod.GoalSamplingFrenetTrajectory = fr_mod.JerkSpaceSamplingFrenetTrajectory
    cfg = F.FopConfig(num_width=GRID["num_width"], num_speed=GRID["num_speed"], num_t=GRID["num_t"], min_t=GRID["min_t"], max_t=GRID["max_t"],
                      highest_speed=GRID["highest_speed"], max_accel=MAX_ACCEL, max_jerk=MAX_JERK)

    class GoalSamplingCostFunction:
        """The builder-specified FOP cost (oracle/fop_oracle.py) on one trajectory object, sequential Python sums."""

        def __init__(self, cost_type, vehicle_info=None, max_road_width=3.5):
            self.half = max_road_width / 2.0

        def cost_total(self, traj, target_speed, t_max, t_min):
            w = cfg.cost_weights
            n = len(traj.s)
            mean = lambda vals: sum(vals, 0.0) / n
            relu2 = lambda v, t: max(abs(v) - t, 0.0) ** 2
            cost = w[1] * (1.0 - (traj.s[-1] - traj.s[0]) / cfg.max_distance)
            cost = cost + w[2] * (mean([d * d for d in traj.d]) / (self.half * self.half))
            cost = cost + w[3] * mean([relu2(v, cfg.thr_lat_accel) for v in traj.d_dd])
            cost = cost + w[4] * mean([relu2(v, cfg.thr_lat_jerk) for v in traj.d_ddd])
            cost = cost + w[5] * mean([relu2(v, cfg.thr_lon_accel) for v in traj.s_dd])
            cost = cost + w[6] * mean([relu2(v, cfg.thr_lon_jerk) for v in traj.s_ddd])
            cost = cost + w[7] * mean([((target_speed - v) / target_speed) ** 2 for v in traj.s_d])
            if t_max > t_min:                                   # traj.horizon: tagged on the candidate below (the reference does not keep Ti)
                cost = cost + w[0] * (1.0 - (traj.horizon - t_min) / (t_max - t_min))
            return cost

    sys.modules["common.cost.cost_function"].GoalSamplingCostFunction = GoalSamplingCostFunction
    sys.path.insert(0, os.path.join(REF, "intersection_local_planner"))
    sys.path.insert(0, REF)
    import frenet_optimal_planner as ref                           # the unmodified reference module
    import dlii_jssp_b200 as J

    # the reference builds the lateral quintic with the horizon Ti but does not keep Ti on the trajectory; the stub quintic
    # records it on the template so that the (builder-specified) time term of the cost can read it
    class QuinticWithHorizon(O.QuinticPolynomial):
        last_T = None

        def __init__(self, xs, vxs, axs, xe, vxe, axe, time):
            super().__init__(xs, vxs, axs, xe, vxe, axe, time)
            QuinticWithHorizon.last_T = float(time)
    ref.QuinticPolynomial = QuinticWithHorizon
    Traj = fr_mod.JerkSpaceSamplingFrenetTrajectory
    Traj.horizon = property(lambda self: self._horizon)
    orig_init = Traj.__init__

    def init_with_horizon(self):
        orig_init(self)
        self._horizon = None
    Traj.__init__ = init_with_horizon
    orig_quartic = F.QuarticPolynomial

    class QuarticTagging(orig_quartic):
        def __init__(self, xs, vxs, axs, vxe, axe, time):
            super().__init__(xs, vxs, axs, vxe, axe, time)
            QuarticTagging.last_T = float(time)
    ref.QuarticPolynomial = QuarticTagging
    # tag every deep-copied candidate with its horizon right before its cost is computed
    import copy as _copy
    orig_deepcopy = _copy.deepcopy

    def tagging_deepcopy(obj, *a, **k):
        new = orig_deepcopy(obj, *a, **k)
        if isinstance(new, Traj):
            new._horizon = QuinticWithHorizon.last_T
        return new
    ref.copy.deepcopy = tagging_deepcopy

    sc = J.scenario.scenario_from_fixture(os.path.join(ROOT, "tests", "golden", "scenario_8_3_4_565.npz"))
    settings = ref.FrenetOptimalPlannerSettings(num_width=GRID["num_width"], num_speed=GRID["num_speed"], num_t=GRID["num_t"],
                                                highest_speed=GRID["highest_speed"], lowest_speed=0.0, min_planning_t=GRID["min_t"],
                                                max_planning_t=GRID["max_t"])
    gold = {"labels": np.array([c[0] for c in CYCLES]), "frenet0": np.array([c[1] for c in CYCLES]), "t": np.array([c[2] for c in CYCLES]),
            "grid": np.array([GRID["num_width"], GRID["num_speed"], GRID["num_t"]]), "limits": np.array([MAX_ACCEL, MAX_JERK])}
    for label, fr, t in CYCLES:
        veh = Vehicle(sc.ego["length"], sc.ego["width"])
        veh.max_accel, veh.max_jerk = MAX_ACCEL, MAX_JERK
        pl = ref.FrenetOptimalPlanner(settings, veh)
        pl.cubic_spline = O.CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1])
        obs = {"test_setting": {"t": t, "dt": sc.dt, "max_t": sc.max_t, "scenario_name": sc.name}}
        now = int(t / sc.dt)
        fs = FrenetState(t=t, s=fr[0], s_d=fr[1], s_dd=fr[2], d=fr[3], d_d=fr[4], d_dd=fr[5])
        with contextlib.redirect_stdout(io.StringIO()):
            ftlist = pl.sampling_frenet_trajectories(fs)            # :315
            sampled = list(ftlist)
            ftlist = pl.calc_global_paths(ftlist)                   # :320
            ftlist = pl.check_constraints(ftlist)                   # :327
            ftlist = pl.check_collisions(ftlist, sc.vehicle_traj, now, obs)   # :329
        min_cost, best = float("inf"), None                          # :334-338
        for ft in ftlist:
            if min_cost >= ft.cost_total:
                min_cost, best = ft.cost_total, ft
        pos = {id(tr): n for n, tr in enumerate(sampled)}
        N = len(sampled)
        feasible = np.array([tr.constraint_passed for tr in sampled])
        coll_pass = np.array([tr.collision_passed for tr in sampled])
        cost = np.array([tr.cost_total for tr in sampled], dtype=np.float64)       # the reference costs every candidate at sampling time
        n_pts = np.array([len(tr.x) for tr in sampled])
        n_own = np.array([len(tr.s) for tr in sampled])
        best_idx = pos[id(best)] if best is not None else -1
        for k, v in (("n", np.array(N)), ("feasible", feasible), ("collision_passed", coll_pass), ("cost", cost), ("n_pts", n_pts),
                     ("n_own", n_own), ("best", np.array(best_idx))):
            gold[f"{label}_{k}"] = v
        if best is not None:
            P = len(best.s)
            rows = np.full((P, len(G.FIELDS)), np.nan)
            for c, f in enumerate(G.FIELDS):
                v = np.asarray(getattr(best, f), dtype=np.float64)[:P]
                rows[:len(v), c] = v
            gold[f"{label}_best_rows"] = rows
        print(f"{label}: N={N} feasible={int(feasible.sum())} survivors={len(ftlist)} best={best_idx} cost={min_cost}")
    np.savez_compressed(out, **gold)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "fop_cycle_golden.npz"))
