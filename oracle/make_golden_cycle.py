#!/usr/bin/env python
"""Generate tests/golden/jssp_cycle_golden.npz by RUNNING the reference's own plan cycle.

TEST INFRASTRUCTURE - runs only in the build container, where /root/reference is mounted; nothing on the GPU box executes
this file.  It imports the *unmodified* reference module

    /root/reference/code/intersection_local_planner/jerk_space_sampling_planner.py

and calls its ``JerkSpaceSamplingPlanner.plan_traj`` (:296-362) - i.e. the reference's own ``sampling_frenet_trajectories``
(with the real ``polyvt_sampling.py``), ``calc_global_paths``, ``check_constraints_curvature``, ``check_collisions`` /
``has_collision``, ``calc_all_trajs_cost`` and the PriorityQueue selection - on plan cycles of the bundled scenario
inputs/8_3_4_565.  What the published tree does not contain is supplied as ``sys.modules`` stubs holding the LEAF classes
only (their formulation is the builder specification of SURVEY.md section 8c, the same one oracle/jssp_oracle.py states):

  shapely.geometry.Polygon, shapely.affinity      convex polygon, translate, rotate about the bbox centre, closed-set
                                                  ``intersects`` by the separating-axis theorem (shapely is not installed)
  scipy.integrate.cumtrapz                        imported by the reference, never called (removed from scipy >= 1.14)
  common.geometry.cubic_spline.CubicSpline2D      natural cubic spline over chord length (oracle.jssp_oracle.CubicSpline2D)
  common.geometry.polynomial.QuinticPolynomial    PythonRobotics quintic (oracle.jssp_oracle.QuinticPolynomial)
  common.cost.cost_function.JerkSpaceSamplingCostFunction   the cost of SURVEY section 8c on one trajectory
  common.scenario.frenet.*                        list-field containers; trajectories compare equal, so that PriorityQueue ties
                                                  fall through to the candidate's position (lowest index first)
  common.vehicle.vehicle.Vehicle                  l, w, max_curvature = 0.2, polygon = centred l x w rectangle
  spatio_temporal_graph_node, common.plot_scripts, matplotlib, tqdm   import-only

So everything ABOVE the leaves - loop structure, sample grid, truncation at the end of the reference line, forward-difference
headings with the last one repeated, un-unwrapped curvature, `i % 2` collision rows, string time keys, +0.1 m inflation,
`final_time_step - time_step_now` window, survivor filtering order, heap selection - is the reference's own executable code.
The golden file pins oracle/jssp_oracle.py (tests/test_oracle.py) and the CUDA path (tests/test_gpu_parity.py) to it.

Usage:  python oracle/make_golden_cycle.py  [out.npz]
"""
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/code"
sys.path.insert(0, ROOT)

# plan cycles recorded: (label, frenet start (s, s_d, s_dd, d, d_d, d_dd), t)
CYCLES = [("t0", (2.93, 5.67, 0.0, 0.6, -0.4, 0.0), 0.0),
          ("t2", (14.0, 5.2, -0.3, 0.1, 0.0, 0.0), 2.0),
          ("t5", (27.0, 4.0, 0.2, 0.0, 0.0, 0.0), 5.0),
          ("t10", (40.0, 6.0, 0.0, -0.2, 0.1, 0.0), 10.0)]
FIELDS = ("t", "s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd", "x", "y", "yaw", "ds", "c", "c_d", "c_dd")


# ---- shapely stand-in -------------------------------------------------------------------------------------------------
class Polygon:
    """Convex polygon given by its vertices (a repeated closing vertex is dropped)."""

    def __init__(self, pts):
        p = [tuple(map(float, q)) for q in pts]
        if len(p) > 1 and p[0] == p[-1]:
            p = p[:-1]
        self.pts = p

    def intersects(self, other) -> bool:      # closed sets: touching counts, like shapely
        for poly in (self.pts, other.pts):
            n = len(poly)
            for i in range(n):
                ex, ey = poly[(i + 1) % n][0] - poly[i][0], poly[(i + 1) % n][1] - poly[i][1]
                ax, ay = -ey, ex
                a = [ax * x + ay * y for x, y in self.pts]
                b = [ax * x + ay * y for x, y in other.pts]
                if max(a) < min(b) or max(b) < min(a):
                    return False
        return True


def _translate(poly, xoff=0.0, yoff=0.0):
    return Polygon([(x + xoff, y + yoff) for x, y in poly.pts])


def _rotate(poly, angle, origin="center", use_radians=False):
    a = angle if use_radians else math.radians(angle)
    xs, ys = [p[0] for p in poly.pts], [p[1] for p in poly.pts]
    cx, cy = (min(xs) + max(xs)) / 2.0, (min(ys) + max(ys)) / 2.0          # shapely's default origin: centre of the bounding box
    c, s = math.cos(a), math.sin(a)
    return Polygon([(cx + c * (x - cx) - s * (y - cy), cy + s * (x - cx) + c * (y - cy)) for x, y in poly.pts])


def install_stubs():
    from oracle import jssp_oracle as O

    def mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        return m

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.collections"):
        mod(name)
    sys.modules["matplotlib.gridspec"].GridSpec = object
    sys.modules["matplotlib.collections"].LineCollection = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    mod("tqdm").tqdm = lambda it, *a, **k: it
    mod("shapely"); mod("shapely.geometry").Polygon = Polygon
    aff = mod("shapely.affinity"); aff.translate = _translate; aff.rotate = _rotate
    sys.modules["shapely"].affinity = aff; sys.modules["shapely"].geometry = sys.modules["shapely.geometry"]
    import scipy.integrate
    if not hasattr(scipy.integrate, "cumtrapz"):
        scipy.integrate.cumtrapz = getattr(scipy.integrate, "cumulative_trapezoid", None)

    class State:
        def __init__(self, t=0.0, x=0.0, y=0.0, yaw=0.0, v=0.0, a=0.0):
            self.t, self.x, self.y, self.yaw, self.v, self.a = t, x, y, yaw, v, a

    class FrenetState:
        def __init__(self, t=0.0, s=0.0, s_d=0.0, s_dd=0.0, s_ddd=0.0, d=0.0, d_d=0.0, d_dd=0.0, d_ddd=0.0):
            self.t, self.s, self.s_d, self.s_dd, self.s_ddd, self.d, self.d_d, self.d_dd, self.d_ddd = t, s, s_d, s_dd, s_ddd, d, d_d, d_dd, d_ddd

        def from_state(self, state, polyline):
            """Builder-specified projection (the class is absent): nearest point of the re-discretised reference line, chord-length
            arc sum, left-hand normal, speed / acceleration split by the heading error (dlii_jssp_b200/frenet.py states the same)."""
            d2 = (polyline[:, 0] - state.x) ** 2 + (polyline[:, 1] - state.y) ** 2
            i = int(np.argmin(d2))
            seg = np.hypot(np.diff(polyline[:, 0]), np.diff(polyline[:, 1]))
            rx, ry, ryaw = polyline[i, 0], polyline[i, 1], polyline[i, 2]
            dx, dy = state.x - rx, state.y - ry
            tx, ty = math.cos(ryaw), math.sin(ryaw)
            self.t = state.t
            self.s = float(seg[:i].sum()) + (dx * tx + dy * ty)
            self.d = -dx * ty + dy * tx
            dyaw = state.yaw - ryaw
            self.s_d, self.d_d = state.v * math.cos(dyaw), state.v * math.sin(dyaw)
            self.s_dd, self.d_dd = state.a * math.cos(dyaw), state.a * math.sin(dyaw)
            self.s_ddd = self.d_ddd = 0.0
            return self

    class JerkSpaceSamplingFrenetTrajectory:
        def __init__(self):
            for f in FIELDS:
                setattr(self, f, [])
            self.path_id = 0
            self.cost_total = float("inf")
            self.constraint_passed = False
            self.collision_passed = False

        # every trajectory compares equal: a PriorityQueue tie falls through to the third tuple element, the position in
        # the survivor list (jerk_space_sampling_planner.py:292) - lowest candidate index first
        def __eq__(self, other):
            return True

        __hash__ = object.__hash__

        def __lt__(self, other):
            return False

        def frenet_state_at_time_step(self, t=1):                    # intersection_planner.py:312
            return FrenetState(t=self.t[t], s=self.s[t], s_d=self.s_d[t], s_dd=self.s_dd[t], s_ddd=self.s_ddd[t],
                               d=self.d[t], d_d=self.d_d[t], d_dd=self.d_dd[t], d_ddd=self.d_ddd[t])

    class Vehicle:
        def __init__(self, length, width, max_curvature=0.2, longitudinal_a_max=8.0, longitudinal_jerk_max=10.0, **_):
            self.l, self.w, self.max_curvature = length, width, max_curvature
            self.max_accel, self.max_jerk = longitudinal_a_max, longitudinal_jerk_max
            self.polygon = Polygon([(length / 2, width / 2), (length / 2, -width / 2), (-length / 2, -width / 2), (-length / 2, width / 2)])

    class JerkSpaceSamplingCostFunction:
        """The cost of SURVEY section 8c, evaluated on one trajectory object with sequential Python sums."""

        def __init__(self, cost_type, vehicle_info=None, max_road_width=3.5):
            from configuration_parameters import parameters, paras_planner
            self.w = paras_planner["planner_JSSP"]["cost_wights"]
            self.max_distance = paras_planner["planner_JSSP"]["max_distance"]
            self.thr = parameters["para_threshold"]
            self.half = max_road_width / 2.0

        def cost_total(self, traj, target_speed):
            w, thr = self.w, self.thr
            n = len(traj.s)
            mean = lambda vals: sum(vals, 0.0) / n
            relu2 = lambda v, t: max(abs(v) - t, 0.0) ** 2
            cost = w[1] * (1.0 - (traj.s[-1] - traj.s[0]) / self.max_distance)
            cost = cost + w[2] * (mean([d * d for d in traj.d]) / (self.half * self.half))
            cost = cost + w[3] * mean([relu2(v, thr["lateral_accel_max"]) for v in traj.d_dd])
            cost = cost + w[4] * mean([relu2(v, thr["lateral_jerk_max"]) for v in traj.d_ddd])
            cost = cost + w[5] * mean([relu2(v, thr["lon_accel_max"]) for v in traj.s_dd])
            cost = cost + w[6] * mean([relu2(v, thr["lon_jerk_max"]) for v in traj.s_ddd])
            cost = cost + w[7] * mean([((target_speed - v) / target_speed) ** 2 for v in traj.s_d])
            traj.cost_total = cost
            return traj

    for name in ("common", "common.cost", "common.geometry", "common.scenario", "common.vehicle", "common.plot_scripts"):
        mod(name)
    mod("common.cost.cost_function").JerkSpaceSamplingCostFunction = JerkSpaceSamplingCostFunction
    mod("common.geometry.cubic_spline").CubicSpline2D = O.CubicSpline2D
    poly = mod("common.geometry.polynomial"); poly.QuinticPolynomial = O.QuinticPolynomial; poly.QuarticPolynomial = object
    fr = mod("common.scenario.frenet")
    fr.State, fr.FrenetState, fr.JerkSpaceSamplingFrenetTrajectory = State, FrenetState, JerkSpaceSamplingFrenetTrajectory
    fr.GoalSamplingFrenetTrajectory = JerkSpaceSamplingFrenetTrajectory
    mod("common.tool")
    sys.modules["common"].tool = sys.modules["common.tool"]
    mod("common.cost.cost_function").GoalSamplingCostFunction = JerkSpaceSamplingCostFunction    # constructed, never used, by the "frenet" wiring
    mod("common.vehicle.vehicle").Vehicle = Vehicle
    mod("spatio_temporal_graph_node").StNode = object
    mod("common.plot_scripts.plot_polyvt_sampling").plot_jerk_sampling_of_best_traj = lambda *a, **k: None
    return State, FrenetState, Vehicle


def main(out):
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not mounted; golden vectors can only be regenerated in the build container")
    State, FrenetState, Vehicle = install_stubs()
    sys.path.insert(0, os.path.join(REF, "intersection_local_planner"))
    sys.path.insert(0, REF)
    import contextlib, io
    import jerk_space_sampling_planner as ref                       # the unmodified reference module
    from oracle import jssp_oracle as O
    import dlii_jssp_b200 as J                                      # only to rebuild the scenario dictionary from its fixture

    sc = J.scenario.scenario_from_fixture(os.path.join(ROOT, "tests", "golden", "scenario_8_3_4_565.npz"))
    settings = ref.JerkSpaceSamplingPlannerSettings(num_width=3, num_jerk=25, delta_t=0.1, max_road_width=3.5, highest_jerk=10.0,
                                                    plan_horizon=5.0, traj_num_max=50)
    gold = {"labels": np.array([c[0] for c in CYCLES]), "frenet0": np.array([c[1] for c in CYCLES]), "t": np.array([c[2] for c in CYCLES])}
    for label, fr, t in CYCLES:
        planner = ref.JerkSpaceSamplingPlanner(settings, Vehicle(sc.ego["length"], sc.ego["width"]))
        planner.cubic_spline = O.CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1])
        obs = {"test_setting": {"t": t, "dt": sc.dt, "max_t": sc.max_t, "scenario_name": sc.name}}
        fs = FrenetState(t=t, s=fr[0], s_d=fr[1], s_dd=fr[2], d=fr[3], d_d=fr[4], d_dd=fr[5])
        with contextlib.redirect_stdout(io.StringIO()):
            best = planner.plan_traj(fs, State(t=t), sc.vehicle_traj, obs)
        sampled = planner.all_path_sampling_trajs[-1]                # every candidate, in sampling order (:124-127)
        survivors = planner.all_trajs[-1]                            # after both checks, with costs
        N = len(sampled)
        pos = {id(tr): n for n, tr in enumerate(sampled)}
        feasible = np.array([tr.constraint_passed for tr in sampled])
        coll_pass = np.array([tr.collision_passed for tr in sampled])        # only evaluated for the feasible ones
        cost = np.full(N, np.nan)
        for tr in survivors:
            cost[pos[id(tr)]] = tr.cost_total
        n_pts = np.array([len(tr.x) for tr in sampled])
        best_idx = pos[id(best)] if best is not None else -1
        gold[f"{label}_n"] = np.array(N)
        gold[f"{label}_feasible"] = feasible
        gold[f"{label}_collision_passed"] = coll_pass
        gold[f"{label}_cost"] = cost
        gold[f"{label}_n_pts"] = n_pts
        gold[f"{label}_best"] = np.array(best_idx)
        if best is not None:
            P = len(best.s)
            rows = np.full((P, len(FIELDS)), np.nan)
            for c, f in enumerate(FIELDS):
                v = np.asarray(getattr(best, f), dtype=np.float64)[:P]      # ft.t holds 102 entries (:118 pre-fills it, the unroll appends)
                rows[:len(v), c] = v
            gold[f"{label}_best_rows"] = rows
        # a deterministic subsample of candidates with their Cartesian paths (pins calc_global_paths on non-selected candidates)
        pick = np.unique(np.linspace(0, N - 1, 24).astype(int))
        xy = np.full((len(pick), 51, 4), np.nan)
        for q, n in enumerate(pick):
            tr = sampled[n]
            k = len(tr.x)
            xy[q, :k, 0], xy[q, :k, 1] = tr.x, tr.y
            if k >= 2:
                xy[q, :k, 2] = tr.yaw
                xy[q, :k - 1, 3] = tr.c
        gold[f"{label}_pick"] = pick
        gold[f"{label}_pick_xyyawc"] = xy
        print(f"{label}: t={t} N={N} feasible={int(feasible.sum())} survivors={len(survivors)} best={best_idx} "
              f"cost={cost[best_idx] if best_idx >= 0 else None}")
    np.savez_compressed(out, **gold)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "jssp_cycle_golden.npz"))
