"""Loop-form CPU port of the reference planner's evaluation stages (float64, one Python object per trajectory).

TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/jssp_oracle.py).  This is the "reference-loop" form named
in SURVEY.md section 8d: the stages run per candidate with scalar Python arithmetic and per-trajectory NumPy calls, in
the order of JerkSpaceSamplingPlanner.plan_traj (jerk_space_sampling_planner.py:296-327), because that is how the
reference spends its time.  It exists for two reasons:

  * to cross-check the vectorised oracle with an independently written implementation of the same semantics;
  * to serve as the CPU baseline ("kind": "port") that bench.py times next to the GPU path.  The reference itself
    cannot be run (its ``common`` package is unpublished and shapely is not installed); the one substitution is a
    separating-axis rectangle test in place of ``shapely.Polygon.intersects`` (jerk_space_sampling_planner.py:268).
"""
from __future__ import annotations

import bisect
import math
from queue import PriorityQueue

import numpy as np

from .jssp_oracle import OracleConfig, QuinticPolynomial, lateral_targets, sample_seeds


class Trajectory:
    __slots__ = ("t", "s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd", "x", "y", "yaw", "ds", "c", "c_d", "c_dd", "path_id",
                 "idx", "cost_total", "constraint_passed", "collision_passed")

    def __init__(self):
        for f in self.__slots__[:16]:
            setattr(self, f, [])
        self.path_id = 0; self.idx = -1; self.cost_total = float("inf")
        self.constraint_passed = False; self.collision_passed = False

    def __lt__(self, other):   # tie rule of the PriorityQueue tuples: lowest candidate index
        return self.idx < other.idx


class ScalarSpline:
    """Scalar evaluation of the oracle's CubicSpline2D tables (bisect + cubic per call, like the reference's class)."""

    def __init__(self, spline):
        self.s = [float(v) for v in spline.s_list]
        self.cx = [list(map(float, col)) for col in (spline.sx.a, spline.sx.b, spline.sx.c, spline.sx.d)]
        self.cy = [list(map(float, col)) for col in (spline.sy.a, spline.sy.b, spline.sy.c, spline.sy.d)]
        self.n = len(self.s)

    def _seg(self, s):
        return min(bisect.bisect_right(self.s, s) - 1, self.n - 2)

    def calc_position(self, s):
        if s < self.s[0] or s > self.s[-1]:
            return None, None
        i = self._seg(s); dx = s - self.s[i]
        (ax, bx, cx, dxx), (ay, by, cy, dyy) = self.cx, self.cy
        return (((ax[i] + bx[i] * dx) + cx[i] * (dx * dx)) + dxx[i] * ((dx * dx) * dx),
                ((ay[i] + by[i] * dx) + cy[i] * (dx * dx)) + dyy[i] * ((dx * dx) * dx))

    def calc_yaw(self, s):
        i = self._seg(s); dx = s - self.s[i]
        tx = (self.cx[1][i] + (2.0 * self.cx[2][i]) * dx) + (3.0 * self.cx[3][i]) * (dx * dx)
        ty = (self.cy[1][i] + (2.0 * self.cy[2][i]) * dx) + (3.0 * self.cy[3][i]) * (dx * dx)
        return math.atan2(ty, tx)


def _st(s0, v0, a0, j, t):
    return s0 + v0 * t + 0.5 * a0 * t * t + (1 / 6.0) * j * t * t * t


def _vt(v0, a0, j, t):
    return v0 + a0 * t + 0.5 * j * t * t


def unroll_one(seed, s0, v0, a0, times, proto: Trajectory) -> Trajectory:
    """One longitudinal profile appended to a copy of the lateral prototype (polyvt_sampling.py:311-385)."""
    tr = Trajectory()
    tr.path_id = proto.path_id
    tr.d, tr.d_d, tr.d_dd, tr.d_ddd = proto.d, proto.d_d, proto.d_dd, proto.d_ddd
    j1, t1, j2, t2, j3, t3, j4, t4, j5, _ = [float(v) for v in seed]
    knots = [(s0, v0, a0)]
    for j, dur in ((j1, t1), (j2, t2), (j3, t3), (j4, t4)):
        sb, vb, ab = knots[-1]
        knots.append((_st(sb, vb, ab, j, dur), _vt(vb, ab, j, dur), ab + j * dur))
    jerks = (j1, j2, j3, j4, j5)
    for t in times:
        if t < t1 - 1e-9:
            k, dt = 0, t
        elif t < t1 + t2 + 1e-9:
            k, dt = 1, t - t1
        elif t < t1 + t2 + t3 + 1e-9:
            k, dt = 2, t - t1 - t2
        elif t < t1 + t2 + t3 + t4 + 1e-9:
            k, dt = 3, t - t1 - t2 - t3
        else:
            k, dt = 4, t - t1 - t2 - t3 - t4
        sb, vb, ab = knots[k]; j = jerks[k]
        tr.t.append(t); tr.s_ddd.append(j); tr.s_dd.append(ab + j * dt)
        tr.s_d.append(_vt(vb, ab, j, dt)); tr.s.append(_st(sb, vb, ab, j, dt))
    return tr


def global_path(tr: Trajectory, sp: ScalarSpline, dt: float):
    """calc_global_paths for one trajectory (jerk_space_sampling_planner.py:134-165)."""
    xs, ys = [], []
    for s, d in zip(tr.s, tr.d):
        ix, iy = sp.calc_position(s)
        if ix is None:
            break
        yaw = sp.calc_yaw(s)
        xs.append(ix + d * math.cos(yaw + math.pi / 2.0))
        ys.append(iy + d * math.sin(yaw + math.pi / 2.0))
    tr.x, tr.y = xs, ys
    if len(xs) >= 2:
        tr.x, tr.y = np.array(xs), np.array(ys)
        xd, yd = np.diff(tr.x), np.diff(tr.y)
        yaw = np.arctan2(yd, xd)
        tr.ds = np.hypot(xd, yd)
        tr.yaw = np.append(yaw, yaw[-1])
        with np.errstate(divide="ignore", invalid="ignore"):
            tr.c = np.divide(np.diff(tr.yaw), tr.ds)
            tr.c_d = np.divide(np.diff(tr.c), dt)
            tr.c_dd = np.divide(np.diff(tr.c_d), dt)


def rect_hit(ex, ey, eyaw, ea, eb, ox, oy, oyaw, oa, ob) -> bool:
    """Closed-set rectangle overlap, scalar (stand-in for the two shapely polygons + intersects, :263-268)."""
    c1, s1, c2, s2 = math.cos(eyaw), math.sin(eyaw), math.cos(oyaw), math.sin(oyaw)
    dx, dy = ox - ex, oy - ey
    cd, sd = abs(c1 * c2 + s1 * s2), abs(c1 * s2 - s1 * c2)
    if abs(dx * c1 + dy * s1) > ea + oa * cd + ob * sd: return False
    if abs(-dx * s1 + dy * c1) > eb + oa * sd + ob * cd: return False
    if abs(dx * c2 + dy * s2) > oa + ea * cd + eb * sd: return False
    if abs(-dx * s2 + dy * c2) > ob + ea * sd + eb * cd: return False
    return True


def has_collision(tr: Trajectory, obstacles, time_step_now, final_time_step, cfg: OracleConfig, n_dict=None):
    """jerk_space_sampling_planner.py:233-272 with early exit; returns (collision, risk)."""
    state, valid, size, prob = obstacles
    O, M, Tt, _ = state.shape
    if (O if n_dict is None else n_dict) <= 0:
        return False, 0.0
    risk = 0.0
    ea, eb = cfg.ego_length / 2.0, cfg.ego_width / 2.0
    for i in range(min(len(tr.x), final_time_step - time_step_now)):
        if i % cfg.check_res:
            continue
        if len(tr.yaw) <= i:          # 1-point trajectory: the reference's bare except reports a collision (:250-254)
            return True, risk
        step = i + time_step_now
        if step >= Tt:
            continue
        ex, ey, eyaw = float(tr.x[i]), float(tr.y[i]), float(tr.yaw[i])
        for o in range(O):
            oa, ob = (size[o, 0] + cfg.obstacle_inflation) / 2.0, (size[o, 1] + cfg.obstacle_inflation) / 2.0
            for m in range(M):
                if not valid[o, m, step]:
                    continue
                ox, oy, oyaw = state[o, m, step]
                if rect_hit(ex, ey, eyaw, ea, eb, ox, oy, oyaw, oa, ob):
                    if prob[o, m] >= cfg.p_min:
                        return True, risk
                    risk += prob[o, m]
    return False, risk


def traj_cost(tr: Trajectory, cfg: OracleConfig, risk=0.0) -> float:
    w = cfg.cost_weights
    n = len(tr.s)
    half = cfg.max_road_width / 2.0
    acc = [0.0] * 6
    for i in range(n):
        acc[0] = acc[0] + tr.d[i] * tr.d[i]
        acc[1] = acc[1] + max(abs(tr.d_dd[i]) - cfg.thr_lat_accel, 0.0) ** 2
        acc[2] = acc[2] + max(abs(tr.d_ddd[i]) - cfg.thr_lat_jerk, 0.0) ** 2
        acc[3] = acc[3] + max(abs(tr.s_dd[i]) - cfg.thr_lon_accel, 0.0) ** 2
        acc[4] = acc[4] + max(abs(tr.s_ddd[i]) - cfg.thr_lon_jerk, 0.0) ** 2
        acc[5] = acc[5] + ((cfg.speed_max - tr.s_d[i]) / cfg.speed_max) ** 2
    cost = w[1] * (1.0 - (tr.s[-1] - tr.s[0]) / cfg.max_distance)
    cost = cost + w[2] * (acc[0] / n / (half * half))
    for k, wk in ((1, w[3]), (2, w[4]), (3, w[5]), (4, w[6]), (5, w[7])):
        cost = cost + wk * (acc[k] / n)
    return cost + cfg.w_risk * risk


def plan_cycle_loop(cfg: OracleConfig, spline, frenet0, obstacles, time_step_now, final_time_step, seeds=None, targets=None,
                    n_dict=None, candidate_subset=None) -> dict:
    """Same inputs/outputs as jssp_oracle.plan_cycle, computed one trajectory at a time.  ``candidate_subset`` restricts
    the work to the given candidate indices (for bounded CPU timing samples)."""
    s0, v0, a0, d0, dd0, ddd0 = [float(v) for v in frenet0]
    if seeds is None:
        seeds = sample_seeds(s0, v0, a0, cfg)
    targets = lateral_targets(cfg) if targets is None else targets
    times = np.linspace(0.0, cfg.plan_horizon, cfg.n_steps + 1).tolist()
    sp = ScalarSpline(spline)
    protos = []
    for pid, di in enumerate(targets):
        q = QuinticPolynomial(d0, dd0, ddd0, float(di), 0.0, 0.0, cfg.plan_horizon)
        p = Trajectory(); p.path_id = pid
        p.d = [float(q.calc_point(t)) for t in times]; p.d_d = [float(q.calc_first_derivative(t)) for t in times]
        p.d_dd = [float(q.calc_second_derivative(t)) for t in times]; p.d_ddd = [float(q.calc_third_derivative(t)) for t in times]
        protos.append(p)
    Ns, N = len(seeds), len(seeds) * len(targets)
    todo = range(N) if candidate_subset is None else candidate_subset
    feasible = {}; collision = {}; cost = {}
    queue = PriorityQueue()
    for n in todo:
        w, k = divmod(n, Ns)
        tr = unroll_one(seeds[k], s0, v0, a0, times, protos[w]); tr.idx = n
        global_path(tr, sp, cfg.delta_t)
        with np.errstate(invalid="ignore"):
            tr.constraint_passed = not any(abs(c) > cfg.max_curvature for c in tr.c)
        hit, risk = has_collision(tr, obstacles, time_step_now, final_time_step, cfg, n_dict)
        tr.collision_passed = not hit
        tr.cost_total = traj_cost(tr, cfg, risk)
        feasible[n], collision[n], cost[n] = tr.constraint_passed, hit, tr.cost_total
        if tr.constraint_passed and tr.collision_passed:
            queue.put((tr.cost_total, tr, n))
    best = -1; best_traj = None
    if not queue.empty():
        _, best_traj, best = queue.get()
    idx = list(todo)
    return dict(index=np.array(idx), feasible=np.array([feasible[n] for n in idx], dtype=bool), collision=np.array([collision[n] for n in idx], dtype=bool),
                cost=np.array([cost[n] for n in idx]), best=best, best_traj=best_traj, n=len(idx))
