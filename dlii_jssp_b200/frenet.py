"""Data contract of the hot path: the containers the reference imports from its (unpublished) ``common.scenario.frenet``
and ``common.vehicle.vehicle`` modules (call sites: jerk_space_sampling_planner.py:27-28, intersection_planner.py:117-142,312).
Field names and meaning follow those call sites so that the reference driver can read a planned trajectory unchanged."""
from __future__ import annotations

import math
from typing import Optional

import numpy as np


class State:
    """Cartesian ego state: State(t, x, y, yaw, v, a) (intersection_planner.py:129-138)."""

    def __init__(self, t: float = 0.0, x: float = 0.0, y: float = 0.0, yaw: float = 0.0, v: float = 0.0, a: float = 0.0):
        self.t, self.x, self.y, self.yaw, self.v, self.a = t, x, y, yaw, v, a

    def __repr__(self):
        return f"State(t={self.t}, x={self.x}, y={self.y}, yaw={self.yaw}, v={self.v}, a={self.a})"


class FrenetState:
    """FrenetState(t, s, s_d, s_dd, s_ddd, d, d_d, d_dd, d_ddd) (intersection_planner.py:76,141,312)."""

    def __init__(self, t=0.0, s=0.0, s_d=0.0, s_dd=0.0, s_ddd=0.0, d=0.0, d_d=0.0, d_dd=0.0, d_ddd=0.0):
        self.t, self.s, self.s_d, self.s_dd, self.s_ddd = t, s, s_d, s_dd, s_ddd
        self.d, self.d_d, self.d_dd, self.d_ddd = d, d_d, d_dd, d_ddd

    def from_state(self, state: State, polyline: np.ndarray):
        """Project a Cartesian state on the re-discretised reference line ``polyline`` [N, 4] = x, y, yaw, curvature
        (intersection_planner.py:140-142, 159-161).  Builder-specified (class absent from the tree): nearest polyline
        point, signed lateral offset by the left-hand normal, speed/acceleration decomposed by the heading error."""
        d2 = (polyline[:, 0] - state.x) ** 2 + (polyline[:, 1] - state.y) ** 2
        i = int(np.argmin(d2))
        seg = np.hypot(np.diff(polyline[:, 0]), np.diff(polyline[:, 1]))
        s_i = float(seg[:i].sum())
        rx, ry, ryaw = polyline[i, 0], polyline[i, 1], polyline[i, 2]
        dx, dy = state.x - rx, state.y - ry
        tx, ty = math.cos(ryaw), math.sin(ryaw)
        self.t = state.t
        self.s = s_i + (dx * tx + dy * ty)
        self.d = -dx * ty + dy * tx
        dyaw = state.yaw - ryaw
        self.s_d = state.v * math.cos(dyaw)
        self.d_d = state.v * math.sin(dyaw)
        self.s_dd = state.a * math.cos(dyaw)
        self.d_dd = state.a * math.sin(dyaw)
        self.s_ddd = 0.0
        self.d_ddd = 0.0
        return self

    def as_tuple(self):
        return (self.s, self.s_d, self.s_dd, self.d, self.d_d, self.d_dd)

    def __repr__(self):
        return f"FrenetState(t={self.t}, s={self.s}, s_d={self.s_d}, s_dd={self.s_dd}, d={self.d}, d_d={self.d_d}, d_dd={self.d_dd})"


class JerkSpaceSamplingFrenetTrajectory:
    """Trajectory container with the reference's field names (list/array per field).
    ``x, y, yaw`` have ``n_pts`` entries, ``ds, c`` n_pts-1, ``c_d`` n_pts-2, ``c_dd`` n_pts-3
    (jerk_space_sampling_planner.py:146-165)."""

    FIELDS = ("t", "s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd", "x", "y", "yaw", "ds", "c", "c_d", "c_dd")

    def __init__(self):
        for f in self.FIELDS:
            setattr(self, f, [])
        self.path_id = 0
        self.idx = -1
        self.cost_total = float("inf")
        self.constraint_passed = False
        self.collision_passed = False

    @classmethod
    def from_rows(cls, rows: np.ndarray, n_pts: int, idx: int, path_id: int, cost: float) -> "JerkSpaceSamplingFrenetTrajectory":
        """rows = [P, 16] float64 as exported by jssp_evaluate (columns = FIELDS)."""
        tr = cls()
        cols = np.ascontiguousarray(rows.T)          # one copy; the fields below are views of it
        n = max(int(n_pts), 0)
        tr.t, tr.s, tr.s_d, tr.s_dd, tr.s_ddd, tr.d, tr.d_d, tr.d_dd, tr.d_ddd = cols[:9]
        tr.x, tr.y, tr.yaw = cols[9, :n], cols[10, :n], (cols[11, :n] if n >= 2 else cols[11, :0])
        tr.ds, tr.c = cols[12, :max(n - 1, 0)], cols[13, :max(n - 1, 0)]
        tr.c_d, tr.c_dd = cols[14, :max(n - 2, 0)], cols[15, :max(n - 3, 0)]
        tr.idx, tr.path_id, tr.cost_total = idx, path_id, float(cost)
        tr.constraint_passed = tr.collision_passed = True
        return tr

    def frenet_state_at_time_step(self, t: int = 1) -> FrenetState:
        """The t-th samples of the Frenet fields (intersection_planner.py:312)."""
        return FrenetState(t=self.t[t], s=self.s[t], s_d=self.s_d[t], s_dd=self.s_dd[t], s_ddd=self.s_ddd[t],
                           d=self.d[t], d_d=self.d_d[t], d_dd=self.d_dd[t], d_ddd=self.d_ddd[t])

    def state_at_time_step(self, t: int = 1) -> State:
        return State(t=self.t[t], x=self.x[t], y=self.y[t], yaw=self.yaw[t], v=self.s_d[t], a=self.s_dd[t])

    # PriorityQueue tuples (cost, traj, id) fall through to the trajectory on equal cost (jerk_space_sampling_planner.py:292)
    def __lt__(self, other):
        return (self.cost_total, self.idx) < (other.cost_total, other.idx)


class Vehicle:
    """Vehicle(length, width, longitudinal_v_max=..., ...) (intersection_planner.py:117-126) with the attributes the
    planner reads: l, w, max_curvature, max_accel, max_jerk (jerk_space_sampling_planner.py:103,175,197,200)."""

    def __init__(self, length: float, width: float, longitudinal_v_max: float = 18.0, longitudinal_a_max: float = 8.0,
                 longitudinal_jerk_max: float = 10.0, lateral_accel_max: float = 1.2, lateral_jerk_max: float = 2.0,
                 centripetal_accel_max: float = 2.0, max_curvature: float = 0.2):
        self.l, self.w = float(length), float(width)
        self.max_speed, self.max_accel, self.max_jerk = longitudinal_v_max, longitudinal_a_max, longitudinal_jerk_max
        self.max_lateral_accel, self.max_lateral_jerk = lateral_accel_max, lateral_jerk_max
        self.max_centripetal_accel = centripetal_accel_max
        self.max_curvature = max_curvature   # builder-specified

    @property
    def polygon(self) -> np.ndarray:
        """Footprint corners in the box-centre frame, clockwise from front-left (jerk_space_sampling_planner.py:218-230)."""
        l, w = self.l, self.w
        return np.array([(l / 2, w / 2), (l / 2, -w / 2), (-l / 2, -w / 2), (-l / 2, w / 2)])
