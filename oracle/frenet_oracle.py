"""CPU oracle of the Cartesian -> Frenet projection (TEST INFRASTRUCTURE).

Restates what the reference calls ``FrenetState().from_state(ego_state, ref_ego_lane_pts)``
(code/intersection_local_planner/intersection_planner.py:140-142, 308-310) on the reference line re-discretised every 0.1 m
(``generate_frenet_frame``, jerk_space_sampling_planner.py:364-371).  The class itself (``common.scenario.frenet.FrenetState``)
is absent from the published tree => PARITY UNPINNED for the projection formula; the builder specification (SURVEY.md 8c /
DESIGN.md section 2) is: nearest polyline sample (first minimum of the squared distance, like ``numpy.argmin``), arc length =
sum of the chord lengths up to that sample plus the tangential offset, lateral offset along the left-hand normal, speed and
acceleration split by the heading error.  Written as scalar loops over this module's own spline sampling, independently of
the product's vectorised host code (dlii_jssp_b200/frenet.py) and of the CUDA kernel (project_states_kernel)."""
from __future__ import annotations

import math

import numpy as np

from . import jssp_oracle as O


def rediscretise(spline: "O.CubicSpline2D", step: float = 0.1):
    """[(x, y, yaw)] at s = 0, step, 2*step, ... < s_end   (np.arange(0, s_list[-1], step), jerk_space_sampling_planner.py:366)."""
    pts = []
    for s in np.arange(0, spline.s_list[-1], step):
        x, y = spline.calc_position(float(s))
        pts.append((float(x), float(y), float(spline.calc_yaw(float(s)))))
    return pts


def project_state(pts, x, y, yaw, v, a):
    """-> (s, s_d, s_dd, d, d_d, d_dd)"""
    best, bi = math.inf, 0
    for i, (px, py, _) in enumerate(pts):
        d2 = (px - x) ** 2 + (py - y) ** 2
        if d2 < best:
            best, bi = d2, i
    chords = np.array([math.hypot(pts[i + 1][0] - pts[i][0], pts[i + 1][1] - pts[i][1]) for i in range(bi)], dtype=np.float64)
    arc = float(chords.sum()) if bi > 0 else 0.0                       # the product accumulates with numpy's sum as well
    rx, ry, ryaw = pts[bi]
    ex, ey = x - rx, y - ry
    tx, ty = math.cos(ryaw), math.sin(ryaw)
    dyaw = yaw - ryaw
    return (arc + (ex * tx + ey * ty), v * math.cos(dyaw), a * math.cos(dyaw), -ex * ty + ey * tx, v * math.sin(dyaw), a * math.sin(dyaw))
