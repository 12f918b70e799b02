#!/usr/bin/env python
"""Generate tests/golden/scenario_8_3_4_565.npz from the reference's bundled scenario.

TEST INFRASTRUCTURE - runs only in the build container (needs /root/reference).  It

  * parses ``inputs/8_3_4_565/8_3_4_565_exam.xosc`` with the REFERENCE's own
    ``ReplayParser._parse_openscenario`` (code/onsite_trans/controller.py:167-221) - imported
    unmodified, with stubs for the absent ``lxml`` / ``shapely`` packages it merely imports -
    so the obstacle dictionary (string time keys, 2-dp rounding) is the genuine one;
  * converts ``8_3_4_565.xodr`` with the reference's in-tree OpenDRIVE converter
    (code/onsite_trans/opendrive2discretenet/**) under an ``lxml -> xml.etree.ElementTree`` shim;
  * finds the ego route the way ``IntersectionPlanner.plan_global_route``
    (intersection_planner.py:165-287) would: start lane = lane containing the ego, goal lane = lane
    containing the goal-box centre, connected through successors (the lane-graph helpers of
    ``common.tool`` are absent from the tree, so a breadth-first search over ``successor`` stands in),
    concatenates the lane centre lines and removes duplicate points - with a 1e-6 m tolerance,
    because adjacent lanes share end points only to ~1e-14 m.

The arrays are stored so that the GPU box (which has no /root/reference) can rebuild the exact
obstacle dictionary and reference line.
"""
import os
import sys
import types
import xml.etree.ElementTree as ET
from collections import deque

import numpy as np

REF = "/root/reference"
NAME = "8_3_4_565"


def _stubs():
    lx = types.ModuleType("lxml"); lx.etree = ET
    sys.modules["lxml"] = lx; sys.modules["lxml.etree"] = ET
    for n in ("shapely", "shapely.geometry"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["shapely.geometry"].Polygon = object
    sys.modules["shapely"].geometry = sys.modules["shapely.geometry"]


def _inside(pt, lane):
    poly = np.concatenate([lane.left_vertices, lane.right_vertices[::-1]], 0)
    x, y = pt; c = False
    for i in range(len(poly)):
        x1, y1 = poly[i]; x2, y2 = poly[(i + 1) % len(poly)]
        if (y1 > y) != (y2 > y) and x < (x2 - x1) * (y - y1) / (y2 - y1) + x1:
            c = not c
    return c


def main(out):
    _stubs()
    sys.path.insert(0, os.path.join(REF, "code"))
    from onsite_trans.controller import ReplayParser  # unmodified reference parser

    base = os.path.join(REF, "inputs", NAME)
    p = ReplayParser()
    p._parse_openscenario(os.path.join(base, NAME + "_exam.xosc"))
    p._parse_opendrive(os.path.join(base, NAME + ".xodr"))
    info = p.replay_info
    ego = info.ego_info
    goal = info.test_setting["goal"]
    lanes = {l.lane_id: l for l in info.road_info.discretelanes}

    start = [k for k, l in lanes.items() if _inside((ego["x"], ego["y"]), l)]
    gx, gy = sum(goal["x"]) / 2, sum(goal["y"]) / 2
    goal_lane = [k for k, l in lanes.items() if _inside((gx, gy), l)]
    assert len(start) == 1 and len(goal_lane) == 1, (start, goal_lane)
    prev = {start[0]: None}; dq = deque(start)
    while dq:
        k = dq.popleft()
        for nx in lanes[k].successor or []:
            if nx in lanes and nx not in prev:
                prev[nx] = k; dq.append(nx)
    route = []; k = goal_lane[0]
    while k is not None:
        route.append(k); k = prev[k]
    route = route[::-1]
    pts = np.concatenate([np.asarray(lanes[k].center_vertices, dtype=float) for k in route], 0)
    keep = [0]
    for i in range(1, len(pts)):
        if np.hypot(*(pts[i] - pts[keep[-1]])) > 1e-6:
            keep.append(i)
    pts = pts[keep]
    print("route", route, len(pts), "pts, length", np.hypot(*np.diff(pts, axis=0).T).sum())

    ids, shape, rows = [], [], []
    for vid, tr in info.vehicle_traj.items():
        ids.append(int(vid)); shape.append((tr["shape"]["length"], tr["shape"]["width"]))
        for key, st in tr.items():
            if key == "shape" or not isinstance(key, str):   # add_vehicle(t=-1, ...) leaves an empty {-1: {}} entry
                continue
            rows.append((int(vid), float(key), st["x"], st["y"], st["v"], st["a"], st["yaw"]))
    pv = np.concatenate([np.concatenate([l.left_vertices, l.right_vertices, l.center_vertices], 0) for l in info.road_info.discretelanes], 0)
    bounds = np.array([pv[:, 0].min(), pv[:, 0].max(), pv[:, 1].min(), pv[:, 1].max()])       # controller.py:279-288 (x_min, x_max, y_min, y_max)
    np.savez_compressed(
        out, bounds=bounds, route_xy=pts, route_lanes=np.array(route), ego=np.array([ego["x"], ego["y"], ego["v"], ego["a"], ego["yaw"], ego["length"], ego["width"]]),
        goal=np.array([goal["x"], goal["y"]]), dt=float(info.test_setting["dt"]), max_t=float(info.test_setting["max_t"]),
        obs_ids=np.array(ids), obs_shape=np.array(shape), obs_rows=np.array(rows))
    print("ego", ego, "goal", goal, "dt", info.test_setting["dt"], "max_t", info.test_setting["max_t"], "obstacles", len(ids), "rows", len(rows))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "..", "tests", "golden", f"scenario_{NAME}.npz"))
