"""Scenario side of the hot path: OnSite obstacle dictionary <-> dense tensors, OpenSCENARIO replay parsing, and the
seeded synthetic workloads named in BASELINE.json (configs 2-5; the reference ships data for config 1 only).

Reference anchors: obstacle dictionary format code/onsite_trans/controller.py:92-107,186-213; time-key lookup
jerk_space_sampling_planner.py:256-260; xosc parsing controller.py:167-221; synthetic shapes SURVEY.md section 8d."""
from __future__ import annotations

import math
import re
import xml.dom.minidom
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence

import numpy as np


# ----------------------------------------------------------------------------------------------------------------
# obstacle dictionary -> tensors
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class ObstacleTensors:
    state: np.ndarray     # [O, M, Tt, 3] float64  x, y, yaw
    valid: np.ndarray     # [O, M, Tt]    uint8
    size: np.ndarray      # [O, 2]        float64  length, width (not inflated)
    prob: np.ndarray      # [O, M]        float64  intention probability of each mode
    n_dict: int           # len(obstacles) as seen by has_collision's early-out (jerk_space_sampling_planner.py:241)
    ids: list = field(default_factory=list)
    version: int = 0      # bump after mutating the arrays in place: the planners' table cache keys on (object, version)

    @property
    def shape(self):
        return self.state.shape[:3]


def time_key(step: int, dt: float) -> str:
    """The reference's lookup key for absolute step ``step`` (jerk_space_sampling_planner.py:256-257)."""
    return str(round(step * dt, 2))


def pack_obstacle_dict(obstacles: Optional[dict], dt: float, n_total: int) -> ObstacleTensors:
    """{id: {"shape": {"length","width"}, "<t>": {"x","y","v","a","yaw"}}} -> tensors with M = 1, prob = 1.
    A state exists at step k iff the dictionary holds the key ``str(round(k*dt, 2))``."""
    obstacles = obstacles or {}
    ids = list(obstacles.keys())
    O = len(ids)
    state = np.zeros((O, 1, n_total, 3)); valid = np.zeros((O, 1, n_total), dtype=np.uint8); size = np.zeros((O, 2))
    keys = [time_key(k, dt) for k in range(n_total)]
    for o, vid in enumerate(ids):
        ob = obstacles[vid]
        size[o] = (ob["shape"]["length"], ob["shape"]["width"])
        for k, key in enumerate(keys):
            st = ob.get(key)
            if st:
                state[o, 0, k] = (st["x"], st["y"], st["yaw"])
                valid[o, 0, k] = 1
    return ObstacleTensors(state, valid, size, np.ones((O, 1)), O, ids)


def tensors_to_dict(obs: ObstacleTensors, dt: float, mode: int = 0) -> dict:
    """Inverse of pack_obstacle_dict for one mode (used to drive the reference-shaped planner interface from synthetic data)."""
    out = {}
    O, M, Tt = obs.shape
    for o in range(O):
        d = {"shape": {"length": float(obs.size[o, 0]), "width": float(obs.size[o, 1])}}
        for k in range(Tt):
            if obs.valid[o, mode, k]:
                x, y, yaw = obs.state[o, mode, k]
                d[time_key(k, dt)] = {"x": float(x), "y": float(y), "v": 0.0, "a": 0.0, "yaw": float(yaw)}
        out[o + 1] = d
    return out


# ----------------------------------------------------------------------------------------------------------------
# replay scenario (OnSite / OpenSCENARIO)
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class ReplayScenario:
    name: str
    ego: dict                      # x, y, v, a, yaw, length, width
    vehicle_traj: dict             # the obstacle dictionary (string time keys)
    goal: dict                     # {"x": [lo, hi], "y": [lo, hi]}
    dt: float
    max_t: float
    route_xy: Optional[np.ndarray] = None   # sparse centre line of the ego route [K, 2]
    bounds: Optional[tuple] = None          # (x_min, x_max, y_min, y_max) of the map's lane vertices (controller.py:279-288); None = no map

    def observation(self) -> dict:
        """Initial observation dict in the reference's schema (code/onsite_trans/observation.py:25-46)."""
        return {"vehicle_info": {"ego": dict(self.ego)}, "light_info": {},
                "test_setting": {"scenario_name": self.name, "scenario_type": "replay", "t": 0.0, "dt": self.dt, "max_t": self.max_t,
                                 "goal": self.goal, "end": -1}}


def _wrap_pi(a: float) -> float:
    return (a + np.pi) % (2 * np.pi) - np.pi


def parse_openscenario(path: str, name: Optional[str] = None) -> ReplayScenario:
    """Replay file -> ReplayScenario with the value conventions of ReplayParser._parse_openscenario
    (controller.py:167-221): first <Dimensions> is the ego, one <Act> per background vehicle (ids 1..K), vertex times
    rounded to 3 dp and used as string keys, x/y/v/a rounded to 2 dp, yaw wrapped to [-pi, pi) and rounded to 3 dp,
    v and a by finite differences with the last value repeated / zero; dt and max_t from the keys (controller.py:124-133)."""
    doc = xml.dom.minidom.parse(path).documentElement
    dims = [(float(n.getAttribute("length")), float(n.getAttribute("width"))) for n in doc.getElementsByTagName("Dimensions")]
    private = doc.getElementsByTagName("Private")[0]
    ego_init, goal_init = private.childNodes[3].data, private.childNodes[5].data        # positional, like controller.py:177-178,215
    ev, ex, ey, eh = [float(tok.split("=")[1]) for tok in ego_init.split(",")]
    ego = {"length": dims[0][0], "width": dims[0][1], "x": ex, "y": ey, "v": abs(ev), "a": 0, "yaw": _wrap_pi(eh)}
    gl = [float(v) for v in re.findall(r"-*\d+\.\d+", goal_init)]
    goal = {"x": gl[:2], "y": gl[2:]}
    traj: Dict[int, dict] = {}
    last_keys = None
    for vid, act in enumerate(doc.getElementsByTagName("Act"), start=1):
        ts, xs, ys, hs = [], [], [], []
        for vtx in act.getElementsByTagName("Vertex"):
            ts.append(round(float(vtx.getAttribute("time")), 3))
            wp = vtx.getElementsByTagName("WorldPosition")[0]
            xs.append(float(wp.getAttribute("x"))); ys.append(float(wp.getAttribute("y")))
            hs.append(_wrap_pi(float(wp.getAttribute("h"))))
        tdiff = np.diff(ts)
        v = list(np.around(np.sqrt(np.diff(xs) ** 2 + np.diff(ys) ** 2) / tdiff, 2))
        v.append(v[-1])
        a = list(np.around(np.diff(v) / tdiff, 2))
        a.append(0.0)
        entry = {"shape": {"length": dims[vid][0], "width": dims[vid][1]}}
        for t, x, y, vv, aa, h in zip(ts, xs, ys, v, a, hs):
            entry[str(t)] = {"x": round(x, 2), "y": round(y, 2), "v": round(vv, 2), "a": round(aa, 2), "yaw": round(h, 3)}
        traj[vid] = entry
        last_keys = [str(t) for t in ts]
    max_t = max(float([k for k in tr if k != "shape"][-1]) for tr in traj.values()) if traj else 0.0
    dt = float(np.around(float(last_keys[-1]) - float(last_keys[-2]), 3)) if last_keys and len(last_keys) > 1 else 0.1
    return ReplayScenario(name or path, ego, traj, goal, dt, max_t)


def scenario_from_fixture(npz_path: str, name: str = "8_3_4_565") -> ReplayScenario:
    """Rebuild the scenario stored by oracle/make_fixture_scenario.py (tests/golden/scenario_*.npz)."""
    z = np.load(npz_path)
    e = z["ego"]
    ego = {"x": float(e[0]), "y": float(e[1]), "v": float(e[2]), "a": float(e[3]), "yaw": float(e[4]), "length": float(e[5]), "width": float(e[6])}
    traj: Dict[int, dict] = {}
    for vid, (l, w) in zip(z["obs_ids"], z["obs_shape"]):
        traj[int(vid)] = {"shape": {"length": float(l), "width": float(w)}}
    for vid, t, x, y, v, a, yaw in z["obs_rows"]:
        traj[int(vid)][str(float(t))] = {"x": float(x), "y": float(y), "v": float(v), "a": float(a), "yaw": float(yaw)}
    goal = {"x": [float(v) for v in z["goal"][0]], "y": [float(v) for v in z["goal"][1]]}
    bounds = tuple(float(v) for v in z["bounds"]) if "bounds" in z.files else None
    return ReplayScenario(name, ego, traj, goal, float(z["dt"]), float(z["max_t"]), np.asarray(z["route_xy"], dtype=np.float64), bounds)


# ----------------------------------------------------------------------------------------------------------------
# synthetic workloads - all seeded, all float64
# ----------------------------------------------------------------------------------------------------------------
def corner_refline(straight: float = 30.0, radius: float = 15.0, step: float = 0.5, origin=(-30.0, -20.0)) -> np.ndarray:
    """straight + 90 degree left arc + straight, sampled every ``step`` metres -> [K, 2]."""
    ox, oy = origin
    n1 = int(round(straight / step))
    pts = [(ox + i * step, oy) for i in range(n1)]
    cx, cy = ox + straight, oy + radius
    na = int(round(radius * math.pi / 2 / step))
    for i in range(na):
        a = -math.pi / 2 + (math.pi / 2) * i / na
        pts.append((cx + radius * math.cos(a), cy + radius * math.sin(a)))
    ex, ey = cx + radius, cy
    pts += [(ex, ey + i * step) for i in range(n1 + 1)]
    return np.array(pts, dtype=np.float64)


def _st(s0, v0, a0, j, t):
    return s0 + v0 * t + 0.5 * a0 * t * t + (1 / 6.0) * j * t * t * t


def _vt(v0, a0, j, t):
    return v0 + a0 * t + 0.5 * j * t * t


def _state_ok(s, v, a, s_prev, a_max, v_max):
    return ~((a > a_max - 1e-9) | (a < -(a_max - 1e-9)) | (v > v_max - 1e-9) | (v < -1e-9) | (s < s_prev - 1e-9))


def random_seeds(n: int, rng: np.random.Generator, s0=0.0, v0=6.0, a0=0.0, total_time=5.0, jerk=10.0, a_max=8.0, v_max=18.0) -> np.ndarray:
    """n seeds drawn i.i.d. from the reference's maneuver grammar with continuous parameters, rejection-filtered by
    check_state after every segment and |j5| <= 0.3*jerk (polyvt_sampling.py:88-113,195-254)."""
    out = np.zeros((0, 10))
    T = total_time
    while len(out) < n:
        m = max(4 * (n - len(out)), 1024)
        j1 = rng.uniform(-jerk, jerk, m)
        t1 = rng.uniform(1.0, 2.5, m)
        t2 = rng.uniform(0.0, 1.0, m) * (T - 2.4 - t1)
        j3 = rng.uniform(0.0, 0.6 * jerk, m) * np.where(j1 > 0, -1.0, 1.0)
        t3 = rng.uniform(0.0, 1.0, m) * (T - 2.0 - t1 - t2)
        t4 = rng.uniform(0.0, 1.0, m) * (T - 1.5 - t1 - t2 - t3)
        s1, v1, a1 = _st(s0, v0, a0, j1, t1), _vt(v0, a0, j1, t1), a0 + j1 * t1
        ok = _state_ok(s1, v1, a1, s0, a_max, v_max)
        s2, v2, a2 = _st(s1, v1, a1, 0.0, t2), _vt(v1, a1, 0.0, t2), a1 + 0.0 * t2
        ok &= _state_ok(s2, v2, a2, s1, a_max, v_max)
        s3, v3, a3 = _st(s2, v2, a2, j3, t3), _vt(v2, a2, j3, t3), a2 + j3 * t3
        ok &= _state_ok(s3, v3, a3, s2, a_max, v_max)
        s4, v4, a4 = _st(s3, v3, a3, 0.0, t4), _vt(v3, a3, 0.0, t4), a3 + 0.0 * t4
        ok &= _state_ok(s4, v4, a4, s3, a_max, v_max)
        t5 = T - t1 - t2 - t3 - t4
        j5 = (0.0 - a4) / t5
        ok &= np.abs(j5) <= 0.3 * jerk
        s5, v5, a5 = _st(s4, v4, a4, j5, t5), _vt(v4, a4, j5, t5), a4 + j5 * t5
        ok &= _state_ok(s5, v5, a5, s4, a_max, v_max)
        z = np.zeros(m)
        cand = np.stack([j1, t1, z, t2, j3, t3, z, t4, j5, t5], 1)[ok]
        out = np.concatenate([out, cand], 0)
    return np.ascontiguousarray(out[:n])


def random_obstacles(O: int, M: int, n_steps_total: int, dt: float, rng: np.random.Generator, box: float = 50.0,
                     clear_xy: Optional[np.ndarray] = None, clear_radius: float = 7.0, clear_steps: int = 16) -> ObstacleTensors:
    """O obstacles x M modes: constant-speed / constant-yaw-rate rollouts that share the start pose and differ in yaw rate.
    An obstacle whose any mode comes within ``clear_radius`` of a ``clear_xy`` point during the first ``clear_steps`` steps
    is redrawn, so that the candidates are not all in collision from the first sample on."""
    state = np.zeros((O, M, n_steps_total, 3))
    size = np.stack([rng.uniform(3.6, 5.4, O), rng.uniform(1.7, 2.0, O)], 1) if O else np.zeros((0, 2))
    prob = rng.dirichlet(np.ones(M), O) if O else np.zeros((0, M))
    for o in range(O):
        for _ in range(1000):
            x0, y0 = rng.uniform(-box, box, 2)
            v = rng.uniform(0.0, 12.0); yaw0 = rng.uniform(-math.pi, math.pi)
            omega = rng.normal(0.0, 0.1, M)
            x = np.full(M, x0); y = np.full(M, y0); yaw = np.full(M, yaw0)
            for k in range(n_steps_total):
                state[o, :, k, 0], state[o, :, k, 1] = x, y
                state[o, :, k, 2] = (yaw + np.pi) % (2 * np.pi) - np.pi
                x = x + v * np.cos(yaw) * dt
                y = y + v * np.sin(yaw) * dt
                yaw = yaw + omega * dt
            if clear_xy is None:
                break
            early = state[o, :, :clear_steps, :2].reshape(-1, 1, 2)
            if np.min(np.hypot(early[..., 0] - clear_xy[None, :, 0], early[..., 1] - clear_xy[None, :, 1])) > clear_radius:
                break
    return ObstacleTensors(state, np.ones((O, M, n_steps_total), dtype=np.uint8), size, prob, O, list(range(1, O + 1)))


@dataclass
class SyntheticCycle:
    name: str
    route_xy: np.ndarray
    seeds: np.ndarray             # [Ns, 10]
    targets: np.ndarray           # [W]
    frenet0: tuple                # s, s_d, s_dd, d, d_d, d_dd
    obstacles: ObstacleTensors
    time_step_now: int
    final_time_step: int
    plan_horizon: float
    ego_length: float = 4.924
    ego_width: float = 1.872

    @property
    def n_candidates(self) -> int:
        return len(self.seeds) * len(self.targets)


def dense_cycle(n_seeds: int = 16384, n_widths: int = 4, O: int = 32, M: int = 3, plan_horizon: float = 5.0, seed: int = 20251019,
                straight: Optional[float] = None, name: str = "C3", box: Optional[float] = None) -> SyntheticCycle:
    """BASELINE config 3 (defaults: 65,536 candidates x 50 steps x 32 obstacles x 3 modes) and, with
    n_seeds=131072, n_widths=8, O=64, plan_horizon=8.0, seed=20251020, config 5."""
    rng = np.random.default_rng(seed)
    P = int(plan_horizon / 0.1) + 1
    straight = straight if straight is not None else (30.0 if plan_horizon <= 5.0 else 60.0)
    route = corner_refline(straight=straight)
    seeds = random_seeds(n_seeds, rng, total_time=plan_horizon)
    targets = np.linspace(-0.75, 0.75, n_widths)
    # constant obstacle density: the box grows with sqrt(O) and with the distance covered in the horizon
    box = box if box is not None else 70.0 * math.sqrt(max(O, 1) / 32.0) * (plan_horizon / 5.0)
    obs = random_obstacles(O, M, P, 0.1, rng, box=box, clear_xy=route[:int(14.0 / 0.5)])
    return SyntheticCycle(name, route, seeds, targets, (0.0, 6.0, 0.0, 0.2, 0.0, 0.0), obs, 0, P + 50, plan_horizon)


def synthetic_intersection(seed: int, dt: float = 0.1, n_cycles: int = 100) -> ReplayScenario:
    """One scenario of the synthetic sweep (BASELINE config 4): a 4-arm intersection, the ego route is a left turn, a
    straight crossing or a right turn (seed % 3), 3..12 background vehicles drive the other movements at constant speed
    and appear / disappear at seeded times."""
    rng = np.random.default_rng(1_000_003 * (seed + 1))
    arm, lane = 45.0, 1.75
    kind = seed % 3

    def movement(entry: int, turn: int, step: float = 0.5) -> np.ndarray:
        # entry arm 0..3 (from west, south, east, north), turn: 0 left, 1 straight, 2 right; built in the west-entry frame
        if turn == 1:
            pts = [(-arm + i * step, -lane) for i in range(int(2 * arm / step) + 1)]
        else:
            # right-hand traffic: eastbound lane y = -lane, northbound lane x = +lane, southbound lane x = -lane
            r = 2.0 * lane + (6.0 if turn == 0 else 2.0)
            sign = 1.0 if turn == 0 else -1.0
            cx, cy = sign * lane - r, -lane + sign * r
            n1 = int((cx + arm) / step)
            pts = [(cx - (n1 - i) * step, -lane) for i in range(n1)]
            na = max(int(r * math.pi / 2 / step), 4)
            for i in range(na):
                a = -sign * math.pi / 2 + sign * (math.pi / 2) * i / na
                pts.append((cx + r * math.cos(a), cy + r * math.sin(a)))
            ex, ey = cx + r, cy
            n2 = int((arm - abs(ey)) / step)
            pts += [(ex, ey + sign * i * step) for i in range(n2 + 1)]
        p = np.array(pts)
        ang = entry * math.pi / 2
        c, s = math.cos(ang), math.sin(ang)
        return np.stack([c * p[:, 0] - s * p[:, 1], s * p[:, 0] + c * p[:, 1]], 1)

    route = movement(0, kind)
    n_total = n_cycles + 22
    O = int(rng.integers(3, 13))
    traj = {}
    for vid in range(1, O + 1):
        path = movement(int(rng.integers(0, 4)), int(rng.integers(0, 3)))
        cum = np.concatenate(([0.0], np.cumsum(np.hypot(*np.diff(path, axis=0).T))))
        speed = rng.uniform(2.0, 10.0)
        s_start = rng.uniform(0.0, 0.5 * cum[-1])
        k0 = int(rng.integers(0, n_total // 2)); k1 = int(rng.integers(k0 + 10, n_total + 1))
        entry = {"shape": {"length": float(rng.uniform(3.6, 5.4)), "width": float(rng.uniform(1.7, 2.0))}}
        for k in range(k0, k1):
            s = s_start + speed * (k - k0) * dt
            if s >= cum[-1]:
                break
            i = min(int(np.searchsorted(cum, s, side="right")) - 1, len(path) - 2)
            u = (s - cum[i]) / (cum[i + 1] - cum[i])
            x, y = path[i] + u * (path[i + 1] - path[i])
            yaw = math.atan2(path[i + 1, 1] - path[i, 1], path[i + 1, 0] - path[i, 0])
            entry[time_key(k, dt)] = {"x": round(float(x), 2), "y": round(float(y), 2), "v": round(speed, 2), "a": 0.0, "yaw": round(yaw, 3)}
        traj[vid] = entry
    yaw0 = math.atan2(route[1, 1] - route[0, 1], route[1, 0] - route[0, 0])
    ego = {"x": float(route[4, 0]), "y": float(route[4, 1]), "v": float(rng.uniform(2.0, 10.0)), "a": 0, "yaw": yaw0, "length": 4.924, "width": 1.872}
    goal = {"x": [float(route[-1, 0]) - 5.0, float(route[-1, 0]) + 5.0], "y": [float(route[-1, 1]) - 1.0, float(route[-1, 1]) + 1.0]}
    return ReplayScenario(f"synth_{seed:04d}", ego, traj, goal, dt, round((n_total - 1) * dt, 3), route)


def perturbed_scenario(base: ReplayScenario, seed: int) -> ReplayScenario:
    """Seeded variant of a replay scenario (BASELINE config 2: 8_3_4_565 + 7 variants): obstacle start times shifted by
    U(-1, 1) s (whole steps), lateral offsets N(0, 0.3 m), ego initial speed U(2, 10)."""
    if seed == 0:
        return base
    rng = np.random.default_rng(7000 + seed)
    traj = {}
    for vid, tr in base.vehicle_traj.items():
        shift = int(round(rng.uniform(-1.0, 1.0) / base.dt))
        off = rng.normal(0.0, 0.3)
        entry = {"shape": dict(tr["shape"])}
        for key, st in tr.items():
            if key == "shape" or not isinstance(key, str):
                continue
            k = int(round(float(key) / base.dt)) + shift
            if k < 0:
                continue
            nx, ny = -math.sin(st["yaw"]), math.cos(st["yaw"])
            entry[time_key(k, base.dt)] = {"x": round(st["x"] + off * nx, 2), "y": round(st["y"] + off * ny, 2), "v": st["v"], "a": st["a"], "yaw": st["yaw"]}
        traj[vid] = entry
    ego = dict(base.ego); ego["v"] = float(rng.uniform(2.0, 10.0))
    return ReplayScenario(f"{base.name}_v{seed}", ego, traj, base.goal, base.dt, base.max_t, base.route_xy, base.bounds)


def multimodal_predictions(sc: ReplayScenario, M: int = 3, seed: int = 0, plan_steps: int = 50) -> ObstacleTensors:
    """Builder-specified stand-in for the DLII predictor's output (the reference publishes no predictor code, SURVEY.md
    section 0.3): per background vehicle M intention hypotheses over the whole scenario - mode 0 is the recorded trajectory
    (lane keeping), mode m > 0 is the recorded trajectory bent about its first state by a constant yaw rate of
    +-0.12 rad/s * ceil(m / 2) (turn hypotheses) - with intention probabilities p0 ~ U(0.5, 0.8) for mode 0 and a
    Dirichlet split of the rest.  Covers final_time_step + plan_steps + 2 steps like the planner's own packing."""
    final = int(sc.max_t / sc.dt)
    n_total = final + plan_steps + 2
    gt = pack_obstacle_dict(sc.vehicle_traj, sc.dt, n_total)
    O = gt.state.shape[0]
    state = np.repeat(gt.state, M, axis=1)
    valid = np.repeat(gt.valid, M, axis=1)
    rng = np.random.default_rng(9000 + seed)
    prob = np.zeros((O, M))
    for o in range(O):
        p0 = rng.uniform(0.5, 0.8) if M > 1 else 1.0
        prob[o, 0] = p0
        if M > 1:
            prob[o, 1:] = rng.dirichlet(np.ones(M - 1)) * (1.0 - p0)
        ks = np.nonzero(gt.valid[o, 0])[0]
        if len(ks) == 0:
            continue
        k0 = ks[0]
        x0, y0 = gt.state[o, 0, k0, :2]
        for m in range(1, M):
            omega = (0.12 if m % 2 else -0.12) * ((m + 1) // 2)
            ang = omega * (ks - k0) * sc.dt
            dx, dy = gt.state[o, 0, ks, 0] - x0, gt.state[o, 0, ks, 1] - y0
            state[o, m, ks, 0] = x0 + np.cos(ang) * dx - np.sin(ang) * dy
            state[o, m, ks, 1] = y0 + np.sin(ang) * dx + np.cos(ang) * dy
            state[o, m, ks, 2] = (gt.state[o, 0, ks, 2] + ang + np.pi) % (2 * np.pi) - np.pi
    return ObstacleTensors(state, valid, gt.size, prob, gt.n_dict, gt.ids)
