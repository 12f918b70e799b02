"""Batched, device-resident replay (SURVEY.md section 8 rows f1 + f2): many independent scenarios advance in lock-step, one
plan cycle per step, and the per-cycle hand-off of the reference driver (code/__main__.py:124-167, env.py:24-30,
intersection_planner.py:312, controller.py:303,338-380) runs on the device, so that a replay sweep costs two kernel launches
per step for the whole batch instead of one host round trip per scenario and cycle.

``pack_scenario`` is the scenario -> tensor packer (obstacle dictionary -> dense tables, route -> spline coefficient table,
initial Frenet state, the clock tables that carry the reference's float/str time-key semantics); ``save_packed`` /
``load_packed`` are its on-disk cache (.npz), so a sweep parses and packs every scenario once.

Results are ``replay.ReplayResult`` objects, identical to what ``replay.run_replay`` returns for the same scenario.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import JsspError, JsspCycleRecord, JsspFopIn, JsspScenarioIn, JsspScenarioResult, FLAG_FORCE_THREAD, FLAG_FORCE_WARP, FLAG_FORCE_GROUP, FLAG_FORCE_TILED, FLAG_LOCKSTEP, FLAG_PERSISTENT, FLAG_CHUNKED, FLAG_BEST_FIRST, FLAG_SMALL_CTAS
from .engine import JsspConfig
from .frenet import FrenetState, State
from .geometry import CubicSpline2D
from .replay import END_RUNNING, ReplayResult, end_status, goal_pose_of
from .scenario import ReplayScenario, pack_obstacle_dict, time_key

RECORD_DTYPE = np.dtype([("best_idx", "<i4"), ("n_candidates", "<i4"), ("best_cost", "<f8"), ("t", "<f8"), ("x", "<f8"), ("y", "<f8"),
                         ("v", "<f8"), ("a", "<f8"), ("yaw", "<f8")])
RESULT_DTYPE = np.dtype([("end", "<f8"), ("cycles", "<i4"), ("seed_overflow", "<i4")])
assert RECORD_DTYPE.itemsize == C.sizeof(JsspCycleRecord) and RESULT_DTYPE.itemsize == C.sizeof(JsspScenarioResult)


@dataclass
class PackedScenario:
    """Everything the device needs to replay one scenario (all float64 / uint8 / int32 NumPy arrays)."""
    name: str
    knot_s: np.ndarray            # [K]
    coef: np.ndarray              # [K-1, 8]
    state: np.ndarray             # [O, M, Tt, 3]
    valid: np.ndarray             # [O, M, Tt]
    size: np.ndarray              # [O, 2]
    prob: np.ndarray              # [O, M]
    n_dict: int
    frenet0: np.ndarray           # [6]
    ego: np.ndarray               # [7] x, y, v, a, yaw, length, width
    dt: float
    max_t: float
    final_time_step: int
    cycle_time: np.ndarray        # [C+1] float64
    cycle_now: np.ndarray         # [C+1] int32
    cycle_end_step: np.ndarray    # [C+1] int32
    goal_pose: np.ndarray         # [3] x, y, yaw
    initial_end: float            # end status before the first cycle (a scenario can be over at t = 0)
    bounds: Optional[np.ndarray] = None       # [4] x_min, x_max, y_min, y_max of the map (None: no map, no bounds test)
    gt_state: Optional[np.ndarray] = None     # ground truth for the end-status test when the planner sees predictions (M > 1)
    gt_valid: Optional[np.ndarray] = None
    gt_size: Optional[np.ndarray] = None
    polyline: Optional[np.ndarray] = None     # [n, 6] x, y, yaw, arc, cos yaw, sin yaw: the re-discretised reference line (kinematics mode)

    @property
    def n_cycles(self) -> int:
        return len(self.cycle_time) - 1

    def c_view(self, kinematics: bool = False) -> JsspScenarioIn:
        """The scenario as the ``jssp_scenario_in`` of the C ABI (built once and kept with the arrays it points into: loading
        a packed scenario again costs one library call)."""
        cache = self.__dict__.setdefault("_cview", {})
        hit = cache.get(kinematics)
        if hit is not None:
            return hit[0]
        p = self
        O, M, Tt, _ = p.state.shape
        keep = [_c(p.knot_s, np.float64), _c(p.coef, np.float64), _c(p.state, np.float64), _c(p.valid, np.uint8), _c(p.size, np.float64),
                _c(p.prob, np.float64), _c(p.gt_state, np.float64), _c(p.gt_valid, np.uint8), _c(p.gt_size, np.float64),
                _c(p.cycle_time, np.float64), _c(p.cycle_now, np.int32), _c(p.cycle_end_step, np.int32), _c(p.polyline, np.float64)]
        ptr = [None if a is None else a.ctypes.data for a in keep]
        si = JsspScenarioIn()
        si.knot_s, si.coef, si.K = ptr[0], ptr[1], len(p.knot_s)
        si.state, si.valid, si.size, si.prob = ptr[2], ptr[3], ptr[4], ptr[5]
        si.O, si.M, si.Tt, si.n_dict = O, M, Tt, int(p.n_dict)
        si.gt_state, si.gt_valid, si.gt_size = ptr[6], ptr[7], ptr[8]
        si.gt_O = 0 if p.gt_state is None else p.gt_state.shape[0]
        for i in range(6):
            si.frenet0[i] = float(p.frenet0[i])
        si.max_t, si.final_time_step, si.n_cycles = p.max_t, p.final_time_step, p.n_cycles
        si.cycle_time, si.cycle_now, si.cycle_end_step = ptr[9], ptr[10], ptr[11]
        si.goal_x, si.goal_y, si.goal_yaw = [float(v) for v in p.goal_pose]
        si.ego_length, si.ego_width = float(p.ego[5]), float(p.ego[6])
        si.ego0[0], si.ego0[1], si.ego0[2], si.ego0[3] = float(p.ego[0]), float(p.ego[1]), float(p.ego[4]), float(p.ego[2])   # x, y, yaw, v
        si.has_bounds = 0 if p.bounds is None else 1
        if p.bounds is not None:
            for i in range(4):
                si.bounds[i] = float(p.bounds[i])
        if kinematics:
            if p.polyline is None:
                raise ValueError(f"scenario {p.name}: kinematics mode needs pack_scenario(..., kinematics=True)")
            from .planner import PURE_PURSUIT
            si.ego_update_mode, si.polyline, si.n_polyline = 1, ptr[12], len(p.polyline)
            si.pp_kv, si.pp_ld0 = PURE_PURSUIT["kv_ld0"]
            si.pp_wheel_base = float(p.ego[5]) / PURE_PURSUIT["coff_wheel_base"]
        cache[kinematics] = (si, keep)
        return si


def clock_tables(dt: float, n_cycles: int, n_total: int):
    """The replay clock as the reference advances it: t <- float(t + dt) (controller.py:303); time_step_now = int(t / dt)
    (jerk_space_sampling_planner.py:300, float truncation included); the end-status test looks the background vehicles up by
    the key str(np.around(t, 3)) (controller.py:345) - resolved here to the step of the packed table with that key."""
    key_to_step = {time_key(k, dt): k for k in range(n_total)}
    times, now, end = np.zeros(n_cycles + 1), np.zeros(n_cycles + 1, dtype=np.int32), np.full(n_cycles + 1, -1, dtype=np.int32)
    t = 0.0
    for c in range(n_cycles + 1):
        times[c], now[c] = t, int(t / dt)
        end[c] = key_to_step.get(str(np.around(t, 3)), -1)
        t = float(t + dt)
    return times, now, end


def polyline_table(ref: np.ndarray) -> np.ndarray:
    """The columns FrenetState.from_state reads of the re-discretised reference line ``ref`` [n, >=3] = x, y, yaw, with the
    arc length up to every point accumulated exactly as from_state does (``seg[:i].sum()``) and the unit tangent by math.cos /
    math.sin, so that the device projection sees the host's numbers."""
    import math
    seg = np.hypot(np.diff(ref[:, 0]), np.diff(ref[:, 1]))
    arc = np.array([float(seg[:i].sum()) for i in range(len(ref))])
    return np.column_stack([ref[:, 0], ref[:, 1], ref[:, 2], arc, [math.cos(v) for v in ref[:, 2]], [math.sin(v) for v in ref[:, 2]]])


def pack_scenario(sc: ReplayScenario, plan_steps: int = 50, max_cycles: Optional[int] = None, predictions=None,
                  kinematics: bool = False) -> PackedScenario:
    """``predictions`` (ObstacleTensors, M >= 1) replaces the ground-truth dictionary as what the planner sees; the
    end-status collision test always uses the dictionary.  ``kinematics`` also packs the re-discretised reference line the
    "kinematics" ego-update mode projects the ego on every cycle."""
    spline = CubicSpline2D(sc.route_xy[:, 0], sc.route_xy[:, 1])
    ref = spline.sample(np.arange(0, spline.s_list[-1], 0.1))                       # generate_frenet_frame, jssp.py:366-371
    e = sc.ego
    fr0 = FrenetState().from_state(State(0.0, e["x"], e["y"], e["yaw"], e["v"], e["a"]), ref).as_tuple()   # intersection_planner.py:309-310
    final = int(sc.max_t / sc.dt)                                                   # jssp.py:244
    n_total = final + plan_steps + 2
    gt = pack_obstacle_dict(sc.vehicle_traj, sc.dt, n_total)
    n_cycles = final + 2 if max_cycles is None else int(max_cycles)
    times, now, end = clock_tables(sc.dt, n_cycles, n_total)
    obs = sc.observation()
    gp = goal_pose_of(sc)
    p = PackedScenario(sc.name, spline.s_list.copy(), spline.coefficient_table(), gt.state, gt.valid, gt.size, gt.prob, gt.n_dict,
                       np.array(fr0, dtype=np.float64), np.array([e["x"], e["y"], e["v"], e["a"], e["yaw"], e["length"], e["width"]], dtype=np.float64),
                       float(sc.dt), float(sc.max_t), final, times, now, end, np.array(gp, dtype=np.float64),
                       float(end_status(obs, sc, None)),                        # _get_initial_observation: no goal pose yet
                       None if sc.bounds is None else np.array(sc.bounds, dtype=np.float64))
    if kinematics:
        p.polyline = polyline_table(ref)
    if predictions is not None:
        assert predictions.state.shape[2] >= n_total, "predictions must cover final_time_step + horizon + 2 steps"
        p.gt_state, p.gt_valid, p.gt_size = gt.state[:, 0].copy(), gt.valid[:, 0].copy(), gt.size
        p.state, p.valid, p.size, p.prob = predictions.state[:, :, :n_total].copy(), predictions.valid[:, :, :n_total].copy(), predictions.size, predictions.prob
        p.n_dict = predictions.n_dict
    return p


_ARRAY_FIELDS = ("knot_s", "coef", "state", "valid", "size", "prob", "frenet0", "ego", "cycle_time", "cycle_now", "cycle_end_step", "goal_pose",
                 "bounds", "gt_state", "gt_valid", "gt_size", "polyline")


def save_packed(path: str, packed: Sequence[PackedScenario]):
    """On-disk cache of packed scenarios (one .npz for a whole shard)."""
    out = {"names": np.array([p.name for p in packed]), "n": np.array(len(packed))}
    for i, p in enumerate(packed):
        for f in _ARRAY_FIELDS:
            v = getattr(p, f)
            if v is not None:
                out[f"{i}/{f}"] = v
        out[f"{i}/scalars"] = np.array([p.n_dict, p.dt, p.max_t, p.final_time_step, p.initial_end], dtype=np.float64)
    np.savez_compressed(path, **out)


def load_packed(path: str) -> List[PackedScenario]:
    z = np.load(path)
    res = []
    for i, name in enumerate(z["names"]):
        sc = z[f"{i}/scalars"]
        kw = {f: (z[f"{i}/{f}"] if f"{i}/{f}" in z.files else None) for f in _ARRAY_FIELDS}
        res.append(PackedScenario(name=str(name), n_dict=int(sc[0]), dt=float(sc[1]), max_t=float(sc[2]), final_time_step=int(sc[3]),
                                  initial_end=float(sc[4]), **kw))
    return res


def _c(a: Optional[np.ndarray], dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


class BatchReplay:
    """Owner of a ``jssp_batch`` handle: capacity for ``max_scenarios`` slots x ``max_cycles`` plan cycles."""

    def __init__(self, config: Optional[JsspConfig] = None, device: int = 0, max_scenarios: int = 64, max_cycles: int = 128, fop=None,
                 max_accel: float = 8.0, max_jerk: float = 10.0, ego_update_mode: str = "planner_expected_value"):
        """``fop`` = a FrenetOptimalPlannerSettings: the batch replays with the FOP baseline planner instead of JSSP (``config`` must
        then be the FOP configuration, see ``fop_batch_config``); ``max_accel`` / ``max_jerk`` = Vehicle.max_accel / max_jerk.
        ``ego_update_mode`` = parameters["sim_config"]["ego_update_mode"]: "planner_expected_value" or "kinematics" (the
        scenarios must then be packed with ``kinematics=True``)."""
        assert ego_update_mode in ("planner_expected_value", "kinematics"), ego_update_mode
        self.ego_update_mode = ego_update_mode
        if not torch.cuda.is_available():
            raise RuntimeError("dlii_jssp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = config or JsspConfig(max_seeds=1024, max_knots=512, max_obstacle_modes=32, max_total_steps=256)
        self.device = torch.device("cuda", device)
        self.S, self.C = int(max_scenarios), int(max_cycles)
        self._h = C.c_void_p()
        cc = self.cfg.to_c()
        rc = self.lib.jssp_batch_create(C.byref(cc), device, self.S, self.C, C.byref(self._h))
        if rc != 0:
            raise JsspError(rc, "jssp_batch_create")
        self._n = 0
        self._packed: List[PackedScenario] = []
        self.fop = fop
        if fop is not None:
            g = JsspFopIn()
            g.num_width, g.num_speed, g.num_t = fop.num_width, fop.num_speed, fop.num_t
            g.min_t, g.max_t, g.lowest_speed, g.highest_speed = fop.min_t, fop.max_t, fop.lowest_speed, fop.highest_speed
            g.max_accel, g.max_jerk = float(max_accel), float(max_jerk)
            self._check(self.lib.jssp_batch_set_fop(self._h, C.byref(g)), "jssp_batch_set_fop")

    def _check(self, rc: int, where: str):
        if rc != 0:
            detail = self.lib.jssp_batch_last_cuda_error(self._h).decode() if rc == -2 else ""
            raise JsspError(rc, where, detail)

    @property
    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    @property
    def launch_count(self) -> int:
        return int(self.lib.jssp_batch_launch_count(self._h))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.jssp_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load(self, packed: Sequence[PackedScenario]):
        """Upload the scenarios into slots 0..len-1 and reset their replay state."""
        assert 1 <= len(packed) <= self.S, (len(packed), self.S)
        stream = self.stream
        kin = self.ego_update_mode == "kinematics"
        for slot, p in enumerate(packed):
            if abs(p.dt - self.cfg.delta_t) > 1e-12:
                raise ValueError(f"scenario {p.name}: dt = {p.dt} but the batch was created for delta_t = {self.cfg.delta_t}")
            si = p.c_view(kin)
            si.n_cycles = min(p.n_cycles, self.C)
            self._check(self.lib.jssp_batch_set_scenario(self._h, slot, C.byref(si), stream), f"jssp_batch_set_scenario[{slot}]")
        self._n, self._packed = len(packed), list(packed)

    def reset(self):
        """Back to the freshly loaded state without uploading anything (device-to-device)."""
        self._check(self.lib.jssp_batch_reset(self._h, self._n, self.stream), "jssp_batch_reset")

    def run(self, n_cycles: Optional[int] = None, kernel: Optional[str] = None, small_ctas: bool = False):
        """(``small_ctas``: "bestfirst" only - 256-thread CTAs also for few scenarios, for callers running several sub-batches at once.)
        Enqueue up to ``n_cycles`` plan cycles per scenario (default: enough for every loaded scenario to finish).  ``kernel``
        None / "lockstep" = the lock-step pipeline (two launches per cycle for the whole batch) with its default evaluation
        kernel, "warp" / "thread" / "group" / "tiled" = lock-step with that kernel, "persistent" = the persistent replay kernel
        (one CTA per scenario loops over its own cycles in ONE launch), "bestfirst" = the persistent kernel with best-first selection
        (candidates checked in the order of a lower bound of their cost until none left can win), "auto" = the pipeline measured
        fastest for this many scenarios (``auto_pipeline``).  Same results.  Asynchronous."""
        if n_cycles is None:
            n_cycles = max(min(p.n_cycles, self.C) for p in self._packed)
        if kernel == "auto":
            kernel = auto_pipeline(self._n, fop=self.fop is not None)
        flags = {None: 0, "warp": FLAG_FORCE_WARP, "thread": FLAG_FORCE_THREAD, "group": FLAG_FORCE_GROUP, "tiled": FLAG_FORCE_TILED, "lockstep": FLAG_LOCKSTEP, "persistent": FLAG_PERSISTENT, "chunked": FLAG_CHUNKED, "bestfirst": FLAG_BEST_FIRST}[kernel]
        if small_ctas and kernel == "bestfirst":
            flags |= FLAG_SMALL_CTAS
        self._check(self.lib.jssp_batch_run(self._h, self._n, int(n_cycles), flags, self.stream), "jssp_batch_run")
        return n_cycles

    def fetch(self):
        """(results [S] structured array, records [S, C] structured array); synchronises."""
        res = np.zeros(self._n, dtype=RESULT_DTYPE)
        rec = np.zeros((self._n, self.C), dtype=RECORD_DTYPE)
        rc = self.lib.jssp_batch_fetch(self._h, self._n, res.ctypes.data, rec.ctypes.data, self.stream)
        if rc != -5:       # JSSP_ERR_SEED_OVERFLOW: the results ARE filled in; the scenarios concerned carry res["seed_overflow"]
            self._check(rc, "jssp_batch_fetch")
        return res, rec

    def best_traj(self, slot: int) -> np.ndarray:
        """[P, 16] float64 rows (columns = _lib.TRAJ_COLUMNS) of the trajectory selected in the last cycle of ``slot``."""
        out = np.zeros((self.cfg.n_points, _lib.JSSP_TRAJ_COLS), dtype=np.float64)
        self._check(self.lib.jssp_batch_best_traj(self._h, slot, out.ctypes.data, self.stream), "jssp_batch_best_traj")
        return out

    def results(self, lists: bool = True) -> List[ReplayResult]:
        """Fetch and decode: one ReplayResult per loaded scenario.  ``lists=False`` leaves ``best_idx`` / ``n_candidates`` /
        ``ego_rows`` as NumPy views of the fetched records ([n], [n], [n, 6]) instead of Python lists of ints / tuples."""
        res, rec = self.fetch()
        rows_all = np.stack([rec[f] for f in ("t", "x", "y", "v", "a", "yaw")], -1)     # [S, C, 6]
        ok_all = ~np.isnan(rec["t"])
        bi_all, nc_all, cyc, ends = rec["best_idx"], rec["n_candidates"], res["cycles"].tolist(), res["end"].tolist()
        out = []
        for s, p in enumerate(self._packed):
            n = cyc[s]
            rows = rows_all[s, :n]
            if not ok_all[s, :n].all():
                rows = rows[ok_all[s, :n]]
            rr = ReplayResult(p.name, ends[s], n, seed_overflow=bool(res["seed_overflow"][s]))
            if lists:
                rr.best_idx, rr.n_candidates, rr.ego_rows = bi_all[s, :n].tolist(), nc_all[s, :n].tolist(), list(map(tuple, rows.tolist()))
            else:
                rr.best_idx, rr.n_candidates, rr.ego_rows = bi_all[s, :n], nc_all[s, :n], rows
            out.append(rr)
        return out


def auto_pipeline(n_scenarios: int, fop: bool = False) -> Optional[str]:
    """The replay pipeline measured fastest on a B200 for ``n_scenarios`` on one GPU: "bestfirst" (persistent kernel with best-first
    selection) from 150 scenarios on - 800: 28.9 vs 37.4 ms, 400: 19.1 vs 23.8 ms, 200: 14.1 vs 14.7 ms - and None (the lock-step
    pipeline) below, where a sweep ends with its heaviest scenario and lock-step spreads that scenario's candidates over many SMs
    (100: 10.6 vs 12.6 ms).  The FOP baseline always replays in lock-step."""
    return "bestfirst" if (not fop and n_scenarios >= 150) else None


class MultiStreamReplay:
    """``k`` BatchReplay handles on ``k`` CUDA streams, the scenarios dealt round-robin: the kernels of the sub-batches interleave on
    the device, so the seed enumeration of one sub-batch overlaps the evaluation of another (measured with k = 2: 800-scenario
    sweep 42.1 -> 38.8 ms, 100 scenarios 130.6 -> 125.6 us per cycle; no further gain for k = 3, 4).  Same interface as
    BatchReplay; work is enqueued relative to the caller's current stream (it is waited for at the start of ``run`` and waits
    for the sub-batches at its end)."""

    def __init__(self, config: Optional[JsspConfig] = None, device: int = 0, max_scenarios: int = 64, max_cycles: int = 128, k: int = 2, **kw):
        self.k = max(1, int(k))
        self.device = torch.device("cuda", device)
        per = (int(max_scenarios) + self.k - 1) // self.k
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.k)]
        self.parts = [BatchReplay(config, device=device, max_scenarios=max(per, 1), max_cycles=max_cycles, **kw) for _ in range(self.k)]
        self.cfg, self.C, self.lib = self.parts[0].cfg, self.parts[0].C, self.parts[0].lib
        self._n = 0
        self._sms = torch.cuda.get_device_properties(self.device).multi_processor_count

    def _small(self) -> bool:
        """Best-first sub-batches that together put more than three scenarios on an SM run 256-thread CTAs (see JSSP_FLAG_SMALL_CTAS)."""
        return self.k > 1 and self._n > 3 * self._sms

    def _each(self, fn):
        cur = torch.cuda.current_stream(self.device)
        out = []
        for bt, st in zip(self.parts[:max(1, min(self.k, self._n))], self.streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                out.append(fn(bt))
        for st in self.streams:
            cur.wait_stream(st)
        return out

    def load(self, packed: Sequence[PackedScenario]):
        self._n = len(packed)
        groups = [list(packed[i::self.k]) for i in range(self.k)]
        cur = torch.cuda.current_stream(self.device)
        for bt, st, g in zip(self.parts, self.streams, groups):
            if g:
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    bt.load(g)
                cur.wait_stream(st)

    def reset(self):
        self._each(lambda bt: bt.reset())

    def run(self, n_cycles: Optional[int] = None, kernel: Optional[str] = None):
        return max(self._each(lambda bt: bt.run(n_cycles, kernel=kernel, small_ctas=self._small())))

    def load_and_run(self, packed: Sequence[PackedScenario], n_cycles: Optional[int] = None, kernel: Optional[str] = None):
        """``load`` + ``run`` sub-batch by sub-batch: the replay of one sub-batch (device) overlaps the upload of the next (host staging +
        H2D copies), so an end-to-end sweep costs about max(upload, replay) instead of their sum."""
        self._n = len(packed)
        groups = [list(packed[i::self.k]) for i in range(self.k)]
        cur = torch.cuda.current_stream(self.device)
        out = [0]
        for bt, st, g in zip(self.parts, self.streams, groups):
            if g:
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    bt.load(g)
                    out.append(bt.run(n_cycles, kernel=kernel, small_ctas=self._small()))
        for st in self.streams:
            cur.wait_stream(st)
        return max(out)

    def results(self, lists: bool = True) -> List[ReplayResult]:
        parts = self._each(lambda bt: bt.results(lists=lists))
        out: List[Optional[ReplayResult]] = [None] * self._n
        for i, res in enumerate(parts):
            out[i::self.k] = res
        return out

    @property
    def launch_count(self) -> int:
        return sum(bt.launch_count for bt in self.parts)

    def close(self):
        for bt in self.parts:
            bt.close()


def fop_batch_config(fop, ego_length: float, ego_width: float, **caps) -> JsspConfig:
    """The jssp_config of a batch that replays with the FOP baseline (same fields FrenetOptimalPlanner sets for its engine)."""
    from .fop import PLANNER_FRENET, PARA_THRESHOLD
    return JsspConfig(delta_t=fop.delta_t, plan_horizon=fop.max_t, max_road_width=fop.max_road_width, num_width=fop.num_width,
                      speed_max=fop.highest_speed, ego_length=ego_length, ego_width=ego_width, obstacle_inflation=0.0,
                      cost_weights=tuple(PLANNER_FRENET["cost_wights"]), max_distance=PLANNER_FRENET["max_distance"],
                      thr_lon_accel=PARA_THRESHOLD["lon_accel_max"], thr_lon_jerk=PARA_THRESHOLD["lon_jerk_max"],
                      thr_lat_accel=PARA_THRESHOLD["lateral_accel_max"], thr_lat_jerk=PARA_THRESHOLD["lateral_jerk_max"], **caps)


def run_batch(scenarios: Sequence, device: int = 0, max_cycles: Optional[int] = None, batch: Optional[BatchReplay] = None,
              kernel: Optional[str] = None, fop=None, max_accel: float = 8.0, max_jerk: float = 10.0,
              ego_update_mode: str = "planner_expected_value", max_seeds: int = 1024) -> List[ReplayResult]:
    """Replay independent scenarios (ReplayScenario or PackedScenario) in lock-step on one GPU; same results as
    ``replay.run_sweep``.  Scenarios that are over before the first cycle are answered on the host.  ``fop`` (a
    FrenetOptimalPlannerSettings) replays with the FOP baseline planner instead of JSSP.  ``max_seeds`` sizes the per-scenario
    seed buffer of an auto-created batch; a scenario whose cycle enumerates more is flagged (``ReplayResult.seed_overflow``),
    the others are unaffected."""
    plan_steps = 50 if fop is None else int(np.ceil(fop.max_t / fop.delta_t))
    packed = [s if isinstance(s, PackedScenario) else pack_scenario(s, plan_steps=plan_steps, max_cycles=max_cycles,
                                                                    kinematics=ego_update_mode == "kinematics") for s in scenarios]
    live = [i for i, p in enumerate(packed) if p.initial_end == END_RUNNING]
    out: List[Optional[ReplayResult]] = [None if i in set(live) else ReplayResult(p.name, p.initial_end, 0) for i, p in enumerate(packed)]
    if live:
        todo = [packed[i] for i in live]
        e = todo[0].ego
        need_c = max(min(p.n_cycles, max_cycles or p.n_cycles) for p in todo)
        if batch is None:
            caps = dict(max_seeds=int(max_seeds), max_knots=max(64, max(len(p.knot_s) for p in todo)),
                        max_obstacle_modes=max(1, max(p.state.shape[0] * p.state.shape[1] for p in todo)),
                        max_total_steps=max(p.state.shape[2] for p in todo))
            cfg = JsspConfig(ego_length=float(e[5]), ego_width=float(e[6]), **caps) if fop is None else fop_batch_config(fop, float(e[5]), float(e[6]), **caps)
            batch = BatchReplay(cfg, device=device, max_scenarios=len(todo), max_cycles=need_c, fop=fop, max_accel=max_accel, max_jerk=max_jerk,
                                ego_update_mode=ego_update_mode)
        t0 = time.perf_counter()
        for lo in range(0, len(todo), batch.S):
            chunk = todo[lo:lo + batch.S]
            batch.load(chunk)
            batch.run(min(need_c, batch.C), kernel=kernel)
            for i, rr in zip(live[lo:lo + batch.S], batch.results()):
                out[i] = rr
        dt_s = time.perf_counter() - t0
        cyc = sum(r.cycles for r in out if r is not None)
        for r in out:
            r.plan_seconds = [dt_s / max(cyc, 1)] * r.cycles        # amortised: cycles of a batch are not timed individually
    return out
