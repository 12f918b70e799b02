"""Thin Python owner of a ``jssp_handle``: device buffers are torch tensors, every compute call goes through the C ABI."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field, asdict
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import JsspError, JsspConfigC, JsspCycleIn, JsspCycleOut, JsspFopIn, JsspImmParams, FLAG_NO_SYNC, FLAG_NO_EXPORT, FLAG_NO_CULL, FLAG_ENUMERATE, FLAG_FORCE_WARP, FLAG_FORCE_THREAD, FLAG_FORCE_GROUP, FLAG_FORCE_TILED, TRAJ_COLUMNS


@dataclass
class JsspConfig:
    """Mirror of ``jssp_config`` (include/jssp_b200.h).  Defaults are the reference's values: planner_JSSP block of
    configuration_parameters.py:52-61, vehicle_para / para_threshold :7-22, max_road_width intersection_planner.py:60;
    ``highest_jerk`` = lon_jerk_max and ``max_curvature`` = 0.2 1/m are the decisions recorded in SURVEY.md (0.4, 8c)."""
    delta_t: float = 0.1
    plan_horizon: float = 5.0
    max_road_width: float = 3.5
    num_width: int = 3
    num_jerk: int = 25
    highest_jerk: float = 10.0
    speed_max: float = 18.0
    lon_accel_max: float = 8.0
    ego_length: float = 4.924
    ego_width: float = 1.872
    max_curvature: float = 0.2
    obstacle_inflation: float = 0.1
    check_res: int = 2
    cost_weights: Sequence[float] = (0.0, 3.0, 5.4, 0.2, 0.1, 0.8, 0.4, 4.0)
    max_distance: float = 140.0
    thr_lon_accel: float = 3.0
    thr_lon_jerk: float = 6.0
    thr_lat_accel: float = 0.5
    thr_lat_jerk: float = 1.0
    p_min: float = 0.0
    w_risk: float = 0.0
    max_seeds: int = 4096
    max_knots: int = 1024
    max_obstacle_modes: int = 256
    max_total_steps: int = 512

    @property
    def n_points(self) -> int:
        """Samples per trajectory.  The reference uses int(T / dt) + 1 for the time grid (polyvt_sampling.py:44) and
        round(T / dt) + 1 for the lateral grid (jerk_space_sampling_planner.py:78-79); pairs where the two disagree (0.3 / 0.1 ->
        2 vs 3) are rejected here and by jssp_create (JSSP_ERR_INVALID_ARG) instead of silently losing a sample."""
        n = int(self.plan_horizon / self.delta_t)
        if n != int(round(self.plan_horizon / self.delta_t)):
            raise ValueError(f"plan_horizon / delta_t = {self.plan_horizon} / {self.delta_t} truncates and rounds to different step counts")
        return n + 1

    def to_c(self) -> JsspConfigC:
        c = JsspConfigC()
        for k, v in asdict(self).items():
            if k == "cost_weights":
                c.cost_weights = (C.c_double * 8)(*[float(x) for x in v])
            else:
                setattr(c, k, v)
        return c


@dataclass
class CycleResult:
    n_candidates: int
    n_seeds: int
    best_idx: int
    best_cost: float
    best_n_pts: int
    best_traj: Optional[np.ndarray]      # [P, 16] float64 (columns = _lib.TRAJ_COLUMNS) or None
    feasible: torch.Tensor               # uint8 [N]  (device)
    collision: torch.Tensor              # uint8 [N]  (device)
    cost: torch.Tensor                   # float32 [N] (device)
    states: Optional[torch.Tensor] = None  # float32 [N, P, 8] (device)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class JsspEngine:
    """One handle per (GPU, stream).  Not thread-safe (like the reference planner object)."""

    def __init__(self, config: Optional[JsspConfig] = None, device: int = 0, max_candidates: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("dlii_jssp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = config or JsspConfig()
        self.device = torch.device("cuda", device)
        self._h = C.c_void_p()
        cc = self.cfg.to_c()
        rc = self.lib.jssp_create(C.byref(cc), device, C.byref(self._h))
        if rc != 0:
            raise JsspError(rc, "jssp_create")
        self.P = self.cfg.n_points
        cap = max_candidates or self.cfg.max_seeds * max(self.cfg.num_width, 1)
        self._cap = 0
        self._views = {}
        self._alloc(cap)
        self._best_host = np.zeros((self.P, _lib.JSSP_TRAJ_COLS), dtype=np.float64)
        self._states: Optional[torch.Tensor] = None
        self._fixed_stream: Optional[int] = None
        # persistent structs of the one-call plan cycle (plan_cycle): only the start state changes between cycles
        self._pc_in, self._pc_out = JsspCycleIn(), JsspCycleOut()
        self._pc_in.flags = FLAG_ENUMERATE

    # -- buffers ------------------------------------------------------------------------------------------------
    def _alloc(self, cap: int):
        if cap <= self._cap:
            return
        self._cap = int(cap)
        self.feasible = torch.zeros(self._cap, dtype=torch.uint8, device=self.device)
        self.collision = torch.zeros(self._cap, dtype=torch.uint8, device=self.device)
        self.cost = torch.zeros(self._cap, dtype=torch.float32, device=self.device)
        if getattr(self, "_pc_out", None) is not None:      # plan_cycle's persistent struct points into the old tensors
            self._pc_out.best_traj_host = None

    def _check(self, rc: int, where: str):
        if rc != 0:
            detail = self.lib.jssp_last_cuda_error(self._h).decode() if rc == -2 else ""
            raise JsspError(rc, where, detail)

    @property
    def stream(self) -> int:
        """The CUDA stream handed to the library: torch's current stream, unless pinned with use_stream()."""
        if self._fixed_stream is not None:
            return self._fixed_stream
        return torch.cuda.current_stream(self.device).cuda_stream

    def use_stream(self, stream: Optional[torch.cuda.Stream]):
        """Pin every following call to ``stream`` (None = follow torch's current stream again); saves the per-call lookup."""
        self._fixed_stream = None if stream is None else stream.cuda_stream

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.jssp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- per-scenario ---------------------------------------------------------------------------------------------
    def set_refline(self, knot_s: np.ndarray, coef: np.ndarray):
        """knot_s [K] float64 (cumulative chord length), coef [K-1, 8] = ax,bx,cx,dx,ay,by,cy,dy per segment."""
        knot_s = np.ascontiguousarray(knot_s, dtype=np.float64)
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        assert coef.shape == (len(knot_s) - 1, 8), coef.shape
        self._check(self.lib.jssp_set_refline(self._h, knot_s.ctypes.data, coef.ctypes.data, len(knot_s), self.stream), "jssp_set_refline")

    def set_obstacles(self, state: np.ndarray, valid: np.ndarray, size: np.ndarray, prob: np.ndarray, n_dict: Optional[int] = None):
        """Host arrays: state [O,M,Tt,3] (x,y,yaw), valid [O,M,Tt], size [O,2] (length,width), prob [O,M]."""
        state = np.ascontiguousarray(state, dtype=np.float64)
        O, M, Tt, _ = state.shape
        valid = np.ascontiguousarray(valid, dtype=np.uint8)
        size = np.ascontiguousarray(size, dtype=np.float64)
        prob = np.ascontiguousarray(prob, dtype=np.float64)
        assert valid.shape == (O, M, Tt) and size.shape == (O, 2) and prob.shape == (O, M)
        self._check(self.lib.jssp_set_obstacles(self._h, state.ctypes.data, valid.ctypes.data, size.ctypes.data, prob.ctypes.data,
                                                O, M, Tt, O if n_dict is None else n_dict, self.stream), "jssp_set_obstacles")

    def set_obstacles_dev(self, state: torch.Tensor, valid: torch.Tensor, size: torch.Tensor, prob: torch.Tensor, n_dict: Optional[int] = None):
        """Same from caller-owned device tensors (float64 / uint8); the tensors must stay alive until the next call."""
        O, M, Tt, _ = state.shape
        assert state.dtype == torch.float64 and valid.dtype == torch.uint8 and size.dtype == torch.float64 and prob.dtype == torch.float64
        assert state.is_contiguous() and valid.is_contiguous() and size.is_contiguous() and prob.is_contiguous()
        self._obs_keep = (state, valid, size, prob)
        self._check(self.lib.jssp_set_obstacles_dev(self._h, state.data_ptr(), valid.data_ptr(), size.data_ptr(), prob.data_ptr(),
                                                    O, M, Tt, O if n_dict is None else n_dict, self.stream), "jssp_set_obstacles_dev")

    # -- per-cycle ------------------------------------------------------------------------------------------------
    def enumerate_seeds(self, s0: float, v0: float, a0: float, sync: bool = True):
        """Device-side PolyVTSampling.sampling + sampling_stop_seeds_supplement.  Returns (n_main, n_stop) when sync."""
        if sync:
            a, b = C.c_int32(0), C.c_int32(0)
            self._check(self.lib.jssp_enumerate_seeds(self._h, s0, v0, a0, self.stream, C.byref(a), C.byref(b)), "jssp_enumerate_seeds")
            return a.value, b.value
        self._check(self.lib.jssp_enumerate_seeds(self._h, s0, v0, a0, self.stream, None, None), "jssp_enumerate_seeds")
        return None

    def get_seeds(self) -> np.ndarray:
        n = C.c_int32(0)
        buf = np.zeros((self.cfg.max_seeds, 10), dtype=np.float64)
        self._check(self.lib.jssp_get_seeds(self._h, buf.ctypes.data, self.cfg.max_seeds, C.byref(n), self.stream), "jssp_get_seeds")
        return buf[: n.value].copy()

    def _cycle_structs(self, frenet0, time_step_now, final_time_step, seeds, targets, flags, want_states, n_cand_cap):
        cin = JsspCycleIn()
        cin.s0, cin.v0, cin.a0, cin.d0, cin.d_d0, cin.d_dd0 = [float(v) for v in frenet0]
        cin.time_step_now, cin.final_time_step = int(time_step_now), int(final_time_step)
        cin.flags = flags
        if targets is not None:
            tg = [float(t) for t in targets]
            assert 1 <= len(tg) <= _lib.JSSP_MAX_WIDTHS
            cin.n_widths = len(tg)
            for i, t in enumerate(tg):
                cin.targets[i] = t
        W = cin.n_widths or self.cfg.num_width
        if seeds is not None:
            assert seeds.is_cuda and seeds.dtype == torch.float64 and seeds.is_contiguous() and seeds.shape[-1] == 10
            cin.seeds_dev = seeds.data_ptr()
            cin.n_seeds = seeds.shape[0]
            n_cap = W * seeds.shape[0]
        else:
            n_cap = W * self.cfg.max_seeds
        self._alloc(n_cap)
        cout = JsspCycleOut()
        cout.feasible_dev, cout.collision_dev, cout.cost_dev = self.feasible.data_ptr(), self.collision.data_ptr(), self.cost.data_ptr()
        if want_states:
            if self._states is None or self._states.shape[0] < n_cap:
                self._states = torch.empty((n_cap, self.P, _lib.JSSP_STATE_COLS), dtype=torch.float32, device=self.device)
            cout.states_dev = self._states.data_ptr()
        cout.best_traj_host = self._best_host.ctypes.data
        return cin, cout

    def evaluate(self, frenet0, time_step_now: int, final_time_step: int, seeds: Optional[torch.Tensor] = None,
                 targets: Optional[Sequence[float]] = None, want_states: bool = False, sync: bool = True, export: bool = True,
                 cull: bool = True, kernel: Optional[str] = None, select_only: bool = False) -> Optional[CycleResult]:
        """One plan cycle.  frenet0 = (s, s_d, s_dd, d, d_d, d_dd); seeds = device float64 [Ns, 10] or None (use the
        enumerated seeds).  With sync=False the kernels are only enqueued; call fetch() later.  ``select_only``: no
        per-candidate masks / costs are requested (NULL pointers), only the selection and the exported trajectory - a
        candidate's walk then stops at its first veto, as the reference's filter chain does; the masks of the returned result
        are stale."""
        flags = (0 if sync else FLAG_NO_SYNC) | (0 if export else FLAG_NO_EXPORT) | (0 if cull else FLAG_NO_CULL)
        flags |= {None: 0, "warp": FLAG_FORCE_WARP, "thread": FLAG_FORCE_THREAD, "group": FLAG_FORCE_GROUP, "tiled": FLAG_FORCE_TILED}[kernel]
        cin, cout = self._cycle_structs(frenet0, time_step_now, final_time_step, seeds, targets, flags, want_states, None)
        if select_only:
            cout.feasible_dev = cout.collision_dev = cout.cost_dev = None
        self._last = (cin, cout, want_states)
        self._check(self.lib.jssp_evaluate(self._h, C.byref(cin), self.stream, C.byref(cout)), "jssp_evaluate")
        return self._result(cout, want_states, export) if sync else None

    def plan_cycle(self, frenet0, time_step_now: int, final_time_step: int) -> JsspCycleOut:
        """One whole plan cycle in ONE library call: device-side seed enumeration from (s, s_d, s_dd), fused evaluation,
        selection and export; blocks until the selected trajectory is in ``best_traj_host``.  Returns the (reused)
        jssp_cycle_out struct: n_candidates, n_seeds, best_idx, best_cost, best_n_pts; masks and costs stay in
        ``self.feasible / collision / cost`` on the device."""
        cin, cout = self._pc_in, self._pc_out
        cin.s0, cin.v0, cin.a0, cin.d0, cin.d_d0, cin.d_dd0 = frenet0
        cin.time_step_now, cin.final_time_step = time_step_now, final_time_step
        if not cout.best_traj_host:
            self._alloc(self.cfg.num_width * self.cfg.max_seeds)
            cout.feasible_dev, cout.collision_dev, cout.cost_dev = self.feasible.data_ptr(), self.collision.data_ptr(), self.cost.data_ptr()
            cout.best_traj_host = self._best_host.ctypes.data
        rc = self.lib.jssp_evaluate(self._h, C.byref(cin), self.stream, C.byref(cout))
        if rc != 0:
            self._check(rc, "jssp_evaluate")
        return cout

    def result_of(self, cout: JsspCycleOut, want_states: bool = False) -> CycleResult:
        """Full CycleResult view (device mask / cost slices) for a struct returned by plan_cycle."""
        return self._result(cout, want_states, True)

    def fetch(self) -> CycleResult:
        cin, cout, want_states = self._last
        self._check(self.lib.jssp_fetch_best(self._h, self.stream, C.byref(cout)), "jssp_fetch_best")
        return self._result(cout, want_states, not (cin.flags & FLAG_NO_EXPORT))

    def _result(self, cout, want_states, export) -> CycleResult:
        n = cout.n_candidates
        traj = self._best_host.copy() if (export and cout.best_idx >= 0) else None
        views = self._views.get(n)                     # tensor slicing costs microseconds per call: keep the views per size
        if views is None or views[3] != self._cap:
            if len(self._views) > 64:
                self._views.clear()
            views = self._views[n] = (self.feasible[:n], self.collision[:n], self.cost[:n], self._cap)
        return CycleResult(n, cout.n_seeds, cout.best_idx, cout.best_cost, cout.best_n_pts, traj, views[0], views[1], views[2],
                           self._states[:n] if want_states else None)

    def best_record(self, index_offset: int = 0) -> torch.Tensor:
        """Device int64[2] view of (orderable cost bits, global candidate index) of the last evaluate: the payload of the
        candidate-split reduction (unsigned order of word 0 == numeric order of the float64 cost)."""
        rec = torch.empty(2, dtype=torch.int64, device=self.device)
        self._check(self.lib.jssp_best_record_dev(self._h, int(index_offset), self.stream, rec.data_ptr()), "jssp_best_record_dev")
        return rec

    # -- candidate split over ranks (jssp_set_split / jssp_allreduce_argmin / jssp_peer_*) ------------------------------------
    def set_split(self, seed_lo: int, n_seeds_global: int):
        """The external seeds of the following evaluate() calls are rows [seed_lo, seed_lo + n) of a cycle with
        ``n_seeds_global`` seeds: the selection record then carries the GLOBAL candidate index.  (0, 0) switches it off."""
        self._check(self.lib.jssp_set_split(self._h, int(seed_lo), int(n_seeds_global)), "jssp_set_split")

    def allreduce_argmin(self, nccl_comm: Optional[int], world: int):
        """NCCL transport: enqueue the all-gather of the 16-byte records + the one-warp lexicographic minimum (asynchronous)."""
        self._check(self.lib.jssp_allreduce_argmin(self._h, nccl_comm, int(world), self.stream), "jssp_allreduce_argmin")

    def peer_export(self, world: int) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._check(self.lib.jssp_peer_export(self._h, int(world), buf), "jssp_peer_export")
        return bytes(buf)

    def peer_connect(self, handles: bytes, world: int, rank: int):
        assert len(handles) == 64 * world
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        self._check(self.lib.jssp_peer_connect(self._h, buf, int(world), int(rank)), "jssp_peer_connect")

    def peer_disconnect(self):
        self.lib.jssp_peer_disconnect(self._h)

    def fetch_global_best(self):
        """(global best index or -1, its float64 cost, rank holding it); synchronises the stream."""
        i, c, r = C.c_int64(-1), C.c_double(0.0), C.c_int32(0)
        self._check(self.lib.jssp_fetch_global_best(self._h, self.stream, C.byref(i), C.byref(c), C.byref(r)), "jssp_fetch_global_best")
        return i.value, c.value, r.value

    def evaluate_fop(self, frenet0, time_step_now: int, final_time_step: int, num_width: int, num_speed: int, num_t: int, min_t: float,
                     max_t: float, lowest_speed: float, highest_speed: float, max_accel: float, max_jerk: float,
                     export: bool = True) -> CycleResult:
        """One plan cycle of the FOP baseline (jssp_fop_evaluate): candidate n = (w * num_t + it) * num_speed + iv."""
        fin = JsspFopIn()
        fin.s0, fin.v0, fin.a0, fin.d0, fin.d_d0, fin.d_dd0 = [float(v) for v in frenet0]
        fin.time_step_now, fin.final_time_step = int(time_step_now), int(final_time_step)
        fin.num_width, fin.num_speed, fin.num_t = int(num_width), int(num_speed), int(num_t)
        fin.flags = 0 if export else FLAG_NO_EXPORT
        fin.min_t, fin.max_t, fin.lowest_speed, fin.highest_speed = float(min_t), float(max_t), float(lowest_speed), float(highest_speed)
        fin.max_accel, fin.max_jerk = float(max_accel), float(max_jerk)
        self._alloc(num_width * num_speed * num_t)
        cout = JsspCycleOut()
        cout.feasible_dev, cout.collision_dev, cout.cost_dev = self.feasible.data_ptr(), self.collision.data_ptr(), self.cost.data_ptr()
        cout.best_traj_host = self._best_host.ctypes.data
        self._check(self.lib.jssp_fop_evaluate(self._h, C.byref(fin), self.stream, C.byref(cout)), "jssp_fop_evaluate")
        return self._result(cout, False, export)

    def project_states(self, states: torch.Tensor, step: float = 0.1) -> torch.Tensor:
        """Batched FrenetState.from_state on the uploaded reference line: states [n, 5] = x, y, yaw, v, a (device, float64)
        -> [n, 6] = s, s_d, s_dd, d, d_d, d_dd (device)."""
        assert states.is_cuda and states.dtype == torch.float64 and states.is_contiguous() and states.shape[-1] == 5
        out = torch.empty((states.shape[0], 6), dtype=torch.float64, device=self.device)
        self._check(self.lib.jssp_project_states(self._h, states.data_ptr(), states.shape[0], float(step), out.data_ptr(), self.stream),
                    "jssp_project_states")
        return out

    @property
    def launch_count(self) -> int:
        return int(self.lib.jssp_launch_count(self._h))

    # -- IMM (secondary kernel) -----------------------------------------------------------------------------------
    def imm_update(self, mean: torch.Tensor, cov: torch.Tensor, mu: torch.Tensor, z: torch.Tensor, dt=0.1, q_pos=0.01, q_vel=0.05,
                   r_pos=0.04, k_lat=0.5, p_stay=0.95, mu_floor=1e-6):
        O, M = mu.shape
        for t in (mean, cov, mu, z):
            assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        prm = JsspImmParams(dt, q_pos, q_vel, r_pos, k_lat, p_stay, mu_floor)
        self._check(self.lib.jssp_imm_update(self._h, C.byref(prm), mean.data_ptr(), cov.data_ptr(), mu.data_ptr(), z.data_ptr(), O, M, self.stream),
                    "jssp_imm_update")

    def predict_modes(self, lanes: torch.Tensor, lane_of: torch.Tensor, s_start: torch.Tensor, speed: torch.Tensor, dt: float, Tt: int):
        L, Pn, _ = lanes.shape
        O, M = lane_of.shape
        assert lanes.dtype == torch.float64 and lane_of.dtype == torch.int32 and s_start.dtype == torch.float64 and speed.dtype == torch.float64
        state = torch.empty((O, M, Tt, 3), dtype=torch.float64, device=self.device)
        valid = torch.empty((O, M, Tt), dtype=torch.uint8, device=self.device)
        self._check(self.lib.jssp_predict_modes(self._h, lanes.data_ptr(), L, Pn, lane_of.data_ptr(), s_start.data_ptr(), speed.data_ptr(),
                                                float(dt), O, M, Tt, state.data_ptr(), valid.data_ptr(), self.stream), "jssp_predict_modes")
        return state, valid
