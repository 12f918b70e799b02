"""ctypes binding of include/jssp_b200.h.  Fails loudly when the library is missing: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

JSSP_MAX_WIDTHS = 16
JSSP_TRAJ_COLS = 16
JSSP_STATE_COLS = 8
FLAG_NO_SYNC = 1
FLAG_NO_EXPORT = 2
FLAG_NO_CULL = 4
FLAG_ENUMERATE = 8
FLAG_FORCE_WARP = 16
FLAG_FORCE_THREAD = 32
FLAG_FORCE_GROUP = 64
FLAG_FORCE_TILED = 128
FLAG_LOCKSTEP = 256
FLAG_PERSISTENT = 512
FLAG_BEST_FIRST = 2048
FLAG_SMALL_CTAS = 4096
FLAG_CHUNKED = 1024

TRAJ_COLUMNS = ("t", "s", "s_d", "s_dd", "s_ddd", "d", "d_d", "d_dd", "d_ddd", "x", "y", "yaw", "ds", "c", "c_d", "c_dd")
STATE_COLUMNS = ("s", "s_d", "s_dd", "d", "x", "y", "yaw", "c")


class JsspConfigC(C.Structure):
    _fields_ = [
        ("delta_t", C.c_double), ("plan_horizon", C.c_double), ("max_road_width", C.c_double),
        ("num_width", C.c_int32), ("num_jerk", C.c_int32), ("highest_jerk", C.c_double),
        ("speed_max", C.c_double), ("lon_accel_max", C.c_double),
        ("ego_length", C.c_double), ("ego_width", C.c_double),
        ("max_curvature", C.c_double), ("obstacle_inflation", C.c_double),
        ("check_res", C.c_int32), ("reserved0", C.c_int32),
        ("cost_weights", C.c_double * 8), ("max_distance", C.c_double),
        ("thr_lon_accel", C.c_double), ("thr_lon_jerk", C.c_double), ("thr_lat_accel", C.c_double), ("thr_lat_jerk", C.c_double),
        ("p_min", C.c_double), ("w_risk", C.c_double),
        ("max_seeds", C.c_int32), ("max_knots", C.c_int32), ("max_obstacle_modes", C.c_int32), ("max_total_steps", C.c_int32),
    ]


class JsspCycleIn(C.Structure):
    _fields_ = [
        ("s0", C.c_double), ("v0", C.c_double), ("a0", C.c_double),
        ("d0", C.c_double), ("d_d0", C.c_double), ("d_dd0", C.c_double),
        ("time_step_now", C.c_int32), ("final_time_step", C.c_int32), ("n_widths", C.c_int32), ("flags", C.c_int32),
        ("targets", C.c_double * JSSP_MAX_WIDTHS),
        ("seeds_dev", C.c_void_p), ("n_seeds", C.c_int32), ("reserved0", C.c_int32),
    ]


class JsspCycleOut(C.Structure):
    _fields_ = [
        ("feasible_dev", C.c_void_p), ("collision_dev", C.c_void_p), ("cost_dev", C.c_void_p), ("states_dev", C.c_void_p),
        ("best_traj_host", C.c_void_p),
        ("n_candidates", C.c_int32), ("best_idx", C.c_int32), ("best_n_pts", C.c_int32), ("n_seeds", C.c_int32),
        ("best_cost", C.c_double),
    ]


class JsspFopIn(C.Structure):
    _fields_ = [("s0", C.c_double), ("v0", C.c_double), ("a0", C.c_double), ("d0", C.c_double), ("d_d0", C.c_double), ("d_dd0", C.c_double),
                ("time_step_now", C.c_int32), ("final_time_step", C.c_int32), ("num_width", C.c_int32), ("num_speed", C.c_int32),
                ("num_t", C.c_int32), ("flags", C.c_int32), ("min_t", C.c_double), ("max_t", C.c_double),
                ("lowest_speed", C.c_double), ("highest_speed", C.c_double), ("max_accel", C.c_double), ("max_jerk", C.c_double)]


class JsspImmParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("q_pos", C.c_double), ("q_vel", C.c_double), ("r_pos", C.c_double),
                ("k_lat", C.c_double), ("p_stay", C.c_double), ("mu_floor", C.c_double)]


class JsspCycleRecord(C.Structure):
    _fields_ = [("best_idx", C.c_int32), ("n_candidates", C.c_int32), ("best_cost", C.c_double),
                ("t", C.c_double), ("x", C.c_double), ("y", C.c_double), ("v", C.c_double), ("a", C.c_double), ("yaw", C.c_double)]


class JsspScenarioResult(C.Structure):
    _fields_ = [("end", C.c_double), ("cycles", C.c_int32), ("seed_overflow", C.c_int32)]


class JsspScenarioIn(C.Structure):
    _fields_ = [
        ("knot_s", C.c_void_p), ("coef", C.c_void_p), ("K", C.c_int32), ("reserved0", C.c_int32),
        ("state", C.c_void_p), ("valid", C.c_void_p), ("size", C.c_void_p), ("prob", C.c_void_p),
        ("O", C.c_int32), ("M", C.c_int32), ("Tt", C.c_int32), ("n_dict", C.c_int32),
        ("gt_state", C.c_void_p), ("gt_valid", C.c_void_p), ("gt_size", C.c_void_p), ("gt_O", C.c_int32), ("reserved1", C.c_int32),
        ("frenet0", C.c_double * 6), ("max_t", C.c_double), ("final_time_step", C.c_int32), ("n_cycles", C.c_int32),
        ("cycle_time", C.c_void_p), ("cycle_now", C.c_void_p), ("cycle_end_step", C.c_void_p),
        ("goal_x", C.c_double), ("goal_y", C.c_double), ("goal_yaw", C.c_double),
        ("ego_length", C.c_double), ("ego_width", C.c_double),
        ("ego0", C.c_double * 4), ("bounds", C.c_double * 4), ("has_bounds", C.c_int32), ("ego_update_mode", C.c_int32),
        ("polyline", C.c_void_p), ("n_polyline", C.c_int32), ("reserved2", C.c_int32),
        ("pp_kv", C.c_double), ("pp_ld0", C.c_double), ("pp_wheel_base", C.c_double),
    ]


# jssp_struct_size's numbering
STRUCTS = (JsspConfigC, JsspCycleIn, JsspCycleOut, JsspFopIn, JsspImmParams, JsspCycleRecord, JsspScenarioResult, JsspScenarioIn)

# every symbol include/jssp_b200.h declares (tests check that the library exports exactly these)
SYMBOLS = {
    "jssp_abi_version": (C.c_int32, []),
    "jssp_build_id": (C.c_char_p, []),
    "jssp_struct_size": (C.c_int32, [C.c_int32]),
    "jssp_strerror": (C.c_char_p, [C.c_int32]),
    "jssp_last_cuda_error": (C.c_char_p, [C.c_void_p]),
    "jssp_default_config": (None, [C.POINTER(JsspConfigC)]),
    "jssp_create": (C.c_int32, [C.POINTER(JsspConfigC), C.c_int32, C.POINTER(C.c_void_p)]),
    "jssp_destroy": (C.c_int32, [C.c_void_p]),
    "jssp_make_grids": (C.c_int32, [C.c_double, C.c_double, C.c_int32, C.c_int32] + [C.c_void_p] * 9 + [C.POINTER(C.c_int32 * 9)]),
    "jssp_quintic_coeffs": (None, [C.c_double] * 7 + [C.POINTER(C.c_double * 6)]),
    "jssp_set_refline": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "jssp_set_obstacles": (C.c_int32, [C.c_void_p] * 5 + [C.c_int32] * 4 + [C.c_void_p]),
    "jssp_set_obstacles_dev": (C.c_int32, [C.c_void_p] * 5 + [C.c_int32] * 4 + [C.c_void_p]),
    "jssp_enumerate_seeds": (C.c_int32, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "jssp_get_seeds": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.c_void_p]),
    "jssp_evaluate": (C.c_int32, [C.c_void_p, C.POINTER(JsspCycleIn), C.c_void_p, C.POINTER(JsspCycleOut)]),
    "jssp_fetch_best": (C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(JsspCycleOut)]),
    "jssp_best_record_dev": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "jssp_set_split": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64]),
    "jssp_nccl_unique_id": (C.c_int32, [C.c_void_p]),
    "jssp_nccl_comm_create": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "jssp_nccl_comm_destroy": (C.c_int32, [C.c_void_p]),
    "jssp_allreduce_argmin": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "jssp_peer_export": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p]),
    "jssp_peer_connect": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "jssp_peer_disconnect": (C.c_int32, [C.c_void_p]),
    "jssp_fetch_global_best": (C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "jssp_launch_count": (C.c_int64, [C.c_void_p]),
    "jssp_measure_fma_peak": (C.c_int32, [C.c_int32, C.c_int32, C.POINTER(C.c_double)]),
    "jssp_imm_update": (C.c_int32, [C.c_void_p, C.POINTER(JsspImmParams)] + [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_void_p]),
    "jssp_predict_modes": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "jssp_fop_evaluate": (C.c_int32, [C.c_void_p, C.POINTER(JsspFopIn), C.c_void_p, C.POINTER(JsspCycleOut)]),
    "jssp_project_states": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p]),
    "jssp_batch_create": (C.c_int32, [C.POINTER(JsspConfigC), C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "jssp_batch_destroy": (C.c_int32, [C.c_void_p]),
    "jssp_batch_last_cuda_error": (C.c_char_p, [C.c_void_p]),
    "jssp_batch_set_fop": (C.c_int32, [C.c_void_p, C.POINTER(JsspFopIn)]),
    "jssp_batch_set_scenario": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(JsspScenarioIn), C.c_void_p]),
    "jssp_batch_reset": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p]),
    "jssp_batch_run": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "jssp_batch_fetch": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "jssp_batch_best_traj": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "jssp_batch_launch_count": (C.c_int64, [C.c_void_p]),
}

_lib = None


def library_path() -> str:
    """In-tree libjssp_b200.so; JSSP_B200_LIB overrides it (A/B runs of two builds on the same box)."""
    return os.environ.get("JSSP_B200_LIB") or _build.LIB


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load libjssp_b200.so, (re)building it in-tree with nvcc when it is absent or was compiled from other sources than the
    ones on disk (build id mismatch).  Raises - never falls back to a CPU path, never loads a stale library silently."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    override = bool(os.environ.get("JSSP_B200_LIB"))
    if not override and _build.needs_build():
        if not build_if_missing:
            raise RuntimeError(f"{path} is missing or stale (build id {_build.library_id()} != sources {_build.source_id()}): "
                               "build it with `python -m dlii_jssp_b200.build` (no CPU fallback exists)")
        _build.build_library()
    elif not os.path.exists(path):
        raise RuntimeError(f"{path} does not exist (no CPU fallback exists)")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError here = the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    if not override and lib.jssp_build_id().decode() != _build.source_id():
        raise RuntimeError("libjssp_b200.so build id does not match the sources on disk; rebuild")
    if lib.jssp_abi_version() != 1:
        raise RuntimeError("libjssp_b200.so ABI version mismatch; rebuild")
    for which, st in enumerate(STRUCTS):
        if lib.jssp_struct_size(which) != C.sizeof(st):
            raise RuntimeError(f"libjssp_b200.so was built from another jssp_b200.h ({st.__name__} differs); rebuild")
    _lib = lib
    return lib


class JsspError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        msg = load().jssp_strerror(status).decode()
        super().__init__(f"{where}: {msg} [{status}]" + (f" ({detail})" if detail else ""))
