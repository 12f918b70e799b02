"""Phase timing of the seed enumeration kernel (profiling build): JSSP_NVCC_EXTRA=-DJSSP_PROFILE_PHASES python -m dlii_jssp_b200.build"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dlii_jssp_b200 as J
eng = J.JsspEngine(J.JsspConfig())
fn = eng.lib.jssp_debug_seed_stamps; fn.restype = C.c_int32; fn.argtypes = [C.c_void_p]
for st in ((2.9, 5.686, 0.0), (10.0, 0.0, 0.0), (20.0, 10.0, 1.0)):
    for _ in range(3):
        n = eng.enumerate_seeds(*st)
    ts = np.zeros(16, dtype=np.uint64); assert fn(ts.ctypes.data) == 0
    t = (ts[:8].astype(np.int64) - int(ts[0])) / 1e3
    print(st, n, "stamps us:", np.round(t, 1), "| zero, L1, L2, L3, sort, sync, emit")
