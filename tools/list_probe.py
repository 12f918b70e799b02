"""Entries of the per-block culled obstacle lists (profiling build: JSSP_NVCC_EXTRA=-DJSSP_PROFILE_PHASES, JSSP_B200_LIB=<that .so>):
    python tools/list_probe.py [C3|C5]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import dlii_jssp_b200 as J
for which in (sys.argv[1:] or ["C3", "C5"]):
    cyc = bench.workload(which)
    eng = bench.make_engine(J, cyc, len(cyc.seeds), 0)
    seeds = torch.from_numpy(cyc.seeds).cuda()
    ts = torch.zeros((1 << 16, 16), dtype=torch.int64, device="cuda")
    fn = eng.lib.jssp_debug_phase_buffer; fn.restype = C.c_int32; fn.argtypes = [C.c_void_p, C.c_void_p]
    assert fn(eng._h, ts.data_ptr()) == 0
    eng.evaluate(cyc.frenet0, cyc.time_step_now, cyc.final_time_step, seeds=seeds, targets=cyc.targets, sync=True)
    t = ts.cpu().numpy()
    used = t[:, 0] > 0
    v = t[used, 15]
    n, ovf = v & 0xffffffff, v >> 32
    rows = (min(int(cyc.plan_horizon / 0.1) + 1, cyc.final_time_step - cyc.time_step_now) + 1) // 2
    print(f"{which}: {used.sum()} blocks, list entries per block: min {n.min()} median {int(np.median(n))} p99 {int(np.quantile(n, 0.99))} max {n.max()} "
          f"(~{np.median(n) / rows:.1f} per row of {rows}); blocks that overflowed: {int((ovf > 0).sum())}", flush=True)
    eng.close()
