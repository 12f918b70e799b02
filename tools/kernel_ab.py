"""Kernel time of one plan cycle per evaluation kernel (tools/, GPU):  python tools/kernel_ab.py [C3|C5] [kernels...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import dlii_jssp_b200 as J

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
kerns = sys.argv[2:] or ["tiled", "group", "thread"]
cyc = bench.workload(name)
div = int(os.environ.get("AB_SEEDS_DIV", "1"))          # a rank's share of the seeds (candidate split over `div` ranks)
eng = bench.make_engine(J, cyc, len(cyc.seeds), 0)
seeds = torch.from_numpy(cyc.seeds[: len(cyc.seeds) // div].copy()).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for kern in kerns:
    for so in (False, True):
        def step():
            eng.evaluate(cyc.frenet0, cyc.time_step_now, cyc.final_time_step, seeds=seeds, targets=cyc.targets, sync=False, kernel=kern, select_only=so)
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        r = eng.fetch()
        print(f"{name} {kern:6s} {'select-only' if so else 'full       '}: {1e3 * sorted(ts)[len(ts) // 2]:8.1f} us  best {r.best_idx} cost {r.best_cost:.12f}", flush=True)
eng.close()
