"""End-to-end sweep (upload + replay + fetch) with and without overlapping the upload of one sub-batch with the replay of another:
    python tools/e2e_overlap_probe.py [n] [kernel] [k ...]"""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dlii_jssp_b200 as J
from dlii_jssp_b200 import batch as B, scenario as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 800
kernel = sys.argv[2] if len(sys.argv) > 2 else "bestfirst"
kernel = None if kernel == "default" else kernel
ks = [int(v) for v in sys.argv[3:]] or [1, 2, 4, 8]
cyc = 100
packed = [B.pack_scenario(S.synthetic_intersection(i), max_cycles=cyc) for i in range(n)]
packed = [p for p in packed if p.initial_end == -1]
cfg = J.JsspConfig(max_seeds=1024, max_knots=max(64, max(len(p.knot_s) for p in packed)), max_obstacle_modes=max(1, max(p.state.shape[0] for p in packed)),
                   max_total_steps=max(p.state.shape[2] for p in packed))
ref = None
for k in ks:
    bt = B.MultiStreamReplay(cfg, max_scenarios=len(packed), max_cycles=cyc, k=k)
    bt.load(packed)
    dev = []
    for _ in range(4):
        bt.reset(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bt.run(cyc, kernel=kernel); e1.record(); torch.cuda.synchronize()
        dev.append(e0.elapsed_time(e1))
    a, b = [], []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        bt.load(packed); bt.run(cyc, kernel=kernel); r1 = bt.results(lists=False)
        a.append(time.perf_counter() - t0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        bt.load_and_run(packed, cyc, kernel=kernel); r2 = bt.results(lists=False)
        b.append(time.perf_counter() - t0)
    sig = hash(tuple(tuple(int(v) for v in r.best_idx) for r in r2)) & 0xffffffff
    print(f"n={len(packed)} kernel={kernel} k={k}: device {min(dev):.2f} ms | e2e load,run,results {1e3 * statistics.median(a):.1f} ms | load_and_run,results {1e3 * statistics.median(b):.1f} ms | selections {sig:08x}", flush=True)
    bt.close()
