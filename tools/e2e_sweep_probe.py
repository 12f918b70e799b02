"""Where the end-to-end time of a replay sweep goes: python tools/e2e_sweep_probe.py [n_scenarios] [cycles]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dlii_jssp_b200 as J
from dlii_jssp_b200 import batch as B, scenario as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 800
cyc = int(sys.argv[2]) if len(sys.argv) > 2 else 100
packed = [B.pack_scenario(S.synthetic_intersection(i), max_cycles=cyc) for i in range(n)]
packed = [p for p in packed if p.initial_end == -1]
cfg = J.JsspConfig(max_seeds=1024, max_knots=max(len(p.knot_s) for p in packed), max_obstacle_modes=max(1, max(p.state.shape[0] for p in packed)),
                   max_total_steps=max(p.state.shape[2] for p in packed))
bt = B.BatchReplay(cfg, max_scenarios=len(packed), max_cycles=cyc)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bt.load(packed); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    bt.run(cyc); t3 = time.perf_counter()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    res, rec = bt.fetch(); t5 = time.perf_counter()
    out = bt.results(); t6 = time.perf_counter()
    print(f"iter {it}: load call {1e3*(t1-t0):.1f} ms (+{1e3*(t2-t1):.1f} until uploaded), run enqueue {1e3*(t3-t2):.1f} (+{1e3*(t4-t3):.1f} until done), "
          f"fetch {1e3*(t5-t4):.1f}, results() {1e3*(t6-t5):.1f} ms", flush=True)
