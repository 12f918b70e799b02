"""Best-first vs lock-step replay: first diverging (scenario, cycle) and the time per kernel.  python tools/bf_debug.py [n] [cycles]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dlii_jssp_b200 as J
from dlii_jssp_b200 import batch as B, scenario as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cyc = int(sys.argv[2]) if len(sys.argv) > 2 else 100
packed = [B.pack_scenario(S.synthetic_intersection(i), max_cycles=cyc) for i in range(n)]
packed = [p for p in packed if p.initial_end == -1]
cfg = J.JsspConfig(max_seeds=1024, max_knots=max(len(p.knot_s) for p in packed), max_obstacle_modes=max(1, max(p.state.shape[0] for p in packed)),
                   max_total_steps=max(p.state.shape[2] for p in packed))
bt = B.BatchReplay(cfg, max_scenarios=len(packed), max_cycles=cyc)
bt.load(packed)
out = {}
for kernel in (None, "bestfirst"):
    bt.reset(); bt.run(cyc, kernel=kernel); torch.cuda.synchronize()
    out[kernel] = bt.results()
bad = 0
for s, (a, b) in enumerate(zip(out[None], out["bestfirst"])):
    if a.best_idx != b.best_idx or a.end != b.end:
        c = next((i for i, (x, y) in enumerate(zip(a.best_idx, b.best_idx)) if x != y), None)
        print(f"scenario {s}: first difference at cycle {c}: default {a.best_idx[c] if c is not None else None} bestfirst {b.best_idx[c] if c is not None else None} "
              f"N {a.n_candidates[c] if c is not None else None}/{b.n_candidates[c] if c is not None else None} ends {a.end} {b.end} cycles {len(a.best_idx)} {len(b.best_idx)}")
        bad += 1
print(f"{bad} of {len(packed)} scenarios differ")
