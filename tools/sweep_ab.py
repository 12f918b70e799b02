"""A/B of the replay sweep (tools/, GPU): persistent kernel (threads by env JSSP_B200_TUNE_SEED_THREADS) vs lock-step; prints ms per sweep and
whether every scenario replays identically.   python tools/sweep_ab.py [n_scenarios=800] [cycles=100]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dlii_jssp_b200 as J
from dlii_jssp_b200 import batch as B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 800
C = int(sys.argv[2]) if len(sys.argv) > 2 else 100
packed = [B.pack_scenario(J.scenario.synthetic_intersection(i), max_cycles=C) for i in range(n)]
packed = [p for p in packed if p.initial_end == -1]
e = packed[0].ego
cfg = J.JsspConfig(ego_length=float(e[5]), ego_width=float(e[6]), max_seeds=1024, max_knots=max(64, max(len(p.knot_s) for p in packed)),
                   max_obstacle_modes=max(1, max(p.state.shape[0] * p.state.shape[1] for p in packed)), max_total_steps=max(p.state.shape[2] for p in packed))
bt = B.BatchReplay(cfg, max_scenarios=len(packed), max_cycles=C)
bt.load(packed)
out = {}
kerns = (("lockstep",) if os.environ.get("SWEEP_AB_NO_PERSISTENT") else ("persistent", "lockstep")) + tuple(sys.argv[3:])
for kern in kerns:
    for _ in range(2):
        bt.reset(); bt.run(C, kernel=kern)
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bt.reset(); bt.run(C, kernel=kern); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res = bt.results()
    out[kern] = res
    print(f"{str(kern):>9}: {min(ts):8.3f} ms per sweep of {len(packed)} scenarios, {sum(r.cycles for r in res)} plan cycles, max cycles {max(r.cycles for r in res)}")
def rows(r):
    return (r.end, r.cycles, list(r.best_idx), list(r.n_candidates))
for kern in kerns:
    if kern == "lockstep":
        continue
    bad = [(a.name, a.cycles, b.cycles, a.end, b.end) for a, b in zip(out[kern], out["lockstep"]) if rows(a) != rows(b)]
    print(f"{kern} vs lockstep: scenarios that differ: {len(bad)}", bad[:8])
    for a, b in zip(out[kern], out["lockstep"]):
        if list(a.best_idx) != list(b.best_idx):
            k = next(i for i, (x, y) in enumerate(zip(a.best_idx, b.best_idx)) if x != y)
            print(" first difference", a.name, "cycle", k, a.best_idx[k], b.best_idx[k], a.n_candidates[k], b.n_candidates[k], "overflow", a.seed_overflow, b.seed_overflow)
            break
