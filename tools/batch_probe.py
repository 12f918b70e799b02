"""Timing probe of the batched replay: python tools/batch_probe.py [n_scenarios] [cycles]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dlii_jssp_b200 as J
from dlii_jssp_b200 import batch as B, replay as R, scenario as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
cyc = int(sys.argv[2]) if len(sys.argv) > 2 else 100
t0 = time.perf_counter(); scs = [S.synthetic_intersection(i) for i in range(n)]; t1 = time.perf_counter()
packed = [B.pack_scenario(s, max_cycles=cyc) for s in scs]; t2 = time.perf_counter()
print(f"generate {t1-t0:.2f}s pack {t2-t1:.2f}s for {n} scenarios")
cfg = J.JsspConfig(max_seeds=1024, max_knots=max(len(p.knot_s) for p in packed), max_obstacle_modes=max(1, max(p.state.shape[0] for p in packed)),
                   max_total_steps=max(p.state.shape[2] for p in packed))
bt = B.BatchReplay(cfg, max_scenarios=n, max_cycles=cyc)
for kernel in (None, "thread", "group", "warp"):
    if kernel == "warp" and n > 256:
        continue
    t3 = time.perf_counter(); bt.load(packed); torch.cuda.synchronize(); t4 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); bt.run(cyc, kernel=kernel); e1.record(); torch.cuda.synchronize(); t5 = time.perf_counter()
    res = bt.results(); t6 = time.perf_counter()
    cycles = sum(r.cycles for r in res); cands = sum(sum(r.n_candidates) for r in res)
    print(f"kernel={kernel}: load {t4-t3:.3f}s run {e0.elapsed_time(e1):.1f} ms (host {t5-t4:.3f}s) fetch {t6-t5:.3f}s | cycles {cycles} candidates {cands} "
          f"-> {cands*51/(e0.elapsed_time(e1)*1e-3):.3e} cand-steps/s, {cycles/(e0.elapsed_time(e1)*1e-3):.0f} cycles/s; ends {sorted(set(r.end for r in res))}")
m = min(n, 8)
t7 = time.perf_counter(); solo = R.run_sweep(scs[:m], max_cycles=cyc); t8 = time.perf_counter()
print(f"run_sweep (one scenario at a time) {m} scenarios: {t8-t7:.3f}s, {sum(r.cycles for r in solo)/(t8-t7):.0f} cycles/s")
ok = all(a.best_idx == b.best_idx and a.ego_rows == b.ego_rows and a.end == b.end for a, b in zip(res[:m], solo))
print("batch == sweep on the first", m, ":", ok)
