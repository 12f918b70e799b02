"""Best-of-3 device time of the batched replay sweep: python tools/sweep_time.py [n_scenarios] [cycles] [kernel ...]
(A/B of library builds through JSSP_B200_LIB.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dlii_jssp_b200 as J
from dlii_jssp_b200 import batch as B, scenario as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 800
cyc = int(sys.argv[2]) if len(sys.argv) > 2 else 100
kernels = sys.argv[3:] or ["default"]
stride = int(os.environ.get("SWEEP_STRIDE", "1"))      # scenario ids 0, stride, 2 * stride, ... (a rank's shard of the 800-scenario sweep)
packed = [B.pack_scenario(S.synthetic_intersection(i * stride), max_cycles=cyc) for i in range(n)]
packed = [p for p in packed if p.initial_end == -1]
cfg = J.JsspConfig(max_seeds=int(os.environ.get("SWEEP_MAX_SEEDS", "1024")), max_knots=max(len(p.knot_s) for p in packed), max_obstacle_modes=max(1, max(p.state.shape[0] for p in packed)),
                   max_total_steps=max(p.state.shape[2] for p in packed))
bt = B.BatchReplay(cfg, max_scenarios=len(packed), max_cycles=cyc)
bt.load(packed)
for kernel in kernels:
    k = None if kernel == "default" else kernel
    best = 1e9
    for _ in range(4):
        bt.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bt.run(cyc, kernel=k); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res = bt.results()
    sig = hash(tuple(tuple(r.best_idx) for r in res)) & 0xffffffff
    print(f"{os.path.basename(os.environ.get('JSSP_B200_LIB', 'product'))} n={len(packed)} kernel={kernel}: {best:.2f} ms ({1e3 * best / cyc:.1f} us/step) selections {sig:08x}", flush=True)
