"""Per-step latency of the batched replay for a rank-sized shard: python tools/latency_probe.py [n_scenarios] [cycles]
Sweeps the seed-kernel shape (JSSP_B200_TUNE_SEED_CLUSTER / _THREADS, read at every run call) and the evaluation kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dlii_jssp_b200 as J
from dlii_jssp_b200 import batch as B, scenario as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cyc = int(sys.argv[2]) if len(sys.argv) > 2 else 100
packed = [B.pack_scenario(S.synthetic_intersection(i), max_cycles=cyc) for i in range(n)]
packed = [p for p in packed if p.initial_end == -1]
cfg = J.JsspConfig(max_seeds=1024, max_knots=max(len(p.knot_s) for p in packed), max_obstacle_modes=max(1, max(p.state.shape[0] for p in packed)),
                   max_total_steps=max(p.state.shape[2] for p in packed))
bt = B.BatchReplay(cfg, max_scenarios=len(packed), max_cycles=cyc)
bt.load(packed)
for cl in ("", "1", "2", "4", "8"):
    for th in ("", "512", "1024"):
        if (cl == "") != (th == ""):
            continue
        for kernel in (None, "thread", "group", "warp"):
            for k, v in (("JSSP_B200_TUNE_SEED_CLUSTER", cl), ("JSSP_B200_TUNE_SEED_THREADS", th)):
                os.environ.pop(k, None)
                if v:
                    os.environ[k] = v
            best = 1e9
            for _ in range(3):
                bt.reset()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); bt.run(cyc, kernel=kernel); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"n={len(packed)} cluster={cl or 'auto'} threads={th or 'auto'} kernel={kernel}: {best:.2f} ms = {1e3*best/cyc:.1f} us/step", flush=True)
