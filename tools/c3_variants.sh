#!/bin/bash
# C3 / C5 kernel time of the tiled kernel under the tools/ tuning overrides (GPU):  tools/c3_variants.sh
run() { python - <<'PY'
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
import bench
import dlii_jssp_b200 as J
for name in ("C3", "C5"):
    cyc = bench.workload(name)
    eng = bench.make_engine(J, cyc, len(cyc.seeds), 0)
    seeds = torch.from_numpy(cyc.seeds).cuda()
    def step(): eng.evaluate(cyc.frenet0, cyc.time_step_now, cyc.final_time_step, seeds=seeds, targets=cyc.targets, sync=False)
    for _ in range(5): step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    r = eng.fetch()
    print(f"  {name}: {1e3 * sorted(ts)[len(ts) // 2]:8.1f} us  best {r.best_idx}", flush=True)
    eng.close()
PY
}
echo "default"; run
echo "no masks"; JSSP_B200_TUNE_NO_MASKS=1 run
for r in 1 3 4; do echo "tile rounds $r"; JSSP_B200_TUNE_TILE_ROUNDS=$r run; echo "tile rounds $r, no masks"; JSSP_B200_TUNE_TILE_ROUNDS=$r JSSP_B200_TUNE_NO_MASKS=1 run; done
