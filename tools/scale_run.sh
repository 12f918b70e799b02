#!/bin/bash
# the driver's scaling run, N = 1, 2, 4, 8 back to back on one 8-GPU box:  gpurun --gpus 8 -- bash tools/scale_run.sh <tag>
tag=${1:-scale}
mkdir -p gpurun_out
: > gpurun_out/${tag}_lines.jsonl
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 800 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_n$n.log 2> gpurun_out/${tag}_n$n.err
  else
    timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_n$n.log 2> gpurun_out/${tag}_n$n.err
  fi
  echo "N=$n rc=$?"
  tail -1 gpurun_out/${tag}_n$n.log >> gpurun_out/${tag}_lines.jsonl
done
python - <<PY
import json
for l in open("gpurun_out/${tag}_lines.jsonl"):
    try: d = json.loads(l)
    except Exception as e: print("bad line", l[:200]); continue
    sw, sp = d.get("sweep") or {}, d.get("split") or {}
    print(d["n_gpus"], "C3 ms", round(d["ms_per_step"], 4), "| sweep ms", round(sw.get("ms_per_step", 0), 3), "e2e", round(sw.get("e2e", {}).get("sweep_ms", 0), 2), sw.get("digest"),
          "| split ms", round(sp.get("ms_per_step", 0), 4), sp.get("best_idx"), sp.get("transport"), "e2e p50", round(sp.get("e2e", {}).get("p50_cycle_ms", 0), 4))
PY
