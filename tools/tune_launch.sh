#!/bin/bash
# occupancy experiment: rebuild the library with different block shapes and time the C3 evaluation kernel
for cfg in "256 2" "256 3" "192 3" "128 4" "128 5" "128 6" "64 8"; do
  set -- $cfg
  JSSP_NVCC_EXTRA="-DJSSP_EVAL_THREADS=$1 -DJSSP_EVAL_MIN_BLOCKS=$2" python dlii_jssp_b200/build.py 2>&1 | grep -E "evaluate_kernelILi2" -A2 | grep -E "spill|registers" | tr '\n' ' '
  echo
  python bench.py --steps 30 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 x $2', 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'step', round(d['ms_per_step'],4))"
done
