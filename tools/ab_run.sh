#!/bin/bash
# same-box A/B of two library builds: tools/ab_run.sh <tag> [other.so]   (base = tools/_build/libbase.so)
tag=$1; other=${2:-tools/_build/libbase.so}
for L in "$other" ""; do
  echo "== lib: ${L:-product}"
  JSSP_B200_LIB=$L python tools/kernel_ab.py C3 tiled 2>&1 | tail -2
  JSSP_B200_LIB=$L python tools/kernel_ab.py C5 tiled 2>&1 | tail -2
  JSSP_B200_LIB=$L python tools/sweep_time.py 800 100 default 2>&1 | tail -1
  JSSP_B200_LIB=$L python tools/sweep_time.py 100 100 default 2>&1 | tail -1
done > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
