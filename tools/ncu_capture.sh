#!/bin/bash
# ncu capture of one kernel of a bench command, as the profiling recipe prescribes (run under gpurun, one GPU):
#   tools/ncu_capture.sh <out-name> <kernel-regex> <launches-to-skip> <bench args...>
# 1) the plain command must exit 0; 2) launch list (gpu__time_duration.sum); 3) --set full capture with source.
# Outputs in gpurun_out/: <name>_plain.log, <name>_launches.csv, <name>.ncu-rep
set -u
name=$1; regex=$2; skip=$3; shift 3
mkdir -p gpurun_out
python bench.py "$@" > gpurun_out/${name}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${name}_launches.csv python bench.py "$@" > gpurun_out/${name}_ncu1.log 2>&1
python bench.py "$@" > gpurun_out/${name}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${regex} -s ${skip} -c 1 -f -o gpurun_out/${name} python bench.py "$@" > gpurun_out/${name}_ncu2.log 2>&1
tail -3 gpurun_out/${name}_ncu2.log
python -c "from dlii_jssp_b200 import build; print('build_id', build.library_id())" | tee gpurun_out/${name}_build_id.txt
