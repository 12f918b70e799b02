#!/bin/bash
# Round-2 verification run (under gpurun, one GPU): GPU tests, smoke, the default bench line, the reference arm, then the ncu
# launch list and the two --set full captures (tiled evaluation kernel, dump kernel) of the C3 command.
#   tools/r02_run.sh [tag]      outputs in gpurun_out/<tag>_*
set -x
tag=${1:-r02}
mkdir -p gpurun_out
python -c "from dlii_jssp_b200 import build; print(build.library_id())" > gpurun_out/${tag}_build_id.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_gputests.log 2>&1; tail -3 gpurun_out/${tag}_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${tag}_bench.log 2>&1; tail -1 gpurun_out/${tag}_bench.log | cut -c1-600
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_ref.log 2>&1; tail -1 gpurun_out/${tag}_ref.log | cut -c1-300
ARGS="--workload C3 --steps 5 --warmup 3 --no-cpu"
timeout 300 python bench.py $ARGS > gpurun_out/${tag}_c3_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c3.csv python bench.py $ARGS > gpurun_out/${tag}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:evaluate_tiled_kernel -s 6 -c 1 -f -o gpurun_out/${tag}_tiled python bench.py $ARGS > gpurun_out/${tag}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dump_states_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_dump python bench.py $ARGS > gpurun_out/${tag}_ncu3.log 2>&1
# the sweep's replay kernel (best-first persistent kernel, one launch per sweep): launch list + one full capture
SW="--workload C4 --steps 3 --warmup 3 --no-cpu"
timeout 600 python bench.py $SW > gpurun_out/${tag}_c4_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_c4.csv python bench.py $SW > gpurun_out/${tag}_ncu4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:replay_persistent_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_replay python bench.py $SW > gpurun_out/${tag}_ncu5.log 2>&1
ls -la gpurun_out | tail -14
