#!/bin/bash
# condense the captures of a tools/r02_run.sh <tag> run into profiles/ (run in the build container, after the library the run used was built):
#   tools/profiles_from_run.sh <tag>
tag=$1
BID=$(cat gpurun_out/${tag}_build_id.txt)
python tools/ncu_summary.py gpurun_out/${tag}_tiled.ncu-rep profiles/r02_evaluate_tiled_kernel evaluate_tiled_kernel $BID > /dev/null
python tools/ncu_summary.py gpurun_out/${tag}_dump.ncu-rep profiles/r02_dump_states_kernel dump_states_kernel $BID > /dev/null
python tools/ncu_summary.py gpurun_out/${tag}_replay.ncu-rep profiles/r02_replay_persistent_kernel_bestfirst replay_persistent_kernel $BID > /dev/null
python tools/ncu_lines.py gpurun_out/${tag}_tiled.ncu-rep 'evaluate_tiled_kernelILi2' 150 > profiles/r02_evaluate_tiled_kernel_lines.txt
python tools/ncu_lines.py gpurun_out/${tag}_dump.ncu-rep 'dump_states_kernel' 60 > profiles/r02_dump_states_kernel_lines.txt
python tools/ncu_lines.py gpurun_out/${tag}_replay.ncu-rep 'replay_persistent_kernelILi512ELb1' 120 > profiles/r02_replay_persistent_kernel_bestfirst_lines.txt
cp gpurun_out/${tag}_launches_c3.csv profiles/r02_launches_bench_c3.csv
cp gpurun_out/${tag}_launches_c4.csv profiles/r02_launches_bench_c4_sweep.csv
tail -1 gpurun_out/${tag}_bench.log > profiles/r02_bench_line_n1.json
tail -1 gpurun_out/${tag}_ref.log > profiles/r02_bench_line_reference_arm.json
echo "profiles/ <- run ${tag}, build ${BID}"
