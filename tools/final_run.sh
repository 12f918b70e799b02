set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/final_c3.log 2>&1; tail -1 gpurun_out/final_c3.log | cut -c1-400
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.log 2>&1; tail -1 gpurun_out/final_ref.log | cut -c1-300
python bench.py --workload C5 --steps 20 --warmup 3 --no-cpu > gpurun_out/final_c5.log 2>&1; tail -1 gpurun_out/final_c5.log | cut -c1-300
python bench.py --workload C4 --no-cpu > gpurun_out/final_c4.log 2>&1; tail -1 gpurun_out/final_c4.log | cut -c1-300
python bench.py --workload C2 --no-cpu > gpurun_out/final_c2.log 2>&1; tail -1 gpurun_out/final_c2.log | cut -c1-300
python bench.py --workload C4 --planner frenet --scenarios 100 --no-cpu > gpurun_out/final_c4fop.log 2>&1; tail -1 gpurun_out/final_c4fop.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"seed_enum|evaluate|pack" -c 200 --csv --log-file gpurun_out/final_launches_c3.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/final_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"evaluate_group_kernel" -s 6 -c 1 -o gpurun_out/final_group_full python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/final_ncu2.log 2>&1
ls -la gpurun_out | tail -12
