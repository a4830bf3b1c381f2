#!/bin/bash
# Round-end evidence run: GPU test suite, bench line, ncu launch list, ncu --set full of the chain kernel, smoke, files bench.
# Outputs land in gpurun_out/; the summaries kept under profiles/ are made from them with tools/launch_list_md.py and tools/ncu_summary.py.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_s2_final.txt 2>&1; tail -2 gpurun_out/gpu_tests_s2_final.txt
python bench.py > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err
tail -1 gpurun_out/bench_s2.json | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_s2.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_s2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:chain_hash -s 1 -c 1 -f -o gpurun_out/prof_chain_s2 python tools/kbench.py 1 8388608 1 3 > gpurun_out/ncu_chain_s2.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python tools/files_bench.py 20000 2 /dev/shm 1.5 > gpurun_out/files_bench_s2.txt 2>&1; tail -7 gpurun_out/files_bench_s2.txt
