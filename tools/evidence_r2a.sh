#!/bin/bash
# Round 2, first GPU call: the new native paths (parity), the bench line, launch list, trim scan numbers.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_gpu_tests.txt
cat gpurun_out/r2a_gpu_tests.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
tail -c 3000 gpurun_out/r2a_bench_n1.json; tail -5 gpurun_out/r2a_bench_n1.err
python tools/trim_bench.py > gpurun_out/r2a_trim_bench.txt 2>&1
cat gpurun_out/r2a_trim_bench.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_bench.csv \
    env B200H_BENCH_NMSG=100000 python bench.py --steps 2 --warmup 1 > gpurun_out/r2a_ncu_bench.log 2>&1
tail -3 gpurun_out/r2a_ncu_bench.log
