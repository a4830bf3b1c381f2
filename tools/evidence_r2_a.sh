#!/bin/bash
# Round 2 evidence, call A (one GPU, short): GPU test suite, bench line, reference arm, smoke.
mkdir -p gpurun_out
set -x
free -g | head -2; nproc; nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests.txt 2>&1; tail -3 gpurun_out/r2_gpu_tests.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 1800 gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_reference_arm.json 2> gpurun_out/r2_reference_arm.err; tail -c 600 gpurun_out/r2_reference_arm.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
