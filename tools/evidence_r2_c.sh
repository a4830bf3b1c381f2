#!/bin/bash
# Round 2 evidence, call C (one GPU, short): the reference's CPU path on real files (C4, C3-v2) and compute-sanitizer
# on the final library.
mkdir -p gpurun_out
set -x
rm -f gpurun_out/r2_j1_cpu_paths.jsonl
timeout 420 python tools/j1_cpu_paths.py c4 c3 > gpurun_out/r2_j1_cpu_paths.log 2>&1; grep '^{' gpurun_out/r2_j1_cpu_paths.log | cut -c1-300
timeout 200 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck.log python tools/sanitize_cases.py > gpurun_out/r2_sanitizer_memcheck.out 2>&1
tail -2 gpurun_out/r2_sanitizer_memcheck.out; tail -2 gpurun_out/r2_sanitizer_memcheck.log
SAN_N=100000 timeout 300 compute-sanitizer --tool racecheck --log-file gpurun_out/r2_sanitizer_racecheck.log python tools/sanitize_cases.py > gpurun_out/r2_sanitizer_racecheck.out 2>&1
tail -2 gpurun_out/r2_sanitizer_racecheck.out; tail -2 gpurun_out/r2_sanitizer_racecheck.log
timeout 200 compute-sanitizer --tool synccheck --log-file gpurun_out/r2_sanitizer_synccheck.log python tools/sanitize_cases.py > gpurun_out/r2_sanitizer_synccheck.out 2>&1
tail -2 gpurun_out/r2_sanitizer_synccheck.out; tail -2 gpurun_out/r2_sanitizer_synccheck.log
