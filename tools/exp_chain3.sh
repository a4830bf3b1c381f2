#!/bin/bash
mkdir -p gpurun_out
{
./tools/outlier_bench 2048 4194304 0 | grep "flags=3"
./tools/outlier_bench 32768 262144 0 | grep "flags=3"
./tools/outlier_bench 64 4194304 0 | grep "flags=3"
./tools/outlier_bench 4 16777216 20000 | grep "flags=[13]"
} > gpurun_out/outlier_bench2.txt 2>&1
echo "== B200H_CHAIN=148 (new policy)" > gpurun_out/sweep_chain_policy.txt
B200H_CHAIN=148 python tools/sweep.py c4 c3 2>&1 | cut -c1-200 >> gpurun_out/sweep_chain_policy.txt
B200H_CHAIN=148 B200H_SWEEP_KMAX=7 python tools/sweep.py c5 2>&1 | cut -c1-200 >> gpurun_out/sweep_chain_policy.txt
cat gpurun_out/outlier_bench2.txt gpurun_out/sweep_chain_policy.txt
