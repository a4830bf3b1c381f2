#!/bin/bash
# One-shot experiment (round 1, session 2): issue-rate microbenchmarks + rotate-by-multiply kernel variants.
mkdir -p gpurun_out
./tools/ubench > gpurun_out/ubench_r1b.txt 2>&1
{
python tools/kbench.py 100000 262144 5 3
for v in 1 2 3; do B200H_LIB=$PWD/build_variants/libb200hash_rotw$v.so python tools/kbench.py 100000 262144 5 3; done
B200H_LIB=$PWD/build_variants/libb200hash_rotw1.so python tools/kbench.py 100000 262144 5 1
python tools/kbench.py 100000 262144 5 1
} > gpurun_out/kbench_rotw.txt 2>&1
B200H_LIB=$PWD/build_variants/libb200hash_rotw3.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "every_length or aligned_and_mixed or golden or single_flags" > gpurun_out/parity_rotw3.txt 2>&1
tail -3 gpurun_out/parity_rotw3.txt; cat gpurun_out/kbench_rotw.txt; tail -40 gpurun_out/ubench_r1b.txt
