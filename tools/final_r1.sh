#!/bin/bash
# Round-end evidence run (session 2): bench both arms, launch list, ncu --set full of the kernels that changed.
mkdir -p gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ref_s2.json 2> gpurun_out/ref_s2.err
python bench.py > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err
tail -1 gpurun_out/bench_s2.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_s2.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch_s2.log 2>&1
# SHA-256-only lane kernel (IMAD.WIDE rotates) and the chain kernel on one long message
timeout 400 ncu --set full --import-source on --clock-control none -k regex:lane_hash -s 2 -c 1 -f -o gpurun_out/prof_lane_sha_s2 python tools/kbench.py 100000 262144 1 1 > gpurun_out/ncu_sha_s2.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:chain_hash -s 2 -c 1 -f -o gpurun_out/prof_chain_s2 python tools/kbench.py 1 8388608 1 3 > gpurun_out/ncu_chain_s2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
python tools/sweep.py c4 c3 2>&1 | cut -c1-220 > gpurun_out/sweep_s2_final.jsonl
B200H_SWEEP_KMAX=8 python tools/sweep.py c5 2>&1 | cut -c1-220 >> gpurun_out/sweep_s2_final.jsonl
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_s2_final.txt 2>&1; tail -2 gpurun_out/gpu_tests_s2_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
