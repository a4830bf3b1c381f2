#!/bin/bash
# Round 2 evidence, call B (one GPU, short): kernel-lever variants, chain-SM yielding, drop-in concurrency, trim scan,
# ncu launch list + full captures of the dominant kernel and of the trim kernels.
mkdir -p gpurun_out
set -x
for v in default build_variants/lib_fusedrot4.so build_variants/lib_fusedrot1.so build_variants/lib_fusedrot2.so; do
  if [ "$v" = default ]; then python tools/kbench.py 100000 262144 5 3; else B200H_LIB=$PWD/$v python tools/kbench.py 100000 262144 5 3; fi
done > gpurun_out/r2_kbench_fusedrot.txt 2>&1; cat gpurun_out/r2_kbench_fusedrot.txt
for y in 1 0; do
  B200H_YIELD_CHAIN_SMS=$y J1_C3_GIB=12.5 J1_C3_FILES=131072 J1_E2E_CAP=0.001 python tools/j1_matrix.py c3 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('yield_chain_sms=$y', d['config'][:40], 'kernel_ms', d['kernel_ms'], 'outliers', d['outliers_on_rank0'])"
done > gpurun_out/r2_yield_chain_sms.txt 2>&1; cat gpurun_out/r2_yield_chain_sms.txt
timeout 300 python tools/dropin_concurrency.py > gpurun_out/r2_dropin_concurrency.txt 2>&1; cut -c1-400 gpurun_out/r2_dropin_concurrency.txt
python tools/trim_bench.py > gpurun_out/r2_trim_bench.txt 2>&1; cat gpurun_out/r2_trim_bench.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:lane_hash -s 2 -c 1 -f -o gpurun_out/r2_prof_lane \
    python tools/kbench.py 100000 262144 1 3 > gpurun_out/r2_ncu_lane.log 2>&1; tail -2 gpurun_out/r2_ncu_lane.log
TRIM_REPS=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:trim_ -s 2 -c 2 -f -o gpurun_out/r2_prof_trim \
    python tools/trim_bench.py blank > gpurun_out/r2_ncu_trim.log 2>&1; tail -2 gpurun_out/r2_ncu_trim.log
B200H_CPU_SAMPLE=512 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/r2_ncu_bench.log 2>&1; tail -c 300 gpurun_out/r2_ncu_bench.log
