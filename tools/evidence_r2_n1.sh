#!/bin/bash
# Round 2, one-GPU evidence run: GPU test suite, bench line (+ reference arm), BASELINE configs C3/C4/C5 with the
# reference's CPU path beside them, drop-in concurrency table, kernel-variant check, ncu launch lists and full captures.
# Outputs land in gpurun_out/; the summaries kept under profiles/ are made from them.
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests.txt 2>&1; tail -3 gpurun_out/r2_gpu_tests.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 1800 gpurun_out/r2_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_reference_arm.json 2> gpurun_out/r2_reference_arm.err; tail -c 600 gpurun_out/r2_reference_arm.json
# kernel lever (one bounded attempt): sigma shifts / Sigma rotates of the FUSED kernel on the FMA pipe
for v in default build_variants/lib_fusedrot4.so build_variants/lib_fusedrot1.so build_variants/lib_fusedrot2.so; do
  if [ "$v" = default ]; then python tools/kbench.py 100000 262144 5 3; else B200H_LIB=$PWD/$v python tools/kbench.py 100000 262144 5 3; fi
done > gpurun_out/r2_kbench_fusedrot.txt 2>&1; cat gpurun_out/r2_kbench_fusedrot.txt
# chain SMs left to the chains vs shared with lane CTAs (round-1 behaviour), on one rank-of-8 share of C3
for y in 1 0; do
  B200H_YIELD_CHAIN_SMS=$y J1_C3_GIB=12.5 J1_C3_FILES=131072 J1_E2E_CAP=0.001 python tools/j1_matrix.py c3 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('yield_chain_sms=$y', d['config'][:40], 'kernel_ms', d['kernel_ms'], 'outliers', d['outliers_on_rank0'])"
done > gpurun_out/r2_yield_chain_sms.txt 2>&1; cat gpurun_out/r2_yield_chain_sms.txt
rm -f gpurun_out/j1_n1.jsonl
python tools/j1_matrix.py c3 c4 c5 cpu > gpurun_out/r2_j1_n1.log 2>&1; cp gpurun_out/j1_n1.jsonl gpurun_out/r2_j1_n1.jsonl; tail -40 gpurun_out/r2_j1_n1.log | cut -c1-300
python tools/dropin_concurrency.py > gpurun_out/r2_dropin_concurrency.txt 2>&1; cat gpurun_out/r2_dropin_concurrency.txt | cut -c1-400
python tools/trim_bench.py > gpurun_out/r2_trim_bench.txt 2>&1; cat gpurun_out/r2_trim_bench.txt
python tools/pump_probe.py > gpurun_out/r2_pump_probe.txt 2>&1; cat gpurun_out/r2_pump_probe.txt
# ncu: launch list of the bench (kernel shares), full captures of the dominant kernel and of the trim scan
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/r2_ncu_bench.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:lane_hash -s 2 -c 1 -f -o gpurun_out/r2_prof_lane \
    python tools/kbench.py 100000 262144 1 3 > gpurun_out/r2_ncu_lane.log 2>&1
TRIM_REPS=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:trim_ -s 2 -c 2 -f -o gpurun_out/r2_prof_trim \
    python tools/trim_bench.py blank > gpurun_out/r2_ncu_trim.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
