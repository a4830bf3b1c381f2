#!/bin/bash
# Round 2, second GPU call: map-pump e2e tuning (window size / windows in flight), compute-sanitizer logs.
set -x
mkdir -p gpurun_out
for cfg in "1073741824 2" "2147483648 2" "4294967296 2" "1073741824 3" "536870912 2" "1073741824 1"; do
  set -- $cfg
  B200H_PUMP_WINDOW_BYTES=$1 B200H_PUMP_WINDOWS_IN_FLIGHT=$2 B200H_CPU_SAMPLE=1024 python bench.py --steps 3 --warmup 2 \
    2>gpurun_out/r2b_bench_$1_$2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('window',$1,'inflight',$2,'e2e',d['e2e']['value'],'pinned',d['e2e_pinned']['value'],'value',d['value'],d['parity'])" | tee -a gpurun_out/r2b_pump_sweep.txt
done
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/r2b_sanitizer_memcheck.log python tools/sanitize_cases.py > gpurun_out/r2b_sanitizer_memcheck.out 2>&1
tail -3 gpurun_out/r2b_sanitizer_memcheck.out; tail -5 gpurun_out/r2b_sanitizer_memcheck.log
SAN_N=100000 timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2b_sanitizer_racecheck.log python tools/sanitize_cases.py > gpurun_out/r2b_sanitizer_racecheck.out 2>&1
tail -3 gpurun_out/r2b_sanitizer_racecheck.out; tail -5 gpurun_out/r2b_sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --log-file gpurun_out/r2b_sanitizer_synccheck.log python tools/sanitize_cases.py > gpurun_out/r2b_sanitizer_synccheck.out 2>&1
tail -3 gpurun_out/r2b_sanitizer_synccheck.out; tail -5 gpurun_out/r2b_sanitizer_synccheck.log
