#!/bin/bash
mkdir -p gpurun_out
{
for n in 4 128 296 512 592; do echo "n=$n x 4 MiB"; ./tools/outlier_bench $n 4194304 0 | grep "mode=chain\|NO"; done
echo "4 x 16 MiB + 20000 small"; ./tools/outlier_bench 4 16777216 20000 | grep "mode=chain\|NO"
} > gpurun_out/outlier_bench4.txt 2>&1
cat gpurun_out/outlier_bench4.txt
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
python tools/sweep.py c4 c3 2>&1 | cut -c1-200 > gpurun_out/sweep_s2_final2.jsonl
B200H_SWEEP_KMAX=8 python tools/sweep.py c5 2>&1 | cut -c1-200 >> gpurun_out/sweep_s2_final2.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/sweep_s2_final2.jsonl"):
    try: d = json.loads(l)
    except Exception: continue
    print(d["config"][:60].ljust(60), d["n"], d["ms"], d["GBps"])
PY
