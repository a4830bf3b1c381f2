#!/bin/bash
mkdir -p gpurun_out
python tools/c4ii_debug.py
python tools/sweep.py c4 c3 2>&1 | cut -c1-200 > gpurun_out/sweep_s2_final3.jsonl
B200H_SWEEP_KMAX=8 python tools/sweep.py c5 2>&1 | cut -c1-200 >> gpurun_out/sweep_s2_final3.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/sweep_s2_final3.jsonl"):
    try: d = json.loads(l)
    except Exception: continue
    print(d["config"][:60].ljust(60), d["n"], d["ms"], d["GBps"])
PY
