#!/bin/bash
mkdir -p gpurun_out
for c in 0 148 296; do
  echo "== B200H_CHAIN=$c"
  B200H_CHAIN=$c python tools/sweep.py c4 c3 2>&1 | cut -c1-220
done > gpurun_out/sweep_chain_c3c4.txt 2>&1
for c in 0 148; do
  echo "== B200H_CHAIN=$c"
  B200H_CHAIN=$c B200H_SWEEP_KMAX=7 python tools/sweep.py c5 2>&1 | cut -c1-200
done > gpurun_out/sweep_chain_c5.txt 2>&1
cat gpurun_out/sweep_chain_c3c4.txt gpurun_out/sweep_chain_c5.txt
