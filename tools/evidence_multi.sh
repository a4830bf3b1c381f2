#!/bin/bash
# usage: evidence_multi.sh N [j1 parts, default "c3 c5"] [tag]  -- bench + j1 matrix on N GPUs of one box
N=$1
PARTS=${2:-"c3 c5"}
TAG=${3:-r2}
set -x
mkdir -p gpurun_out
rm -f gpurun_out/j1_n$N.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
  > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
tail -c 2500 gpurun_out/${TAG}_bench_n$N.json; tail -3 gpurun_out/${TAG}_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/j1_matrix.py $PARTS \
  > gpurun_out/${TAG}_j1_n$N.log 2>&1
cp gpurun_out/j1_n$N.jsonl gpurun_out/${TAG}_j1_n$N.jsonl
tail -30 gpurun_out/${TAG}_j1_n$N.log | cut -c1-400
