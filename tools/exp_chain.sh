#!/bin/bash
# single long message: lane kernel vs chain kernel, plus one ncu capture of the chain kernel (SHA-only)
mkdir -p gpurun_out
{
python tools/kbench.py 1 16777216 2 1
python tools/kbench.py 1 16777216 2 2
python tools/kbench.py 1 16777216 2 3
B200H_CHAIN=1 python tools/kbench.py 1 16777216 2 1
B200H_CHAIN=1 python tools/kbench.py 1 16777216 2 2
B200H_CHAIN=1 python tools/kbench.py 1 16777216 2 3
} > gpurun_out/kbench_chain1.txt 2>&1
cat gpurun_out/kbench_chain1.txt
B200H_CHAIN=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:chain_hash -c 1 -f -o gpurun_out/prof_chain_sha python tools/kbench.py 1 4194304 1 1 > gpurun_out/ncu_chain.log 2>&1
tail -3 gpurun_out/ncu_chain.log
