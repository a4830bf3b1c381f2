#!/bin/bash
# usage: tools/exp_variants.sh <out-name> lib1 lib2 ...   (kernel-only timing of library variants, fused flags=3)
out=gpurun_out/$1; shift
mkdir -p gpurun_out
{
python tools/kbench.py 100000 262144 5 ${KFLAGS:-3}
for v in "$@"; do B200H_LIB=$PWD/build_variants/$v python tools/kbench.py 100000 262144 5 ${KFLAGS:-3}; done
} > $out 2>&1
cat $out
