#!/bin/bash
# ncu capture of the NTT kernels of one LDE + interpolation at 2^20: tools/ncu_ntt.sh <tag> [ZK_LIB path]
tag=$1; lib=$2
[ -n "$lib" ] && export ZK_LIB=$lib
ncu --set full --clock-control none -k regex:ntt -c 8 -o gpurun_out/r2_ntt_$tag -f python tools/ntt_one.py 20 1 > gpurun_out/r2_ntt_$tag.log 2>&1
tail -1 gpurun_out/r2_ntt_$tag.log
