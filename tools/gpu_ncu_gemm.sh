#!/bin/bash
# ncu --set full of one launch of each block GEMM (qkv, proj, fc1, fc2) at the bench's chunk size.  usage: tools/gpu_ncu_gemm.sh <tag> [precisions]
mkdir -p gpurun_out
export PRECS=${2:-f16f8}
python tools/gemm_one.py 4096 qkv proj fc1 fc2 > gpurun_out/plain_gemm_one_$1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -o gpurun_out/prof_gemm_$1 -f python tools/gemm_one.py 4096 qkv proj fc1 fc2 > gpurun_out/ncu_gemm_$1.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_gemm_$1.log
