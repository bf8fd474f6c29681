#!/bin/bash
# ncu --set full of the fc1 / fc2 GEMM in the three precisions (one launch each).  usage: tools_gpu_ncu_gemm.sh <tag>
mkdir -p gpurun_out
export RIBCA_LIB=$PWD/multiplexed_image_annotator_b200/build/libribca_e8.so
python tools_gemm_one.py 4096 fc1 fc2 > gpurun_out/plain_gemm_one_$1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -o gpurun_out/prof_gemm3_$1 -f python tools_gemm_one.py 4096 fc1 fc2 > gpurun_out/ncu_gemm3_$1.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_gemm3_$1.log
