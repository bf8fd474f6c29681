#!/bin/bash
# Bench + profiles on the GPU box.  usage: tools_gpu_bench.sh [tag]
# (ncu runs use the 1024^2 configuration: the full-size launch list costs ~19 GPU-minutes)
tag=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
SMALL="python bench.py --size 1024 --steps 1 --warmup 1 --no-cpu-baseline"
$SMALL > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$tag.csv $SMALL > gpurun_out/ncu_list_$tag.log 2>&1
echo "ncu list exit $?"
$SMALL > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 4 -o gpurun_out/prof_gemm_$tag -f $SMALL > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu gemm exit $?"
$SMALL > gpurun_out/plain3_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 2 -c 2 -o gpurun_out/prof_attn_$tag -f $SMALL > gpurun_out/ncu_attn_$tag.log 2>&1
echo "ncu attn exit $?"
