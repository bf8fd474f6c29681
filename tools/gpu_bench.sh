#!/bin/bash
# Bench + profiles on the GPU box.  usage: tools/gpu_bench.sh [tag]
# (the launch list uses the 1024^2 configuration: a full-size list costs ~19 GPU-minutes; the --set full captures use
#  one launch of each block GEMM at the bench's chunk size (4096 cells, M = 413696) and one attention launch)
tag=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "reference arm exit $?"; tail -c 400 gpurun_out/bench_ref_$tag.json
SMALL="python bench.py --size 1024 --steps 1 --warmup 1 --quick"
$SMALL > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$tag.csv $SMALL > gpurun_out/ncu_list_$tag.log 2>&1
echo "ncu list exit $?"
export PRECS=f16f8
python tools/gemm_one.py 4096 qkv proj fc1 fc2 > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -o gpurun_out/prof_gemm_$tag -f python tools/gemm_one.py 4096 qkv proj fc1 fc2 > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu gemm exit $?"
MID="python bench.py --size 2048 --steps 1 --warmup 1 --quick"
$MID > gpurun_out/plain3_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 14 -c 1 -o gpurun_out/prof_attn_$tag -f $MID > gpurun_out/ncu_attn_$tag.log 2>&1
echo "ncu attn exit $?"
