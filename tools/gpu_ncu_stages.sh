#!/bin/bash
# ncu --set full of one launch of each stage 1-3 / 5 / LayerNorm kernel (HBM-side evidence).  usage: tools/gpu_ncu_stages.sh <tag>
tag=$1
mkdir -p gpurun_out
MID="python bench.py --size 2048 --steps 1 --warmup 1 --quick"
$MID > gpurun_out/plain_stages_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_stages_$tag.log; exit 1; }
cap() {   # name regex skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/prof_$1_$tag -f $MID > gpurun_out/ncu_$1_$tag.log 2>&1
  echo "ncu $1 exit $?"
}
cap fir fir_kernel 40
cap select select_hist_kernel 20
cap cellstats "cell_stats_kernel" 1
cap patches build_patches_kernel 1
cap layernorm layernorm_split_kernel 10
cap merge merge_votes_kernel 1
