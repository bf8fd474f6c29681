#!/bin/bash
# ncu --set full capture of one kernel family.  usage: tools/gpu_ncu.sh <kernel-regex> <tag> [skip] [count]
mkdir -p gpurun_out
SMALL="python bench.py --size 1024 --steps 1 --warmup 1 --quick"
$SMALL > gpurun_out/plain_$2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-20} -c ${4:-2} -o gpurun_out/prof_$2 -f $SMALL > gpurun_out/ncu_$2.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_$2.log
