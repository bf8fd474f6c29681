#!/bin/bash
# 2-GPU validation: bit-identity of the sharded paths + the bench's strong / batch records under torchrun.  usage: tools/gpu_n2.sh <tag> [bench args]
tag=${1:-r02n2}; shift
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > gpurun_out/multi_gpu_check_$tag.log 2>&1; echo "multi_gpu_check exit $?"; tail -3 gpurun_out/multi_gpu_check_$tag.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n${N}_$tag.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n${N}_$tag.json").read().strip().split("\n")[-1])
    print(json.dumps({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "reevaluation", "strong", "batch")}, indent=1)[:5000])
except Exception as e:
    print("no bench line:", e)
PY
