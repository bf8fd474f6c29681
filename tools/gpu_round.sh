#!/bin/bash
# One GPU-box round trip: every stage's parity tests in separate processes (a hang in one cannot take
# the others down), then the smoke test.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/$name.log; }
: > gpurun_out/summary.txt
run stages python -m pytest tests/test_gpu_stages.py -q -m gpu --timeout 300 -p no:cacheprovider
run gemm python -m pytest tests/test_gpu_networks.py -q -m gpu --timeout 300 -p no:cacheprovider -s -k "gemm or layernorm or attention"
run networks python -m pytest tests/test_gpu_networks.py -q -m gpu --timeout 300 -p no:cacheprovider -s -k "vit or mae"
run e2e python -m pytest tests/test_gpu_networks.py -q -m gpu --timeout 300 -p no:cacheprovider -s -k "annotator"
run fullsize python -m pytest tests/test_gpu_fullsize.py -q -m gpu --timeout 400 -p no:cacheprovider
run smoke python __graft_entry__.py smoke
cat gpurun_out/summary.txt
