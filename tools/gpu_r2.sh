#!/bin/bash
# Round-2 GPU round trip: new exact-label tests verbosely, the whole GPU suite, smoke, then the bench.  usage: tools/gpu_r2.sh <tag> [bench args]
tag=${1:-r02a}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_$tag.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_exact.py -q -m gpu -s -p no:cacheprovider > gpurun_out/exact_$tag.log 2>&1; echo "exact tests exit $?"; grep -E "refine|fp32 path|two models|passed|failed|Error|error" gpurun_out/exact_$tag.log | tail -20
timeout 1200 python -m pytest tests/ -q -m gpu -p no:cacheprovider -x > gpurun_out/all_gpu_$tag.log 2>&1; echo "pytest -m gpu exit $?"; tail -5 gpurun_out/all_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_$tag.log
timeout 1500 python bench.py "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; tail -3 gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().split("\n")[-1])
    keep = {k: d.get(k) for k in ("value", "ms_per_step", "e2e", "e2e_annotator", "reevaluation", "parity_sample", "parity_population", "strong", "batch", "clocks")}
    keep["roofline"] = {k: d["roofline"][k] for k in ("achieved", "frac", "issued_frac", "kernel_ms_per_step", "share_of_step")}
    keep["stages"] = {k: v.get("ms_per_step") for k, v in d["stages"].items()}
    keep["cpu_baseline"] = None if not d.get("cpu_baseline") else {k: d["cpu_baseline"].get(k) for k in ("value", "cores", "sample")}
    print(json.dumps(keep, indent=1)[:6000])
except Exception as e:
    print("no bench line:", e)
PY
