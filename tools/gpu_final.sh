#!/bin/bash
# the driver's round-end sequence on one box: GPU tests, smoke, reference arm, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/final_gpu_tests.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/final_smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "reference arm exit $?"
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"; tail -2 gpurun_out/final_bench.err
python - <<'PY'
import json
r = json.load(open("gpurun_out/final_bench_ref.json")); print("ref", r["value"], r["unit"], r["cpu_baseline"]["cores"])
d = json.load(open("gpurun_out/final_bench.json"))
print("b200", d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"])
print(d["roofline"]["frac"], d["roofline"]["issued_frac"], d["roofline"]["traffic"])
print(d["parity_sample"]); print(d["parity_population"], d["cpu_baseline"]["reference_gpu_path"]["stage4_ms"])
PY
