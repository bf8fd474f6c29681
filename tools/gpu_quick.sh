#!/bin/bash
# quick iteration: selected tests + short bench without the CPU baseline.  usage: tools/gpu_quick.sh "<pytest -k expr>" [bench args]
mkdir -p gpurun_out
k=$1; shift
timeout 400 python -m pytest tests/test_gpu_networks.py -q -m gpu -k "$k" -p no:cacheprovider -x 2>&1 | tail -4
python bench.py --steps 2 --warmup 2 --quick "$@" > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_q.json")); r=d["roofline"]; o=r["other_kernels"]
    print(f"cells/s {d['value']:.0f}  e2e {d['e2e']['value']:.0f}  ms/step {d['ms_per_step']:.1f} | gemm {r['kernel_ms_per_step']:.1f} ms issued {r['issued_frac']:.3f} alg {r['frac']:.3f} | attn {o['attention_kernel']['ms_per_step']:.1f} ms | patches {o['build_patches_kernel']['ms_per_step']:.1f} ms | clocks {d['clocks'].get('sm_mhz')} {d['clocks'].get('reasons')}")
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_q.err").read()[-2000:])
PY
