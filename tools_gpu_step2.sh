#!/bin/bash
./tools_gpu_round.sh > gpurun_out/round.log 2>&1; cat gpurun_out/summary.txt
grep -E "max\|dprob\||flips|agreement" gpurun_out/networks.log gpurun_out/e2e.log gpurun_out/smoke.log | head -40
B=multiplexed_image_annotator_b200
timeout 300 python tools_gemm_ab.py 4096 $B/libribca_b200.so $B/build/libribca_s6.so > gpurun_out/gemm_ab2.log 2>&1; tail -8 gpurun_out/gemm_ab2.log
for p in f16f8 bf16x3; do
python bench.py --steps 2 --warmup 2 --no-cpu-baseline --precision $p > gpurun_out/bench_q_$p.json 2>gpurun_out/bench_q_$p.err
python - $p <<'PY'
import json, sys
p = sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/bench_q_{p}.json")); r=d["roofline"]; o=r["other_kernels"]
    print(f"{p}: cells/s {d['value']:.0f}  e2e {d['e2e']['value']:.0f}  ms/step {d['ms_per_step']:.1f} | gemm {r['kernel_ms_per_step']:.1f} ms issued {r['issued_frac']:.3f} alg {r['frac']:.3f} | attn {o['attention_kernel']['ms_per_step']:.1f} ms | LN {d['stages']['4_layernorm']['ms_per_step']:.1f} | clocks {d['clocks'].get('sm_mhz')} {d['clocks'].get('reasons')}")
except Exception as e:
    print("bench failed", e); print(open(f"gpurun_out/bench_q_{p}.err").read()[-2000:])
PY
done
