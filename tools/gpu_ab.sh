#!/bin/bash
# A/B of bench flags on one box.  usage: tools/gpu_ab.sh <tag> "<pytest -k expr or empty>" "<flags A>" "<flags B>" ...
tag=$1; k=$2; shift 2
mkdir -p gpurun_out
if [ -n "$k" ]; then timeout 600 python -m pytest tests/test_gpu_networks.py tests/test_gpu_exact.py -q -m gpu -k "$k" -p no:cacheprovider -x 2>&1 | tail -4; fi
i=0
for flags in "$@"; do
  python bench.py --quick --steps 3 --warmup 3 $flags > gpurun_out/bench_${tag}_$i.json 2> gpurun_out/bench_${tag}_$i.err || tail -5 gpurun_out/bench_${tag}_$i.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${tag}_$i.json").read().strip().split("\n")[-1]); r = d["roofline"]; st = d["stages"]
    print(f"[$flags] cells/s {d['value']:.0f} e2e {d['e2e']['value']:.0f} ms/step {d['ms_per_step']:.1f} serial {r.get('serial_step_ms', 0):.1f} | "
          + " ".join(f"{k[2:]} {v['ms_per_step']:.1f}" for k, v in st.items()) + f" | clocks {d['clocks'].get('sm_mhz')} {d['clocks'].get('reasons')}")
except Exception as e:
    print("bench failed", e)
PY
  i=$((i+1))
done
