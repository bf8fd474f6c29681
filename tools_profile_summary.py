#!/usr/bin/env python
"""Summarise gpurun_out/ ncu artefacts into profiles/ (tracked).  usage: tools_profile_summary.py TAG"""
import collections, csv, re, subprocess, sys

tag = sys.argv[1]
out = open(f"profiles/{tag}_summary.md", "w")

def emit(s=""):
    print(s); out.write(s + "\n")

# ---- launch list (ncu --metrics gpu__time_duration.sum): shares per kernel ---------------------------
try:
    lines = [l for l in open(f"gpurun_out/launches_{tag}.csv") if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"<.*", "", re.sub(r"\(.*", "", r["Kernel Name"])).replace("void ", "")
        v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r["Metric Unit"]]
        agg[name][0] += 1; agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    emit(f"# {tag}: ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, "
         "`bench.py --size 1024 --steps 1 --warmup 1 --no-cpu-baseline`; cold-cache serialised times: compare SHARES)\n")
    emit("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        emit(f"| {k[:70]} | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f}% |")
    emit(f"| all | {sum(v[0] for v in agg.values())} | {tot:.2f} | 100% |\n")
except FileNotFoundError:
    emit(f"(no launch list for {tag})")

# ---- ncu --set full captures ---------------------------------------------------------------------------
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "lts__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]
import glob
for rep in sorted(glob.glob(f"gpurun_out/prof_*_{tag}.ncu-rep")):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    emit(f"## {rep.split('/')[-1]} (`ncu --set full --clock-control none`)\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        emit(f"**{name[:100]}**\n")
        for w in WANT:
            for i, h in enumerate(hdr):
                if h.endswith(w):
                    emit(f"- {w} = {r[i]} {units[i]}")
                    break
        emit()
out.close()
