#!/usr/bin/env python
"""Summarise gpurun_out/ ncu artefacts into profiles/ (tracked).  usage: tools/profile_summary.py TAG"""
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
         "`bench.py --size 1024 --steps 1 --warmup 1 --quick`; cold-cache serialised times: compare SHARES)\n")
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
WANT += ["l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.per_cycle_active",
         "sm__cycles_elapsed.avg.per_second", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
import glob, json
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
gemm_rows = []
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
        if "gemm_tcgen05" in name and rep.endswith(f"prof_gemm_{tag}.ncu-rep"):
            def val(metric):
                for i, h in enumerate(hdr):
                    if h.endswith(metric):
                        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
                return None
            gemm_rows.append({"dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
                              "l2_to_sm_bytes": val("l1tex__m_xbar2l1tex_read_bytes.sum"),
                              "duration_ms": val("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[hdr.index([h for h in hdr if h.endswith("gpu__time_duration.sum")][0])], 1.0)})
if gemm_rows:
    # one launch each of the qkv / proj / fc1 / fc2 GEMMs of a vit_l block at the bench's chunk size: their mean is the
    # per-launch DRAM traffic of the step's launch mix (each occurs once per block and chunk)
    n = len(gemm_rows)
    traffic = {"source": f"ncu --set full, gpurun_out/prof_gemm_{tag}.ncu-rep (tools/gemm_one.py 4096 qkv proj fc1 fc2, f16f8)",
               "launches": gemm_rows,
               "mean_dram_bytes_per_launch": sum(g["dram_read_bytes"] + g["dram_write_bytes"] for g in gemm_rows) / n,
               "mean_l2_to_sm_bytes_per_launch": sum(g["l2_to_sm_bytes"] for g in gemm_rows) / n}
    json.dump(traffic, open(f"profiles/gemm_traffic_{tag}.json", "w"), indent=1)
    emit(f"GEMM traffic per launch (mean of {n} captures): DRAM {traffic['mean_dram_bytes_per_launch'] / 1e9:.3f} GB, "
         f"L2 -> SM {traffic['mean_l2_to_sm_bytes_per_launch'] / 1e9:.3f} GB -> profiles/gemm_traffic_{tag}.json")
out.close()
