#!/usr/bin/env python
"""Benchmark of RIBCA's per-cell annotation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size S] [--workload c2|c3]

A "step" is one pass of the whole hot path (normalise -> cell statistics -> patches -> vit_l -> merge / threshold /
count, with the margin-guarded re-evaluation that makes the labels exact) over one synthetic 15-marker S x S image
(default S = 4096, BASELINE.json configs[1], ~51.8 k cells, immune_full panel).  Metric: cells/sec.
  value          inputs already resident in HBM when the timed region starts
  e2e            the same through HotPath.run with HOST (pinned) image + mask: H2D copies and the D2H read of
                 labels / confidences / counts are inside the timed region
  e2e_annotator  (N = 1) the reference-shaped API: Annotator(...).preprocess() + predict() from the .npy files on disk
With N > 1 (torchrun, one rank per GPU) the headline `value` is WEAK scaling: every rank annotates its own image and the
per-type counts are all-reduced over NCCL.  Two more records ride on the same line:
  strong   BASELINE configs[3] shape: ONE image (default 8192^2, ~207 k cells) in pinned host memory, cells sharded by
           contiguous range over the ranks, stage 1 split by channel, one gather of labels / confidences; per-phase
           device times and the SHA-256 of all labels + confidences (identical at every N)
  batch    BASELINE configs[4]: 64 images of 2048^2 in four marker-file groups (full / structure / nerve /
           structure + nerve), images round-robin over the ranks through HotPath.run_batch
--impl reference times the CPU oracle (the reference's algorithm on the host cores, torch threads = all cores) on a
bounded crop of the same workload; the in-line `cpu_baseline` of the b200 arm uses the same crop rule.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cells/sec end-to-end (RIBCA per-cell annotation hot path)"
FLOP_PER_CELL = {"immune_full": 9.96e9, "immune_extended": 4.48e9, "immune_base": 2.55e9, "structure": 2.55e9,
                 "nerve_cell": 0.67e9}            # SURVEY 2b, algorithmic (no padding, no split passes)
C3_INDEX, C3_PRESENT = [0, 1, 2, 3, 4, -1, 5], [0, 1, 2, 3, 4, 6]     # CD45,CD20,CD4,CD8,DAPI,CD3 -> immune_base, CD11c missing


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--precision", default=os.environ.get("RIBCA_PRECISION", "f16f8"))
    ap.add_argument("--chunk", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--interleave", action="store_true", help="two halves of every chunk on two streams (ribca_set_interleave; measured neutral, profiles/r02_interleave.md)")
    ap.add_argument("--quick", action="store_true", help="iteration mode: no CPU baseline, annotator, strong or batch record")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3"],
                    help="c2: 15-marker full panel -> vit_l (headline, BASELINE configs[1]); c3: 6-marker basic panel with "
                         "CD11c missing -> MAE imputer + vit_s (configs[2], secondary)")
    ap.add_argument("--strong-size", type=int, default=8192, help="edge of the ONE image of the strong-scaling record (0 = skip)")
    ap.add_argument("--strong-steps", type=int, default=2)
    ap.add_argument("--batch-images", type=int, default=64, help="images of the batch-CSV record (0 = skip)")
    ap.add_argument("--batch-size", type=int, default=2048)
    ap.add_argument("--no-annotator", action="store_true", help="skip the e2e_annotator leg")
    ap.add_argument("--ln-fold", type=int, default=None, help="A/B: RIBCA_LN_FOLD mask (0 = separate LayerNorm kernels [default], 1 = norm1 folded into "
                                                              "qkv, 3 = norm1 and norm2 folded; profiles/r02_lnfold.md)")
    args = ap.parse_args()
    if args.ln_fold is not None:
        os.environ["RIBCA_LN_FOLD"] = str(args.ln_fold)
    if args.quick:
        args.no_cpu_baseline = args.no_annotator = True
        args.strong_size = args.batch_images = 0
    return args


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().split("\n") if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any("Active" in r[5 + k] and "Not" not in r[5 + k] for r in rows)]
        loaded = [float(r[1]) for r in rows if float(r[3]) > 300] or sm
        out.update(sm_mhz=float(np.median(loaded)), sm_max_mhz=float(rows[0][2]), reasons=reasons, samples=len(rows),
                   power_w_max=max(float(r[3]) for r in rows))
        return out


# ------------------------------------------------------------------------------------------------
def make_scene(size, seed, device, channels=15, grid=18):
    from multiplexed_image_annotator_b200 import synth
    mask = synth.synth_mask(size, size, grid=grid, seed=seed, device=device)
    img = synth.synth_image(mask, channels, seed=seed)
    return img, mask


def sample_edge(steps, warmup):
    """Edge of the CPU sample crop: the SAME rule for `--impl reference` and the in-line cpu_baseline, sized so that
    (steps + warmup) oracle passes end within a few minutes (~26 cells/s/8 cores probed)."""
    total = steps + warmup
    return 512 if total <= 12 else (384 if total <= 24 else 256)


def cpu_oracle_rate(img_u16, mask_i32, sd, edge, threads, panel="immune_full", index=None, mae_sd=None, present=None):
    """cells/sec of the reference algorithm (oracle) on a crop of the workload, host cores only; the stages are the calls of
    orc.annotate_image, timed one by one (per-stage CPU rates next to the total)."""
    from oracle import ribca_oracle as orc
    torch.set_num_threads(threads)
    crop_i = np.ascontiguousarray(img_u16[:, :edge, :edge])
    crop_m = np.ascontiguousarray(mask_i32[:edge, :edge])
    model = orc.make_vit(panel)
    model.load_state_dict(sd)
    index = list(range(crop_i.shape[0])) if index is None else index
    mae = None
    if mae_sd is not None:
        mae = orc.make_mae(panel)
        mae.load_state_dict(mae_sd)
    t = [time.perf_counter()]
    img = orc.normalize(crop_i, 0.3, 99.8); t.append(time.perf_counter())
    stats = orc.cell_stats(crop_m); t.append(time.perf_counter())
    pt, inten, wins = orc.build_patches(img, crop_m, index, stats, 30); t.append(time.perf_counter())
    if mae is not None:
        pt = orc.impute(mae, pt, present)
    t_imp = time.perf_counter()
    probs = orc.vit_probs(model, pt, 128); t.append(time.perf_counter())
    labels, conf = orc.merge_by_voting({panel: probs}, 0.3, None); t.append(time.perf_counter())
    dt = t[-1] - t[0]
    n = len(labels)
    mpx_ch = crop_i.shape[0] * edge * edge / 1e6
    stages = {"1_normalize_ms_per_Mpx_channel": 1e3 * (t[1] - t[0]) / mpx_ch, "2_cell_stats_us_per_pixel": 1e6 * (t[2] - t[1]) / (edge * edge),
              "3_build_patches_ms_per_cell": 1e3 * (t[3] - t[2]) / max(n, 1), "4_networks_ms_per_cell": 1e3 * (t[4] - t[3]) / max(n, 1),
              "4a_of_which_imputer_ms_per_cell": 1e3 * (t_imp - t[3]) / max(n, 1),
              "5_merge_us_per_cell": 1e6 * (t[5] - t[4]) / max(n, 1)}
    res = {"labels": labels, "confidence": conf, "probs": {panel: probs}, "stages": stages}
    return n / dt, res, dt


def workload_models(args, dev=None):
    """(panel, channel index, n_markers, vit state, mae state or None) of the selected workload."""
    from multiplexed_image_annotator_b200 import weights
    if args.workload == "c2":
        return "immune_full", list(range(15)), 15, weights.random_vit_state("immune_full", seed=7), None
    return "immune_base", C3_INDEX, 6, weights.random_vit_state("immune_base", seed=7), weights.random_mae_state("immune_base", seed=7)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from multiplexed_image_annotator_b200 import synth
    threads = os.cpu_count() or 1
    total = args.steps + args.warmup
    edge = sample_edge(args.steps, args.warmup)
    panel, index, n_markers, sd, mae_sd = workload_models(args)
    img, mask = make_scene(max(edge, 512), 2, "cpu", n_markers)
    img_u16, mask_i32 = synth.to_uint16(img), mask.numpy()
    rates, cells = [], 0
    for s in range(total):
        r, res, dt = cpu_oracle_rate(img_u16, mask_i32, sd, edge, threads, panel, index, mae_sd, C3_PRESENT if mae_sd else None)
        cells = len(res["labels"])
        if s >= args.warmup:
            rates.append((cells, dt))
    n = sum(c for c, _ in rates)
    t = sum(d for _, d in rates)
    val = n / t
    what = "immune_full panel / vit_l" if args.workload == "c2" else "immune_base panel, CD11c imputed by the MAE, vit_s"
    sample = f"{edge}x{edge} crop of the {n_markers}-marker scene, {cells} cells per step, oracle (numpy/scipy/torch fp32) preprocess+predict"
    line = json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * t / max(len(rates), 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload.upper()} sample: synthetic {n_markers}-marker image, {what}, {sample}"},
        "cpu_baseline": {"value": val, "unit": "cells/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(line, file=_JSON_OUT(), flush=True)


def decision_report(ref_probs, got_probs, lab_ref, lab_got, margin_ref):
    """Label parity with its sensitivity (SURVEY 8d 'always report'): final-label agreement, pre-threshold argmax agreement,
    how many cells sit within 1e-3 / 1e-4 of a decision boundary of the checker, and the checker's top-2 gap histogram."""
    top = torch.topk(ref_probs, 2, dim=1).values
    gap = (top[:, 0] - top[:, 1]).double().cpu().numpy()
    edges = [0, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1.0000001]
    hist = np.histogram(gap, bins=edges)[0].tolist()
    n = int(lab_ref.numel())
    differ = int((lab_ref != lab_got).sum().item())
    am = int((ref_probs.argmax(1) != got_probs.argmax(1)).sum().item())
    m = margin_ref.double().cpu().numpy()
    return {"cells": n, "labels_differ": differ, "label_agreement": 1.0 - differ / max(n, 1),
            "argmax_differ_pre_threshold": am, "argmax_agreement_pre_threshold": 1.0 - am / max(n, 1),
            "max_abs_dprob": float((ref_probs - got_probs).abs().max().item()),
            "cells_within_1e-3_of_a_decision_boundary": int((m < 1e-3).sum()), "cells_within_1e-4_of_a_decision_boundary": int((m < 1e-4).sum()),
            "cells_labelled_others_by_threshold": int((lab_ref == 17).sum().item()),
            "top2_gap_histogram": {"bin_edges": edges[:-1] + [1.0], "cells": hist}}


def _JSON_OUT():
    """The process's real stdout (see the bottom of the file): the one JSON line goes there."""
    return globals().get("_json_out", sys.__stdout__)


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)
    import torch.distributed as dist
    from multiplexed_image_annotator_b200 import _lib, engine, exact, ops, synth, weights
    from multiplexed_image_annotator_b200.cell_type_annotation.model import ALL_TYPES, merge_on_device
    from multiplexed_image_annotator_b200.pipeline import HotPath

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner to fd 1 when the communicator is created,
        # so fd 1 points at stderr until that has happened
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    L = _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- workload: one synthetic scene per rank, host copy in pinned memory ----------------------------
    S = args.size
    panel, index, n_markers, sd, mae_sd = workload_models(args)
    img_i32, mask_d = make_scene(S, 2 + rank, dev, n_markers)
    img_host = torch.from_numpy(synth.to_uint16(img_i32)).pin_memory()
    mask_host = mask_d.cpu().pin_memory()
    del img_i32
    img_dev = img_host.to(dev)
    imputers = {}
    if mae_sd is not None:       # markers CD45,CD20,CD4,CD8,DAPI,CD3 -> immune_base index [0,1,2,3,4,-1,5], strict=False, infer=True
        mae = engine.MaeEngine(panel, mae_sd, dev, precision=args.precision)
        imputers = {panel: (mae, C3_PRESENT)}
    eng = engine.VitEngine(panel, sd, dev, precision=args.precision, max_cells_per_call=args.chunk)
    hp = HotPath({panel: index}, {panel: eng}, imputers, chunk_cells=args.chunk, device=dev, shard_cells=False)
    ops.set_interleave(args.interleave)
    # head calibration: spread the label histogram of the random-init classifier.  Rank 0's statistics serve
    # every rank, so that all ranks hold the SAME model (the strong-scaling record shards one image over them).
    warm = hp.run(img_dev, mask_d, to_host=False, keep_probs=True)
    n_cells = warm.n_cells
    norm = ops.normalize(img_dev, 0.3, 99.8)
    (p256,), _, _ = ops.build_patches(norm, mask_d, ops.channel_min(norm), warm.cells, [index], 0, min(256, n_cells))
    if imputers:
        imputers[panel][0].impute(p256, imputers[panel][1])
    _, logits = eng.forward(p256, return_logits=True)
    mean_logits = logits.mean(0)
    if world > 1:
        dist.broadcast(mean_logits, src=0)
    sd_cal = weights.calibrate_head(sd, mean_logits.cpu().numpy(), 20.0)
    eng.set_head(sd_cal["head.weight"], sd_cal["head.bias"])
    del norm, p256, warm

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launch_count()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), out, ops.launch_count() - l0

    def step_resident():
        r = hp.run(img_dev, mask_d, to_host=False)
        if world > 1:
            dist.all_reduce(r.counts)                          # the one collective: per-type counts
        return r

    def step_e2e():
        r = hp.run(img_host, mask_host, to_host=True)
        if world > 1:
            c = r.counts.to(dev)
            dist.all_reduce(c)
            r.counts = c.cpu()
        return r

    sampler = ClockSampler(local) if rank == 0 else None
    ms_res, res, launches = timed(step_resident, args.steps, args.warmup)
    ms_e2e, res_e2e, _ = timed(step_e2e, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else {}

    cells_all = torch.tensor([n_cells], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(cells_all)
    total_cells = float(cells_all.item())
    value = total_cells * args.steps / (ms_res / 1000)
    e2e_value = total_cells * args.steps / (ms_e2e / 1000)
    refine = res.refine.as_dict()
    refine["share_of_step"] = sum(refine["level_ms"]) / (ms_res / args.steps)

    # ---- roofline of the dominant kernel, measured live with CUDA events around every launch -----------
    import ctypes as C
    # (the profiled step runs the SERIAL schedule - the library switches the two-stream interleave off while profiling - so
    #  every span holds one kernel alone; shares are quoted against that step's own time, `serial_step_ms`)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.ribca_profile_begin()
    pe0.record()
    hp.run(img_dev, mask_d, to_host=False)
    pe1.record()
    nc = 8
    ms = (C.c_double * nc)(); ln = (C.c_longlong * nc)(); wk = (C.c_double * nc)()
    _lib.check(L.ribca_profile_end(ms, ln, wk, nc), "ribca_profile_end")
    torch.cuda.synchronize()
    serial_step_ms = pe0.elapsed_time(pe1)
    pk, pk_src = peaks()
    # tensor time in units of one bf16 pass over K: bf16x3 = 3; f16f8 = fp16 pass + e4m3 pass over 2K at twice the rate = 2
    passes = {"bf16x3": 3, "f16f8": 2}.get(args.precision, 1)
    gemm_tf = wk[0] / (ms[0] / 1000) / 1e12 if ms[0] > 0 else 0.0
    peak_tf = pk["bf16_tflops_sustained"]
    # DRAM / L2 traffic of the dominant kernel per launch: from the committed ncu --set full capture (profiles/), if any
    traffic, l2_note = None, None
    import glob
    tfiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "gemm_traffic_*.json")))
    if tfiles and args.workload == "c2" and S >= 4096:
        tj = json.load(open(tfiles[-1]))
        traffic = tj["mean_dram_bytes_per_launch"]
        l2_note = {"mean_l2_to_sm_bytes_per_launch": tj["mean_l2_to_sm_bytes_per_launch"], "source": os.path.basename(tfiles[-1]),
                   "note": "f16f8 main loop is bound by L2 -> SM delivery (~6300 B/clk chip-wide): 64 algorithmic FLOP per L2 byte at 256 x 256 pair tiles"}
    roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": gemm_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": gemm_tf / peak_tf, "traffic": traffic, "l2": l2_note, "peak_source": pk_src + ", sustained bf16",
                "launches_per_step": int(ln[0]), "avg_launch_ms": ms[0] / max(ln[0], 1), "kernel_ms_per_step": ms[0],
                "algorithmic_flop_per_step": wk[0], "tensor_passes": passes, "issued_frac": passes * gemm_tf / peak_tf,
                "share_of_step": ms[0] / serial_step_ms, "serial_step_ms": serial_step_ms,
                "schedule_note": "spans are measured on the serial schedule (one kernel at a time, like the ncu launch list); serial_step_ms is "
                                 "that profiled step's own time (it carries the event overhead of ~700 spans)",
                "note": "GEMM launches of the profiled step include the bf16x3 re-evaluation of the boundary cells (3 passes); "
                        "issued_frac counts them as the default precision's passes",
                "other_kernels": {
                    "attention_kernel": {"ms_per_step": ms[1], "launches": int(ln[1]), "tflops_fp32": wk[1] / (ms[1] / 1000) / 1e12 if ms[1] > 0 else 0},
                    "build_patches_kernel": {"ms_per_step": ms[2], "launches": int(ln[2]), "achieved_gbs": wk[2] / (ms[2] / 1000) / 1e9 if ms[2] > 0 else 0,
                                             "frac_of_hbm": (wk[2] / (ms[2] / 1000) / 1e9 / pk["hbm_gbs"]) if ms[2] > 0 else 0}}}

    hbm = pk["hbm_gbs"]

    def hbm_stage(i):
        gbs = wk[i] / (ms[i] / 1000) / 1e9 if ms[i] > 0 else 0.0
        return {"ms_per_step": ms[i], "launches": int(ln[i]), "algorithmic_bytes": wk[i], "achieved_gbs": gbs, "frac_of_hbm": gbs / hbm}

    stages = {   # every stage against its roofline (SURVEY 8d work definitions), same profiled step
        "1_normalize": dict(hbm_stage(3), note="FP64-pipe bound: 161-tap separable FIR in scipy's exact order (~650 flop per 8 algorithmic bytes)"),
        "2_cell_stats": hbm_stage(4),
        "3_build_patches": hbm_stage(2),
        "4_gemm_tcgen05": {"ms_per_step": ms[0], "launches": int(ln[0]), "algorithmic_tflops": gemm_tf, "frac_of_tensor_sustained": gemm_tf / peak_tf,
                           "issued_frac": passes * gemm_tf / peak_tf},
        "4_attention_tc": {"ms_per_step": ms[1], "launches": int(ln[1]), "algorithmic_tflops": wk[1] / (ms[1] / 1000) / 1e12 if ms[1] > 0 else 0.0},
        "4_layernorm": hbm_stage(5),
        "4_sgemm_fp32_reevaluation": {"ms_per_step": ms[7], "launches": int(ln[7]), "tflops_fp32": wk[7] / (ms[7] / 1000) / 1e12 if ms[7] > 0 else 0.0},
        "5_merge": hbm_stage(6),
    }

    # ---- CPU baseline + label agreement on a bounded sample (rank 0, N = 1 only) -------------------------
    cpu_baseline, agreement = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import ribca_oracle as orc                         # checker / baseline only
        threads = os.cpu_count() or 1
        edge = min(sample_edge(args.steps, args.warmup), S)          # the crop rule of --impl reference
        img_u16 = img_host.numpy()
        rate, ores, dt = cpu_oracle_rate(img_u16, mask_host.numpy(), sd_cal, edge, threads, panel, index, mae_sd, C3_PRESENT if mae_sd else None)
        crop = hp.run(np.ascontiguousarray(img_u16[:, :edge, :edge]), np.ascontiguousarray(mask_host.numpy()[:edge, :edge]),
                      to_host=True, keep_probs=True)
        oprobs = torch.from_numpy(ores["probs"][panel]).to(dev)
        olab, _, _, omargin = merge_on_device({panel: oprobs}, 0.3, None, want_margin=True)
        agreement = decision_report(oprobs, crop.probs[panel], olab, crop.label.to(dev), omargin)
        agreement["labels_present"] = sorted(set(ores["labels"]))
        agreement["names_equal"] = crop.names() == list(ores["labels"])
        agreement["reevaluation"] = crop.refine.as_dict()
        st = ores["stages"]
        full_s = (st["1_normalize_ms_per_Mpx_channel"] * n_markers * S * S / 1e6 + st["2_cell_stats_us_per_pixel"] * S * S / 1e3
                  + (st["3_build_patches_ms_per_cell"] + st["4_networks_ms_per_cell"]) * n_cells + st["5_merge_us_per_cell"] * n_cells / 1e3) / 1e3
        cpu_baseline = {"value": rate, "unit": "cells/s", "cores": threads, "kind": "port",
                        "sample": f"{edge}x{edge} crop of the same scene, {len(ores['labels'])} cells, {dt:.1f} s, "
                                  "oracle (numpy/scipy/torch fp32) preprocess+predict",
                        "stage_rates": st, "extrapolated_full_workload_s": full_s}

    # ---- still the baseline leg (the only place the oracle runs): the reference's own GPU path for stage 4 (SURVEY 8d "second
    #      baseline") = the oracle's torch modules in eager fp32 on this GPU, bs = 128 slices as cta/model.py:397-406, which also
    #      serves as the checker of the step's labels over the whole population
    ref_gpu = None
    if cpu_baseline is not None:
        # strict fp32 like the reference's CPU path: cuDNN would otherwise run the patch-embedding conv in TF32, which alone
        # costs the stock GPU path ~1.5e-3 of probability (profiles/precision_population_r01.json)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        ref_model = orc.make_vit(panel)
        ref_model.load_state_dict(sd_cal)
        ref_model = ref_model.to(dev).eval()
        full = hp.run(img_dev, mask_d, to_host=False, keep_probs=True)
        norm = ops.normalize(img_dev, 0.3, 99.8)
        mn = ops.channel_min(norm)
        ref_probs, t_ref = [], 0.0
        with torch.no_grad():
            for a in range(0, n_cells, 8192):
                (pt,), _, _ = ops.build_patches(norm, mask_d, mn, full.cells, [index], a, min(8192, n_cells - a))
                if imputers:       # the checker classifies the SAME imputed inputs (the imputer has its own parity tests)
                    imputers[panel][0].impute(pt, imputers[panel][1], precision="bf16x3")
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                outs = [torch.softmax(ref_model(pt[b:b + 128]), dim=1) for b in range(0, len(pt), 128)]
                e1.record()
                torch.cuda.synchronize()
                t_ref += e0.elapsed_time(e1)
                ref_probs.append(torch.cat(outs))
        ref_probs = torch.cat(ref_probs)
        lab_ref, _, _, margin_ref = merge_on_device({panel: ref_probs}, 0.3, None, want_margin=True)
        pop = decision_report(ref_probs, full.probs[panel], lab_ref, full.label, margin_ref)
        pop["checker"] = "oracle torch modules, eager strict fp32 on this GPU, all cells of the step"
        pop["reevaluation"] = full.refine.as_dict()
        ref_gpu = {"what": "oracle classifier (plain torch.nn, eager fp32, TF32 off for matmul and cuDNN) forward + softmax on this GPU in bs = 128 slices, "
                           "patches resident in HBM (no per-batch H2D / D2H, which the reference also pays)",
                   "cells": n_cells, "stage4_ms": t_ref, "stage4_cells_per_s": n_cells / (t_ref / 1000),
                   "b200_stage4_ms": ms[0] + ms[1] + ms[5], "full_population_parity": pop}
        del ref_model, ref_probs, norm, full
        cpu_baseline["reference_gpu_path"] = {k: v for k, v in ref_gpu.items() if k != "full_population_parity"}

    # ---- the reference-shaped API end to end: Annotator(...).preprocess() + predict() from files on disk (N = 1) ---------
    e2e_annotator = None
    if rank == 0 and world == 1 and not args.no_annotator:
        e2e_annotator = annotator_leg(args, img_host, mask_host, panel, sd_cal, mae_sd, n_cells, res_e2e)

    del img_dev
    torch.cuda.empty_cache()
    strong = strong_record(args, dev, world, rank, eng, hp, barrier, max_over_ranks) if args.strong_size and args.workload == "c2" else None
    batch = batch_record(args, dev, world, rank, eng, barrier, max_over_ranks) if args.batch_images and args.workload == "c2" else None

    if rank == 0:
        hist = np.bincount(res_e2e.label.numpy(), minlength=18)
        what = (f"C2: synthetic 15-marker {S}x{S} uint16 image + int32 mask per GPU, {n_cells} cells, "
                f"immune_full panel -> vit_l (random-init, calibrated head), blur 0.3, amax 99.8, confidence 0.3"
                if args.workload == "c2" else
                f"C3: synthetic 6-marker {S}x{S} image (basic panel, CD11c missing), {n_cells} cells, strict=False infer=True "
                f"-> MAE imputer (L=7, keep 6) + vit_s, blur 0.3, amax 99.8, confidence 0.3")
        out = {
            "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": f"{args.precision} tensor-core passes, fp32 accumulate (stages 1-3 f32/f64 exact order); boundary cells re-evaluated in bf16x3 / fp32",
            "data": "synthetic",
            "config": {"workload": what + f"; CPU arms run a {sample_edge(args.steps, args.warmup)}^2 crop of the same scene",
                       "cells_per_gpu": n_cells, "chunk_cells": args.chunk, "precision": args.precision,
                       "schedule": "two halves of every chunk interleaved on two streams (ribca_set_interleave)" if args.interleave else "serial: one stream, one kernel at a time",
                       "exact_labels": {"levels": hp.exact_labels, "eps": [exact.EPS1, exact.EPS2]},
                       "l2": f"inputs ({img_host.numel() * 2 / 1e6:.0f} MB image + {mask_host.numel() * 4 / 1e6:.0f} MB mask) exceed the 126 MB L2",
                       "parallelism": f"{world} x (one image per GPU), all-reduce of 18 counts"},
            "e2e": {"value": e2e_value, "unit": "cells/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(img_host.numel() * 2 + mask_host.numel() * 4),
                    "d2h_bytes_per_step": int(n_cells * 5 + 18 * 8)},
            "e2e_annotator": e2e_annotator,
            "gpu_launches": int(launches),
            "build": {k: _lib.build_info().get(k) for k in ("built_at_utc", "host", "nvcc", "sources_match", "gpu_visible_at_build")},
            "clocks": clocks, "roofline": roofline, "stages": stages, "reevaluation": refine,
            "cpu_baseline": cpu_baseline, "parity_sample": agreement,
            "parity_population": None if ref_gpu is None else ref_gpu["full_population_parity"],
            "label_histogram": {ALL_TYPES[k]: int(v) for k, v in enumerate(hist) if v},
            "strong": strong, "batch": batch,
        }
        print(json.dumps(out), file=_JSON_OUT(), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def annotator_leg(args, img_host, mask_host, panel, sd_cal, mae_sd, n_cells, res_e2e):
    """Annotator(...).preprocess() + predict() (what main.py:19-20 / gui_api.py:23-24 call) on the bench scene stored as
    .npy files: wall clock around the two calls, CUDA-synchronised, file read and Python result assembly included."""
    from multiplexed_image_annotator_b200 import synth
    from multiplexed_image_annotator_b200.cell_type_annotation import markerImputer as bimp
    from multiplexed_image_annotator_b200.cell_type_annotation import model as bmodel
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        os.chdir(tmp)
        try:
            np.save("img.npy", img_host.numpy())
            np.save("mask.npy", mask_host.numpy())
            markers = synth.FULL_PANEL_MARKERS if args.workload == "c2" else ["CD45", "CD20", "CD4", "CD8", "DAPI", "CD3"]
            synth.write_marker_file("markers.txt", markers)
            with open("images.csv", "w") as f:
                f.write("image_path,mask_path\nimg.npy,mask.npy\n")
            bmodel.register_state(panel, sd_cal)
            if mae_sd is not None:
                bimp.register_state(panel, mae_sd)
            times, ann = [], None
            for _ in range(2):                                    # first pass warms the allocator and the page cache
                del ann
                ann = bmodel.Annotator("markers.txt", "images.csv", "cuda", "./", "bench", args.workload == "c2", True, -1, True,
                                       0.3, 99.8, 0.3, 30, None, n_jobs=0)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ann.preprocess()
                t1 = time.perf_counter()
                ann.predict(128)
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                times.append((t1 - t0, t2 - t1))
            same = bool(np.array_equal(ann.labels_index[0], res_e2e.label.numpy()))
            pre_s, pred_s = times[-1]
            return {"value": n_cells / (pre_s + pred_s), "unit": "cells/s", "preprocess_s": pre_s, "predict_s": pred_s,
                    "what": "Annotator(markers.txt, images.csv, 'cuda', ...).preprocess() + .predict(128): .npy image + mask read from "
                            "disk (tmpfs), stages 1-5, labels / confidences / probabilities / intensities back on the host as the "
                            "reference's attributes; wall clock, second of two runs",
                    "labels_equal_hot_path": same, "file_bytes": int(img_host.numel() * 2 + mask_host.numel() * 4)}
        finally:
            os.chdir(cwd)


def strong_record(args, dev, world, rank, eng, hp_weak, barrier, max_over_ranks):
    """BASELINE configs[3] shape on N GPUs: ONE image in pinned host memory, sharded by cell range."""
    import torch.distributed as dist
    from multiplexed_image_annotator_b200 import synth
    from multiplexed_image_annotator_b200.pipeline import HotPath
    S = args.strong_size
    img_i32, mask_d = make_scene(S, 4, dev, 15, grid=18)
    img_host = torch.from_numpy(synth.to_uint16(img_i32)).pin_memory()          # the same bytes on every rank = one file
    mask_host = mask_d.cpu().pin_memory()
    del img_i32, mask_d
    torch.cuda.empty_cache()
    hp = HotPath(hp_weak.panels, hp_weak.models, chunk_cells=args.chunk, device=dev, shard_cells=True)
    hp.run(img_host, mask_host, to_host=True)                                   # warm-up (allocator, NCCL channels)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.strong_steps):
        res = hp.run(img_host, mask_host, to_host=True, time_phases=True)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.strong_steps
    phases = res.phases.ms()
    names = sorted(phases)
    t = torch.tensor([phases[k] for k in names], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    st = res.refine
    rv = torch.tensor(st.reevaluated + st.relabelled, device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(rv)
    digest = hashlib.sha256(res.label.numpy().tobytes() + res.confidence.numpy().tobytes()).hexdigest()
    n = res.n_cells
    c = img_host.shape[0]
    own = len(range(rank, c, world)) if world > 1 else c
    return {"what": f"ONE synthetic 15-marker {S}x{S} uint16 image + int32 mask in pinned host memory, {n} cells, cells sharded by "
                    f"contiguous range over {world} rank(s); stage 1 split by channel with NCCL broadcasts of the normalised planes; "
                    "one all-gather of labels / confidences + all-reduce of counts; results on the host of every rank",
            "scaling": "strong", "n_gpus": world, "cells": n, "steps": args.strong_steps, "ms_per_step": ms,
            "value": n / (ms / 1000), "unit": "cells/s",
            "phases_ms_max_over_ranks": {k: float(v) for k, v in zip(names, t.tolist())},
            "h2d_bytes_per_rank": int(own * img_host[0].numel() * 2 + mask_host.numel() * 4),
            "nccl_bytes": {"plane_broadcasts": int(c * img_host[0].numel() * 4) if world > 1 else 0,
                           "gather": int(n * 5 + 18 * 8) if world > 1 else 0},
            "reevaluated_cells_level1_level2": rv.tolist()[:2], "relabelled_level1_level2": rv.tolist()[2:],
            "labels_confidences_sha256": digest, "counts": res.counts.tolist(),
            "max_mem_gb": torch.cuda.max_memory_allocated() / 1e9}


def batch_record(args, dev, world, rank, eng_full, barrier, max_over_ranks):
    """BASELINE configs[4]: a batch-processing CSV of images in four marker-file groups (one marker file per CSV is the
    reference's constraint), images round-robin over the ranks (HotPath.run_batch)."""
    from multiplexed_image_annotator_b200 import engine, synth, weights
    from multiplexed_image_annotator_b200.parallel import image_owner
    from multiplexed_image_annotator_b200.pipeline import HotPath
    S, per_group = args.batch_size, max(args.batch_images // 4, 1)
    struct = engine.VitEngine("structure", weights.random_vit_state("structure", seed=8), dev, max_cells_per_call=args.chunk)
    nerve = engine.VitEngine("nerve_cell", weights.random_vit_state("nerve_cell", seed=9), dev, max_cells_per_call=args.chunk)
    groups = [   # (name, markers, panels -> channel index, models)
        ("full15_branch5", 15, {"immune_full": list(range(15))}, {"immune_full": eng_full}),
        ("structure7_branch6", 7, {"structure": list(range(7))}, {"structure": struct}),
        ("nerve3_branch7", 3, {"nerve_cell": [0, 1, 2]}, {"nerve_cell": nerve}),
        ("structure_plus_gfap_branch3", 8, {"structure": list(range(7)), "nerve_cell": [0, 6, 7]}, {"structure": struct, "nerve_cell": nerve}),
    ]
    work = []
    h2d = 0
    for g, (name, ch, panels, models) in enumerate(groups):
        items = []
        for i in range(per_group):
            if image_owner(i, world) != rank:
                items.append(None)
                continue
            img, m = make_scene(S, 100 + 16 * g + i, dev, ch)
            items.append((torch.from_numpy(synth.to_uint16(img)).pin_memory(), m.cpu().pin_memory()))
            h2d += items[-1][0].numel() * 2 + items[-1][1].numel() * 4
        work.append((name, HotPath(panels, models, chunk_cells=args.chunk, device=dev), items))
    torch.cuda.empty_cache()

    def one_pass():
        return [(name, hp.run_batch(items, to_host=True)) for name, hp, items in work]

    one_pass()                                                  # warm-up
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = one_pass()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    cells = {name: int(sum(r.n_cells for r in rs)) for name, rs in out}
    h = hashlib.sha256()
    for _, rs in out:
        for r in rs:
            h.update(r.label.numpy().tobytes()); h.update(r.confidence.numpy().tobytes())
    total = sum(cells.values())
    return {"what": f"{per_group * 4} synthetic {S}x{S} images in four batch CSVs of {per_group} (full-15 -> vit_l, structure-7, "
                    "DAPI/CD45/GFAP -> nerve, structure + GFAP -> structure & nerve vote), image i of a CSV on rank i mod N, "
                    "pinned host images in, every image's labels / confidences / counts on the host of every rank",
            "n_gpus": world, "images": per_group * 4, "cells": total, "cells_per_group": cells, "ms": ms,
            "value": total / (ms / 1000), "unit": "cells/s", "h2d_bytes_this_rank": int(h2d),
            "labels_confidences_sha256": h.hexdigest()}


if __name__ == "__main__":
    # stdout carries exactly ONE line, the JSON record: everything the legs print on the way (the reference-shaped Annotator announces
    # its panels with print(), like the reference) goes to stderr
    _json_out = sys.stdout
    sys.stdout = sys.stderr
    main()
