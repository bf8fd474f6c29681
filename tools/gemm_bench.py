#!/usr/bin/env python
"""Per-shape timing of the split-bf16 GEMM (CUDA events, L2 flushed by the operand sizes)."""
import sys, torch
sys.path.insert(0, ".")
from multiplexed_image_annotator_b200 import ops
dev = "cuda"
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
M = cells * 101
D = 576
shapes = [("qkv", 3 * D, D, ops.EPI_STORE_SPLIT), ("proj", D, D, ops.EPI_RESIDUAL), ("fc1", 4 * D, D, ops.EPI_GELU),
          ("fc2", D, 4 * D, ops.EPI_RESIDUAL), ("embed", D, 240, ops.EPI_STORE)]
g = torch.Generator(device=dev).manual_seed(0)
tot = 0.0
for name, N, K, epi in shapes:
    af = torch.randn((M, K), generator=g, device=dev)
    wf = torch.randn((N, K), generator=g, device=dev) * 0.05
    t = ops.weight_log2_scale(float(wf.abs().max().item()))
    ops_in = {"bf16": (ops.split_bf16(af), ops.split_bf16(wf)),
              "f16f8": (ops.split_planes(af, ops.FMT_F16F8), ops.split_planes(wf, ops.FMT_F16F8, True, t))}
    del af, wf
    b = torch.randn(N, generator=g, device=dev)
    out = None
    if epi == ops.EPI_RESIDUAL or epi == ops.EPI_STORE:
        out = torch.zeros((M, N), device=dev)
    else:
        out = torch.empty((2, M, N), dtype=torch.bfloat16, device=dev)
    for prec in ("bf16x3", "f16f8", "bf16x1"):
        a, w = ops_in["f16f8" if prec == "f16f8" else "bf16"]
        if out.dim() == 3:
            out = out.view(torch.int16 if (prec == "f16f8" and epi == ops.EPI_GELU) else torch.bfloat16)
        for _ in range(3):
            ops.gemm(a, w, b, None, epi, out=out, precision=prec, w_log2_scale=t + 8)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            ops.gemm(a, w, b, None, epi, out=out, precision=prec, w_log2_scale=t + 8)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        passes = {"bf16x3": 3, "f16f8": 2, "bf16x1": 1}[prec]
        tf = 2.0 * M * N * K / ms / 1e9
        print(f"{name:6s} M={M} N={N:5d} K={K:5d} {prec}: {ms:7.3f} ms  algorithmic {tf:7.1f} TF/s  issued {tf*passes:7.1f} TF/s")
        if prec == "bf16x3" and name != "embed": tot += ms
    del a, w, out, ops_in
print(f"sum of the four block GEMMs (bf16x3): {tot:.3f} ms per layer-chunk -> {tot*12*52000/cells:.0f} ms per 52k-cell step")
