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
    a = ops.split_bf16(torch.randn((M, K), generator=g, device=dev))
    w = ops.split_bf16(torch.randn((N, K), generator=g, device=dev) * 0.05)
    b = torch.randn(N, generator=g, device=dev)
    out = None
    if epi == ops.EPI_RESIDUAL or epi == ops.EPI_STORE:
        out = torch.zeros((M, N), device=dev)
    else:
        out = torch.empty((2, M, N), dtype=torch.bfloat16, device=dev)
    for prec in ("bf16x3", "bf16x1"):
        for _ in range(3):
            ops.gemm(a, w, b, None, epi, out=out, precision=prec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            ops.gemm(a, w, b, None, epi, out=out, precision=prec)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        passes = 3 if prec == "bf16x3" else 1
        tf = 2.0 * M * N * K / ms / 1e9
        print(f"{name:6s} M={M} N={N:5d} K={K:5d} {prec}: {ms:7.3f} ms  algorithmic {tf:7.1f} TF/s  issued {tf*passes:7.1f} TF/s")
        if prec == "bf16x3" and name != "embed": tot += ms
    del a, w, out
print(f"sum of the four block GEMMs (bf16x3): {tot:.3f} ms per layer-chunk -> {tot*12*52000/cells:.0f} ms per 52k-cell step")
