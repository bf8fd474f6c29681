#!/usr/bin/env python
"""One launch of each block GEMM shape per precision (for ncu captures).  usage: [PRECS=f16f8,bf16x3] tools/gemm_one.py <cells> [shape ...]"""
import os, sys, torch
sys.path.insert(0, ".")
from multiplexed_image_annotator_b200 import ops
cells = int(sys.argv[1]); want = sys.argv[2:] or ["fc1"]
dev = "cuda"; M, D = cells * 101, 576
shapes = {"qkv": (3 * D, D, ops.EPI_STORE_SPLIT), "proj": (D, D, ops.EPI_RESIDUAL), "fc1": (4 * D, D, ops.EPI_GELU), "fc2": (D, 4 * D, ops.EPI_RESIDUAL)}
g = torch.Generator(device=dev).manual_seed(0)
for name in want:
    N, K, epi = shapes[name]
    af = torch.randn((M, K), generator=g, device=dev); wf = torch.randn((N, K), generator=g, device=dev) * 0.05
    t = ops.weight_log2_scale(float(wf.abs().max().item()))
    b = torch.randn(N, generator=g, device=dev)
    split = epi in (ops.EPI_GELU, ops.EPI_STORE_SPLIT)
    out = torch.zeros((2, M, N), dtype=torch.int16, device=dev) if split else torch.zeros((M, N), device=dev)
    for prec in os.environ.get("PRECS", "bf16x3,f16f8,bf16x1").split(","):
        if prec == "f16f8":
            a, w = ops.split_planes(af, ops.FMT_F16F8), ops.split_planes(wf, ops.FMT_F16F8, True, t)
        else:
            a, w = ops.split_bf16(af), ops.split_bf16(wf)
        o = out.view(torch.bfloat16) if (split and not (prec == "f16f8" and epi == ops.EPI_GELU)) else out
        ops.gemm(a, w, b, None, epi, out=o, precision=prec, w_log2_scale=t + 8)
        torch.cuda.synchronize()
        print(name, prec, "done", flush=True)
