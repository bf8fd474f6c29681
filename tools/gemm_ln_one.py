#!/usr/bin/env python
"""One launch of each LayerNorm-fold GEMM variant (for ncu).  usage: tools/gemm_ln_one.py <cells> <variant> ...
variants: qkv qkv+ln_in fc1 fc1+ln_in proj proj+ln_out fc2 fc2+ln_out"""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from multiplexed_image_annotator_b200 import _lib, ops

cells = int(sys.argv[1])
dev = "cuda"
M, D = cells * 101, 576
g = torch.Generator(device=dev).manual_seed(0)
prec, fmt = "f16f8", ops.FMT_F16F8
x = torch.randn((M, D), generator=g, device=dev)
xa = ops.split_planes(x, fmt)
# statistics of x as a producer epilogue would leave them (6 slots of 96 columns)
xs = x.view(M, 6, 96)
stats = torch.zeros((M, 8, 2), device=dev)
stats[:, :6, 0] = xs.sum(2)
stats[:, :6, 1] = (xs * xs).sum(2)
shapes = {"qkv": (3 * D, D, ops.EPI_STORE_SPLIT), "fc1": (4 * D, D, ops.EPI_GELU), "proj": (D, D, ops.EPI_RESIDUAL), "fc2": (D, 4 * D, ops.EPI_RESIDUAL)}
for v in sys.argv[2:]:
    name, _, mode = v.partition("+")
    N, K, epi = shapes[name]
    a = xa if K == D else ops.split_planes(torch.randn((M, K), generator=g, device=dev), fmt)
    wf = torch.randn((N, K), generator=g, device=dev) * 0.05
    t = ops.weight_log2_scale(float(wf.abs().max().item()))
    w = ops.split_planes(wf, fmt, True, t)
    b, c1 = torch.randn(N, generator=g, device=dev), torch.randn(N, generator=g, device=dev)
    if mode == "ln_in":
        out = ops.gemm_ln(a, w, b, None, epi, precision=prec, w_log2_scale=t + 8, stats_in=stats, c1=c1, slots_in=6)
    elif mode == "ln_out":
        out = ops.gemm_ln(a, w, b, None, ops.EPI_RESIDUAL_LN, out=x.clone(), precision=prec, w_log2_scale=t + 8)
    else:
        out = ops.gemm(a, w, b, None, epi, out=x.clone() if epi == ops.EPI_RESIDUAL else None, precision=prec, w_log2_scale=t + 8)
    torch.cuda.synchronize()
    print(v, "ok")
