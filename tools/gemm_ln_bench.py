#!/usr/bin/env python
"""Per-GEMM cost of the LayerNorm-folded flow against the plain epilogues + LayerNorm kernels, interleaved rounds in one process.
usage: tools/gemm_ln_bench.py <cells> [lib.so ...]   (f16f8, vit_l shapes)"""
import ctypes as C, statistics, sys, torch
sys.path.insert(0, ".")
from multiplexed_image_annotator_b200 import _lib, ops

cells = int(sys.argv[1])
paths = sys.argv[2:] or [_lib.LIB_PATH]
dev = "cuda"
M, D = cells * 101, 576
g = torch.Generator(device=dev).manual_seed(0)
st = torch.cuda.current_stream().cuda_stream
prec, fmt = "f16f8", ops.FMT_F16F8
x = torch.randn((M, D), generator=g, device=dev)
stats = torch.zeros((M, 8, 2), device=dev)
xa = ops.split_planes(x, fmt)
gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)


def planes(n):
    return torch.zeros((2, M, n), dtype=torch.int16, device=dev)


cases = []     # (name, N, K, epilogue, ln_in, ln_out)
for name, N, K, epi in (("qkv", 3 * D, D, ops.EPI_STORE_SPLIT), ("fc1", 4 * D, D, ops.EPI_GELU)):
    cases += [(name, N, K, epi, False, False), (name + "+ln_in", N, K, epi, True, False)]
for name, N, K in (("proj", D, D), ("fc2", D, 4 * D)):
    cases += [(name, N, K, ops.EPI_RESIDUAL, False, False), (name + "+ln_out", N, K, ops.EPI_RESIDUAL_LN, False, True)]
for p in paths:
    h = C.CDLL(p)
    fn = h.ribca_gemm_ln
    fn.restype, fn.argtypes = _lib.SIGNATURES["ribca_gemm_ln"]
    res = {}
    for name, N, K, epi, ln_in, ln_out in cases:
        a = ops.split_planes(torch.randn((M, K), generator=g, device=dev), fmt)
        wf = torch.randn((N, K), generator=g, device=dev) * 0.05
        t = ops.weight_log2_scale(float(wf.abs().max().item()))
        w = ops.split_planes(wf, fmt, True, t)
        b, c1 = torch.randn(N, generator=g, device=dev), torch.randn(N, generator=g, device=dev)
        split = epi in (ops.EPI_GELU, ops.EPI_STORE_SPLIT)
        out = planes(N) if split else torch.zeros((M, N), device=dev)
        ln = _lib.LnFold()
        ln.eps = 1e-6
        if ln_in:
            ln.stats_in, ln.c1, ln.slots_in = stats.data_ptr(), c1.data_ptr(), 6
        if ln_out:
            ln.stats_out = stats.data_ptr()

        def call():
            rc = fn(a.data_ptr(), M * K, w.data_ptr(), N * K, M, N, K, b.data_ptr(), None, 0, epi, None if split else out.data_ptr(),
                    out.data_ptr() if split else (xa.data_ptr() if ln_out else None), M * N, ops.PRECISION[prec], t + 8, C.byref(ln), st)
            assert rc == 0, (rc, name)
        for _ in range(3):
            call()
        ts = []
        for r in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(6):
                call()
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 6)
        res[name] = statistics.median(ts)
        print(f"{p.split('/')[-1]:22s} {name:12s} N={N:5d} K={K:5d}: {res[name]:6.3f} ms (min {min(ts):6.3f})", flush=True)
        del a, w, out
    ts = []
    for r in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(6):
            ops.layernorm_split(x, gamma, beta, 1e-6, fmt)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 6)
    ln_ms = statistics.median(ts)
    plain = res["qkv"] + res["proj"] + res["fc1"] + res["fc2"] + 2 * ln_ms
    fold = res["qkv+ln_in"] + res["proj+ln_out"] + res["fc1+ln_in"] + res["fc2+ln_out"]
    print(f"{p.split('/')[-1]:22s} layernorm {ln_ms:.3f} ms | block: plain + 2 LN {plain:.3f} ms, folded {fold:.3f} ms -> "
          f"{(fold - plain) * 12 * 51984 / cells:+.0f} ms per C2 step")
