#!/usr/bin/env python
"""A/B timing of libribca build variants on the block GEMM shapes: the variants are loaded side by side in one process
and timed in interleaved rounds (the GPU is power-capped, so back-to-back runs of different processes are not comparable).
usage: tools/gemm_ab.py <cells> <lib.so> [<lib.so> ...]"""
import ctypes as C, statistics, sys, torch
sys.path.insert(0, ".")
from multiplexed_image_annotator_b200 import _lib, ops

cells = int(sys.argv[1])
paths = sys.argv[2:]
libs = []
for p in paths:
    h = C.CDLL(p)
    fn = h.ribca_gemm_splitbf16
    fn.restype, fn.argtypes = _lib.SIGNATURES["ribca_gemm_splitbf16"]
    libs.append((p.split("/")[-1], fn, h))
dev = "cuda"
M, D = cells * 101, 576
shapes = [("qkv", 3 * D, D, ops.EPI_STORE_SPLIT), ("proj", D, D, ops.EPI_RESIDUAL), ("fc1", 4 * D, D, ops.EPI_GELU),
          ("fc2", D, 4 * D, ops.EPI_RESIDUAL)]
g = torch.Generator(device=dev).manual_seed(0)
st = torch.cuda.current_stream().cuda_stream
tot = {}
for name, N, K, epi in shapes:
    af = torch.randn((M, K), generator=g, device=dev)
    wf = torch.randn((N, K), generator=g, device=dev) * 0.05
    t = ops.weight_log2_scale(float(wf.abs().max().item()))
    ops_in = {"bf16x3": (ops.split_bf16(af), ops.split_bf16(wf)), "bf16x1": None,
              "f16f8": (ops.split_planes(af, ops.FMT_F16F8), ops.split_planes(wf, ops.FMT_F16F8, True, t))}
    ops_in["bf16x1"] = ops_in["bf16x3"]
    del af, wf
    b = torch.randn(N, generator=g, device=dev)
    split = epi in (ops.EPI_GELU, ops.EPI_STORE_SPLIT)
    out = torch.zeros((2, M, N), dtype=torch.int16, device=dev) if split else torch.zeros((M, N), device=dev)
    for prec in ("bf16x3", "f16f8", "bf16x1"):
        a, w = ops_in[prec]
        def call(fn):
            rc = fn(a.data_ptr(), M * K, w.data_ptr(), N * K, M, N, K, b.data_ptr(), None, 0, epi,
                    None if split else out.data_ptr(), out.data_ptr() if split else None, M * N, ops.PRECISION[prec], t + 8, st)
            assert rc == 0, rc
        times = {n: [] for n, _, _ in libs}
        for n, fn, _ in libs:
            for _ in range(3):
                call(fn)
        torch.cuda.synchronize()
        for r in range(7):
            for n, fn, _ in libs:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(8):
                    call(fn)
                e1.record(); torch.cuda.synchronize()
                times[n].append(e0.elapsed_time(e1) / 8)
        line = f"{name:5s} N={N:5d} K={K:5d} {prec:7s}:"
        for n, _, _ in libs:
            med = statistics.median(times[n])
            tot[(n, prec)] = tot.get((n, prec), 0.0) + med
            line += f"  {n}: {med:6.3f} (min {min(times[n]):6.3f})"
        print(line, flush=True)
    del out, ops_in
for (n, prec), v in sorted(tot.items()):
    print(f"sum of the four block GEMMs {n:24s} {prec:7s}: {v:7.3f} ms -> {v * 12 * 51984 / cells:7.0f} ms per C2 step")
