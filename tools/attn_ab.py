#!/usr/bin/env python
"""A/B timing of libribca build variants on the classifier attention (ribca_attention_tc), interleaved rounds in one process.
usage: tools/attn_ab.py <cells> <lib.so> [<lib.so> ...]   (ONE=1: a single launch of the last library, for ncu)"""
import ctypes as C, os, statistics, sys, torch
sys.path.insert(0, ".")
from multiplexed_image_annotator_b200 import _lib, ops

cells = int(sys.argv[1])
libs = []
for p in sys.argv[2:]:
    h = C.CDLL(p)
    fn = h.ribca_attention_tc
    fn.restype, fn.argtypes = _lib.SIGNATURES["ribca_attention_tc"]
    libs.append((p.split("/")[-1], fn, h))
dev = "cuda"
tokens, heads, hd = 101, 12, int(os.environ.get("HD", 48))
hdp = (hd + 15) // 16 * 16
M, W = cells * tokens, 3 * heads * hdp
g = torch.Generator(device=dev).manual_seed(0)
qkv = torch.zeros((M, 3, heads, hdp), device=dev)
qkv[..., :hd] = torch.randn((M, 3, heads, hd), generator=g, device=dev)
qs = ops.split_bf16(qkv.reshape(M, W))
del qkv
out = torch.zeros((2, M, heads * hd), dtype=torch.int16, device=dev)
st = torch.cuda.current_stream().cuda_stream
def call(fn):
    rc = fn(qs.data_ptr(), M * W, cells, tokens, heads, hd, out.data_ptr(), M * heads * hd, ops.FMT_F16F8, st)
    assert rc == 0, rc
if os.environ.get("ONE"):
    call(libs[-1][1]); torch.cuda.synchronize(); call(libs[-1][1]); torch.cuda.synchronize(); sys.exit(0)
res = {}
for n, fn, _ in libs:
    call(fn); call(fn)
    torch.cuda.synchronize()
    res[n] = out.clone()
times = {n: [] for n, _, _ in libs}
for r in range(9):
    for n, fn, _ in libs:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            call(fn)
        e1.record(); torch.cuda.synchronize()
        times[n].append(e0.elapsed_time(e1) / 8)
bytes_alg = M * W * 4 + M * heads * hd * 4
for n, _, _ in libs:
    med = statistics.median(times[n])
    same = bool(torch.equal(res[n], res[libs[0][0]]))
    print(f"attention cells={cells} hd={hd} {n:28s}: {med:6.3f} ms (min {min(times[n]):6.3f})  {bytes_alg / med / 1e6:7.1f} GB/s algorithmic  "
          f"-> {med * 12 * 51984 / cells:6.1f} ms per C2 step  bits equal to first: {same}")
