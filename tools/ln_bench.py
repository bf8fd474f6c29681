import sys, statistics, torch
sys.path.insert(0, ".")
from multiplexed_image_annotator_b200 import ops
M, D = 4096 * 101, 576
x = torch.randn((M, D), device="cuda")
g, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
for _ in range(3): ops.layernorm_split(x, g, b, 1e-6, ops.FMT_F16F8)
ts = []
for r in range(7):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.layernorm_split(x, g, b, 1e-6, ops.FMT_F16F8)
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 10)
print(f"layernorm M={M} D={D}: {statistics.median(ts):.4f} ms (min {min(ts):.4f}) -> {M*D*8/statistics.median(ts)/1e6:.0f} GB/s")
