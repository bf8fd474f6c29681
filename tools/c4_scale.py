#!/usr/bin/env python
"""Scale check at BASELINE configs[3] size: synthetic 15-marker S x S image (default 20000^2, 20-px grid, ~1 M cells),
whole hot path device-resident.  One GPU, or - under torchrun - the same image sharded by cell range over the ranks
(every rank holds the image and runs stages 1-2, stages 3-5 on its range, one all-gather of labels / confidences at the end:
STRONG scaling).  Prints cells/s (device time, max over ranks) and checks size-independent invariants.
    python tools/c4_scale.py [S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29536 tools/c4_scale.py [S]"""
import hashlib, json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from multiplexed_image_annotator_b200 import engine, ops, synth, weights
from multiplexed_image_annotator_b200.pipeline import HotPath

S = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank() if world > 1 else 0
t0 = time.time()
mask = synth.synth_mask(S, S, grid=20, seed=4, device=dev)
img = synth.synth_image(mask, 15, seed=4).to(torch.uint16)
torch.cuda.synchronize()
if rank == 0:
    print(f"scene {S}x{S}: {time.time()-t0:.1f} s, image {img.numel()*2/1e9:.1f} GB, mask {mask.numel()*4/1e9:.1f} GB", flush=True)
panel = "immune_full"
sd = weights.random_vit_state(panel, seed=7)
eng = engine.VitEngine(panel, sd, dev, max_cells_per_call=4096)
hp = HotPath({panel: list(range(15))}, {panel: eng}, chunk_cells=4096, device=dev, shard_cells=world > 1)
# calibrate the head on a corner crop so the label histogram is spread (identical on every rank)
norm = ops.normalize(img[:, :1024, :1024].contiguous(), 0.3, 99.8)
m1 = mask[:1024, :1024].contiguous()
cells1 = ops.cell_stats(m1)
(p256,), _, _ = ops.build_patches(norm, m1, ops.channel_min(norm), cells1, [list(range(15))], 0, 256)
_, logits = eng.forward(p256, return_logits=True)
cal = weights.calibrate_head(sd, logits.mean(0).cpu().numpy(), 20.0)
eng.set_head(cal["head.weight"], cal["head.bias"])
del norm, p256
hp.run(img[:, :2048, :2048].contiguous(), mask[:2048, :2048].contiguous(), to_host=False)        # warm-up (allocator, NCCL)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
res = hp.run(img, mask, to_host=False)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms.item())
n = res.n_cells
cells = res.cells
assert int(res.counts.sum()) == n and res.label.numel() == n
assert int(cells.count.sum()) == int((mask > 0).sum())
c = cells.centroids()
assert bool(((c[:, 0] >= cells.bbox[:, 0]) & (c[:, 0] <= cells.bbox[:, 1]) & (c[:, 1] >= cells.bbox[:, 2]) & (c[:, 1] <= cells.bbox[:, 3])).all())
assert torch.equal(res.counts, torch.bincount(res.label.long(), minlength=18))
digest = hashlib.sha256(res.label.cpu().numpy().tobytes() + res.confidence.cpu().numpy().tobytes()).hexdigest()[:16]
if rank == 0:
    out = {"size": S, "cells": n, "n_gpus": world, "scaling": "strong (one image, cell ranges)" if world > 1 else "single GPU",
           "ms": ms, "cells_per_s": n / (ms / 1000), "labels": int((res.counts > 0).sum()), "counts": res.counts.tolist(),
           "labels_confidences_sha256": digest, "max_mem_gb": torch.cuda.max_memory_allocated() / 1e9, "precision": eng.precision}
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
