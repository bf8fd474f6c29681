#!/usr/bin/env python
"""2-GPU check (run under torchrun on a multi-GPU box, not collected by pytest):
the cell-range sharded HotPath (every rank holds the image, owns a contiguous range of cells, one
all_gather of labels / confidences + all_reduce of counts over NCCL) must reproduce the unsharded
single-rank result bit-for-bit.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiplexed_image_annotator_b200 import engine, synth, weights          # noqa: E402
from multiplexed_image_annotator_b200.pipeline import HotPath                # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()

mask = synth.synth_mask(1500, 1300, seed=11, device=dev)
img = torch.from_numpy(synth.to_uint16(synth.synth_image(mask, 10, seed=11))).to(dev)
idx = {"immune_extended": list(range(10))}
sd = weights.random_vit_state("immune_extended", seed=5)
eng = engine.VitEngine("immune_extended", sd, dev)
single = HotPath(idx, {"immune_extended": eng}, device=dev, shard_cells=False, chunk_cells=1000)
ref = single.run(img, mask, to_host=False, keep_probs=True)
cal = weights.calibrate_head(sd, torch.log(ref.probs["immune_extended"][:256]).mean(0).cpu().numpy(), 20.0)   # any fixed head change
eng.set_head(cal["head.weight"], cal["head.bias"])
ref = single.run(img, mask, to_host=False)
sharded = HotPath(idx, {"immune_extended": eng}, device=dev, shard_cells=True, chunk_cells=1000).run(img, mask, to_host=False)
assert sharded.n_cells == ref.n_cells
assert torch.equal(sharded.label, ref.label), "labels differ between sharded and single-rank runs"
assert torch.equal(sharded.confidence, ref.confidence), "confidences differ"
assert torch.equal(sharded.counts, ref.counts), "counts differ"
# batch mode: three images round-robin over the ranks; every rank ends with every image's labels = the single-rank results
scenes = []
for k in range(3):
    m = synth.synth_mask(700 + 100 * k, 900, seed=40 + k, device=dev)
    scenes.append((torch.from_numpy(synth.to_uint16(synth.synth_image(m, 10, seed=40 + k))).to(dev), m))
batch = HotPath(idx, {"immune_extended": eng}, device=dev, chunk_cells=1000).run_batch(scenes, to_host=False)
for (im, m), got in zip(scenes, batch):
    want = single.run(im, m, to_host=False)
    assert got.n_cells == want.n_cells and torch.equal(got.label, want.label) and torch.equal(got.confidence, want.confidence)
    assert torch.equal(got.counts, want.counts)
# the reference-shaped Annotator under 2 ranks (stage 1 split by channel, cells split by range): same labels as one rank
import tempfile
import numpy as np
from multiplexed_image_annotator_b200.cell_type_annotation import model as bmodel
from multiplexed_image_annotator_b200.cell_type_annotation.markerParse import MarkerParser
with tempfile.TemporaryDirectory() as tmp:
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        np.save("img.npy", img.cpu().numpy()); np.save("mask.npy", mask.cpu().numpy())
        synth.write_marker_file("markers.txt", list(MarkerParser(strict=True).panels["immune_extended"]))
        with open("images.csv", "w") as f:
            f.write("image_path,mask_path\nimg.npy,mask.npy\n")
        cal_sd = dict(sd); cal_sd["head.weight"], cal_sd["head.bias"] = cal["head.weight"], cal["head.bias"]
        bmodel.register_state("immune_extended", cal_sd)
        ann = bmodel.Annotator("markers.txt", "images.csv", "cuda", "./", "mg", True, True, -1, True, 0.3, 99.8, 0.3, 30, None, n_jobs=0)
        ann.preprocess()
        ann.predict(128)
        assert np.array_equal(ann.labels_index[0], ref.label.cpu().numpy()), "Annotator labels differ under 2 ranks"
    finally:
        os.chdir(cwd)
ok = torch.ones(1, device=dev)
dist.all_reduce(ok)
if rank == 0:
    print(f"multi-gpu check ok: world={world} cells={ref.n_cells} labels={len(torch.unique(ref.label))} "
          f"counts={ref.counts.tolist()}")
dist.destroy_process_group()
