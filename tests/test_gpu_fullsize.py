"""Parity at BASELINE.json's full single-GPU size (15 x 4096 x 4096, ~52 k cells) through
size-independent properties and independently computed references:

  stage 2  every cell's bbox / coordinate sums / area against torch scatter-reductions of the same mask
  stage 1  one full 4096^2 channel bit-exact against the CPU oracle (scipy), range / -1 properties for all
  stage 3  a random sample of cells bit-exact against the oracle; window integers for every cell
  stage 4  chunking invariance: the probabilities of a cell do not depend on the batch it is computed in
           (bit-for-bit), rows sum to 1
  stage 5  counts add up to the number of cells; labels of 'Others' carry confidence -1 or the max prob
  N ranks  the cell-range sharding (parallel.shard_range) reproduces the single-range result bit-for-bit
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ribca_oracle as orc                                  # checker only
from multiplexed_image_annotator_b200 import engine, ops, parallel, synth, weights
from multiplexed_image_annotator_b200.cell_type_annotation.model import merge_on_device, OTHERS

DEV = "cuda"
S = 4096


@pytest.fixture(scope="module")
def scene():
    mask = synth.synth_mask(S, S, seed=2, device=DEV)
    img = synth.synth_image(mask, 15, seed=2)
    img_u16 = torch.from_numpy(synth.to_uint16(img)).to(DEV)
    del img
    norm = ops.normalize(img_u16, 0.3, 99.8)
    cells = ops.cell_stats(mask)
    return {"mask": mask, "img_u16": img_u16, "norm": norm, "cells": cells, "min": ops.channel_min(norm)}


def test_cell_stats_full_size_against_torch_reductions(scene):
    mask, cells = scene["mask"], scene["cells"]
    flat = mask.reshape(-1).long()
    n_ids = int(flat.max()) + 1
    rows = torch.arange(S, device=DEV).view(-1, 1).expand(S, S).reshape(-1)
    cols = torch.arange(S, device=DEV).view(1, -1).expand(S, S).reshape(-1)
    count = torch.bincount(flat, minlength=n_ids)
    ids = torch.nonzero(count[1:] > 0).reshape(-1) + 1
    assert cells.n == len(ids) > 50000
    assert torch.equal(cells.ids.long(), ids)
    assert torch.equal(cells.count.long(), count[ids])
    sum_r = torch.zeros(n_ids, dtype=torch.int64, device=DEV).scatter_add_(0, flat, rows)
    sum_c = torch.zeros(n_ids, dtype=torch.int64, device=DEV).scatter_add_(0, flat, cols)
    assert torch.equal(cells.sums, torch.stack([sum_r[ids], sum_c[ids]], 1))
    big = torch.full((n_ids,), 1 << 30, dtype=torch.int64, device=DEV)
    rmin = big.clone().scatter_reduce_(0, flat, rows, "amin"); rmax = (-big).scatter_reduce_(0, flat, rows, "amax")
    cmin = big.clone().scatter_reduce_(0, flat, cols, "amin"); cmax = (-big).scatter_reduce_(0, flat, cols, "amax")
    want = torch.stack([rmin[ids], rmax[ids], cmin[ids], cmax[ids]], 1)
    assert torch.equal(cells.bbox.long(), want)
    assert int(cells.count.sum()) == int((mask > 0).sum())


def test_cell_pixels_full_size_against_a_stable_sort(scene):
    """CSR pixel lists of the 4096^2 mask == stable sort of the foreground pixels by label (raster order inside a cell)."""
    mask, cells = scene["mask"], scene["cells"]
    off, rows, cols = ops.cell_pixels(mask, cells)
    flat = mask.reshape(-1)
    fg = torch.nonzero(flat > 0).reshape(-1)                     # raster order
    order = torch.sort(flat[fg].long(), stable=True).indices
    lin = fg[order]
    assert torch.equal(rows.long(), lin // S) and torch.equal(cols.long(), lin % S)
    assert torch.equal(off[1:] - off[:-1], cells.count.long()) and int(off[-1]) == len(fg)
    assert torch.equal(mask[rows.long(), cols.long()].long(), torch.repeat_interleave(cells.ids.long(), cells.count.long()))


def test_normalize_full_size(scene):
    norm = scene["norm"]
    assert norm.shape == (15, S, S) and float(norm.min()) >= -1.0 and float(norm.max()) <= 1.0
    assert torch.all(norm.amax(dim=(1, 2)) == 1.0)                       # max > 25 everywhere in this scene
    want = orc.normalize(scene["img_u16"][3:4].cpu().numpy(), 0.3, 99.8)     # one full channel through scipy / numpy
    got = norm[3:4].cpu().numpy()
    assert np.array_equal(got, want), f"{(got != want).sum()} of {got.size} pixels differ"


def test_patches_full_size_sample(scene):
    cells, norm, mask = scene["cells"], scene["norm"], scene["mask"]
    index = list(range(15))
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(cells.n, size=48, replace=False))
    pick[0], pick[-1] = 0, cells.n - 1                                      # corner cells
    host_img, host_mask = norm.cpu().numpy(), mask.cpu().numpy()
    st = {"ids": cells.ids.cpu().numpy(), "bbox": cells.bbox.cpu().numpy()}
    want, want_int, want_win = orc.build_patches(host_img, host_mask, index, st, cells=pick.tolist())
    for j, k in enumerate(pick.tolist()):
        (pt,), avg, win = ops.build_patches(norm, mask, scene["min"], cells, [index], k, 1, want_intensity=True, want_windows=True)
        assert np.array_equal(win.cpu().numpy()[0], want_win[j])
        assert np.array_equal(pt.cpu().numpy()[0], want[j]), f"cell index {k}"
        np.testing.assert_allclose((avg.cpu().numpy()[0] + 1) / 2, want_int[j], rtol=1e-12, atol=1e-14)
    # window integers of every cell (utils.py:227-235) against a vectorised host computation
    _, _, wins = ops.build_patches(norm, mask, scene["min"], cells, [], 0, cells.n, want_windows=True)
    bb = st["bbox"].astype(np.int64)
    r0 = np.maximum((bb[:, 0] + bb[:, 1]) // 2 - 20, 0); c0 = np.maximum((bb[:, 2] + bb[:, 3]) // 2 - 20, 0)
    want_w = np.stack([r0, np.minimum(r0 + 40, S), c0, np.minimum(c0 + 40, S)], 1)
    assert np.array_equal(wins.cpu().numpy(), want_w)


def test_network_chunking_invariance_and_merge_properties(scene):
    cells, norm, mask = scene["cells"], scene["norm"], scene["mask"]
    n = 3000
    (patches,), _, _ = ops.build_patches(norm, mask, scene["min"], cells, [list(range(15))], 1000, n)
    sd = weights.random_vit_state("immune_full", seed=7)
    big = engine.VitEngine("immune_full", sd, DEV, max_cells_per_call=4096)
    small = engine.VitEngine("immune_full", sd, DEV, max_cells_per_call=777)
    _, logits = big.forward(patches[:256], return_logits=True)
    cal = weights.calibrate_head(sd, logits.mean(0).cpu().numpy(), 20.0)
    for e in (big, small):
        e.set_head(cal["head.weight"], cal["head.bias"])
    p_big, p_small = big.forward(patches), small.forward(patches)
    assert torch.equal(p_big, p_small), "a cell's probabilities depend on the batch it was computed in"
    assert torch.equal(p_big, big.forward(patches)), "non-deterministic forward"
    assert float((p_big.sum(1) - 1).abs().max()) < 1e-5
    label, conf, counts = merge_on_device({"immune_full": p_big}, 0.3, None)
    assert int(counts.sum()) == n and torch.equal(counts, torch.bincount(label.long(), minlength=18))
    top, arg = p_big.max(1)
    relabelled = conf == -1
    assert torch.all(label[relabelled] == OTHERS) and torch.all(top[relabelled] < np.float32(0.3))
    assert torch.equal(conf[~relabelled], top[~relabelled])
    assert len(torch.unique(label)) >= 6                    # calibrated head: a spread label histogram, not one class
    # cell-range sharding: concatenating the ranks' ranges reproduces the single-range result bit-for-bit
    parts = []
    for r in range(3):
        lo, hi = parallel.shard_range(n, r, 3)
        parts.append(big.forward(patches[lo:hi]))
    assert torch.equal(torch.cat(parts), p_big)
