"""GPU parity tests for stages 1-3 and 5: the CUDA path (through the C ABI) against the golden
fixtures produced by the reference and against the CPU oracle on seeded inputs.
Bit-exact for integer work AND for the float32 stage outputs (the kernels follow scipy's / numpy's
operation order); 1e-12 relative for the float64 per-cell mean intensities (different summation tree)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ribca_oracle as orc                                  # checker only
from multiplexed_image_annotator_b200 import ops, synth                  # noqa: E402

DEV = "cuda"


def _npz(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _report_mismatch(name, got, want):
    bad = np.argwhere(got != want)
    msg = f"{name}: {len(bad)} / {got.size} elements differ"
    if len(bad):
        i = tuple(bad[0])
        msg += f"; first at {i}: got {got[i]!r} want {want[i]!r}; max|d| {np.abs(got.astype(np.float64) - want).max():.3e}"
    return msg


# ------------------------------------------------------------------------------------------------
# stage 2
# ------------------------------------------------------------------------------------------------
def _check_stats(mask_np):
    want = orc.cell_stats(mask_np)
    tab = ops.cell_stats(torch.from_numpy(mask_np.astype(np.int32)).to(DEV))
    assert tab.n == len(want["ids"])
    assert np.array_equal(tab.ids.cpu().numpy(), want["ids"])
    assert np.array_equal(tab.bbox.cpu().numpy(), want["bbox"]), _report_mismatch("bbox", tab.bbox.cpu().numpy(), want["bbox"])
    assert np.array_equal(tab.sums.cpu().numpy(), np.stack([want["sum_r"], want["sum_c"]], 1))
    assert np.array_equal(tab.count.cpu().numpy(), want["count"])
    assert np.array_equal(tab.centroids().cpu().numpy(), orc.centroids(want))       # float64, bit-exact
    idx = tab.id_to_index.cpu().numpy()
    assert np.array_equal(np.nonzero(idx >= 0)[0], want["ids"])
    # CSR pixel lists (= cell_pos_dict): every cell's rows / cols in raster order, cells in ascending id order
    off, rows, cols = (t.cpu().numpy() for t in ops.cell_pixels(torch.from_numpy(mask_np.astype(np.int32)).to(DEV), tab))
    rr, cc = np.nonzero(mask_np)
    order = np.argsort(mask_np[rr, cc], kind="stable")
    assert np.array_equal(off, np.concatenate([[0], np.cumsum(want["count"])]))
    assert np.array_equal(rows, rr[order]) and np.array_equal(cols, cc[order])
    return tab


def test_cell_pixels_match_the_reference_dict(golden_dir):
    """cell_pos_dict[i][id] == (row list, col list) of the reference's _cell_pos_dict, through the lazy mapping."""
    from multiplexed_image_annotator_b200.cell_type_annotation.preprocess import ImageProcessor
    mask = _npz(golden_dir, "cells_example1_crop.npz")["mask"].astype(np.int32)
    want = orc.cell_pos_dict(mask)
    dev = torch.from_numpy(mask).to(DEV)
    pos = ImageProcessor._positions(mask, ops.cell_stats(dev), dev)
    assert [int(k) for k in pos] == [int(k) for k in want]
    for cid in want:
        assert pos[cid] == want[cid]
        r, c = pos.centroid(cid)
        assert r == np.mean(want[cid][0]) and c == np.mean(want[cid][1])


@pytest.mark.parametrize("fixture", ["cells_example2.npz", "cells_example1_crop.npz"])
def test_cell_stats_reference_golden(golden_dir, fixture):
    g = _npz(golden_dir, fixture)
    tab = _check_stats(g["mask"].astype(np.int32))
    t = g["table"]
    assert np.array_equal(tab.ids.cpu().numpy(), t[:, 0])
    assert np.array_equal(tab.bbox.cpu().numpy(), t[:, 1:5])
    assert np.array_equal(tab.sums.cpu().numpy(), t[:, 5:7])
    assert np.array_equal(tab.count.cpu().numpy(), t[:, 7])


@pytest.mark.parametrize("shape", [(1, 1), (7, 13), (64, 64), (257, 1031), (1000, 1200)])
def test_cell_stats_ragged_masks(shape):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    h, w = shape
    mask = np.zeros(shape, np.int32)
    # sparse non-contiguous labels, 1-pixel cells, cells touching every border
    labels = rng.choice(np.arange(1, 50000), size=max(1, h * w // 40), replace=False)
    for lab in labels:
        r, c = rng.integers(0, h), rng.integers(0, w)
        rh, cw = rng.integers(1, 9), rng.integers(1, 9)
        mask[r:r + rh, c:c + cw] = lab
    mask[0, 0] = 49999 + 1
    mask[h - 1, w - 1] = 7
    _check_stats(mask)


def test_cell_stats_large_synthetic():
    mask = synth.synth_mask(2048, 2048, seed=5).numpy()
    tab = _check_stats(mask)
    assert tab.n > 12000


def test_cell_stats_empty_mask():
    tab = ops.cell_stats(torch.zeros((33, 47), dtype=torch.int32, device=DEV))
    assert tab.n == 0 and tab.ids.numel() == 0


# ------------------------------------------------------------------------------------------------
# stage 1
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,blur,amax", [("b03_a998", 0.3, 99.8), ("b0_a100", 0, 100), ("b1_a100", 1, 100),
                                           ("b04_a95", 0.4, 95.0)])
def test_normalize_reference_golden(golden_dir, tag, blur, amax):
    g = _npz(golden_dir, "normalize.npz")
    got = ops.normalize(torch.from_numpy(g["img"]).to(DEV), blur, amax).cpu().numpy()
    assert np.array_equal(got, g[tag]), _report_mismatch(tag, got, g[tag])
    got = ops.normalize(torch.from_numpy(g["img_f32"]).to(DEV), blur, amax).cpu().numpy()
    assert np.array_equal(got, g["f32_" + tag]), _report_mismatch("f32_" + tag, got, g["f32_" + tag])


@pytest.mark.parametrize("shape,dtype", [((3, 300, 517), np.uint16), ((2, 64, 64), np.uint8), ((1, 700, 90), np.float32)])
def test_normalize_vs_oracle(shape, dtype):
    rng = np.random.default_rng(11)
    c, h, w = shape
    mask = synth.synth_mask(h, w, seed=3)
    img = synth.synth_image(mask, c, seed=3).numpy()
    if dtype == np.uint8:
        img = (img // 16).clip(0, 255)
    img = img.astype(dtype)
    want = orc.normalize(img, 0.3, 99.8)
    got, stats = ops.normalize(torch.from_numpy(img).to(DEV), 0.3, 99.8, return_stats=True)
    got = got.cpu().numpy()
    assert np.array_equal(got, want), _report_mismatch("normalize", got, want)
    mn = ops.channel_min(torch.from_numpy(want).to(DEV)).cpu().numpy()
    assert np.array_equal(mn, want.min(axis=(1, 2)))


# ------------------------------------------------------------------------------------------------
# stage 3
# ------------------------------------------------------------------------------------------------
def _gpu_patches(image, mask, index, n=None):
    img = torch.from_numpy(np.ascontiguousarray(image)).to(DEV)
    m = torch.from_numpy(np.ascontiguousarray(mask.astype(np.int32))).to(DEV)
    tab = ops.cell_stats(m)
    mn = ops.channel_min(img)
    outs, avg, wins = ops.build_patches(img, m, mn, tab, [index], 0, n, want_intensity=True, want_windows=True)
    return outs[0].cpu().numpy(), (avg.cpu().numpy() + 1) / 2, wins.cpu().numpy(), tab


@pytest.mark.parametrize("tag,maskkey", [("synth", "mask"), ("synth_q3", "mask"), ("real", "mask_real")])
def test_patches_reference_golden(golden_dir, tag, maskkey):
    g = _npz(golden_dir, "patches.npz")
    mask = g[maskkey]
    image = g["img_norm"][:, : mask.shape[0], : mask.shape[1]]
    pt, inten, wins, tab = _gpu_patches(image, mask, g[tag + "_index"].tolist())
    assert np.array_equal(tab.ids.cpu().numpy(), g[tag + "_ids"])
    assert np.array_equal(wins, g[tag + "_windows"]), _report_mismatch("windows", wins, g[tag + "_windows"])
    n = len(g[tag + "_patches"])
    assert np.array_equal(pt[:n], g[tag + "_patches"]), _report_mismatch("patches", pt[:n], g[tag + "_patches"])
    np.testing.assert_allclose(inten, g[tag + "_intensity"], rtol=1e-12, atol=1e-14)


def test_patches_vs_oracle_border_cells():
    h, w = 97, 131                                    # windows truncated at every border
    mask = synth.synth_mask(h, w, grid=14, seed=9, r_lo=3, r_hi=7).numpy()
    mask[0:3, 0:4] = 9001                             # corner cells
    mask[h - 2:, w - 3:] = 9002
    mask[50, 60] = 9003                               # 1-pixel cell
    img = synth.to_uint16(synth.synth_image(torch.from_numpy(mask), 6, seed=9))
    norm = orc.normalize(img, 0.3, 99.8)
    index = [5, 0, -1, 3, 1]
    want, w_int, w_win = orc.build_patches(norm, mask, index)
    pt, inten, wins, _ = _gpu_patches(norm, mask, index)
    assert np.array_equal(wins, w_win)
    assert np.array_equal(pt, want), _report_mismatch("patches", pt, want)
    np.testing.assert_allclose(inten, w_int, rtol=1e-12, atol=1e-14)


def test_patches_multi_panel_and_chunks():
    mask = synth.synth_mask(200, 200, seed=4).numpy()
    img = orc.normalize(synth.to_uint16(synth.synth_image(torch.from_numpy(mask), 8, seed=4)), 0.3, 99.8)
    it = torch.from_numpy(img).to(DEV)
    mt = torch.from_numpy(mask).to(DEV)
    tab = ops.cell_stats(mt)
    mn = ops.channel_min(it)
    panels = [[0, 1, 2], [7, -1, 5, 4, -1, 3, 2], [6]]
    full, _, _ = ops.build_patches(it, mt, mn, tab, panels)
    for p, chans in enumerate(panels):
        want, _, _ = orc.build_patches(img, mask, chans, cells=range(0, 30))
        assert np.array_equal(full[p][:30].cpu().numpy(), want)
    part, _, _ = ops.build_patches(it, mt, mn, tab, panels, cell_begin=17, n_cells=40)
    for p in range(3):
        assert torch.equal(part[p], full[p][17:57])


@pytest.mark.parametrize("cell_size", [15, 20, 40, 45, 60])
def test_patches_other_cell_sizes_reference_golden(golden_dir, cell_size):
    """cell_size != 30: window edge int(40 * cell_size / 30), anti-aliased nearest resize to 40 x 40."""
    g = _npz(golden_dir, "cellsize.npz")
    img = torch.from_numpy(g["img_norm"]).to(DEV)
    m = torch.from_numpy(g["mask"]).to(DEV)
    tab = ops.cell_stats(m)
    mn = ops.channel_min(img)
    (pt,), avg, _ = ops.build_patches(img, m, mn, tab, [[4, -1, 2, 0]], want_intensity=True, cell_size=cell_size)
    want = g[f"cs{cell_size}_patches"]
    got = pt.cpu().numpy()[: len(want)]
    assert np.array_equal(got, want), _report_mismatch(f"cell_size {cell_size}", got, want)
    np.testing.assert_allclose((avg.cpu().numpy() + 1) / 2, g[f"cs{cell_size}_intensity"], rtol=1e-12, atol=1e-14)


# ------------------------------------------------------------------------------------------------
# stage 5
# ------------------------------------------------------------------------------------------------
def test_merge_reference_golden(golden_dir):
    from multiplexed_image_annotator_b200.cell_type_annotation.model import merge_on_device, ALL_TYPES
    cases = json.load(open(os.path.join(golden_dir, "merge.json")))
    checked = 0
    for case in cases:
        probs = {k: torch.tensor(v, dtype=torch.float32, device=DEV) for k, v in case["probs"].items()}
        if "raises" in case:
            with pytest.raises(KeyError):
                merge_on_device(probs, case["confidence"], case["ctc"])
            continue
        label, conf, counts = merge_on_device(probs, case["confidence"], case["ctc"])
        names = [ALL_TYPES[i] for i in label.cpu().tolist()]
        assert names == case["labels"], (case["panels"], case["confidence"])
        assert np.array_equal(conf.cpu().numpy().astype(np.float64), np.array(case["conf"]))
        want_counts = np.bincount([ALL_TYPES.index(n) for n in case["labels"]], minlength=18)
        assert np.array_equal(counts.cpu().numpy(), want_counts)
        checked += 1
    assert checked >= 30


# ------------------------------------------------------------------------------------------------
# degenerate inputs through the in-memory pipeline
# ------------------------------------------------------------------------------------------------
def test_hot_path_degenerate_inputs():
    from multiplexed_image_annotator_b200 import engine, weights
    from multiplexed_image_annotator_b200.pipeline import HotPath
    eng = engine.VitEngine("nerve_cell", weights.random_vit_state("nerve_cell", seed=1), DEV)
    hp = HotPath({"nerve_cell": [0, 1, 2]}, {"nerve_cell": eng}, device=DEV, shard_cells=False)
    rng = np.random.default_rng(3)
    # no cells at all
    img = rng.integers(0, 3000, (3, 64, 80)).astype(np.uint16)
    res = hp.run(img, np.zeros((64, 80), np.int32))
    assert res.n_cells == 0 and res.label.numel() == 0 and int(res.counts.sum()) == 0
    # one single-pixel cell in an image smaller than the 40 x 40 patch; an all-zero channel (-> -1 everywhere)
    img = rng.integers(0, 3000, (3, 23, 31)).astype(np.uint16)
    img[1] = 0
    mask = np.zeros((23, 31), np.int32)
    mask[22, 30] = 77
    res = hp.run(img, mask, keep_probs=True)
    assert res.n_cells == 1 and int(res.counts.sum()) == 1
    norm = orc.normalize(img, 0.3, 99.8)
    assert (norm[1] == -1).all()
    want, _, _ = orc.build_patches(norm, mask, [0, 1, 2])
    ref = orc.make_vit("nerve_cell")
    ref.load_state_dict(weights.random_vit_state("nerve_cell", seed=1))
    assert np.abs(res.probs["nerve_cell"].cpu().numpy() - orc.vit_probs(ref, want)).max() < 1e-3


# ------------------------------------------------------------------------------------------------
# spatial statistics: GPU k-NN and the two reductions against scikit-learn / the reference's loops
# ------------------------------------------------------------------------------------------------
def _sk_knn(xy, k):
    from sklearn.neighbors import NearestNeighbors
    return NearestNeighbors(n_neighbors=k, algorithm="ball_tree").fit(xy).kneighbors(xy)


@pytest.mark.parametrize("n,k,layout", [(3000, 25, "uniform"), (5000, 201, "uniform"), (4000, 25, "clustered"), (300, 201, "uniform"),
                                        (2000, 10, "line")])
def test_knn_matches_sklearn_ball_tree(n, k, layout):
    rng = np.random.default_rng(n + k)
    if layout == "uniform":
        xy = rng.random((n, 2)) * [4096.0, 3000.0]
    elif layout == "clustered":            # dense blobs + empty space: many empty grid rings
        c = rng.random((8, 2)) * 5000
        xy = c[rng.integers(0, 8, n)] + rng.normal(0, 20.0, (n, 2))
    else:                                  # degenerate extent in y
        xy = np.stack([rng.random(n) * 1e4, np.full(n, 7.25) + rng.random(n) * 1e-3], 1)
    dist, idx = _sk_knn(xy, k)
    got_idx, got_d = ops.knn_2d(torch.from_numpy(xy).to(DEV), k, return_distance=True)
    got_idx, got_d = got_idx.cpu().numpy(), got_d.cpu().numpy()
    assert np.array_equal(got_idx[:, 0], np.arange(n))                        # self first (distance 0)
    np.testing.assert_allclose(got_d, dist, rtol=1e-14, atol=0)
    same = got_idx == idx
    if not same.all():                     # only exact distance ties may be ordered differently
        rows = np.nonzero(~same.all(1))[0]
        for r in rows:
            assert np.array_equal(np.sort(got_d[r]), np.sort(dist[r]))
            assert set(got_idx[r][got_d[r] < got_d[r, -1]]) == set(idx[r][dist[r] < dist[r, -1]])


def test_neighbourhood_matrix_and_compositions_match_the_reference_loops(tmp_path):
    from multiplexed_image_annotator_b200.cell_type_annotation import spatial_methods as sm
    rng = np.random.default_rng(5)
    n, n_types = 2500, 7
    xy = rng.random((n, 2)) * 2048
    types = rng.integers(0, n_types, n)
    # reference loops (cta/spatial_methods.py:35-45 and 150-176) on scikit-learn's neighbours
    _, idx = _sk_knn(xy, 25)
    want = np.zeros((n_types, n_types))
    for j in range(n):
        for kk in idx[j][1:]:
            want[types[j], types[kk]] += 1
    got = sm.neighborhood_matrix(xy, types, n_types, 25, DEV)
    assert np.array_equal(got, want)
    _, idx = _sk_knn(xy, 201)
    idx = idx[:, 1:]
    comp = np.zeros((n, 8 * n_types))
    for j in range(n):
        for l, m in enumerate(sm.REGION_LEVELS):
            h = np.bincount(types[idx[j, :m]], minlength=n_types).astype(np.float64)
            comp[j, l * n_types:(l + 1) * n_types] = h / h.sum()
    assert np.array_equal(sm.neighbor_compositions(xy, types, DEV), comp)
    # the CSV of neighborhood_analysis: reference text format, rows normalised
    names = np.array([f"T{t}" for t in range(n_types)])
    rows = [{"Column": [x], "Row": [y], "Cell type": int(t), "Cell ID": j} for j, ((x, y), t) in enumerate(zip(xy, types))]
    m = sm.neighborhood_analysis([rows], n_neighbors=25, cell_types=names, integrate=True, normalize=True, batch_id="b",
                                 result_dir=str(tmp_path), device=DEV)
    np.testing.assert_array_equal(m, want / want.sum(1, keepdims=True))
    text = open(tmp_path / "b_integrated_neighborhood.csv").read().split("\n")
    assert text[0] == "cell_type," + "".join(f"T{t}," for t in range(n_types))
    assert text[1] == "T0," + "".join(f"{v:.3f}," for v in m[0])
    with pytest.raises(ValueError):
        ops.knn_2d(torch.from_numpy(xy[:10]).to(DEV), 25)                     # sklearn: n_neighbors > n_samples


def test_normalize_from_host_overlapped_upload_is_identical():
    """The e2e entry uploads the stack channel by channel on a side stream and normalises each channel as it arrives:
    same bits as normalising the resident stack (per-channel statistics are independent)."""
    mask = synth.synth_mask(300, 260, seed=4)
    img = torch.from_numpy(synth.to_uint16(synth.synth_image(mask, 5, seed=4)))
    want = ops.normalize(img.to(DEV), 0.3, 99.8)
    for host in (img, img.pin_memory()):
        got = ops.normalize_from_host(host, torch.device(DEV, torch.cuda.current_device()), 0.3, 99.8)
        torch.cuda.synchronize()
        assert torch.equal(got, want)
