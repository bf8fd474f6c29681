"""Pin the CPU oracle (oracle/ribca_oracle.py) against the golden fixtures that
tests/golden/make_golden.py produced by running the unmodified reference."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ribca_oracle as orc
from multiplexed_image_annotator_b200 import weights


def _npz(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


@pytest.mark.parametrize("fixture", ["cells_example2.npz", "cells_example1_crop.npz"])
def test_cell_stats_match_reference_pixel_lists(golden_dir, fixture):
    g = _npz(golden_dir, fixture)
    st = orc.cell_stats(g["mask"].astype(np.int32))
    tab = g["table"]
    assert np.array_equal(st["ids"], tab[:, 0])
    assert np.array_equal(st["bbox"], tab[:, 1:5])
    assert np.array_equal(st["sum_r"], tab[:, 5])
    assert np.array_equal(st["sum_c"], tab[:, 6])
    assert np.array_equal(st["count"], tab[:, 7])


def test_cell_pos_dict_lists(golden_dir):
    g = _npz(golden_dir, "cells_example1_crop.npz")
    mask = g["mask"]
    d = orc.cell_pos_dict(mask)
    assert list(d.keys()) == sorted(d.keys()) == g["table"][:, 0].tolist()
    for cid in list(d)[:40]:
        rr, cc = np.nonzero(mask == cid)
        assert d[cid] == (rr.tolist(), cc.tolist())


def test_marker_parser(golden_dir, tmp_path):
    cases = json.load(open(os.path.join(golden_dir, "markers.json")))
    for name, case in cases.items():
        f = tmp_path / f"{name}.txt"
        f.write_text("\n".join(case["markers"]) + "\n")
        if "raises" in case:
            with pytest.raises(TypeError):
                orc.parse_markers(str(f), case["strict"])
            continue
        assert orc.parse_markers(str(f), case["strict"]) == case["indices"], name


@pytest.mark.parametrize("tag,blur,amax", [("b03_a998", 0.3, 99.8), ("b0_a100", 0, 100), ("b1_a100", 1, 100),
                                           ("b04_a95", 0.4, 95.0)])
def test_normalize_bit_exact(golden_dir, tag, blur, amax):
    g = _npz(golden_dir, "normalize.npz")
    assert np.array_equal(orc.normalize(g["img"], blur, amax), g[tag])
    assert np.array_equal(orc.normalize(g["img_f32"], blur, amax), g["f32_" + tag])


@pytest.mark.parametrize("tag,maskkey", [("synth", "mask"), ("synth_q3", "mask"), ("real", "mask_real")])
def test_patches_bit_exact(golden_dir, tag, maskkey):
    g = _npz(golden_dir, "patches.npz")
    mask = g[maskkey]
    image = g["img_norm"][:, : mask.shape[0], : mask.shape[1]]
    st = orc.cell_stats(mask)
    assert np.array_equal(st["ids"], g[tag + "_ids"])
    n = len(g[tag + "_patches"])
    pt, inten, wins = orc.build_patches(image, mask, g[tag + "_index"].tolist(), st, cells=range(n))
    assert np.array_equal(pt, g[tag + "_patches"])
    assert np.array_equal(inten, g[tag + "_intensity"][:n])
    assert np.array_equal(wins, g[tag + "_windows"][:n])
    # windows for every cell
    h, w = mask.shape
    allw = np.array([orc.crop_window(b, 40, h, w) for b in st["bbox"]], np.int32)
    assert np.array_equal(allw, g[tag + "_windows"])
    # soft mask alone
    for k in range(len(g[tag + "_smooth"])):
        x0, x1, y0, y1 = allw[k]
        mp = np.zeros((40, 40)); mp[: x1 - x0, : y1 - y0] = mask[x0:x1, y0:y1]
        assert np.array_equal(orc.soft_mask(mp, st["ids"][k]), g[tag + "_smooth"][k])


@pytest.mark.parametrize("cell_size", [15, 20, 40, 45, 60])
def test_patches_other_cell_sizes(golden_dir, cell_size):
    g = _npz(golden_dir, "cellsize.npz")
    st = orc.cell_stats(g["mask"])
    pt, inten, _ = orc.build_patches(g["img_norm"], g["mask"], [4, -1, 2, 0], st, cell_size=cell_size, cells=range(24))
    assert np.array_equal(pt, g[f"cs{cell_size}_patches"])
    assert np.array_equal(inten, g[f"cs{cell_size}_intensity"][:24])


@pytest.mark.parametrize("panel", ["immune_base", "nerve_cell"])
def test_vit_forward(golden_dir, panel):
    g = _npz(golden_dir, "vit.npz")
    model = orc.make_vit(panel)
    model.load_state_dict(weights.random_vit_state(panel, seed=1))
    with torch.no_grad():
        logits = model(torch.from_numpy(g[panel + "_x"])).numpy()
    np.testing.assert_allclose(logits, g[panel + "_logits"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(orc.vit_probs(model, g[panel + "_x"]), g[panel + "_probs"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("panel", ["immune_base", "immune_extended"])
def test_mae_impute(golden_dir, panel):
    g = _npz(golden_dir, "mae.npz")
    model = orc.make_mae(panel)
    model.load_state_dict(weights.random_mae_state(panel, seed=1))
    x = g[panel + "_x"]
    present = g[panel + "_present"].tolist()
    out = orc.impute(model, x, present)
    ref = g[panel + "_out"]
    for k in present:                                      # present channels come back untouched
        assert np.array_equal(out[:, k], x[:, k])
        assert np.array_equal(ref[:, k], x[:, k])
    np.testing.assert_allclose(out, ref, rtol=0, atol=5e-6)
    missing = [k for k in range(x.shape[1]) if k not in present]
    assert np.abs(ref[:, missing] + 1).max() > 1e-3        # really imputed, not the -1 fill


def test_merge_by_voting_all_branches(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "merge.json")))
    assert len(cases) >= 30
    for case in cases:
        preds = {k: np.asarray(v, np.float32) for k, v in case["probs"].items()}
        if "raises" in case:
            with pytest.raises(KeyError):
                orc.merge_by_voting(preds, case["confidence"], case["ctc"])
            continue
        labels, conf = orc.merge_by_voting(preds, case["confidence"], case["ctc"])
        assert labels == case["labels"], case["panels"]
        assert [isinstance(c, int) for c in conf] == case["conf_is_int"]
        assert np.array_equal(np.array([float(c) for c in conf]), np.array(case["conf"]))
        assert orc.unique_cell_types([labels]).tolist() == case["cell_types"]


@pytest.mark.parametrize("tag,strict", [("full", True), ("impute", False), ("struct_nerve", True)])
def test_end_to_end_labels(golden_dir, tag, strict, tmp_path):
    g = _npz(golden_dir, "e2e.npz")
    markers = [str(m) for m in g[tag + "_markers"]]
    mf = tmp_path / "m.txt"
    mf.write_text("\n".join(markers) + "\n")
    indices = orc.parse_markers(str(mf), strict)
    models, imputers = {}, {}
    for panel in orc.predicted_panels(indices):
        sd = weights.calibrate_head(weights.random_vit_state(panel, seed=2), g[f"{tag}_meanlogits_{panel}"], 20.0)
        models[panel] = orc.make_vit(panel)
        models[panel].load_state_dict(sd)
        if -1 in indices[panel]:
            imputers[panel] = orc.make_mae(panel)
            imputers[panel].load_state_dict(weights.random_mae_state(panel, seed=2))
    res = orc.annotate_image(g[tag + "_img"], g[tag + "_mask"], indices, models, imputers, bs=32)
    for panel in res["patches"]:
        tol = 0 if panel not in imputers else 5e-6
        np.testing.assert_allclose(res["patches"][panel], g[f"{tag}_patches_{panel}"], rtol=0, atol=tol)
        np.testing.assert_allclose(res["probs"][panel], g[f"{tag}_probs_{panel}"], rtol=0, atol=2e-6)
    assert res["labels"] == g[tag + "_labels"].tolist()
    np.testing.assert_allclose(np.array([float(c) for c in res["confidence"]]), g[tag + "_conf"], rtol=0, atol=2e-6)
    assert np.array_equal(res["intensity"], g[tag + "_intensity"])


def test_c1_reference_example_configuration(golden_dir, tmp_path):
    """BASELINE configs[0]: the reference's example mask (1850 cells) + examples/markers.txt (vit_m + vit_s, merge branch 2) on
    the seeded synthetic 17-marker image; the oracle against the unmodified reference's CPU run (fixture c1.npz)."""
    import torch
    from multiplexed_image_annotator_b200 import synth
    g = _npz(golden_dir, "c1.npz")
    mask = g["mask"]
    markers = [str(m) for m in g["markers"]]
    img = synth.to_uint16(synth.synth_image(torch.from_numpy(mask), len(markers), seed=1))
    assert int(img.astype(np.int64).sum()) == int(g["image_checksum"][0])          # the generator reproduces the fixture's image
    mf = tmp_path / "m.txt"
    mf.write_text("\n".join(markers) + "\n")
    indices = orc.parse_markers(str(mf), True)
    assert orc.predicted_panels(indices) == ["immune_extended", "structure"]
    models = {}
    for panel in orc.predicted_panels(indices):
        models[panel] = orc.make_vit(panel)
        models[panel].load_state_dict(weights.calibrate_head(weights.random_vit_state(panel, seed=2), g[f"meanlogits_{panel}"], 20.0))
    # the networks on a 300-cell slice keep the CPU suite short; stages 1-3 and 5 run on all 1850 cells
    img_n = orc.normalize(img, 0.3, 99.8)
    stats = orc.cell_stats(mask)
    assert len(stats["ids"]) == 1850 == len(g["labels"])
    sl = slice(700, 1000)
    for panel in models:
        pt, _, _ = orc.build_patches(img_n, mask, indices[panel], stats)
        np.testing.assert_allclose(orc.vit_probs(models[panel], pt[sl], 128), g[f"probs_{panel}"][sl], rtol=0, atol=2e-6)
    labels, conf = orc.merge_by_voting({p: g[f"probs_{p}"] for p in models}, 0.3, None)
    assert labels == g["labels"].tolist()
    assert np.array_equal(np.array([float(c) for c in conf]), g["conf"])
    assert orc.unique_cell_types([labels]).tolist() == g["cell_types"].tolist()
