#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run by hand in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

The reference's own .py files are imported through oracle/refshim.py (third-party packages that are
missing offline replaced by the restatements of oracle/standins.py) and executed on small seeded
inputs; inputs and outputs are stored as compressed .npz / .json so that the tests never need
/root/reference.  Fixtures:

  cells_example2.npz   reference golden vector results/test_annotation_1.csv (pixel lists of 582
                       cells) reduced to bbox / sums / count, plus examples/example_2_cell_mask.png
  markers.json         MarkerParser.parse on examples/markers.txt and on synthetic marker lists
  normalize.npz        ImageProcessor._normalize on small uint16 stacks (several blur / amax)
  patches.npz          ImageProcessor._img2patches (crop_cell + smooth + resize + channel select)
  cellsize.npz         the same for cell_size 15 / 20 / 40 / 45 / 60 (patch resampling)
  vit.npz              reference vit_s / vit_tiny forward + softmax on a few patches (seeded weights)
  mae.npz              MarkerImputer.impute on a few cells (seeded weights)
  merge.json           Annotator.merge_by_voting for every reachable branch
  e2e.npz              Annotator.preprocess() + predict() on a small synthetic image
"""
from __future__ import annotations

import ast
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

from oracle.refshim import load_reference, REFERENCE_ROOT          # noqa: E402
from multiplexed_image_annotator_b200 import synth, weights       # noqa: E402

ref = load_reference()


class _NullLogger:
    def log(self, *_a, **_k):
        pass

    def log_all_hyperparameters(self, *_a, **_k):
        pass


def _processor(blur=0.3, amax=99.8, cell_size=30, infer=True, device="cpu"):
    p = object.__new__(ref.preprocess.ImageProcessor)
    p.blur, p.amax, p.scale, p.infer, p.device, p.n_jobs = blur, amax, cell_size / 30.0, infer, device, 0
    p.logger = _NullLogger()
    return p


# ------------------------------------------------------------------------------------------------
def golden_cells():
    from PIL import Image
    mask = np.array(Image.open(os.path.join(REFERENCE_ROOT, "examples/example_2_cell_mask.png")))
    rows = []
    import re
    with open(os.path.join(REFERENCE_ROOT, "results/test_annotation_1.csv")) as f:
        next(f)
        for line in f:                      # id, type, conf, [rows], [cols], Region k (lists unquoted)
            cid = int(line.split(",", 1)[0])
            lists = re.findall(r"\[[^\]]*\]", line)
            r = ast.literal_eval(lists[0]); c = ast.literal_eval(lists[1])
            rows.append((cid, min(r), max(r), min(c), max(c), sum(r), sum(c), len(r)))
            # the CSV's lists must be exactly this mask's pixels in raster order
            rr, cc = np.nonzero(mask == cid)
            assert rr.tolist() == r and cc.tolist() == c, cid
    arr = np.array(rows, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "cells_example2.npz"), mask=mask, table=arr)
    # second real mask (1850 cells) through the reference's own per-pixel loop
    mask1 = np.array(Image.open(os.path.join(REFERENCE_ROOT, "examples/example_1_cell_mask.png")))
    crop = mask1[100:260, 200:380].astype(np.int32)
    d = _processor()._cell_pos_dict(crop, 0)
    tab = np.array([(k, min(r), max(r), min(c), max(c), sum(r), sum(c), len(r)) for k, (r, c) in d.items()], np.int64)
    np.savez_compressed(os.path.join(OUT, "cells_example1_crop.npz"), mask=crop, table=tab)
    print("cells:", len(rows), "+", len(tab))


def golden_markers():
    cases = {
        "examples": (open(os.path.join(REFERENCE_ROOT, "examples/markers.txt")).read().split("\n"), True),
        "examples_loose": (open(os.path.join(REFERENCE_ROOT, "examples/markers.txt")).read().split("\n"), False),
        "full15": (synth.FULL_PANEL_MARKERS, True),
        "base_minus_cd11c_strict": (["CD45", "CD20", "CD4", "CD8", "DAPI", "CD3"], True),
        "base_minus_cd11c_loose": (["CD45", "CD20", "CD4", "CD8", "DAPI", "CD3"], False),
        "full_minus3_loose": ([m for m in synth.FULL_PANEL_MARKERS if m not in ("CD15", "CD138", "FoxP3")], False),
        "full_minus4_loose": ([m for m in synth.FULL_PANEL_MARKERS if m not in ("CD15", "CD138", "FoxP3", "CD56")], False),
        "aliases": (["DNA", "SMActin", "CD31", "CytoKeratin", "Vimentin", "Ki67", "CD45", "CHGA"], True),
        "alias_truncated": (["DNA", "aSMA", "CD31", "CK", "Vim", "Ki67", "CD45"], False),
        "nerve_only": (["GFAP", "DAPI", "CD45"], True),
        "structure_nerve": (synth.STRUCTURE_MARKERS + ["GFAP"], True),
        "nothing": (["foo", "bar"], True),
        "single": (["DAPI"], False),
    }
    out = {}
    for name, (markers, strict) in cases.items():
        markers = [m for m in markers if m]
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            f.write("\n".join(markers) + "\n")
        p = ref.markerParse.MarkerParser(strict=strict, logger=_NullLogger())
        try:
            p.parse(f.name)
        except TypeError as e:              # a one-line marker file is a 0-d array in the reference
            out[name] = {"markers": markers, "strict": strict, "raises": "TypeError"}
            continue
        finally:
            os.unlink(f.name)
        out[name] = {"markers": markers, "strict": strict, "indices": p.indices,
                     "flags": [p.immune_base, p.immune_extended, p.immune_full, p.struct, p.nerve],
                     "parsed_markers": [str(m) for m in p.markers], "n_markers": p.n_markers}
    json.dump(out, open(os.path.join(OUT, "markers.json"), "w"), indent=1)
    print("markers:", {k: v.get("flags", v.get("raises")) for k, v in out.items()})


def _small_scene(h, w, c, seed, grid=18):
    mask = synth.synth_mask(h, w, grid=grid, seed=seed)
    img = synth.to_uint16(synth.synth_image(mask, c, seed=seed))
    return img, mask.numpy().astype(np.int32)


def golden_normalize():
    img, _ = _small_scene(150, 131, 5, 11)
    img[1] = 0                                   # all-zero channel -> -1
    img[2] = (img[2] // 200).astype(np.uint16)   # dim channel: percentile <= 20, max < 25
    img[3, :, :60] = 0
    big = np.random.default_rng(5).gamma(2.0, 40.0, (2, 40, 300)).astype(np.float32)   # float input, H < radius
    out = {"img": img, "img_f32": big}
    for tag, blur, amax in (("b03_a998", 0.3, 99.8), ("b0_a100", 0, 100), ("b1_a100", 1, 100), ("b04_a95", 0.4, 95.0)):
        out[tag] = _processor()._normalize(img.copy(), blur=blur, amax=amax)
        out["f32_" + tag] = _processor()._normalize(big.copy(), blur=blur, amax=amax)
    np.savez_compressed(os.path.join(OUT, "normalize.npz"), **out)
    print("normalize:", {k: v.shape for k, v in out.items() if k.startswith("b")})


def golden_patches():
    img, mask = _small_scene(120, 140, 9, 21)
    # one real-mask region as well (irregular touching cells)
    from PIL import Image
    real = np.array(Image.open(os.path.join(REFERENCE_ROOT, "examples/example_1_cell_mask.png")))[300:400, 250:360].astype(np.int32)
    proc = _processor()
    norm = proc._normalize(img.copy(), blur=0.3, amax=99.8)
    out = {"img_norm": norm, "mask": mask, "mask_real": real}
    for tag, m, index in (("synth", mask, [4, 1, 2, 3, 0, 5, 6]), ("synth_q3", mask, [0, -1, 2, -1, 4, 5, 8]),
                          ("real", real, [8, 7, 6, 5, 4, 3, 2, 1, 0])):
        image = norm[:, : m.shape[0], : m.shape[1]]
        d = proc._cell_pos_dict(m, 0)
        with tempfile.TemporaryDirectory() as tmp:
            inten = proc._img2patches(image, m, index, d, None, id="g", save_path=tmp, save_tensor=True,
                                      int_full=True, batch_size=10000)
            pt = torch.load(os.path.join(tmp, "g_batch_0.pt")).numpy()
        wins = []
        for cid in d:                                            # utils.py:227-235 window integers
            xm = (min(d[cid][0]) + max(d[cid][0])) // 2
            x0 = int(max(xm - 20.0, 0)); x1 = int(min(x0 + 40, image.shape[1]))
            ym = (min(d[cid][1]) + max(d[cid][1])) // 2
            y0 = int(max(ym - 20.0, 0)); y1 = int(min(y0 + 40, image.shape[2]))
            wins.append((x0, x1, y0, y1))
        keep = slice(0, 48)
        out[tag + "_index"] = np.array(index)
        out[tag + "_ids"] = np.array(list(d.keys()), np.int32)
        out[tag + "_patches"] = pt[keep]
        out[tag + "_intensity"] = inten
        out[tag + "_windows"] = np.array(wins, np.int32)
        # the soft mask alone, straight from utils.smooth, for the first cells
        sm = []
        for cid, (x0, x1, y0, y1) in list(zip(d, wins))[:16]:
            mp = np.zeros((40, 40)); mp[: x1 - x0, : y1 - y0] = m[x0:x1, y0:y1]
            sm.append(ref.utils.smooth(mp, cid))
        out[tag + "_smooth"] = np.array(sm)
    np.savez_compressed(os.path.join(OUT, "patches.npz"), **out)
    print("patches:", {k: v.shape for k, v in out.items() if k.endswith("_patches")})


def golden_cellsize():
    """_img2patches for cell_size != 30 (patch edge int(40 * cell_size / 30), anti-aliased nearest resize)."""
    img, mask = _small_scene(150, 170, 5, 41)
    norm = _processor()._normalize(img.copy(), blur=0.3, amax=99.8)
    out = {"img_norm": norm, "mask": mask}
    for cs in (15, 20, 40, 45, 60):
        proc = _processor(cell_size=cs)
        d = proc._cell_pos_dict(mask, 0)
        with tempfile.TemporaryDirectory() as tmp:
            inten = proc._img2patches(norm, mask, [4, -1, 2, 0], d, None, id="g", save_path=tmp, save_tensor=True,
                                      int_full=True, batch_size=10000)
            pt = torch.load(os.path.join(tmp, "g_batch_0.pt")).numpy()
        out[f"cs{cs}_patches"] = pt[:24]
        out[f"cs{cs}_intensity"] = inten
    np.savez_compressed(os.path.join(OUT, "cellsize.npz"), **out)
    print("cellsize:", {k: v.shape for k, v in out.items() if k.endswith("_patches")})


def _ref_vit(panel, sd):
    ctor = {"immune_base": ref.model.vit_s, "immune_extended": ref.model.vit_m, "immune_full": ref.model.vit_l,
            "structure": ref.model.vit_s, "nerve_cell": ref.model.vit_tiny}[panel]
    s = weights.VIT_SPECS[panel]
    m = ctor(img_size=40, in_chans=s.in_chans, num_classes=len(s.classes), drop_path_rate=0.1, global_pool=False)
    m.load_state_dict(sd)
    return m.eval()


def golden_vit():
    out = {}
    g = torch.Generator().manual_seed(3)
    for panel in ("immune_base", "nerve_cell"):
        s = weights.VIT_SPECS[panel]
        sd = weights.random_vit_state(panel, seed=1)
        x = torch.rand((6, s.in_chans, 40, 40), generator=g) * 2 - 1
        with torch.no_grad():
            logits = _ref_vit(panel, sd)(x)
        out[panel + "_x"] = x.numpy()
        out[panel + "_logits"] = logits.numpy()
        out[panel + "_probs"] = torch.softmax(logits, dim=1).numpy()
    np.savez_compressed(os.path.join(OUT, "vit.npz"), **out)
    print("vit:", {k: v.shape for k, v in out.items() if "logits" in k})


def golden_mae():
    out = {}
    g = torch.Generator().manual_seed(4)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for panel, present in (("immune_base", [0, 1, 2, 3, 4, 6]), ("immune_extended", [0, 1, 2, 4, 5, 6, 7, 9])):
                s = weights.MAE_SPECS[panel]
                sd = weights.random_mae_state(panel, seed=1)
                weights.save_checkpoint(sd, os.path.join(weights.MODEL_DIR, s.ckpt))
                imp = ref.markerImputer.MarkerImputer(present, "cpu", panel)
                x = torch.rand((5, s.channels, 40, 40), generator=g) * 2 - 1
                for k in range(s.channels):
                    if k not in present:
                        x[:, k] = -1
                out[panel + "_x"] = x.numpy().copy()
                out[panel + "_present"] = np.array(present)
                out[panel + "_out"] = imp.impute(x.clone(), 64).numpy()
                os.remove(os.path.join(weights.MODEL_DIR, s.ckpt))
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "mae.npz"), **out)
    print("mae:", {k: v.shape for k, v in out.items() if k.endswith("_out")})


def _bare_annotator(confidence, ctc):
    a = object.__new__(ref.model.Annotator)
    for name in ("annotations", "confidence", "immune_annotations", "struct_annotations", "nerve_annotations",
                 "immune_base_pred", "immune_extended_pred", "immune_full_pred", "struct_pred", "nerve_pred"):
        setattr(a, name, [])
    a.confidence_thresh = confidence
    a.extra_cell_types = False
    a.cell_type_confidence = ctc
    return a


def golden_merge():
    rng = np.random.default_rng(9)
    names = {p: list(s.classes) for p, s in weights.VIT_SPECS.items()}
    base_ctc = {k: -1 for k in ref.model.Annotator.__init__.__code__.co_consts if False} or None
    all_types = ["B cell", "CD4 T cell", "CD8 T cell", "Dendritic cell", "Regulatory T cell", "Granulocyte cell",
                 "Mast cell", "M1 macrophage cell", "M2 macrophage cell", "Natural killer cell", "Plasma cell",
                 "Endothelial cell", "Epithelial cell", "Stroma cell", "Smooth muscle", "Proliferating/tumor cell",
                 "Nerve cell", "Others"]

    def probs(panel, n, temp):
        z = rng.normal(size=(n, len(names[panel]))).astype(np.float32) * temp
        e = np.exp(z - z.max(1, keepdims=True))
        p = (e / e.sum(1, keepdims=True)).astype(np.float32)
        p[0] = 1.0 / p.shape[1]                                  # exact tie across all classes
        p[1, :2] = p[1, :2].mean()                               # tie between two classes
        return p

    cases = []
    combos = [("immune_full",), ("immune_extended",), ("immune_base",), ("structure",), ("nerve_cell",),
              ("immune_full", "structure"), ("immune_base", "structure"), ("immune_extended", "structure", "nerve_cell"),
              ("structure", "nerve_cell"), ("immune_full", "nerve_cell"), ("immune_base", "nerve_cell"),
              ("immune_full", "structure", "nerve_cell")]
    for combo in combos:
        for conf, ctc_over in ((0.3, {}), (0.55, {"B cell": 1, "Proliferating/tumor cell": 1}),
                               (0.2, {"CD4 T cell": 0.0, "Stroma cell": 0.9, "Others": 0.5, "Nerve cell": 0.7})):
            ctc = {k: -1 for k in all_types}
            ctc.update(ctc_over)
            n = 40
            p = {panel: probs(panel, n, 1.5) for panel in combo}
            p[combo[0]][2, 0] = np.float32(conf)                 # probability exactly at the threshold (Q13)
            a = _bare_annotator(conf, ctc)
            for panel in combo:
                lst = [{names[panel][i]: row[i] for i in range(len(row))} for row in p[panel]]
                if panel.startswith("immune"):
                    getattr(a, panel + "_pred").append(lst)
                    a.immune_annotations.append(lst)
                elif panel == "structure":
                    a.struct_pred.append(lst); a.struct_annotations.append(lst)
                else:
                    a.nerve_pred.append(lst); a.nerve_annotations.append(lst)
            rec = {"panels": list(combo), "confidence": conf, "ctc": ctc,
                   "probs": {k: v.tolist() for k, v in p.items()}}
            try:
                a.merge_by_voting()
                rec["labels"] = a.annotations[0]
                rec["conf"] = [float(c) for c in a.confidence[0]]
                rec["conf_is_int"] = [isinstance(c, int) for c in a.confidence[0]]
                a.annotations_copy = a.annotations
                rec["cell_types"] = [str(x) for x in _unique_types(a)]
            except KeyError as e:
                rec["raises"] = "KeyError:" + str(e.args[0])
            cases.append(rec)
    json.dump(cases, open(os.path.join(OUT, "merge.json"), "w"))
    print("merge:", len(cases), "cases;", sum("raises" in c for c in cases), "raise")


def _unique_types(a):
    ct = a._get_unique_cell_types()
    ct = np.delete(ct, np.where(ct == "Others"))
    return np.append(ct, "Others")


def golden_e2e():
    """Full reference run (Annotator.preprocess + predict + export_annotations) on a small image."""
    out = {}
    cwd = os.getcwd()
    for tag, markers, strict, infer, seed in (
            ("full", synth.FULL_PANEL_MARKERS, True, True, 31),
            ("impute", ["CD45", "CD20", "CD4", "CD8", "DAPI", "CD3"], False, True, 32),
            ("struct_nerve", synth.STRUCTURE_MARKERS + ["GFAP"], True, True, 33)):
        img, mask = _small_scene(110, 128, len(markers), seed)
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            try:
                np.save("img.npy", img); np.save("mask.npy", mask)
                synth.write_marker_file("markers.txt", markers)
                with open("images.csv", "w") as f:
                    f.write("image_path,mask_path\nimg.npy,mask.npy\n")
                # seeded + calibrated weights for every model the run will load
                mean_logits = {}
                for panel, s in weights.VIT_SPECS.items():
                    weights.save_checkpoint(weights.random_vit_state(panel, seed=2), os.path.join(weights.MODEL_DIR, s.ckpt))
                for panel, s in weights.MAE_SPECS.items():
                    if tag == "impute" and panel == "immune_base":
                        weights.save_checkpoint(weights.random_mae_state(panel, seed=2), os.path.join(weights.MODEL_DIR, s.ckpt))
                ann = ref.model.Annotator("markers.txt", "images.csv", "cpu", "./", "g", strict, infer, -1, True,
                                          0.3, 99.8, 0.3, 30, None, n_jobs=0)
                ann.preprocess()
                # calibrate the heads on this image's own patches (SURVEY 8d recipe), then predict
                ann.load_models()
                for panel, attr in (("immune_base", "immune_base_model"), ("immune_extended", "immune_extended_model"),
                                    ("immune_full", "immune_full_model"), ("structure", "struct_model"),
                                    ("nerve_cell", "nerve_model")):
                    f = os.path.join("tmp", f"g_0_{panel}_batch_0.pt")
                    if not os.path.exists(f):
                        continue
                    x = torch.load(f)
                    model = getattr(ann, attr)
                    with torch.no_grad():
                        mean_logits[panel] = model(x[:64]).mean(0).numpy()
                    sd = weights.calibrate_head(weights.random_vit_state(panel, seed=2), mean_logits[panel], 20.0)
                    model.load_state_dict(sd)
                    if not (tag == "full" and panel != "immune_full"):      # keep the fixture small
                        out[f"{tag}_patches_{panel}"] = x.numpy()
                ann.predict(32)
                ann.export_annotations()
                out[tag + "_img"] = img; out[tag + "_mask"] = mask
                out[tag + "_markers"] = np.array(markers)
                out[tag + "_labels"] = np.array(ann.annotations[0])
                out[tag + "_conf"] = np.array([float(c) for c in ann.confidence[0]], np.float64)
                out[tag + "_cell_types"] = np.array([str(c) for c in ann.cell_types])
                out[tag + "_intensity"] = np.asarray(ann.preprocessor.intensity_full[0])
                out[tag + "_csv"] = np.array(open("results/g_annotation_0.csv").read())
                for panel, v in mean_logits.items():
                    out[f"{tag}_meanlogits_{panel}"] = v
                for panel, attr in (("immune_base", "immune_base_pred"), ("immune_extended", "immune_extended_pred"),
                                    ("immune_full", "immune_full_pred"), ("structure", "struct_pred"), ("nerve_cell", "nerve_pred")):
                    lst = getattr(ann, attr)
                    if lst:
                        out[f"{tag}_probs_{panel}"] = np.array([[d[k] for k in weights.VIT_SPECS[panel].classes] for d in lst[0]], np.float32)
                ann.logger.close()
            finally:
                os.chdir(cwd)
        print("e2e", tag, "cells", len(out[tag + "_labels"]), dict(zip(*np.unique(out[tag + "_labels"], return_counts=True))))
    np.savez_compressed(os.path.join(OUT, "e2e.npz"), **out)


def golden_c1():
    """BASELINE configs[0]: the reference's own example mask (examples/example_1_cell_mask.png, 1850 cells) and
    marker list (examples/markers.txt: immune_base + immune_extended + structure apply -> vit_m + vit_s, merge branch 2),
    device='cpu'.  examples/example_1.tif is missing from the checkout, so the image is the seeded synthetic 17 x 600 x 600
    stack painted on that mask (synth.synth_image(mask, 17, seed=1)); only the mask and the outputs are stored."""
    from PIL import Image
    mask = np.array(Image.open(os.path.join(REFERENCE_ROOT, "examples/example_1_cell_mask.png"))).astype(np.int32)
    markers = [m for m in open(os.path.join(REFERENCE_ROOT, "examples/markers.txt")).read().split("\n") if m]
    img = synth.to_uint16(synth.synth_image(torch.from_numpy(mask), len(markers), seed=1))
    out = {"mask": mask, "markers": np.array(markers), "image_checksum": np.array([int(img.astype(np.int64).sum())], np.int64)}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            np.save("img.npy", img); np.save("mask.npy", mask)
            synth.write_marker_file("markers.txt", markers)
            with open("images.csv", "w") as f:
                f.write("image_path,mask_path\nimg.npy,mask.npy\n")
            for panel, s in weights.VIT_SPECS.items():
                weights.save_checkpoint(weights.random_vit_state(panel, seed=2), os.path.join(weights.MODEL_DIR, s.ckpt))
            ann = ref.model.Annotator("markers.txt", "images.csv", "cpu", "./", "c1", True, True, -1, True, 0.3, 99.8, 0.3, 30, None, n_jobs=0)
            ann.preprocess()
            ann.load_models()
            for panel, attr in (("immune_extended", "immune_extended_model"), ("structure", "struct_model")):
                x = torch.load(os.path.join("tmp", f"c1_0_{panel}_batch_0.pt"))
                model = getattr(ann, attr)
                with torch.no_grad():
                    ml = model(x[:256]).mean(0).numpy()
                out[f"meanlogits_{panel}"] = ml
                model.load_state_dict(weights.calibrate_head(weights.random_vit_state(panel, seed=2), ml, 20.0))
            ann.predict(128)
            ann.export_annotations()
            out["labels"] = np.array(ann.annotations[0])
            out["conf"] = np.array([float(c) for c in ann.confidence[0]], np.float64)
            out["cell_types"] = np.array([str(c) for c in ann.cell_types])
            out["csv"] = np.array(open("results/c1_annotation_0.csv").read())
            for panel, attr in (("immune_extended", "immune_extended_pred"), ("structure", "struct_pred")):
                out[f"probs_{panel}"] = np.array([[d[k] for k in weights.VIT_SPECS[panel].classes] for d in getattr(ann, attr)[0]], np.float32)
            ann.logger.close()
        finally:
            os.chdir(cwd)
    print("c1 cells", len(out["labels"]), dict(zip(*np.unique(out["labels"], return_counts=True))))
    np.savez_compressed(os.path.join(OUT, "c1.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cells", "markers", "normalize", "patches", "cellsize", "vit", "mae", "merge", "e2e", "c1"]
    for name in which:
        globals()["golden_" + name]()
