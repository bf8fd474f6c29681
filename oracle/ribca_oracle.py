"""TEST INFRASTRUCTURE - CPU oracle for RIBCA's per-cell annotation hot path.

A numpy / scipy / plain-torch restatement of the reference algorithm (sun-huangqingbo/
multiplexed-image-annotator), each function citing the reference file:line it follows
(paths relative to the reference root; `cta/` = src/multiplexed_image_annotator/cell_type_annotation/).
Only `tests/` (incl. its GPU check scripts), `__graft_entry__.smoke()` and `bench.py`'s baseline leg (`cpu_baseline`, which also
runs these torch modules in eager fp32 on the GPU as the checker of the full population, and `--impl reference`) may import this
module; the product package and the `tools_*` scripts never do (the product fails loudly without its CUDA library).

Pinning (see tests/golden/make_golden.py, tests/test_oracle_golden.py):
  * cell statistics are pinned by the reference's own golden vector
    results/test_annotation_1.csv <-> examples/example_2_cell_mask.png (582 cells);
  * every stage is additionally pinned against outputs of the UNMODIFIED reference .py files run
    in the build container through oracle/refshim.py (fixtures under tests/golden/).
PARITY UNPINNED for the third-party numerics the reference delegates to scikit-image and timm
(soft-mask Gaussians, resize, ViT/MAE blocks): those packages are absent offline, so both the
reference run and this oracle use the restatements in oracle/standins.py.
"""
from __future__ import annotations

import math
from functools import partial

import numpy as np
import scipy.ndimage as ndi
import torch
import torch.nn as nn

from . import standins

# ----------------------------------------------------------------------------------------------
# a1. panel matching                                             (cta/markerParse.py:4-117)
# ----------------------------------------------------------------------------------------------
PANELS = {
    "immune_base": ["CD45", "CD20", "CD4", "CD8", "DAPI", "CD11c", "CD3"],
    "immune_extended": ["DAPI", "CD3", "CD4", "CD8", "CD11c", "CD20", "CD45", "CD68", "CD163", "CD56"],
    "immune_full": ["DAPI", "CD3", "CD4", "CD8", "CD11c", "CD15", "CD20", "CD45", "CD56", "CD68",
                    "CD138", "CD163", "FoxP3", "Granzyme B", "Trypase"],
    "structure": ["DAPI", "aSMA", "CD31", "PanCK", "Vimentin", "Ki67", "CD45"],
    "nerve_cell": ["DAPI", "CD45", "GFAP"],
}
MISSING_BUDGET = {"immune_base": 1, "immune_extended": 2, "immune_full": 3, "structure": 1, "nerve_cell": 0}
ALIASES = {"DNA": "DAPI", "DPAI-02": "DAPI", "CD16": "CD15", "CD38": "CD138", "CD79": "CD20",
           "CHGA": "GFAP", "SMActin": "aSMA", "CD3e": "CD3", "CK": "PanCK", "CytoKeratin": "PanCK",
           "Cytokeratin": "PanCK", "Cytokeratin-19": "PanCK", "panCK": "PanCK"}


def parse_markers(marker_file: str, strict: bool = True) -> dict:
    """markerParse.py:62-117.  Returns {panel: list[int] | None}.

    Keeps the fixed-width numpy string array of np.loadtxt, so an alias longer than the widest
    marker in the file is truncated on assignment (markerParse.py:64,80-82).
    """
    arr = np.loadtxt(marker_file, delimiter=",", dtype=str)
    if arr.ndim == 0:
        raise TypeError("iteration over a 0-d array")      # markerParse.py:67 on a one-line file
    for i in range(len(arr)):
        if arr[i] in ALIASES and ALIASES[arr[i]] not in arr:
            arr[i] = ALIASES[arr[i]]
    names = list(arr)
    out = {}
    for panel, wanted in PANELS.items():
        matched, n_missing, ok = [], 0, True
        for m in wanted:
            if m in names:
                matched.append(names.index(m))
            elif not strict and len(wanted) > 3:
                matched.append(-1)
                n_missing += 1
                if n_missing > MISSING_BUDGET[panel]:
                    ok = False
                    break
            else:
                ok = False
                break
        # the reference tests `if matched:` (markerParse.py:98) - an empty list is also "not applied"
        out[panel] = matched if (ok and matched) else None
    return out


# ----------------------------------------------------------------------------------------------
# a3. normalisation                                              (cta/preprocess.py:214-239)
# ----------------------------------------------------------------------------------------------
def normalize(img: np.ndarray, blur=0, amax=100) -> np.ndarray:
    """preprocess.py:214-239 - per channel: sigma-20 background subtraction (bg clamped to 125),
    optional Gaussian blur, upper-percentile clip when the percentile exceeds 20, then
    2*x/max(25, max) - 1; a channel with no positive pixel becomes -1."""
    out = img.astype(np.float32)
    for c in range(out.shape[0]):
        ch = out[c]
        bg = ndi.gaussian_filter(ch, sigma=20)
        bg = np.where(bg > 125, 125, bg)
        ch = np.clip(ch - bg, 0, None)
        if blur:
            ch = ndi.gaussian_filter(ch, sigma=blur)
        if not (ch > 0).any():
            out[c] = -1
            continue
        t = np.percentile(ch, amax)
        if t > 20:
            ch = np.clip(ch, 0, t)
        out[c] = 2 * (ch / max(25, np.max(ch))) - 1
    return out


# ----------------------------------------------------------------------------------------------
# a4. per-cell statistics                   (cta/preprocess.py:159-211, cta/utils.py:272-290)
# ----------------------------------------------------------------------------------------------
def cell_pos_dict(mask: np.ndarray) -> dict:
    """preprocess.py:166-181 restated with a stable sort: {id: (rows, cols)} in raster order,
    ids ascending, 0 = background.  Same content as the reference's per-pixel Python loop."""
    flat = mask.ravel()
    order = np.argsort(flat, kind="stable")
    ids, start = np.unique(flat[order], return_index=True)
    bounds = list(start[1:]) + [flat.size]
    w = mask.shape[1]
    out = {}
    for cid, s, e in zip(ids, start, bounds):
        if cid == 0:
            continue
        px = order[s:e]
        out[mask.dtype.type(cid)] = ((px // w).tolist(), (px % w).tolist())
    return out


def cell_stats(mask: np.ndarray) -> dict:
    """Integer summary of cell_pos_dict that every consumer derives (utils.py:227,232 bbox =
    min/max of the lists; model.py:785-786 centroid = np.mean of the lists; area = len).
    Returns ids (ascending), bbox [rmin, rmax, cmin, cmax], sum_r, sum_c, count."""
    m = np.asarray(mask)
    rr, cc = np.nonzero(m)
    lab = m[rr, cc].astype(np.int64)
    ids, inv = np.unique(lab, return_inverse=True)
    n = len(ids)
    rmin = np.full(n, np.iinfo(np.int32).max, np.int64); rmax = np.full(n, -1, np.int64)
    cmin = rmin.copy(); cmax = rmax.copy()
    np.minimum.at(rmin, inv, rr); np.maximum.at(rmax, inv, rr)
    np.minimum.at(cmin, inv, cc); np.maximum.at(cmax, inv, cc)
    return {
        "ids": ids.astype(np.int32),
        "bbox": np.stack([rmin, rmax, cmin, cmax], 1).astype(np.int32),
        "sum_r": np.bincount(inv, rr, n).astype(np.int64),      # float64 weights: exact below 2^53
        "sum_c": np.bincount(inv, cc, n).astype(np.int64),
        "count": np.bincount(inv, minlength=n).astype(np.int32),
    }


def centroids(stats: dict) -> np.ndarray:
    """model.py:785-786: np.mean(list of ints) = exact integer sum / count in float64."""
    n = stats["count"].astype(np.float64)
    return np.stack([stats["sum_r"].astype(np.float64) / n, stats["sum_c"].astype(np.float64) / n], 1)


# ----------------------------------------------------------------------------------------------
# a6/a7. crop window, soft mask, patch                        (cta/utils.py:226-270)
# ----------------------------------------------------------------------------------------------
def crop_window(bbox_row, patch: int, h: int, w: int):
    """utils.py:227-235: bbox-centre anchored window, clamped at 0, truncated at the far border."""
    rmin, rmax, cmin, cmax = (int(v) for v in bbox_row)
    xm = (rmin + rmax) // 2
    x0 = int(max(xm - patch / 2, 0))
    x1 = int(min(x0 + patch, h))
    ym = (cmin + cmax) // 2
    y0 = int(max(ym - patch / 2, 0))
    y1 = int(min(y0 + patch, w))
    return x0, x1, y0, y1


def soft_mask(mask_patch: np.ndarray, cid) -> np.ndarray:
    """utils.py:255-270 `smooth`: cell indicator + 4 disk dilations + 6 Gaussians of the
    dilations (sigma 1 | 1,2 | 1,2,3), float32 running sum, /11, / max(s + 1e-6)."""
    m = mask_patch == cid
    s = m.astype(np.float32)
    terms = 1
    for j in range(1, 5):
        d = standins.dilation(m, standins.disk(j))
        s += d.astype(np.float32)
        terms += 1
        for i in range(j - 1):
            s += standins.gaussian(d, sigma=1 + i)
            terms += 1
    s /= terms
    s /= np.max(s + 1e-6)
    return s


def crop_cell(img_zero: np.ndarray, mask: np.ndarray, min_val: np.ndarray, bbox_row, cid, patch: int):
    """utils.py:226-253: zero-padded window copy, multiply by the soft mask, add the channel
    minimum back (float64), per-channel mean over every labelled pixel of the window."""
    c, h, w = img_zero.shape
    x0, x1, y0, y1 = crop_window(bbox_row, patch, h, w)
    iz = np.zeros((c, patch, patch))
    mp = np.zeros((patch, patch))
    iz[:, : x1 - x0, : y1 - y0] = img_zero[:, x0:x1, y0:y1]
    mp[: x1 - x0, : y1 - y0] = mask[x0:x1, y0:y1]
    marker = iz * soft_mask(mp, cid) + min_val
    sel = mp > 0
    avg = np.array([np.mean(marker[k][sel]) for k in range(c)])
    return marker, avg, (x0, x1, y0, y1)


def select_channels(patch: np.ndarray, channel_index) -> np.ndarray:
    """preprocess.py:110-120 incl. quirk Q3: only the FIRST -1 becomes a -1-filled plane; any
    further -1 stays in the index list and numpy-selects the last image channel."""
    idx = list(channel_index)
    if -1 in idx:
        k = idx.index(-1)
        rest = np.delete(np.asarray(idx), k)
        sel = patch[rest]
        return np.concatenate((sel[:k], -np.ones_like(sel[0:1]), sel[k:]), axis=0)
    return patch[np.asarray(idx)]


def build_patches(image: np.ndarray, mask: np.ndarray, channel_index, stats: dict | None = None,
                  cell_size: float = 30, cells=None):
    """preprocess.py:76-151 `_img2patches` without the disk spill: returns
    (patches float32 (N, C_panel, 40, 40), intensity (N, C_img) float64 in [0,1], windows (N,4))."""
    stats = stats or cell_stats(mask)
    min_val = np.min(image, axis=(1, 2), keepdims=True)          # preprocess.py:153-157
    img_zero = image - min_val
    p = int(40 * (cell_size / 30.0))                              # preprocess.py:67,78
    sel = range(len(stats["ids"])) if cells is None else cells
    out = np.zeros((len(sel), len(channel_index), 40, 40))
    inten = np.zeros((len(sel), image.shape[0]))
    wins = np.zeros((len(sel), 4), np.int32)
    for j, k in enumerate(sel):
        marker, avg, win = crop_cell(img_zero, mask, min_val, stats["bbox"][k], stats["ids"][k], p)
        marker = standins.resize(marker, (marker.shape[0], 40, 40), order=0, anti_aliasing=True)
        out[j] = select_channels(marker, channel_index)
        inten[j] = avg
        wins[j] = win
    return torch.tensor(out, dtype=torch.float32).numpy(), (inten + 1) / 2, wins


# ----------------------------------------------------------------------------------------------
# a10. classifier zoo                                           (cta/model.py:31-88,188-239)
# ----------------------------------------------------------------------------------------------
class _RefViT(standins.VisionTransformer):
    """model.py:31-64 with global_pool=False: output = head(norm(x)[:, 0])."""

    def __init__(self, **kw):
        super().__init__(**kw)
        self.global_pool = False

    def forward_features(self, x):
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1) + self.pos_embed
        x = self.blocks(x)
        return self.norm(x)[:, 0]


VIT_ZOO = {   # name: (embed_dim, in_chans, classes)       model.py:66-88,190-230
    "immune_base": (288, 7, 5),
    "immune_extended": (384, 10, 8),
    "immune_full": (576, 15, 12),
    "structure": (288, 7, 6),
    "nerve_cell": (144, 3, 2),
}
CLASS_NAMES = {   # model.py:247-252,266-270,284-287,309-312,334
    "immune_full": ["CD4 T cell", "CD8 T cell", "Dendritic cell", "B cell", "M1 macrophage cell",
                    "M2 macrophage cell", "Regulatory T cell", "Granulocyte cell", "Plasma cell",
                    "Natural killer cell", "Mast cell", "Others"],
    "immune_extended": ["CD4 T cell", "CD8 T cell", "Dendritic cell", "B cell", "M1 macrophage cell",
                        "M2 macrophage cell", "Natural killer cell", "Others"],
    "immune_base": ["B cell", "CD4 T cell", "CD8 T cell", "Others", "Dendritic cell"],
    "structure": ["Stroma cell", "Smooth muscle", "Endothelial cell", "Epithelial cell",
                  "Proliferating/tumor cell", "Others"],
    "nerve_cell": ["Nerve cell", "Others"],
}
VOTE_ORDER = ["CD4 T cell", "CD8 T cell", "Dendritic cell", "B cell", "M1 macrophage cell",    # utils.py:143-146
              "M2 macrophage cell", "Regulatory T cell", "Granulocyte cell", "Plasma cell",
              "Natural killer cell", "Mast cell", "Stroma cell", "Smooth muscle", "Endothelial cell",
              "Epithelial cell", "Proliferating/tumor cell", "Nerve cell"]


def make_vit(panel: str) -> nn.Module:
    d, c, k = VIT_ZOO[panel]
    return _RefViT(img_size=40, patch_size=4, in_chans=c, num_classes=k, embed_dim=d, depth=12,
                   num_heads=12, mlp_ratio=4, qkv_bias=True,
                   norm_layer=partial(nn.LayerNorm, eps=1e-6)).eval()


@torch.no_grad()
def vit_probs(model: nn.Module, patches, bs: int = 128) -> np.ndarray:
    """model.py:397-406: forward in `bs` slices, softmax(dim=1), float32."""
    x = torch.as_tensor(patches, dtype=torch.float32)
    out = [torch.softmax(model(x[i:i + bs]), dim=1) for i in range(0, len(x), bs)]
    return torch.cat(out).numpy() if out else np.zeros((0, model.num_classes), np.float32)


# ----------------------------------------------------------------------------------------------
# a9. marker imputer                                            (cta/markerImputer.py:69-329)
# ----------------------------------------------------------------------------------------------
MAE_GRID = {"immune_base": (1, 7), "immune_extended": (2, 5), "immune_full": (3, 5)}   # :262-274


def sincos_2d(dim: int, grid_hw) -> np.ndarray:
    """markerImputer.py:11-65: fixed 2-D sin-cos table with a zero cls row, (1 + gh*gw, dim)."""
    gh, gw = grid_hw
    ys, xs = np.meshgrid(np.arange(gh, dtype=np.float32), np.arange(gw, dtype=np.float32), indexing="ij")

    def one(d, pos):
        omega = np.arange(d // 2, dtype=np.float32)
        omega /= d / 2.0
        omega = 1.0 / 10000 ** omega
        o = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(o), np.cos(o)], axis=1)

    # the reference meshgrids with w first and encodes grid[0] (= x) in the first half
    emb = np.concatenate([one(dim // 2, xs), one(dim // 2, ys)], axis=1)
    return np.concatenate([np.zeros([1, dim]), emb], axis=0)


class _RefMAE(nn.Module):
    """markerImputer.py:69-255 at the sizes of :280-284 (enc 768x12x12h, dec 512x8x8h, one token
    per 40x40 channel tile).  The constant 0.1/0.8 noise makes `random_masking` a fixed compaction:
    kept tokens = present channels in ascending order (tie order is irrelevant: attention is
    permutation-equivariant and the positional term is added before the gather)."""

    def __init__(self, grid_hw):
        super().__init__()
        gh, gw = grid_hw
        norm = partial(nn.LayerNorm, eps=1e-6)
        self.grid = grid_hw
        self.patch_embed = standins.PatchEmbed((40 * gh, 40 * gw), 40, 1, 768)
        n = gh * gw
        self.cls_token = nn.Parameter(torch.zeros(1, 1, 768))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, 768), requires_grad=False)
        self.blocks = nn.ModuleList([standins.Block(768, 12, 4, qkv_bias=True, norm_layer=norm) for _ in range(12)])
        self.norm = norm(768)
        self.decoder_embed = nn.Linear(768, 512, bias=True)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, 512))
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, n + 1, 512), requires_grad=False)
        self.decoder_blocks = nn.ModuleList([standins.Block(512, 8, 4, qkv_bias=True, norm_layer=norm) for _ in range(8)])
        self.decoder_norm = norm(512)
        self.decoder_pred = nn.Linear(512, 1600, bias=True)

    def forward(self, tiles, present):
        """tiles (B, L, 1600) row-major 40x40 channel tiles; present = sorted kept positions."""
        w = self.patch_embed.proj.weight.reshape(768, 1600)
        x = tiles @ w.t() + self.patch_embed.proj.bias + self.pos_embed[:, 1:, :]
        x = x[:, present, :]
        x = torch.cat(((self.cls_token + self.pos_embed[:, :1, :]).expand(x.shape[0], -1, -1), x), dim=1)
        for blk in self.blocks:
            x = blk(x)
        x = self.decoder_embed(self.norm(x))
        full = self.mask_token.repeat(x.shape[0], tiles.shape[1], 1).clone()
        full[:, present, :] = x[:, 1:, :]
        x = torch.cat([x[:, :1, :], full], dim=1) + self.decoder_pos_embed
        for blk in self.decoder_blocks:
            x = blk(x)
        return self.decoder_pred(self.decoder_norm(x))[:, 1:, :]


def make_mae(panel: str) -> nn.Module:
    return _RefMAE(MAE_GRID[panel]).eval()


@torch.no_grad()
def impute(model: nn.Module, patches: np.ndarray, present, bs: int = 64) -> np.ndarray:
    """markerImputer.py:294-329: present channels are returned unchanged, missing ones are
    replaced by the decoder's prediction for that tile."""
    x = torch.as_tensor(patches, dtype=torch.float32).clone()
    n, c = x.shape[:2]
    present = sorted(int(p) for p in present)
    missing = [k for k in range(c) if k not in present]
    for i in range(0, n, bs):
        tiles = x[i:i + bs].reshape(-1, c, 1600)
        pred = model(tiles, present)
        for k in missing:
            x[i:i + bs, k] = pred[:, k].reshape(-1, 40, 40)
    return x.numpy()


# ----------------------------------------------------------------------------------------------
# a12. vote merge / threshold                          (cta/model.py:481-636, utils.py:143-146)
# ----------------------------------------------------------------------------------------------
ALL_TYPES = ["B cell", "CD4 T cell", "CD8 T cell", "Dendritic cell", "Regulatory T cell",    # model.py:97-99
             "Granulocyte cell", "Mast cell", "M1 macrophage cell", "M2 macrophage cell",
             "Natural killer cell", "Plasma cell", "Endothelial cell", "Epithelial cell", "Stroma cell",
             "Smooth muscle", "Proliferating/tumor cell", "Nerve cell", "Others"]


def merge_by_voting(preds: dict, confidence: float, cell_type_confidence: dict | None = None):
    """model.py:481-636 for one image.  `preds` maps panel name -> (N, classes) float32 softmax
    output for the models that ran ("immune_*" at most one, "structure", "nerve_cell").
    Returns (list[str] labels, list conf) where conf is np.float32 or the int -1, as the reference.

    Branch order of the reference's elif chain: full+struct+nerve (raises KeyError 'Others', Q1) >
    immune+struct > struct+nerve > immune+nerve > single immune > single struct > single nerve.
    """
    ctc = cell_type_confidence or {k: -1 for k in ALL_TYPES}
    immune = next((p for p in ("immune_full", "immune_extended", "immune_base") if p in preds), None)
    has_s, has_n = "structure" in preds, "nerve_cell" in preds
    if immune == "immune_full" and has_s and has_n:
        raise KeyError("Others")                      # model.py:488-491: vote dict has no 'Others'
    if immune and has_s:
        used = [immune, "structure"]
    elif has_s and has_n:
        used = ["structure", "nerve_cell"]
    elif immune and has_n:
        used = [immune, "nerve_cell"]
    elif immune:
        used = [immune]
    elif has_s:
        used = ["structure"]
    elif has_n:
        used = ["nerve_cell"]
    else:
        raise ValueError("No predictions to merge")
    n = len(preds[used[0]])
    labels, conf = [], []
    if len(used) == 1:
        names = CLASS_NAMES[used[0]]
        p = preds[used[0]]
        for j in range(n):
            k = int(np.argmax(p[j]))                  # first maximum in class-index order (Q8)
            best = names[k]
            thr = ctc[best] if ctc[best] > 0 else confidence
            if best != "Others" and p[j][k] < thr:
                labels.append("Others"); conf.append(-1)
            else:
                labels.append(best); conf.append(p[j][k])
        return labels, conf
    for j in range(n):
        vote = {k: 0 for k in VOTE_ORDER}
        others = []
        for panel in used:
            names = CLASS_NAMES[panel]
            for k, name in enumerate(names):
                if name == "Others":
                    others.append(preds[panel][j][k])
                else:
                    vote[name] += preds[panel][j][k]
        best = max(vote, key=vote.get)                # first maximum in VOTE_ORDER (Q8)
        thr = min(*others, confidence) if ctc[best] < 0 else ctc[best]
        if vote[best] < thr:
            labels.append("Others"); conf.append(-1)
        else:
            labels.append(best); conf.append(vote[best])
    return labels, conf


def unique_cell_types(all_labels) -> np.ndarray:
    """model.py:455-458,678-686: sorted unique labels, 'Others' moved last."""
    names = set()
    for lab in all_labels:
        names.update(lab)
    arr = np.sort(np.array(list(names)))
    arr = np.delete(arr, np.where(arr == "Others"))
    return np.append(arr, "Others")


# ----------------------------------------------------------------------------------------------
# a2/a11. sequencing for one image                 (cta/preprocess.py:241-290, model.py:431-453)
# ----------------------------------------------------------------------------------------------
def predicted_panels(indices: dict) -> list:
    """model.py:246-349: full elif extended elif base, then structure, then nerve."""
    out = []
    for p in ("immune_full", "immune_extended", "immune_base"):
        if indices.get(p):
            out.append(p)
            break
    for p in ("structure", "nerve_cell"):
        if indices.get(p):
            out.append(p)
    return out


def annotate_image(image, mask, indices: dict, models: dict, imputers: dict | None = None, *,
                   normalization=True, blur=0.3, amax=99.8, confidence=0.3, cell_type_confidence=None,
                   infer=True, bs=128, cell_size=30):
    """End-to-end oracle for one image: preprocess.py:241-290 then model.py:431-453.
    Only the panels that `predict` consumes are cropped (the reference also crops the unused
    immune panels and deletes them, Q5).  Returns a dict of every intermediate."""
    img = normalize(image, blur, amax) if normalization else image
    stats = cell_stats(mask)
    res = {"image": img, "stats": stats, "patches": {}, "probs": {}}
    for panel in predicted_panels(indices):
        idx = indices[panel]
        pt, inten, wins = build_patches(img, mask, idx, stats, cell_size)
        if infer and -1 in idx and panel.startswith("immune"):            # preprocess.py:268
            pt = impute(imputers[panel], pt, [i for i, v in enumerate(idx) if v != -1])
        res["patches"][panel] = pt
        res["windows"] = wins
        res.setdefault("intensity", inten)     # avg_int covers all image channels: panel-independent
        res["probs"][panel] = vit_probs(models[panel], pt, bs)
    res["labels"], res["confidence"] = merge_by_voting(res["probs"], confidence, cell_type_confidence)
    return res
