"""`ImageProcessor`: stages 1-3 of the hot path on the device, behind the reference's interface
(cta/preprocess.py:25-290: constructor arguments, `transform()`, and the public attributes
`masks`, `cell_pos_dict`, `intensity_full` that downstream reference code reads).

What changed under the interface: the image stack is normalised by ribca_normalize, the mask is
reduced to per-cell bbox / sums / area by ribca_cell_stats, and the 40x40 model inputs are built by
ribca_build_patches straight into device batches - they are handed to `Annotator.predict` in HBM
instead of being spilled to `<main_dir>/tmp/*.pt` (the tmp directory is still created and wiped so
`clear_tmp()` keeps working).  Only the panels `predict` consumes are cropped (the reference also
crops the unused immune panels and deletes them afterwards, quirk Q5).
"""
from __future__ import annotations

import os
from collections.abc import Mapping

import numpy as np
import pandas as pd
import torch

from .. import ops
from ..io import is_npy, is_tiff, iter_npy_planes, iter_tiff_planes, read_image, read_mask
from ..parallel import shard_range, world
from .markerImputer import MarkerImputer

PATCH_CACHE_BYTES = int(os.environ.get("RIBCA_PATCH_CACHE_BYTES", 48 << 30))      # all images of a batch together
IMAGE_CACHE_BYTES = int(os.environ.get("RIBCA_IMAGE_CACHE_BYTES", 64 << 30))      # resident normalised stacks + masks
CHUNK_CELLS = int(os.environ.get("RIBCA_CHUNK_CELLS", 8192))


class CellPositions(Mapping):
    """`cell_pos_dict[i]` of the reference ({id: (rows, cols)} in raster order, ids ascending,
    cta/preprocess.py:159-181) backed by the device cell table: keys, bbox, centroid and area come
    from the table; the pixel lists are CSR arrays built on the device by ribca_cell_pixels the first time a
    cell's lists are asked for (one kernel + one D2H copy for the whole image), then sliced per cell."""

    def __init__(self, mask_host: np.ndarray, ids: np.ndarray, bbox: np.ndarray, sums: np.ndarray, count: np.ndarray,
                 csr_source=None):
        self._mask, self.ids, self.bbox, self.sums, self.count = mask_host, ids, bbox, sums, count
        self._index = None
        self._csr_source = csr_source      # () -> (offsets, rows, cols) device tensors, or None (host window scan)
        self._csr = None

    def csr(self):
        """(offsets int64 (n+1,), rows int32, cols int32) numpy arrays of every cell's pixel list."""
        if self._csr is None:
            if self._csr_source is not None:
                self._csr = tuple(t.cpu().numpy() for t in self._csr_source())
            else:
                off = np.zeros(len(self.ids) + 1, dtype=np.int64)
                np.cumsum(self.count, out=off[1:])
                rr, cc = np.nonzero(self._mask)                                 # raster order
                order = np.argsort(self._mask[rr, cc], kind="stable")            # group by label, raster order kept
                self._csr = (off, rr[order].astype(np.int32), cc[order].astype(np.int32))
        return self._csr

    def __len__(self):
        return len(self.ids)

    def __iter__(self):
        return iter(self._mask.dtype.type(i) for i in self.ids)

    def _row(self, key):
        if self._index is None:
            self._index = {int(v): k for k, v in enumerate(self.ids)}
        return self._index[int(key)]

    def __contains__(self, key):
        try:
            self._row(key)
            return True
        except (KeyError, TypeError, ValueError):
            return False

    def __getitem__(self, key):
        k = self._row(key)
        off, rows, cols = self.csr()
        return rows[off[k]:off[k + 1]].tolist(), cols[off[k]:off[k + 1]].tolist()

    def centroid(self, key):
        """(mean row, mean col) = np.mean of the lists, without materialising them."""
        k = self._row(key)
        return float(self.sums[k, 0]) / float(self.count[k]), float(self.sums[k, 1]) / float(self.count[k])


class ImageProcessor(object):
    def __init__(self, csv_path, parser, main_path, device, batch_id='', infer=True, normalization=True, blur=0,
                 amax=100, cell_size=30, logger=None, n_jobs=0) -> None:
        df = pd.read_csv(csv_path)
        self.image_paths = df['image_path']
        self.mask_paths = df['mask_path']
        assert len(self.image_paths) == len(self.mask_paths)
        self.logger = logger
        self._n_images = len(self.image_paths)
        self.logger.log("Number of images: {}.".format(self._n_images))
        self.main_dir = main_path
        self.save_path = os.path.join(self.main_dir, "tmp")
        self.batch_id = batch_id
        self.normalization = normalization
        self.blur = blur
        self.amax = amax
        self.parser = parser
        self.cell_pos_dict = []
        self.intensity_full = []
        os.makedirs(self.save_path, exist_ok=True)
        if world()[0] == 0:                      # ranks sharing one main_dir: rank 0 owns the filesystem side effects
            for f in os.listdir(self.save_path):
                fp = os.path.join(self.save_path, f)
                try:
                    if os.path.isfile(fp):
                        os.remove(fp)
                except FileNotFoundError:
                    pass
        self.infer = infer
        self.masks = []
        if str(device).startswith("cpu"):
            raise RuntimeError("the B200 build has no CPU path: pass device='cuda' (the reference's CPU path is the oracle)")
        self.device = torch.device(device if str(device) != "cuda" else f"cuda:{torch.cuda.current_device()}")
        self.cell_size = cell_size
        self.scale = cell_size / 30.0
        if not 8 <= int(40 * self.scale) <= 80:
            raise ValueError("cell_size must give a patch edge int(40 * cell_size / 30) in [8, 80] (cell_size 6..60)")
        self.n_jobs = n_jobs
        # device-side state per image
        self.images_dev, self.masks_dev, self.cells, self.min_val = [], [], [], []
        self.patches = []          # per image: {panel: tensor} for this rank's cell range, or None (streamed)
        self.cell_range = []       # per image: (lo, hi) cells owned by this rank
        self._resident = []        # per image: stack + mask stay on the device (inside IMAGE_CACHE_BYTES)
        self._patch_bytes = 0      # bytes of cached patches over all images
        self._image_bytes = 0      # bytes of resident stacks + masks over all images
        self._imputers = {}
        self.logger.log("\n")
        self.logger.log("Starting image processing...")

    # ---- panels ------------------------------------------------------------------------------------
    def predicted_panels(self):
        """Panels `Annotator.predict` consumes: full > extended > base, then structure, nerve
        (reference model.py:246-349)."""
        idx = self.parser.indices
        out = [p for p in ("immune_full", "immune_extended", "immune_base") if idx.get(p)][:1]
        out += [p for p in ("structure", "nerve_cell") if idx.get(p)]
        return out

    def _imputer_for(self, panel):
        """preprocess.py:267-281: imputation only for immune panels with a missing marker and infer=True."""
        index = self.parser.indices[panel]
        if not self.infer or -1 not in index or panel == "structure" or panel == "nerve":
            return None
        if panel not in self._imputers:
            present = [i for i, x in enumerate(index) if x != -1]
            self._imputers[panel] = MarkerImputer(present, self.device, panel)
            print("Imputer for {} is created".format(panel))
            msg = "Imputer for {} is created. Marker(s) ".format(panel)
            for k, ch in enumerate(index):
                if ch == -1:
                    msg += "{} ".format(self.parser.panels[panel][k])
            self.logger.log(msg + "are imputed.")
        return self._imputers[panel]

    # ---- reference-named stage entry points (device tensors in, device tensors out) -----------------
    def _normalize(self, img, blur=0, amax=100):
        """cta/preprocess.py:214-239 on the device.  Accepts a numpy array or a CUDA tensor."""
        t = torch.from_numpy(np.ascontiguousarray(img)).to(self.device) if isinstance(img, np.ndarray) else img
        return ops.normalize(t.contiguous(), blur, amax)

    def _cell_pos_dict(self, mask, n_jobs=0):
        """cta/preprocess.py:159-211: returns the lazy mapping (n_jobs is accepted and ignored)."""
        host = np.ascontiguousarray(mask.cpu().numpy() if isinstance(mask, torch.Tensor) else mask)
        dev = torch.from_numpy(host.astype(np.int32)).to(self.device)
        tab = ops.cell_stats(dev)
        return self._positions(host, tab, dev)

    @staticmethod
    def _positions(mask_host, tab, mask_dev=None):
        """mask_dev: the device mask, or a callable returning it (so that a released mask is uploaded again on demand)."""
        get = mask_dev if callable(mask_dev) else (lambda: mask_dev)
        src = (lambda: ops.cell_pixels(get(), tab)) if mask_dev is not None else None
        return CellPositions(mask_host, tab.ids.cpu().numpy(), tab.bbox.cpu().numpy(), tab.sums.cpu().numpy(),
                             tab.count.cpu().numpy(), src)

    def _move_image_range(self, image):
        """cta/preprocess.py:153-157."""
        mn = ops.channel_min(image)
        return mn.view(-1, 1, 1), image - mn.view(-1, 1, 1)

    def patch_chunks(self, i, panels=None, chunk=None, want_intensity=False):
        """Yield (lo, hi, {panel: (n, C_p, 40, 40) float32 cuda}, avg_int or None) over this rank's
        cells of image i - `_img2patches` (cta/preprocess.py:76-151) without the disk spill."""
        panels = self.predicted_panels() if panels is None else panels
        chunk = chunk or CHUNK_CELLS
        lo, hi = self.cell_range[i]
        idx = [self.parser.indices[p] for p in panels]
        for a in range(lo, hi, chunk):
            b = min(a + chunk, hi)
            outs, avg, _ = ops.build_patches(self.image_dev(i), self.mask_dev(i), self.min_val[i], self.cells[i], idx,
                                             a, b - a, want_intensity=want_intensity, cell_size=self.cell_size)
            batch = dict(zip(panels, outs))
            for p in panels:
                imp = self._imputer_for(p)
                if imp is not None:
                    imp.impute(batch[p], 64)
            yield a, b, batch, avg

    def patches_of_cells(self, i, sel, panels, precision=None):
        """Model inputs of the cells `sel` (int64 device tensor, indices into this rank's range of image i) for
        `panels`, imputed at `precision`: what exact.refine_labels re-evaluates."""
        lo, _ = self.cell_range[i]
        sub = self.cells[i].subset(sel + lo)
        cached = self.patches[i]
        rebuilt = [p for p in panels if cached is None or self._imputer_for(p) is not None]     # imputed inputs start from the raw crop
        batch = {p: cached[p][sel] for p in panels if p not in rebuilt}
        if rebuilt:
            idx = [self.parser.indices[p] for p in rebuilt]
            outs, _, _ = ops.build_patches(self.image_dev(i), self.mask_dev(i), self.min_val[i], sub, idx, 0, sub.n,
                                           cell_size=self.cell_size)
            batch.update(zip(rebuilt, outs))
        for p in rebuilt:
            imp = self._imputer_for(p)
            if imp is not None:
                imp.impute(batch[p], 64, precision=precision)
        return batch

    # ---- device residency of a batch (the reference bounds memory by spilling to tmp/*.pt, preprocess.py:132-135) -------
    def _load_image(self, i):
        """Decode image i and run stage 1: the normalised float32 stack on the device."""
        import struct
        rank, nranks = world()
        image_path = self.image_paths[i]
        img_dev = None
        if self.normalization and nranks == 1 and (is_tiff(image_path) or is_npy(image_path)):
            try:                                          # decode, upload and stage 1 overlapped plane by plane
                planes = iter_tiff_planes(image_path) if is_tiff(image_path) else iter_npy_planes(image_path)
                img_dev = ops.normalize_from_planes(planes, self.device, self.blur, self.amax)
            except (ValueError, KeyError, struct.error, OSError):
                img_dev = None                            # a file flavour the plane readers do not take: whole-file decode below
        if img_dev is not None:
            return img_dev
        image = read_image(image_path)
        if self.normalization and nranks > 1:
            from ..pipeline import normalize_over_ranks          # channels split over the ranks, planes broadcast
            return normalize_over_ranks(image, self.device, self.blur, self.amax, rank, nranks)
        if self.normalization:
            return ops.normalize_from_host(torch.from_numpy(image), self.device, self.blur, self.amax)
        img_dev = torch.from_numpy(image).to(self.device, non_blocking=True)
        return img_dev if img_dev.dtype == torch.float32 else img_dev.to(torch.float32)

    def image_dev(self, i):
        """Normalised stack of image i on the device; an image that was released to stay inside RIBCA_IMAGE_CACHE_BYTES is
        decoded and normalised again (same kernels, same bits) and stays until `release(i)`."""
        if self.images_dev[i] is None:
            self.images_dev[i] = self._load_image(i)
        return self.images_dev[i]

    def mask_dev(self, i):
        if self.masks_dev[i] is None:
            self.masks_dev[i] = torch.from_numpy(self.masks[i]).to(self.device, non_blocking=True)
        return self.masks_dev[i]

    def release(self, i):
        """Drop image i's device stack and mask if it is not one of the images the cache budget keeps resident."""
        if not self._resident[i]:
            self.images_dev[i] = None
            self.masks_dev[i] = None

    # ---- the reference's driver ----------------------------------------------------------------------
    def transform(self):
        """cta/preprocess.py:241-290.  Device memory of a batch is bounded: the patch cache is ONE budget over all images
        (RIBCA_PATCH_CACHE_BYTES, default 48 GiB; images that do not fit are cropped again chunk by chunk in predict) and so
        are the resident normalised stacks + masks (RIBCA_IMAGE_CACHE_BYTES, default 64 GiB; an image beyond it is released
        after its statistics are taken and re-read from its file when predict needs it) - peak memory is one image plus
        the two budgets, however long the CSV is."""
        rank, nranks = world()
        for i, (image_path, mask_path) in enumerate(zip(self.image_paths, self.mask_paths)):
            mask = read_mask(mask_path)                       # 2-D, int32 (preprocess.py:246-250)
            mask_dev = torch.from_numpy(mask).to(self.device, non_blocking=True)
            img_dev = self._load_image(i)
            self.masks.append(mask)
            tab = ops.cell_stats(mask_dev)
            self.images_dev.append(img_dev)
            self.masks_dev.append(mask_dev)
            self.cells.append(tab)
            self.min_val.append(ops.channel_min(img_dev))
            self.cell_pos_dict.append(self._positions(mask, tab, lambda i=i: self.mask_dev(i)))
            self.cell_range.append(shard_range(tab.n, rank, nranks))
            panels = self.predicted_panels()
            lo, hi = self.cell_range[i]
            per_cell = sum(len(self.parser.indices[p]) for p in panels) * 40 * 40 * 4
            keep = self._patch_bytes + (hi - lo) * per_cell <= PATCH_CACHE_BYTES
            inten = torch.empty((hi - lo, img_dev.shape[0]), dtype=torch.float64, device=self.device)
            kept = {p: [] for p in panels}
            for a, b, batch, avg in self.patch_chunks(i, panels if keep else [], want_intensity=True):
                inten[a - lo:b - lo] = avg
                for p in batch:
                    kept[p].append(batch[p])
            self.patches.append({p: torch.cat(v) if v else torch.empty((0, len(self.parser.indices[p]), 40, 40), device=self.device)
                                 for p, v in kept.items()} if keep else None)
            if keep:
                self._patch_bytes += (hi - lo) * per_cell
            # preprocess.py:144-149,284-285: (avg_int + 1) / 2, all image channels, first applied panel
            self.intensity_full.append(self._gather_rows((inten + 1) / 2, tab.n, lo, hi).cpu().numpy())
            img_bytes = img_dev.numel() * 4 + mask_dev.numel() * 4
            self._resident.append(self._image_bytes + img_bytes <= IMAGE_CACHE_BYTES)
            if self._resident[i]:
                self._image_bytes += img_bytes
            del img_dev, mask_dev
            self.release(i)

    @staticmethod
    def _gather_rows(local, n_total, lo, hi):
        from ..parallel import all_gather_rows
        return all_gather_rows(local, n_total, lo, hi)
