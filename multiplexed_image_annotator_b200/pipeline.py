"""In-memory entry point of the hot path: decoded image + label mask in, labels / confidences /
per-type counts out.  This is the call `bench.py` times end to end (host buffers in, host results
out) and the sequence `Annotator.preprocess()` + `Annotator.predict()` runs per image
(reference cta/preprocess.py:241-290, cta/model.py:431-453), without the file and plotting layers.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import exact, ops
from .cell_type_annotation.model import ALL_TYPES, merge_on_device
from .engine import MaeEngine, VitEngine
from .parallel import all_gather_rows, all_reduce_sum, image_owner, shard_range, share_image_results, world


@dataclass
class HotPathResult:
    n_cells: int
    label: torch.Tensor            # uint8 (n,) index into ALL_TYPES
    confidence: torch.Tensor       # float32 (n,), -1 where re-labelled "Others"
    counts: torch.Tensor           # int64 (18,)
    probs: dict = field(default_factory=dict)
    cells: object = None
    margin: torch.Tensor = None    # float32 (n,) decision margins of this rank's cells (after re-evaluation)
    refine: object = None          # exact.RefineStats of this rank's cells
    phases: object = None          # _Phases (run(..., time_phases=True)): device time of every phase of the step

    def names(self):
        return [ALL_TYPES[k] for k in self.label.cpu().tolist()]


class _Phases:
    """CUDA-event marks at the phase boundaries of one HotPath.run (bench.py's per-phase times of the strong-scaling
    record).  mark(name) closes the phase `name`; ms() synchronises and returns {name: milliseconds}."""

    def __init__(self, enabled: bool):
        self.ev = None
        if enabled:
            self.ev = []
            self.mark("start")

    def mark(self, name: str):
        if self.ev is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.ev.append((name, e))

    def ms(self) -> dict:
        if not self.ev:
            return {}
        self.ev[-1][1].synchronize()
        out = {}
        for (_, a), (name, b) in zip(self.ev[:-1], self.ev[1:]):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out


def normalize_over_ranks(image, device, blur, amax, rank, nranks, phases: _Phases | None = None) -> torch.Tensor:
    """Stage 1 of ONE image shared by the ranks: its channels are independent (reference preprocess.py:218-238), so rank r
    uploads and normalises channels c = r (mod N) only and each finished float32 plane is broadcast over NVLink from its
    owner - instead of every rank uploading and normalising the whole stack.  Same kernel on the same data: bit-identical."""
    import torch.distributed as dist
    if isinstance(image, np.ndarray):
        image = torch.from_numpy(image)
    c, h, w = image.shape
    out = torch.empty((c, h, w), dtype=torch.float32, device=device)
    # every rank walks the planes in the same order; the owner normalises its plane right before it posts the broadcast.  The
    # broadcasts are asynchronous (NCCL's own stream, ordered after the kernels queued so far), so the stage-1 kernels of this
    # rank's NEXT own plane run while the finished planes travel: upload, FP64 filter and NVLink traffic overlap.
    pending = []
    for k in range(c):
        if k % nranks == rank:
            plane = image[k:k + 1]
            plane = plane.to(device, non_blocking=True) if not plane.is_cuda else plane.contiguous()
            ops.normalize(plane, blur, amax, out=out[k:k + 1])
        pending.append(dist.broadcast(out[k], src=k % nranks, async_op=True))
    if phases is not None:
        phases.mark("1_upload_normalize_own_channels")
    for work in pending:
        work.wait()                       # stream-level wait: the consumer kernels queue behind the last plane
    if phases is not None:
        phases.mark("1b_plane_broadcasts")
    return out


class HotPath:
    """Stages 1-5 for one image.  `panels` maps panel name -> channel index list (MarkerParser.indices
    restricted to the panels that predict consumes); `models` maps panel -> VitEngine; `imputers`
    maps panel -> (MaeEngine, present positions) for panels with a missing marker."""

    def __init__(self, panels: dict, models: dict, imputers: dict | None = None, *, normalization=True, blur=0.3,
                 amax=99.8, confidence=0.3, cell_type_confidence=None, chunk_cells=4096, device="cuda",
                 shard_cells=True, cell_size=30, shard_stage1=True, exact_labels: int | None = None):
        self.panels, self.models, self.imputers = dict(panels), models, imputers or {}
        self.normalization, self.blur, self.amax = normalization, blur, amax
        self.confidence, self.ctc = confidence, cell_type_confidence
        self.chunk = chunk_cells
        self.device = torch.device(device)
        self.shard_cells = shard_cells
        self.shard_stage1 = shard_stage1          # with shard_cells and > 1 rank: channels of stage 1 split over the ranks
        self.cell_size = cell_size
        self.exact_labels = exact.LEVELS if exact_labels is None else exact_labels     # levels of margin-guarded re-evaluation
        for p in self.panels:
            if not isinstance(models.get(p), VitEngine):
                raise ValueError(f"no classifier engine for panel {p}")

    def _normalize_sharded(self, image, rank, nranks, phases=None):
        return normalize_over_ranks(image, self.device, self.blur, self.amax, rank, nranks, phases)

    def _to_device(self, a):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(a)
        return a if a.is_cuda else a.to(self.device, non_blocking=True)

    @torch.no_grad()
    def run(self, image, mask, to_host: bool = True, keep_probs: bool = False, time_phases: bool = False) -> HotPathResult:
        ph = _Phases(time_phases)
        msk = self._to_device(mask)
        if msk.dtype != torch.int32:
            msk = msk.to(torch.int32)
        host_img = torch.from_numpy(image) if isinstance(image, np.ndarray) else image
        rank, nranks = world() if self.shard_cells else (0, 1)
        if self.normalization and nranks > 1 and self.shard_stage1:
            img = self._normalize_sharded(host_img, rank, nranks, ph)
        elif self.normalization and not host_img.is_cuda:
            img = ops.normalize_from_host(host_img, self.device, self.blur, self.amax)     # upload hidden behind stage 1
        else:
            img = self._to_device(host_img)
            if self.normalization:
                img = ops.normalize(img, self.blur, self.amax)
            elif img.dtype != torch.float32:
                img = img.to(torch.float32)
        ph.mark("1_upload_normalize")
        cells = ops.cell_stats(msk)
        mn = ops.channel_min(img)
        ph.mark("2_cell_stats")
        lo, hi = shard_range(cells.n, rank, nranks)
        names = list(self.panels)
        idx = [self.panels[p] for p in names]
        parts = {p: [] for p in names}
        for a in range(lo, hi, self.chunk):
            b = min(a + self.chunk, hi)
            outs, _, _ = ops.build_patches(img, msk, mn, cells, idx, a, b - a, cell_size=self.cell_size)
            for p, t in zip(names, outs):
                if p in self.imputers:
                    eng, present = self.imputers[p]
                    eng.impute(t, present)
                parts[p].append(self.models[p].forward(t))
        probs = {p: (torch.cat(v) if v else torch.empty((0, len(self.models[p].spec.classes)), device=self.device))
                 for p, v in parts.items()}
        ph.mark("3_4_patches_networks")

        def forward_cells(sel, precision):
            # model inputs of the selected cells of this rank's range, rebuilt (stage 3 is cheap) and run at `precision`
            sub = cells.subset(sel + lo)
            outs, _, _ = ops.build_patches(img, msk, mn, sub, idx, 0, sub.n, cell_size=self.cell_size)
            res = {}
            for p, t in zip(names, outs):
                if p in self.imputers:
                    eng, present = self.imputers[p]
                    eng.impute(t, present, precision=precision if precision in eng.PRECISIONS else "bf16x3")
                res[p] = self.models[p].forward(t, precision=precision)
            return res

        label, conf, counts, margin, stats = exact.refine_labels(
            probs, lambda pr, want_margin=True: merge_on_device(pr, self.confidence, self.ctc, want_margin=want_margin),
            forward_cells, levels=self.exact_labels)
        ph.mark("5_merge_reevaluate")
        if nranks > 1:
            label = all_gather_rows(label, cells.n, lo, hi)
            conf = all_gather_rows(conf, cells.n, lo, hi)
            counts = all_reduce_sum(counts)
            ph.mark("6_gather")
        if to_host:
            label, conf, counts = label.cpu(), conf.cpu(), counts.cpu()     # D2H ends the step (synchronises)
            ph.mark("7_results_to_host")
        return HotPathResult(cells.n, label, conf, counts, probs if keep_probs else {}, cells, margin, stats, ph)

    @torch.no_grad()
    def run_batch(self, items, to_host: bool = True) -> list:
        """A batch of (image, mask) pairs (the batch-processing CSV): image i is annotated whole by rank i % world - no
        collective on the data path - and its owner broadcasts labels / confidences / counts at the end, so every rank
        returns every image's HotPathResult (cells / probs only on the owner).  A rank may pass None for the images it
        does not own (it never touches them)."""
        rank, nranks = world()
        shard_cells, self.shard_cells = self.shard_cells, False
        try:
            own = [self.run(*item, to_host=False) if image_owner(i, nranks) == rank else None for i, item in enumerate(items)]
        finally:
            self.shard_cells = shard_cells
        shared = share_image_results([None if r is None else (r.label, r.confidence, r.counts) for r in own], self.device)
        out = []
        for r, (label, conf, counts) in zip(own, shared):
            if to_host:
                label, conf, counts = label.cpu(), conf.cpu(), counts.cpu()
            out.append(HotPathResult(label.shape[0], label, conf, counts, {} if r is None else r.probs, None if r is None else r.cells))
        return out
