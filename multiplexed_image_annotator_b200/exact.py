"""Exact labels by construction: margin-guarded re-evaluation of the cells near a decision boundary.

The reference runs its networks in fp32 (cta/model.py:397-406) and labels every cell by an argmax plus a threshold
compare (cta/model.py:481-636).  The tensor-core forward of this build is faster than fp32 and within 1e-3 of its
probabilities, but a cell whose decision margin is smaller than that error could come out with another label.  So the
labels are made independent of the fast path's rounding:

  level 0  every cell: default precision (f16f8, measured max |dprob| 1.7e-4); stage 5 also returns the decision margin
           of every cell (include/ribca_b200.h: ribca_merge_votes)
  level 1  cells with margin < EPS1 (default 1e-3 = the probability tolerance itself, > 5x the level-0 error bound):
           patches rebuilt for those cells only, forward in bf16x3 (three tensor-core passes, max |dprob| 7.7e-5)
  level 2  cells whose bf16x3 margin is still < EPS2 (default 3e-4, ~4x the level-1 bound): plain fp32 on the FP32 pipe
           (csrc/stage4_fp32.cu) - the reference's own arithmetic

A cell that is not re-evaluated has a margin of more than twice the error bound of the level that decided it, so exact
arithmetic gives the same (label, re-labelled?) outcome; the cells that reach level 2 are decided by fp32 like the
reference.  What remains is only what no fp32 implementation decides reproducibly: cells whose margin is below fp32
rounding (~1e-6).  The re-evaluated cells' probabilities replace the fast ones, so the reported confidences improve too.
The error bounds are MEASURED on random-init networks (the trained checkpoints are not available offline), so the scheme checks its
own premise on every run: the re-evaluated cells are a sample on which both precisions are known, and `observed_error[lvl]` = max
|dprob| between level `lvl` and the level above it over that sample.  If the fast pass is seen to be off by more than EPS1 / 2 (5e-4;
e.g. a checkpoint with activations outside the fp16 / e4m3 ranges of the f16f8 planes), a RuntimeWarning is raised and EVERY cell is
re-evaluated at level 1 - the result then has bf16x3 accuracy, at bf16x3 cost (`RIBCA_EXACT_GUARD=0` turns the escalation off).
Per-cell results do not depend on which other cells share a batch (every GEMM row accumulates on its own in a fixed
order), hence the refined labels are identical for any sharding of the cells over GPUs.
"""
from __future__ import annotations

import os
import warnings
from dataclasses import dataclass, field

import torch

EPS1 = float(os.environ.get("RIBCA_EXACT_EPS1", 1e-3))
EPS2 = float(os.environ.get("RIBCA_EXACT_EPS2", 3e-4))
LEVELS = int(os.environ.get("RIBCA_EXACT_LABELS", 2))          # 0 = off, 1 = bf16x3 only, 2 = bf16x3 then fp32
LEVEL_PRECISION = ("bf16x3", "fp32")
GUARD = os.environ.get("RIBCA_EXACT_GUARD", "1") != "0"          # escalate when the observed level-0 error breaks the premise


@dataclass
class RefineStats:
    cells: int = 0
    reevaluated: list = field(default_factory=lambda: [0, 0])       # cells sent to level 1 / level 2
    relabelled: list = field(default_factory=lambda: [0, 0])        # of those, how many changed label
    eps: tuple = (EPS1, EPS2)
    events: list = field(default_factory=list)                      # (level, start, end) CUDA events of the re-evaluations
    observed_error: list = field(default_factory=lambda: [0.0, 0.0])  # max |dprob| level k vs level k + 1 on the re-evaluated cells
    escalated: bool = False                                         # the guard sent every cell to level 1

    def ms(self):
        """Device time of the two re-evaluation levels (synchronises on their events)."""
        out = [0.0, 0.0]
        for lvl, a, b in self.events:
            b.synchronize()
            out[lvl] += a.elapsed_time(b)
        return out

    def add(self, other: "RefineStats"):
        self.cells += other.cells
        for k in range(2):
            self.reevaluated[k] += other.reevaluated[k]
            self.relabelled[k] += other.relabelled[k]

    def as_dict(self):
        return {"cells": self.cells, "eps": list(self.eps), "level1_bf16x3_cells": self.reevaluated[0],
                "level2_fp32_cells": self.reevaluated[1], "level1_relabelled": self.relabelled[0],
                "level2_relabelled": self.relabelled[1], "level_ms": self.ms(),
                "observed_error_level0_level1": list(self.observed_error), "escalated": self.escalated}


@torch.no_grad()
def refine_labels(probs: dict, merge, forward_cells, levels: int | None = None, eps=None, chunk: int = 2048,
                  guard: bool | None = None):
    """probs: {panel: (n, classes) float32 CUDA tensor} of the fast pass - updated IN PLACE for the re-evaluated cells.
    merge(probs_dict, want_margin=True) -> (label, conf, counts, margin) is stage 5 on the device.
    forward_cells(idx int64 CUDA tensor, precision) -> {panel: (len(idx), classes)} rebuilds the model inputs of those
    cells and runs every voting model at `precision`.
    Returns (label uint8 (n,), conf float32 (n,), counts int64 (18,), margin float32 (n,), RefineStats)."""
    levels = LEVELS if levels is None else levels
    eps = (EPS1, EPS2) if eps is None else eps
    guard = GUARD if guard is None else guard
    label, conf, counts, margin = merge(probs, want_margin=True)
    stats = RefineStats(cells=int(label.shape[0]), eps=tuple(eps))

    def reevaluate(idx, lvl):
        """cells idx at LEVEL_PRECISION[lvl]: probabilities, labels, confidences, margins replaced; -> max |dprob| seen"""
        worst = torch.zeros((), dtype=torch.float32, device=label.device)
        for a in range(0, idx.numel(), chunk):
            part = idx[a:a + chunk]
            new = forward_cells(part, LEVEL_PRECISION[lvl])
            for p, t in new.items():
                worst = torch.maximum(worst, (t - probs[p][part]).abs().max())
                probs[p][part] = t
            l2, c2, _, m2 = merge({p: probs[p][part] for p in probs}, want_margin=True)
            label[part], conf[part], margin[part] = l2, c2, m2
        return worst

    for lvl in range(min(levels, 2)):
        idx = torch.nonzero(margin < eps[lvl]).flatten()           # one host sync per level: the count sizes the batch
        if idx.numel() == 0:
            break
        stats.reevaluated[lvl] = int(idx.numel())
        before = label[idx].clone()
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        worst = reevaluate(idx, lvl)
        stats.observed_error[lvl] = float(worst.item())
        if lvl == 0 and guard and stats.observed_error[0] > 0.5 * eps[0]:
            # the premise "fast-pass error << EPS1" does not hold for these weights / inputs: every cell goes to level 1
            warnings.warn(f"ribca exact labels: the {stats.reevaluated[0]} re-evaluated cells show a fast-pass error of "
                          f"{stats.observed_error[0]:.2e} > EPS1 / 2 = {0.5 * eps[0]:.1e}; re-evaluating all {stats.cells} cells in "
                          f"{LEVEL_PRECISION[0]} (set RIBCA_PRECISION=bf16x3 to make that the fast pass)", RuntimeWarning)
            done = torch.zeros(stats.cells, dtype=torch.bool, device=label.device)
            done[idx] = True
            rest = torch.nonzero(~done).flatten()
            before = torch.cat([before, label[rest].clone()])
            idx = torch.cat([idx, rest])
            stats.observed_error[0] = max(stats.observed_error[0], float(reevaluate(rest, 0).item()))
            stats.reevaluated[0] = int(idx.numel())
            stats.escalated = True
        ev[1].record()
        stats.events.append((lvl, ev[0], ev[1]))
        stats.relabelled[lvl] = int((label[idx] != before).sum().item())
    if stats.relabelled[0] or stats.relabelled[1]:
        counts = torch.bincount(label.long(), minlength=18).to(torch.int64)
    return label, conf, counts, margin, stats
