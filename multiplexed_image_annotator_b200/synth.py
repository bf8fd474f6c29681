"""Synthetic tissue images, label masks and marker lists for tests and benchmarks.

The reference ships no runnable inputs offline (both example TIFFs are missing,
reference .MISSING_LARGE_BLOBS:1-2), so every measured workload is generated here, following the
recipe of SURVEY.md section 8(d): disc cells on a jittered square grid, Gamma(2, 40) background plus
per-cell per-channel Bernoulli(0.3) * U(500, 4000) signal, clipped to uint16.
Everything is vectorised torch so the 20000 x 20000 configuration can be generated on the GPU.
"""
from __future__ import annotations

import numpy as np
import torch

FULL_PANEL_MARKERS = ["DAPI", "CD3", "CD4", "CD8", "CD11c", "CD15", "CD20", "CD45", "CD56", "CD68",
                      "CD138", "CD163", "FoxP3", "Granzyme B", "Trypase"]
BASE_PANEL_MARKERS = ["CD45", "CD20", "CD4", "CD8", "DAPI", "CD11c", "CD3"]
STRUCTURE_MARKERS = ["DAPI", "aSMA", "CD31", "PanCK", "Vimentin", "Ki67", "CD45"]
NERVE_MARKERS = ["DAPI", "CD45", "GFAP"]


def synth_mask(height: int, width: int, grid: int = 18, seed: int = 2, device="cpu",
               r_lo: int = 5, r_hi: int = 8, jitter: int = 2) -> torch.Tensor:
    """int32 (H, W) label mask: one disc per grid square, ids 1..N in raster order of squares."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    gh, gw = (height + grid - 1) // grid, (width + grid - 1) // grid
    cy = (torch.arange(gh) * grid + grid // 2).view(gh, 1) + torch.randint(-jitter, jitter + 1, (gh, gw), generator=g)
    cx = (torch.arange(gw) * grid + grid // 2).view(1, gw) + torch.randint(-jitter, jitter + 1, (gh, gw), generator=g)
    rad = torch.randint(r_lo, r_hi + 1, (gh, gw), generator=g)
    cy, cx, rad = cy.to(device), cx.to(device), rad.to(device)
    ys = torch.arange(height, device=device)
    xs = torch.arange(width, device=device)
    sq_y = (ys // grid).view(-1, 1).expand(height, width)
    sq_x = (xs // grid).view(1, -1).expand(height, width)
    dy = ys.view(-1, 1) - cy[sq_y, sq_x]
    dx = xs.view(1, -1) - cx[sq_y, sq_x]
    inside = dy * dy + dx * dx <= rad[sq_y, sq_x] ** 2
    ids = (sq_y * gw + sq_x + 1).to(torch.int32)
    return torch.where(inside, ids, torch.zeros_like(ids))


def synth_image(mask: torch.Tensor, n_channels: int, seed: int = 2, p_on: float = 0.3,
                amp_lo: float = 500.0, amp_hi: float = 4000.0, bg_scale: float = 40.0) -> torch.Tensor:
    """(C, H, W) image on mask.device, values in uint16 range stored as int32 (see `to_uint16`)."""
    dev = mask.device
    g = torch.Generator(device=dev).manual_seed(seed)
    h, w = mask.shape
    n_ids = int(mask.max().item()) + 1
    on = torch.rand((n_ids, n_channels), generator=g, device=dev) < p_on
    amp = torch.rand((n_ids, n_channels), generator=g, device=dev) * (amp_hi - amp_lo) + amp_lo
    table = torch.where(on, amp, torch.zeros_like(amp))
    table[0] = 0
    out = torch.empty((n_channels, h, w), dtype=torch.int32, device=dev)
    idx = mask.long()
    for c in range(n_channels):
        u = torch.rand((2, h, w), generator=g, device=dev).clamp_min_(1e-12)
        bg = -bg_scale * (u[0].log() + u[1].log())             # Gamma(2, bg_scale)
        out[c] = (bg + table[:, c][idx]).clamp_(0, 65535).to(torch.int32)
    return out


def to_uint16(img: torch.Tensor) -> np.ndarray:
    """Host uint16 numpy view of a synth_image result."""
    return img.cpu().numpy().astype(np.uint16)


def write_marker_file(path: str, markers) -> str:
    with open(path, "w") as f:
        f.write("\n".join(markers) + "\n")
    return path
