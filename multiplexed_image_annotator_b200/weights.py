"""Model shapes, checkpoint I/O and random-init weights for the RIBCA classifier / imputer zoo.

Checkpoint format is the reference's: `torch.load(path, weights_only=False)["model"]` = a timm
state dict (reference model.py:189-231, markerImputer.py:260-285) with keys
`cls_token, pos_embed, patch_embed.proj.*, blocks.N.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}.*,
norm.*, head.*` (+ `mask_token, decoder_*` for the imputer).  The trained checkpoints are not
available offline (reference download_models.py), so benchmarks and tests use random-init weights
of the same architectures, produced here with plain torch ops.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import numpy as np
import torch

MODEL_DIR = "src/multiplexed_image_annotator/cell_type_annotation/models"     # CWD-relative, as the reference


@dataclass(frozen=True)
class VitSpec:
    name: str           # panel key used by MarkerParser / tmp tensor names
    ckpt: str           # checkpoint file name under MODEL_DIR
    dim: int
    in_chans: int
    classes: tuple
    depth: int = 12
    heads: int = 12
    patch: int = 4
    img: int = 40

    @property
    def tokens(self) -> int:
        return (self.img // self.patch) ** 2 + 1


# class index -> name tables: reference model.py:247-252, 266-270, 284-287, 309-312, 334
VIT_SPECS = {
    "immune_base": VitSpec("immune_base", "immune_base.pth", 288, 7,
                           ("B cell", "CD4 T cell", "CD8 T cell", "Others", "Dendritic cell")),
    "immune_extended": VitSpec("immune_extended", "immune_extended.pth", 384, 10,
                               ("CD4 T cell", "CD8 T cell", "Dendritic cell", "B cell", "M1 macrophage cell",
                                "M2 macrophage cell", "Natural killer cell", "Others")),
    "immune_full": VitSpec("immune_full", "immune_full.pth", 576, 15,
                           ("CD4 T cell", "CD8 T cell", "Dendritic cell", "B cell", "M1 macrophage cell",
                            "M2 macrophage cell", "Regulatory T cell", "Granulocyte cell", "Plasma cell",
                            "Natural killer cell", "Mast cell", "Others")),
    "structure": VitSpec("structure", "struct.pth", 288, 7,
                         ("Stroma cell", "Smooth muscle", "Endothelial cell", "Epithelial cell",
                          "Proliferating/tumor cell", "Others")),
    "nerve_cell": VitSpec("nerve_cell", "nerve.pth", 144, 3, ("Nerve cell", "Others")),
}


@dataclass(frozen=True)
class MaeSpec:
    name: str
    ckpt: str
    grid: tuple          # (rows, cols) of 40x40 channel tiles; reference markerImputer.py:262-274
    enc_dim: int = 768
    enc_depth: int = 12
    enc_heads: int = 12
    dec_dim: int = 512
    dec_depth: int = 8
    dec_heads: int = 8

    @property
    def channels(self) -> int:
        return self.grid[0] * self.grid[1]


MAE_SPECS = {
    "immune_base": MaeSpec("immune_base", "immune_base_impute.pth", (1, 7)),
    "immune_extended": MaeSpec("immune_extended", "immune_extended_impute.pth", (2, 5)),
    "immune_full": MaeSpec("immune_full", "immune_full_impute.pth", (3, 5)),
}


def _trunc(gen, shape, std=0.02):
    t = torch.empty(shape)
    return torch.nn.init.trunc_normal_(t, std=std, a=-2 * std, b=2 * std, generator=gen)


def _block(sd, prefix, dim, gen):
    sd[f"{prefix}.norm1.weight"] = torch.ones(dim) + 0.05 * torch.randn(dim, generator=gen)
    sd[f"{prefix}.norm1.bias"] = 0.05 * torch.randn(dim, generator=gen)
    sd[f"{prefix}.attn.qkv.weight"] = _trunc(gen, (3 * dim, dim), 0.05)
    sd[f"{prefix}.attn.qkv.bias"] = 0.02 * torch.randn(3 * dim, generator=gen)
    sd[f"{prefix}.attn.proj.weight"] = _trunc(gen, (dim, dim), 0.03)
    sd[f"{prefix}.attn.proj.bias"] = 0.02 * torch.randn(dim, generator=gen)
    sd[f"{prefix}.norm2.weight"] = torch.ones(dim) + 0.05 * torch.randn(dim, generator=gen)
    sd[f"{prefix}.norm2.bias"] = 0.05 * torch.randn(dim, generator=gen)
    sd[f"{prefix}.mlp.fc1.weight"] = _trunc(gen, (4 * dim, dim), 0.03)
    sd[f"{prefix}.mlp.fc1.bias"] = 0.02 * torch.randn(4 * dim, generator=gen)
    sd[f"{prefix}.mlp.fc2.weight"] = _trunc(gen, (dim, 4 * dim), 0.03)
    sd[f"{prefix}.mlp.fc2.bias"] = 0.02 * torch.randn(dim, generator=gen)


def random_vit_state(panel: str, seed: int = 0) -> dict:
    """Random-init timm-format state dict of the panel's classifier (non-zero biases and
    non-unit LayerNorm affine so that every term of the forward is exercised by parity tests)."""
    s = VIT_SPECS[panel]
    g = torch.Generator().manual_seed(seed * 1000 + s.dim + s.in_chans)
    sd = {}
    sd["cls_token"] = 0.02 * torch.randn((1, 1, s.dim), generator=g)
    sd["pos_embed"] = _trunc(g, (1, s.tokens, s.dim), 0.02)
    k = s.in_chans * s.patch * s.patch
    sd["patch_embed.proj.weight"] = (torch.rand((s.dim, s.in_chans, s.patch, s.patch), generator=g) * 2 - 1) / math.sqrt(k)
    sd["patch_embed.proj.bias"] = (torch.rand(s.dim, generator=g) * 2 - 1) / math.sqrt(k)
    for i in range(s.depth):
        _block(sd, f"blocks.{i}", s.dim, g)
    sd["norm.weight"] = torch.ones(s.dim) + 0.05 * torch.randn(s.dim, generator=g)
    sd["norm.bias"] = 0.05 * torch.randn(s.dim, generator=g)
    sd["head.weight"] = _trunc(g, (len(s.classes), s.dim), 0.02)
    sd["head.bias"] = torch.zeros(len(s.classes))
    return sd


def calibrate_head(sd: dict, mean_logits, gain: float = 20.0) -> dict:
    """SURVEY section 8(d) weights recipe: plain random init labels >=99% of cells with one class, so
    spread the label histogram by centring the logits on a calibration batch and scaling the head:
    head.weight *= gain, head.bias = gain * (head.bias - mean_logits)."""
    out = dict(sd)
    m = torch.as_tensor(np.asarray(mean_logits), dtype=torch.float32)
    out["head.bias"] = gain * (sd["head.bias"] - m)
    out["head.weight"] = gain * sd["head.weight"]
    return out


def sincos_table(dim: int, grid) -> torch.Tensor:
    """Fixed 2-D sin-cos positional table with a zero cls row, (1, 1 + gh*gw, dim): the values a
    trained imputer checkpoint carries in pos_embed / decoder_pos_embed
    (reference markerImputer.py:11-65: first half encodes the column index, second half the row)."""
    gh, gw = grid
    rows = torch.arange(gh, dtype=torch.float32).view(gh, 1).expand(gh, gw).reshape(-1)
    cols = torch.arange(gw, dtype=torch.float32).view(1, gw).expand(gh, gw).reshape(-1)
    quarter = dim // 4
    omega = 1.0 / (10000.0 ** (torch.arange(quarter, dtype=torch.float32) / quarter))
    parts = []
    for pos in (cols, rows):
        ang = pos[:, None] * omega[None, :]
        parts += [ang.sin(), ang.cos()]
    tab = torch.cat(parts, dim=1)
    return torch.cat([torch.zeros(1, dim), tab], dim=0).unsqueeze(0)


def random_mae_state(panel: str, seed: int = 0) -> dict:
    s = MAE_SPECS[panel]
    g = torch.Generator().manual_seed(seed * 1000 + 7 * s.channels)
    sd = {}
    sd["cls_token"] = 0.02 * torch.randn((1, 1, s.enc_dim), generator=g)
    sd["pos_embed"] = sincos_table(s.enc_dim, s.grid)
    sd["patch_embed.proj.weight"] = (torch.rand((s.enc_dim, 1, 40, 40), generator=g) * 2 - 1) / 40.0
    sd["patch_embed.proj.bias"] = (torch.rand(s.enc_dim, generator=g) * 2 - 1) / 40.0
    for i in range(s.enc_depth):
        _block(sd, f"blocks.{i}", s.enc_dim, g)
    sd["norm.weight"] = torch.ones(s.enc_dim) + 0.05 * torch.randn(s.enc_dim, generator=g)
    sd["norm.bias"] = 0.05 * torch.randn(s.enc_dim, generator=g)
    sd["decoder_embed.weight"] = _trunc(g, (s.dec_dim, s.enc_dim), 0.03)
    sd["decoder_embed.bias"] = 0.02 * torch.randn(s.dec_dim, generator=g)
    sd["mask_token"] = 0.02 * torch.randn((1, 1, s.dec_dim), generator=g)
    sd["decoder_pos_embed"] = sincos_table(s.dec_dim, s.grid)
    for i in range(s.dec_depth):
        _block(sd, f"decoder_blocks.{i}", s.dec_dim, g)
    sd["decoder_norm.weight"] = torch.ones(s.dec_dim) + 0.05 * torch.randn(s.dec_dim, generator=g)
    sd["decoder_norm.bias"] = 0.05 * torch.randn(s.dec_dim, generator=g)
    sd["decoder_pred.weight"] = _trunc(g, (1600, s.dec_dim), 0.03)
    sd["decoder_pred.bias"] = 0.02 * torch.randn(1600, generator=g)
    return sd


def save_checkpoint(sd: dict, path: str) -> None:
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    torch.save({"model": sd}, path)


def load_checkpoint(path: str) -> dict:
    """reference model.py:191: torch.load(path, weights_only=False)["model"]."""
    return torch.load(path, map_location="cpu", weights_only=False)["model"]
