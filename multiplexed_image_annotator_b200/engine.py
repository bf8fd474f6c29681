"""Device-resident classifier / imputer engines over ribca_vit_forward and ribca_mae_impute.

A timm-format state dict (reference checkpoint layout, model.py:191, markerImputer.py:261,285) is
packed once into two device blobs - fp32 parameters and the GEMM weights as two 16-bit operand planes
(bf16 {hi, lo}, or fp16 + e4m3 pairs for the default "f16f8" precision; csrc/common.cuh) - plus a
C descriptor of element offsets (include/ribca_b200.h: ribca_vit_desc / ribca_mae_desc).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib, ops
from .weights import MAE_SPECS, VIT_SPECS, MaeSpec, VitSpec


class _Packer:
    def __init__(self):
        self.f32, self.f32_off = [], 0
        self.mat, self.mat_off = [], 0
        self.folds = []          # (matrix offset, rows, cols, gamma): matrices that take W * diag(gamma) in the tensor-core packings

    @staticmethod
    def _pad(t, mult=64):
        n = t.numel()
        pad = (-n) % mult
        flat = t.reshape(-1).to(torch.float32)
        return torch.cat([flat, flat.new_zeros(pad)]) if pad else flat

    def add_f32(self, t) -> int:
        off = self.f32_off
        p = self._pad(t)
        self.f32.append(p)
        self.f32_off += p.numel()
        return off

    def add_mat(self, t) -> int:
        off = self.mat_off
        p = self._pad(t)
        self.mat.append(p)
        self.mat_off += p.numel()
        return off

    def finish(self, device):
        """-> (fp32 parameter blob on the device, fp32 GEMM matrices on the HOST: packed per format on demand)."""
        return torch.cat(self.f32).to(device), torch.cat(self.mat)


def pack_matrices(mats_host: torch.Tensor, fmt: int, device, folds=()):
    """GEMM matrices -> (device operand blob, w_log2_scale) in `fmt`: two 16-bit planes (2, total) for the tensor-core
    formats, the plain fp32 blob for RIBCA_PLANES_F32 (the FP32-pipe re-evaluation path).  `folds`: matrices that are packed
    as W * diag(gamma) in the tensor-core formats (LayerNorm folded into its consumer GEMM, ribca_ln_fold)."""
    if fmt == ops.FMT_F32:
        return mats_host.to(device), 0
    if folds:
        mats_host = mats_host.clone()
        for off, rows, cols, gamma in folds:
            mats_host[off:off + rows * cols].view(rows, cols).mul_(gamma[None, :])
    mats = mats_host.to(device)
    if fmt == ops.FMT_BF16:
        return ops.split_bf16(mats), 0
    t = ops.weight_log2_scale(float(mats.abs().max().item()))
    return ops.split_planes(mats, fmt, w_role=True, log2_scale=t), t + 8


class _Packs:
    """The packed weights of one network per operand format, built lazily: the default precision at construction, the
    higher-precision formats the first time a re-evaluation asks for them (pipeline.refine_labels)."""

    def __init__(self, desc, mats_host, plane_elems, device, folds=(), fold_mask: int = 0):
        self._desc, self._mats, self._plane, self._device, self._folds = desc, mats_host, plane_elems, device, tuple(folds)
        self._fold_mask = fold_mask
        self._by_fmt = {}

    def get(self, precision: str):
        """-> (descriptor for this format, operand blob)."""
        fmt = ops.plane_format(precision)
        if fmt not in self._by_fmt:
            d = type(self._desc)()
            C.memmove(C.byref(d), C.byref(self._desc), C.sizeof(d))
            blob, d.w_log2_scale = pack_matrices(self._mats, fmt, self._device, self._folds)
            d.plane_format, d.split_plane = fmt, self._plane
            if hasattr(d, "ln_folded"):
                d.ln_folded = self._fold_mask if (self._folds and fmt != ops.FMT_F32) else 0
            self._by_fmt[fmt] = (d, blob)
        return self._by_fmt[fmt]


def _pad_heads(t: torch.Tensor, heads: int) -> torch.Tensor:
    """qkv weight (3D, D) / bias (3D,) -> head-padded (3*heads*hdp, ...) with zero rows, hdp = ceil16(hd)."""
    d3 = t.shape[0]
    hd = d3 // (3 * heads)
    hdp = (hd + 15) // 16 * 16
    if hdp == hd:
        return t
    v = t.reshape(3, heads, hd, *t.shape[1:])
    out = v.new_zeros((3, heads, hdp) + tuple(t.shape[1:]))
    out[:, :, :hd] = v
    return out.reshape(3 * heads * hdp, *t.shape[1:])


def _fold_vectors(w, b, gamma, beta):
    """LayerNorm(x; gamma, beta) @ w.T + b = rstd * (x @ (w * gamma).T - mean * c1) + c2 -> (c1, c2), fp64 sums."""
    wg = (w * gamma[None, :]).to(torch.float32)                 # the matrix that is packed (pack_matrices applies the same product)
    c1 = wg.double().sum(1).to(torch.float32)
    c2 = (b.double() + w.double() @ beta.double()).to(torch.float32)
    return c1, c2


def _pack_block(pk: _Packer, sd, prefix: str, desc: _lib.BlockDesc, heads: int, fold: int = 0):
    desc.ln1_g = pk.add_f32(sd[f"{prefix}.norm1.weight"]); desc.ln1_b = pk.add_f32(sd[f"{prefix}.norm1.bias"])
    desc.ln2_g = pk.add_f32(sd[f"{prefix}.norm2.weight"]); desc.ln2_b = pk.add_f32(sd[f"{prefix}.norm2.bias"])
    desc.qkv_b = pk.add_f32(_pad_heads(sd[f"{prefix}.attn.qkv.bias"], heads)); desc.proj_b = pk.add_f32(sd[f"{prefix}.attn.proj.bias"])
    desc.fc1_b = pk.add_f32(sd[f"{prefix}.mlp.fc1.bias"]); desc.fc2_b = pk.add_f32(sd[f"{prefix}.mlp.fc2.bias"])
    desc.qkv_w = pk.add_mat(_pad_heads(sd[f"{prefix}.attn.qkv.weight"], heads)); desc.proj_w = pk.add_mat(sd[f"{prefix}.attn.proj.weight"])
    desc.fc1_w = pk.add_mat(sd[f"{prefix}.mlp.fc1.weight"]); desc.fc2_w = pk.add_mat(sd[f"{prefix}.mlp.fc2.weight"])
    g1, b1, g2, b2 = (sd[f"{prefix}.norm{i}.{n}"] for i in (1, 2) for n in ("weight", "bias"))
    dim = g1.numel()
    if fold & 1:
        c1, c2 = _fold_vectors(sd[f"{prefix}.attn.qkv.weight"], sd[f"{prefix}.attn.qkv.bias"], g1, b1)
        desc.qkv_c1 = pk.add_f32(_pad_heads(c1, heads)); desc.qkv_c2 = pk.add_f32(_pad_heads(c2, heads))
        qkv_rows = _pad_heads(sd[f"{prefix}.attn.qkv.weight"], heads).shape[0]
        pk.folds.append((desc.qkv_w, qkv_rows, dim, g1.to(torch.float32)))
    if fold & 2:
        c1, c2 = _fold_vectors(sd[f"{prefix}.mlp.fc1.weight"], sd[f"{prefix}.mlp.fc1.bias"], g2, b2)
        desc.fc1_c1 = pk.add_f32(c1); desc.fc1_c2 = pk.add_f32(c2)
        pk.folds.append((desc.fc1_w, sd[f"{prefix}.mlp.fc1.weight"].shape[0], dim, g2.to(torch.float32)))


class _Workspace:
    """Grow-only device scratch shared by the engines of one process."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = None
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self.buf


_WS = _Workspace()


class VitEngine:
    """One panel's ViT classifier (reference model.py:31-88) resident on a CUDA device."""

    def __init__(self, spec: VitSpec | str, state_dict: dict, device="cuda", precision: str = ops.DEFAULT_PRECISION,
                 max_cells_per_call: int = 4096):
        self.spec = VIT_SPECS[spec] if isinstance(spec, str) else spec
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("VitEngine needs a CUDA device: the B200 path has no CPU fallback")
        self.precision = precision
        self.max_cells = max_cells_per_call
        s = self.spec
        sd = {k: v.detach().to("cpu", torch.float32) for k, v in state_dict.items()}
        want = (1, s.tokens, s.dim)
        if tuple(sd["pos_embed"].shape) != want:
            raise ValueError(f"pos_embed {tuple(sd['pos_embed'].shape)} does not match {want}")
        pk = _Packer()
        d = _lib.VitDesc()
        d.dim, d.heads, d.depth, d.in_chans, d.classes, d.tokens = s.dim, s.heads, s.depth, s.in_chans, len(s.classes), s.tokens
        pos = sd["pos_embed"][0]
        table = pos + sd["patch_embed.proj.bias"][None, :]
        table[0] = pos[0] + sd["cls_token"].reshape(-1)
        d.embed_table = pk.add_f32(table)
        d.embed_w = pk.add_mat(sd["patch_embed.proj.weight"].reshape(s.dim, -1))
        d.norm_g = pk.add_f32(sd["norm.weight"]); d.norm_b = pk.add_f32(sd["norm.bias"])
        d.head_w = pk.add_f32(sd["head.weight"]); d.head_b = pk.add_f32(sd["head.bias"])
        # LayerNorm folded into its consumer GEMM (ribca_ln_fold): bit 0 = norm1 into qkv (fc2 / the patch embedding leave the
        # planes + row statistics), bit 1 = norm2 into fc1 (proj leaves them).  Opt-in: measured neutral (mask 1) to -1 % (mask 3)
        # on the power-capped B200s of this pool - the epilogue work it adds to the GEMMs costs what the LayerNorm kernels
        # it removes cost (profiles/r02_lnfold.md)
        fold = int(os.environ.get("RIBCA_LN_FOLD", "0")) & 3
        for i in range(s.depth):
            _pack_block(pk, sd, f"blocks.{i}", d.blocks[i], s.heads, fold=fold)
        self.wf32, mats = pk.finish(self.device)
        self.ln_fold = fold
        self._packs = _Packs(d, mats, pk.mat_off, self.device, pk.folds, fold)
        self.desc, self.wsplit = self._packs.get(precision)

    def set_head(self, weight: torch.Tensor, bias: torch.Tensor):
        """Replace the classification head in place (used by the calibration recipe)."""
        k, dim = len(self.spec.classes), self.spec.dim
        self.wf32[self.desc.head_w: self.desc.head_w + k * dim] = weight.reshape(-1).to(self.device, torch.float32)
        self.wf32[self.desc.head_b: self.desc.head_b + k] = bias.reshape(-1).to(self.device, torch.float32)

    @torch.no_grad()
    def forward(self, patches: torch.Tensor, return_logits: bool = False, precision: str | None = None):
        """patches (n, C, 40, 40) float32 on the device -> softmax probabilities (n, classes)."""
        ops._need_cuda(patches)
        n, c = patches.shape[0], patches.shape[1]
        if c != self.spec.in_chans or patches.shape[2:] != (40, 40) or patches.dtype != torch.float32:
            raise ValueError(f"expected (n, {self.spec.in_chans}, 40, 40) float32 patches, got {tuple(patches.shape)} {patches.dtype}")
        L = _lib.lib()
        k = len(self.spec.classes)
        probs = torch.empty((n, k), dtype=torch.float32, device=self.device)
        logits = torch.empty((n, k), dtype=torch.float32, device=self.device) if return_logits else None
        prec = ops.PRECISION[precision or self.precision]
        desc, wsplit = self._packs.get(precision or self.precision)
        for i in range(0, n, self.max_cells):
            m = min(self.max_cells, n - i)
            ws_bytes = L.ribca_vit_workspace_bytes(C.byref(desc), m)
            ws = _WS.get(ws_bytes, self.device)
            _lib.check(L.ribca_vit_forward(C.byref(desc), ops._ptr(self.wf32), ops._ptr(wsplit),
                                           ops._ptr(patches[i:i + m]), m, ops._ptr(probs[i:i + m]),
                                           ops._ptr(logits[i:i + m]) if logits is not None else 0,
                                           ops._ptr(ws), ws.numel(), prec, ops._stream()), "ribca_vit_forward")
        return (probs, logits) if return_logits else probs


class MaeEngine:
    """One panel's MAE marker imputer (reference markerImputer.py:69-329) resident on a CUDA device."""
    PRECISIONS = ("f16f8", "bf16x3", "bf16x1", "bf16", "simt")

    def __init__(self, spec: MaeSpec | str, state_dict: dict, device="cuda", precision: str = ops.DEFAULT_PRECISION,
                 max_cells_per_call: int = 8192):
        self.spec = MAE_SPECS[spec] if isinstance(spec, str) else spec
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("MaeEngine needs a CUDA device: the B200 path has no CPU fallback")
        self.precision = precision
        self.max_cells = max_cells_per_call
        s = self.spec
        sd = {k: v.detach().to("cpu", torch.float32) for k, v in state_dict.items()}
        pk = _Packer()
        d = _lib.MaeDesc()
        d.channels = s.channels
        d.enc_dim, d.enc_heads, d.enc_depth = s.enc_dim, s.enc_heads, s.enc_depth
        d.dec_dim, d.dec_heads, d.dec_depth = s.dec_dim, s.dec_heads, s.dec_depth
        d.embed_w = pk.add_mat(sd["patch_embed.proj.weight"].reshape(s.enc_dim, 1600))
        d.embed_bias = pk.add_f32(sd["patch_embed.proj.bias"])
        d.cls_token = pk.add_f32(sd["cls_token"]); d.pos_embed = pk.add_f32(sd["pos_embed"])
        d.norm_g = pk.add_f32(sd["norm.weight"]); d.norm_b = pk.add_f32(sd["norm.bias"])
        d.dec_embed_w = pk.add_mat(sd["decoder_embed.weight"]); d.dec_embed_b = pk.add_f32(sd["decoder_embed.bias"])
        d.mask_token = pk.add_f32(sd["mask_token"]); d.dec_pos_embed = pk.add_f32(sd["decoder_pos_embed"])
        d.dec_norm_g = pk.add_f32(sd["decoder_norm.weight"]); d.dec_norm_b = pk.add_f32(sd["decoder_norm.bias"])
        d.pred_w = pk.add_mat(sd["decoder_pred.weight"]); d.pred_b = pk.add_f32(sd["decoder_pred.bias"])
        for i in range(s.enc_depth):
            _pack_block(pk, sd, f"blocks.{i}", d.enc_blocks[i], s.enc_heads)
        for i in range(s.dec_depth):
            _pack_block(pk, sd, f"decoder_blocks.{i}", d.dec_blocks[i], s.dec_heads)
        self.wf32, mats = pk.finish(self.device)
        self._packs = _Packs(d, mats, pk.mat_off, self.device)
        self.desc, self.wsplit = self._packs.get(precision)

    @torch.no_grad()
    def impute(self, patches: torch.Tensor, present, precision: str | None = None) -> torch.Tensor:
        """In place: channels not listed in `present` (ascending panel positions) are replaced by the
        decoder's prediction; present channels are left untouched.  Returns `patches`."""
        ops._need_cuda(patches)
        n, c = patches.shape[0], patches.shape[1]
        if c != self.spec.channels or patches.dtype != torch.float32:
            raise ValueError(f"expected (n, {self.spec.channels}, 40, 40) float32 patches")
        present = sorted(int(p) for p in present)
        L = _lib.lib()
        arr = (C.c_int * len(present))(*present)
        prec = ops.PRECISION[precision or self.precision]
        desc, wsplit = self._packs.get(precision or self.precision)
        for i in range(0, n, self.max_cells):
            m = min(self.max_cells, n - i)
            ws_bytes = L.ribca_mae_workspace_bytes(C.byref(desc), m)
            ws = _WS.get(ws_bytes, self.device)
            _lib.check(L.ribca_mae_impute(C.byref(desc), ops._ptr(self.wf32), ops._ptr(wsplit),
                                          ops._ptr(patches[i:i + m]), m, arr, len(present), ops._ptr(ws), ws.numel(),
                                          prec, ops._stream()), "ribca_mae_impute")
        return patches
