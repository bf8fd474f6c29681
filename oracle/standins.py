"""TEST INFRASTRUCTURE - CPU restatements of the third-party calls on RIBCA's hot path.

The reference (`/root/reference`, read-only) imports scikit-image and timm, neither of which is
installed in this image.  This module restates, with scipy.ndimage and plain
torch.nn, exactly the calls the reference makes:

  * skimage.morphology.disk / dilation      (reference utils.py:260)
  * skimage.filters.gaussian                (reference utils.py:265)
  * skimage.transform.resize                (reference preprocess.py:106)
  * timm.models.vision_transformer.{VisionTransformer, PatchEmbed, Block}
                                            (reference model.py:14,31-64, markerImputer.py:7,80-103)

It is used in two places only: `oracle/refshim.py` registers these as stand-in modules so the
reference's own .py files import unmodified in the build container, and `oracle/ribca_oracle.py`
calls them for its CPU restatement.  Nothing under the product package imports this file.

PARITY UNPINNED for these functions: scikit-image / timm sources are neither under
/root/reference nor installed, so the restatements follow their published behaviour
(skimage 0.19+ `resize`/`gaussian`/`dilation`, timm 0.9-1.0.14 `VisionTransformer`).
"""
from __future__ import annotations

import math
from functools import partial

import numpy as np
import scipy.ndimage as ndi
import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# scikit-image
# ----------------------------------------------------------------------------------------------
def disk(radius: int, dtype=np.uint8) -> np.ndarray:
    """skimage.morphology.disk: footprint of pixels with dx^2 + dy^2 <= r^2."""
    r = int(radius)
    yy, xx = np.mgrid[-r:r + 1, -r:r + 1]
    return (xx * xx + yy * yy <= r * r).astype(dtype)


def dilation(image: np.ndarray, footprint: np.ndarray) -> np.ndarray:
    """skimage.morphology.dilation for a symmetric footprint.

    scikit-image mirrors the footprint and calls scipy's grey dilation; pixels outside the image
    never raise the maximum for a disk footprint (reflecting a convex symmetric footprint only
    re-visits pixels that are already closer), so bool in -> bool out.
    """
    fp = np.asarray(footprint)[::-1, ::-1] != 0
    out = ndi.grey_dilation(np.asarray(image).astype(np.uint8), footprint=fp, mode="reflect")
    return out.astype(np.asarray(image).dtype)


def gaussian(image: np.ndarray, sigma=1, mode="nearest", cval=0, truncate=4.0) -> np.ndarray:
    """skimage.filters.gaussian: float64 conversion, scipy gaussian_filter, mode 'nearest'."""
    img = np.asarray(image)
    if img.dtype == bool:
        img = img.astype(np.float64)
    elif img.dtype.kind != "f":
        raise TypeError("stand-in covers only the bool/float inputs the reference passes")
    else:
        img = img.astype(np.float64, copy=False)
    return ndi.gaussian_filter(img, sigma, mode=mode, cval=cval, truncate=truncate)


def resize(image: np.ndarray, output_shape, order=0, anti_aliasing=True, preserve_range=True,
           mode="reflect") -> np.ndarray:
    """skimage.transform.resize as the reference calls it (order 0, anti_aliasing, preserve_range).

    skimage >= 0.19: optional Gaussian pre-filter with sigma = max(0, (factor - 1) / 2) using
    scipy mode 'mirror', then scipy.ndimage.zoom(order, mode='mirror', grid_mode=True), then clip
    to the input range.  When the output shape equals the input shape this is the identity.
    """
    image = np.asarray(image, dtype=np.float64)
    in_shape = np.asarray(image.shape, dtype=np.float64)
    out_shape = tuple(int(s) for s in output_shape)
    factors = in_shape / np.asarray(out_shape, dtype=np.float64)
    if np.all(factors == 1):
        return image.copy()
    lo, hi = image.min(), image.max()
    filtered = image
    if anti_aliasing:
        sigma = np.maximum(0, (factors - 1) / 2)
        if np.any(sigma > 0):
            filtered = ndi.gaussian_filter(image, sigma, mode="mirror", cval=0)
    zoom = [1.0 / f for f in factors]
    out = ndi.zoom(filtered, zoom, order=order, mode="mirror", cval=0, grid_mode=True)
    return np.clip(out, lo, hi)


# ----------------------------------------------------------------------------------------------
# timm
# ----------------------------------------------------------------------------------------------
def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class PatchEmbed(nn.Module):
    """timm.layers.PatchEmbed: strided conv, flatten(2).transpose(1, 2)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None,
                 flatten=True, bias=True, **_):
        super().__init__()
        self.img_size = _pair(img_size)
        self.patch_size = _pair(patch_size)
        self.grid_size = (self.img_size[0] // self.patch_size[0], self.img_size[1] // self.patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size, bias=bias)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        _, _, h, w = x.shape
        assert (h, w) == self.img_size, f"input {(h, w)} != img_size {self.img_size}"
        x = self.proj(x)
        if self.flatten:
            x = x.flatten(2).transpose(1, 2)
        return self.norm(x)


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, **_):
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.q_norm = nn.Identity()
        self.k_norm = nn.Identity()
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)

    def forward(self, x):
        b, n, c = x.shape
        qkv = self.qkv(x).reshape(b, n, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        x = F.scaled_dot_product_attention(q, k, v)
        x = x.transpose(1, 2).reshape(b, n, c)
        return self.proj_drop(self.proj(x))


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features, **_):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.drop1 = nn.Dropout(0.0)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden_features, in_features)
        self.drop2 = nn.Dropout(0.0)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


class Block(nn.Module):
    """timm pre-LN transformer block (no layer-scale, drop-path inactive in eval)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, norm_layer=nn.LayerNorm, **_):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias)
        self.ls1 = nn.Identity()
        self.drop_path1 = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.ls2 = nn.Identity()
        self.drop_path2 = nn.Identity()

    def forward(self, x):
        x = x + self.drop_path1(self.ls1(self.attn(self.norm1(x))))
        x = x + self.drop_path2(self.ls2(self.mlp(self.norm2(x))))
        return x


class VisionTransformer(nn.Module):
    """timm.models.vision_transformer.VisionTransformer, the subset the reference subclasses.

    Constructed with timm's default global_pool='token' (the reference subclass does not forward
    its own `global_pool` kwarg), so `fc_norm` is Identity and `norm` is a LayerNorm.
    """

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, global_pool="token",
                 embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0, qkv_bias=True,
                 drop_path_rate=0.0, norm_layer=None, **_):
        super().__init__()
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.num_classes = num_classes
        self.global_pool = global_pool
        self.num_features = self.embed_dim = embed_dim
        self.num_prefix_tokens = 1
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        n = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, n + 1, embed_dim) * 0.02)
        self.pos_drop = nn.Dropout(0.0)
        self.blocks = nn.Sequential(*[
            Block(embed_dim, num_heads, mlp_ratio, qkv_bias=qkv_bias, norm_layer=norm_layer)
            for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.fc_norm = nn.Identity()
        self.head_drop = nn.Dropout(0.0)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self._timm_init()

    def _timm_init(self):
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward_features(self, x):
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1) + self.pos_embed
        x = self.blocks(self.pos_drop(x))
        return self.norm(x)

    def forward_head(self, x, pre_logits: bool = False):
        if self.global_pool:
            x = x[:, self.num_prefix_tokens:].mean(dim=1) if self.global_pool == "avg" else x[:, 0]
        x = self.head_drop(self.fc_norm(x))
        return x if pre_logits else self.head(x)

    def forward(self, x):
        return self.forward_head(self.forward_features(x))
