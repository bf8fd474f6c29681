"""Typed Python wrappers over the C ABI (include/ribca_b200.h).

PyTorch is used here only to own device memory and streams; every computation is a call into
libribca_b200.so with raw device pointers.  The small host-side plans (Gaussian tap weights,
np.percentile order statistics) are computed with the same numpy expressions scipy / numpy use,
which is what makes the device results bit-identical to the reference's CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

PATCH = 40
MAX_PANEL_CH = 16
GAUSS_STRIDE = 16
PRECISION = {"bf16x3": 0, "bf16": 1, "bf16x1": 1, "simt": 2, "f16f8": 3, "fp32": 4}
DEFAULT_PRECISION = "f16f8"
FMT_BF16, FMT_F16F8, FMT_F32 = 0, 1, 2          # ribca_plane_format


def plane_format(precision: str) -> int:
    code = PRECISION[precision]
    return FMT_F16F8 if code == 3 else (FMT_F32 if code == 4 else FMT_BF16)


def weight_log2_scale(max_abs: float) -> int:
    """t of the f16f8 weight packing: the largest power of two with max|w| * 2^t <= 128, so that the e4m3 copies
    (|.| <= 448) and the fp16 main plane (w * 2^(t+8) <= 32768) stay finite."""
    import math
    if not max_abs > 0.0 or not math.isfinite(max_abs):
        return 0
    return int(max(-20, min(30, math.floor(math.log2(128.0 / max_abs)))))
EPI_STORE, EPI_RESIDUAL, EPI_GELU, EPI_STORE_SPLIT, EPI_STORE_LN, EPI_RESIDUAL_LN = 0, 1, 2, 3, 4, 5
_DTYPES = {torch.uint8: 0, torch.uint16: 1, torch.float32: 2, torch.int32: 3}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
            raise RuntimeError("libribca_b200 operates on contiguous CUDA tensors only (no CPU fallback)")


def launch_count() -> int:
    return int(_lib.lib().ribca_launch_count())


def set_interleave(on: bool) -> None:
    """Two-way interleave of the classifier forward (ribca_set_interleave): halves of a call on two streams, so the HBM-bound
    LayerNorm / im2col kernels of one half run under the tensor-core GEMM of the other.  Opt-in (neutral under the
    power cap, profiles/r02_interleave.md); bit-identical results."""
    _lib.check(_lib.lib().ribca_set_interleave(1 if on else 0), "ribca_set_interleave")


# ------------------------------------------------------------------------------------------------
# host-side plans
# ------------------------------------------------------------------------------------------------
def gaussian_half_kernel(sigma, truncate: float = 4.0):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius) from the centre outwards:
    returns (w[0..r], r) with r = int(truncate * sigma + 0.5)."""
    radius = int(truncate * float(sigma) + 0.5)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[::-1][radius:], dtype=np.float64), radius


def percentile_plan(n: int, amax, dtype=np.float32):
    """(k_lo, k_hi, gamma) of np.percentile(x, amax) for a float32 array of n elements, method
    'linear' (numpy/lib/_function_base_impl.py: percentile -> _quantile -> _get_indexes/_get_gamma).
    The quantile and the virtual index are float32 because numpy divides by `a.dtype.type(100)`."""
    q = np.asanyarray(np.true_divide(amax, dtype(100)))
    vi = np.asanyarray((n - 1) * q)
    prev = np.asanyarray(np.floor(vi))
    nxt = np.asanyarray(prev + 1)
    if vi >= n - 1:
        prev = nxt = np.asanyarray(-1.0)
    if vi < 0:
        prev = nxt = np.asanyarray(0.0)
    prev_i, nxt_i = prev.astype(np.intp), nxt.astype(np.intp)
    gamma = np.asanyarray(vi - prev_i, dtype=vi.dtype)
    k_lo, k_hi = int(prev_i), int(nxt_i)
    if k_lo < 0:
        k_lo += n
    if k_hi < 0:
        k_hi += n
    return k_lo, k_hi, float(gamma)


def _dptr(arr: np.ndarray):
    return arr.ctypes.data_as(C.POINTER(C.c_double))


# ------------------------------------------------------------------------------------------------
# stage 1
# ------------------------------------------------------------------------------------------------
def normalize_from_host(img_host: torch.Tensor, device, blur=0.3, amax=99.8) -> torch.Tensor:
    """normalize() of a HOST (ideally pinned) stack with the upload hidden behind the kernels: the channels are copied one
    by one on a side stream and each is normalised as soon as it has arrived (the per-channel statistics of
    _normalize are independent, reference preprocess.py:218-238)."""
    if img_host.is_cuda:
        return normalize(img_host, blur, amax)
    if img_host.dtype not in _DTYPES:
        raise TypeError(f"unsupported image dtype {img_host.dtype}")
    c, h, w = img_host.shape
    main = torch.cuda.current_stream(device)
    side = _side_stream(device)
    raw = torch.empty((c, h, w), dtype=img_host.dtype, device=device)
    out = torch.empty((c, h, w), dtype=torch.float32, device=device)
    side.wait_stream(main)                          # `raw` was allocated on the main stream
    events = []
    with torch.cuda.stream(side):
        for k in range(c):
            raw[k].copy_(img_host[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
            events.append(ev)
    for k in range(c):
        main.wait_event(events[k])
        normalize(raw[k:k + 1], blur, amax, out=out[k:k + 1])
    raw.record_stream(side)
    return out


def normalize_from_planes(planes, device, blur=0.3, amax=99.8) -> torch.Tensor:
    """normalize() fed by a decoder: `planes` yields (k, host stack (C, H, W)) as soon as plane k of the (pinned) stack is
    complete (io.iter_tiff_planes).  Plane k is uploaded on the side stream and normalised while the decoder reads plane k + 1:
    disk read, H2D copy and stage 1 overlap."""
    main = torch.cuda.current_stream(device)
    side = _side_stream(device)
    raw = out = None
    for k, stack in planes:
        host = torch.from_numpy(stack) if isinstance(stack, np.ndarray) else stack
        if host.dtype not in _DTYPES:
            host = host.to(torch.float32)
        if raw is None:
            c, h, w = host.shape
            raw = torch.empty((c, h, w), dtype=host.dtype, device=device)
            out = torch.empty((c, h, w), dtype=torch.float32, device=device)
            side.wait_stream(main)
        with torch.cuda.stream(side):
            raw[k].copy_(host[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
        main.wait_event(ev)
        normalize(raw[k:k + 1], blur, amax, out=out[k:k + 1])
    if raw is None:
        raise ValueError("normalize_from_planes: the decoder produced no plane")
    side.synchronize()            # only copies run there: the decoder may release its (pinned) buffer once this returns
    raw.record_stream(side)
    return out


_SIDE = {}


def _side_stream(device):
    key = torch.device(device).index
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device)
    return _SIDE[key]


def normalize(img: torch.Tensor, blur=0.3, amax=99.8, return_stats: bool = False, out: torch.Tensor | None = None):
    """ImageProcessor._normalize on the device.  img: (C, H, W) uint8 / uint16 / int32 / float32."""
    _need_cuda(img)
    if img.dtype not in _DTYPES:
        raise TypeError(f"unsupported image dtype {img.dtype}")
    c, h, w = img.shape
    L = _lib.lib()
    w_bg, r_bg = gaussian_half_kernel(20)
    if blur:
        w_bl, r_bl = gaussian_half_kernel(blur)
    else:
        w_bl, r_bl = np.zeros(1), -1
    k_lo, k_hi, gamma = percentile_plan(h * w, amax)
    if out is None:
        out = torch.empty((c, h, w), dtype=torch.float32, device=img.device)
    elif out.shape != img.shape or out.dtype != torch.float32 or not out.is_contiguous() or not img.is_contiguous():
        raise ValueError("normalize: `out` must be a contiguous float32 tensor of the image's shape")
    stats = torch.empty((c, 4), dtype=torch.float32, device=img.device)
    ws_bytes = L.ribca_normalize_workspace_bytes(c, h, w)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=img.device)
    _lib.check(L.ribca_normalize(_ptr(img), _DTYPES[img.dtype], c, h, w, _dptr(w_bg), r_bg, _dptr(w_bl), r_bl,
                                 k_lo, k_hi, gamma, _ptr(out), _ptr(stats), _ptr(ws), ws_bytes, _stream()),
               "ribca_normalize")
    return (out, stats) if return_stats else out


def channel_min(img: torch.Tensor) -> torch.Tensor:
    _need_cuda(img)
    c = img.shape[0]
    out = torch.empty(c, dtype=torch.float32, device=img.device)
    _lib.check(_lib.lib().ribca_channel_min(_ptr(img), c, img[0].numel(), _ptr(out), _stream()), "ribca_channel_min")
    return out


# ------------------------------------------------------------------------------------------------
# stage 2
# ------------------------------------------------------------------------------------------------
class CellTable:
    """Compacted per-cell statistics on the device (ids ascending)."""

    def __init__(self, ids, bbox, sums, count, id_to_index, n, max_id):
        self.ids, self.bbox, self.sums, self.count = ids, bbox, sums, count
        self.id_to_index, self.n, self.max_id = id_to_index, n, max_id

    def subset(self, idx: torch.Tensor) -> "CellTable":
        """The table rows `idx` (int64 device tensor) as a table of their own - the cells of a re-evaluation batch."""
        return CellTable(self.ids[idx].contiguous(), self.bbox[idx].contiguous(), self.sums[idx].contiguous(),
                         self.count[idx].contiguous(), self.id_to_index, int(idx.numel()), self.max_id)

    def centroids(self) -> torch.Tensor:
        """np.mean of the pixel lists: exact integer sums, one float64 division (model.py:785-786)."""
        return self.sums.to(torch.float64) / self.count.to(torch.float64).unsqueeze(1)


def cell_stats(mask: torch.Tensor) -> CellTable:
    """_cell_pos_dict reduced to (ids, bbox, coordinate sums, area).  mask: (H, W) int32 on the device."""
    _need_cuda(mask)
    if mask.dtype != torch.int32 or mask.dim() != 2:
        raise TypeError("mask must be a 2-D int32 tensor")
    L = _lib.lib()
    h, w = mask.shape
    dev = mask.device
    mm = torch.empty(2, dtype=torch.int32, device=dev)
    _lib.check(L.ribca_mask_minmax(_ptr(mask), mask.numel(), _ptr(mm), _stream()), "ribca_mask_minmax")
    lo, hi = (int(v) for v in mm.tolist())         # the one host sync of stage 2: table sizes depend on it
    if lo < 0:
        raise ValueError("negative cell labels are not supported")
    if hi > (1 << 27):
        raise ValueError(f"largest label {hi} exceeds the dense-table limit 2^27; relabel the mask")
    n_ids = hi + 1
    bbox = torch.empty((n_ids, 4), dtype=torch.int32, device=dev)
    sums = torch.empty((n_ids, 2), dtype=torch.int64, device=dev)
    count = torch.empty(n_ids, dtype=torch.int32, device=dev)
    _lib.check(L.ribca_cell_stats(_ptr(mask), h, w, hi, _ptr(bbox), _ptr(sums), _ptr(count), _stream()), "ribca_cell_stats")
    ids = torch.empty(n_ids, dtype=torch.int32, device=dev)
    cbbox = torch.empty((n_ids, 4), dtype=torch.int32, device=dev)
    csums = torch.empty((n_ids, 2), dtype=torch.int64, device=dev)
    ccount = torch.empty(n_ids, dtype=torch.int32, device=dev)
    id2idx = torch.empty(n_ids, dtype=torch.int32, device=dev)
    n_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = L.ribca_compact_workspace_bytes(hi)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(L.ribca_compact_cells(_ptr(bbox), _ptr(sums), _ptr(count), hi, _ptr(ids), _ptr(cbbox), _ptr(csums),
                                     _ptr(ccount), _ptr(id2idx), _ptr(n_dev), _ptr(ws), ws_bytes, _stream()),
               "ribca_compact_cells")
    n = int(n_dev.item())
    return CellTable(ids[:n], cbbox[:n], csums[:n], ccount[:n], id2idx, n, hi)


def cell_pixels(mask: torch.Tensor, cells: CellTable):
    """The pixel lists of cell_pos_dict as CSR arrays on the device: (offsets int64 (n+1,), rows int32, cols int32);
    cell j owns [offsets[j], offsets[j+1]) in raster order (reference preprocess.py:159-181)."""
    _need_cuda(mask)
    h, w = mask.shape
    dev = mask.device
    offsets = torch.zeros(cells.n + 1, dtype=torch.int64, device=dev)
    if cells.n:
        torch.cumsum(cells.count, 0, out=offsets[1:])
    total = int(offsets[-1].item())
    rows = torch.empty(total, dtype=torch.int32, device=dev)
    cols = torch.empty(total, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().ribca_cell_pixels(_ptr(mask), h, w, _ptr(cells.ids), _ptr(cells.bbox), _ptr(offsets), cells.n,
                                            _ptr(rows), _ptr(cols), _stream()), "ribca_cell_pixels")
    return offsets, rows, cols


# ------------------------------------------------------------------------------------------------
# stage 3
# ------------------------------------------------------------------------------------------------
_GAUSS = None


def _gauss_table() -> np.ndarray:
    global _GAUSS
    if _GAUSS is None:
        tab = np.zeros((3, GAUSS_STRIDE), dtype=np.float64)
        for s in (1, 2, 3):
            wts, r = gaussian_half_kernel(s)
            tab[s - 1, : r + 1] = wts
        _GAUSS = tab
    return _GAUSS


def resize_plan(patch_edge: int):
    """Host plan of skimage.transform.resize((C, P, P) -> (C, 40, 40), order=0, anti_aliasing=True):
    (source index per output row/column, anti-alias half kernel, radius).  The index arithmetic is
    scipy.ndimage.zoom's (grid_mode=True, order 0): cc = (k + 0.5) * zoom - 0.5, idx = floor(cc + 0.5),
    with zoom = P / 40 in float64."""
    zoom = np.divide(np.array([patch_edge]), np.array([PATCH]), out=np.ones(1, dtype=np.float64))[0]
    idx = []
    for k in range(PATCH):
        cc = float(k)
        cc += 0.5
        cc *= zoom
        cc -= 0.5
        idx.append(int(np.floor(cc + 0.5)))
    sigma = max(0.0, (patch_edge / PATCH - 1) / 2)
    if sigma > 0:
        w, r = gaussian_half_kernel(sigma)
    else:
        w, r = np.zeros(1), -1
    return np.asarray(idx, dtype=np.int32), np.ascontiguousarray(w, dtype=np.float64), r


def build_patches(img: torch.Tensor, mask: torch.Tensor, min_val: torch.Tensor, cells: CellTable, panels,
                  cell_begin: int = 0, n_cells: int | None = None, want_intensity: bool = False,
                  want_windows: bool = False, cell_size=30):
    """crop_cell + smooth + channel select for cells [cell_begin, cell_begin + n_cells).
    panels: list of channel-index lists (may contain -1).  Returns (list of (n, C_p, 40, 40) float32
    tensors, avg_int (n, C_img) float64 or None, windows (n, 4) int32 or None)."""
    _need_cuda(img, mask, min_val)
    L = _lib.lib()
    c_img, h, w = img.shape
    n = cells.n - cell_begin if n_cells is None else n_cells
    dev = img.device
    npan = len(panels)
    if npan > 3:
        raise ValueError("at most 3 panels per call")
    n_ch = (C.c_int * max(npan, 1))(*[len(p) for p in panels])
    idx = (C.c_int * (max(npan, 1) * MAX_PANEL_CH))()
    outs = []
    out_ptrs = (C.c_void_p * max(npan, 1))()
    for p, chans in enumerate(panels):
        if len(chans) > MAX_PANEL_CH:
            raise ValueError("panel too wide")
        for k, ch in enumerate(chans):
            idx[p * MAX_PANEL_CH + k] = int(ch)
        t = torch.empty((n, len(chans), PATCH, PATCH), dtype=torch.float32, device=dev)
        outs.append(t)
        out_ptrs[p] = t.data_ptr()
    avg = torch.empty((n, c_img), dtype=torch.float64, device=dev) if want_intensity else None
    wins = torch.empty((n, 4), dtype=torch.int32, device=dev) if want_windows else None
    g = _gauss_table()
    edge = int(PATCH * (cell_size / 30.0))                    # reference preprocess.py:67,78
    if edge == PATCH:
        _lib.check(L.ribca_build_patches(_ptr(img), _ptr(mask), c_img, h, w, _ptr(min_val), _ptr(cells.ids), _ptr(cells.bbox),
                                         cell_begin, n, npan, n_ch, idx, out_ptrs, _dptr(g), _ptr(avg), _ptr(wins), _stream()),
                   "ribca_build_patches")
    else:
        src, w_aa, r_aa = resize_plan(edge)
        _lib.check(L.ribca_build_patches_resized(_ptr(img), _ptr(mask), c_img, h, w, _ptr(min_val), _ptr(cells.ids),
                                                 _ptr(cells.bbox), cell_begin, n, npan, n_ch, idx, out_ptrs, _dptr(g), edge,
                                                 src.ctypes.data_as(C.POINTER(C.c_int)), _dptr(w_aa), r_aa, _ptr(avg), _ptr(wins),
                                                 _stream()), "ribca_build_patches_resized")
    return outs, avg, wins


# ------------------------------------------------------------------------------------------------
# stage 4 primitives
# ------------------------------------------------------------------------------------------------
def split_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 tensor -> (2, *shape) bf16 planes {hi, lo} with x ~= hi + lo."""
    _need_cuda(x)
    out = torch.empty((2,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().ribca_split_bf16(_ptr(x), x.numel(), _ptr(out[0]), _ptr(out[1]), _stream()), "ribca_split_bf16")
    return out


def split_planes(x: torch.Tensor, fmt: int = FMT_F16F8, w_role: bool = False, log2_scale: int = 0) -> torch.Tensor:
    """fp32 tensor -> (2, *shape) 16-bit operand planes in `fmt` (A role, or W role scaled by 2^log2_scale)."""
    _need_cuda(x)
    x = x.contiguous()
    out = torch.empty((2,) + tuple(x.shape), dtype=torch.bfloat16 if fmt == FMT_BF16 else torch.int16, device=x.device)
    _lib.check(_lib.lib().ribca_split_planes(_ptr(x), x.numel(), fmt, int(bool(w_role)), int(log2_scale), _ptr(out[0]),
                                             _ptr(out[1]), _stream()), "ribca_split_planes")
    return out


def gemm(a_split: torch.Tensor, w_split: torch.Tensor, bias=None, row_table=None, epilogue=EPI_STORE,
         out=None, precision="bf16x3", w_log2_scale: int = 0):
    """out (+)= A . W^T with A (2, M, K), W (2, N, K) operand planes; see ribca_gemm_splitbf16."""
    _need_cuda(a_split, w_split, bias, row_table, out)
    _, m, k = a_split.shape
    _, n, k2 = w_split.shape
    assert k == k2
    dev = a_split.device
    out_f32 = out_split = None
    if epilogue in (EPI_GELU, EPI_STORE_SPLIT):
        f8 = plane_format(precision) == FMT_F16F8 and epilogue == EPI_GELU
        out_split = out if out is not None else torch.empty((2, m, n), dtype=torch.int16 if f8 else torch.bfloat16, device=dev)
    else:
        out_f32 = out if out is not None else torch.empty((m, n), dtype=torch.float32, device=dev)
    period = row_table.shape[0] if row_table is not None else 0
    _lib.check(_lib.lib().ribca_gemm_splitbf16(_ptr(a_split), m * k, _ptr(w_split), n * k, m, n, k, _ptr(bias), _ptr(row_table),
                                               period, epilogue, _ptr(out_f32), _ptr(out_split), m * n, PRECISION[precision],
                                               int(w_log2_scale), _stream()), "ribca_gemm_splitbf16")
    return out_split if epilogue in (EPI_GELU, EPI_STORE_SPLIT) else out_f32


def gemm_ln(a_split: torch.Tensor, w_split: torch.Tensor, bias=None, row_table=None, epilogue=EPI_STORE, out=None,
            precision="bf16x3", w_log2_scale: int = 0, stats_in=None, c1=None, slots_in: int = 0, eps: float = 1e-6):
    """ribca_gemm_ln: the GEMMs around a folded LayerNorm.
    Producer (epilogue EPI_STORE_LN / EPI_RESIDUAL_LN): returns (x fp32 (M, N) [`out` += for the residual form], planes of x
    (2, M, N) in the format of `precision`, stats (M, LN_SLOTS, 2), filled slots).
    Consumer (stats_in, c1, slots_in given; bias = c2): like gemm() with v = rstd * (acc - mean * c1) + c2."""
    _need_cuda(a_split, w_split, bias, row_table, out, stats_in, c1)
    _, m, k = a_split.shape
    _, n, k2 = w_split.shape
    assert k == k2
    dev = a_split.device
    L = _lib.lib()
    ln = _lib.LnFold()
    ln.stats_in, ln.c1, ln.slots_in, ln.eps = _ptr(stats_in) or None, _ptr(c1) or None, int(slots_in), float(eps)
    ln_out = epilogue in (EPI_STORE_LN, EPI_RESIDUAL_LN)
    out_f32 = out_split = stats = None
    fmt = plane_format(precision)
    if ln_out:
        out_f32 = out if out is not None else torch.empty((m, n), dtype=torch.float32, device=dev)
        out_split = _planes((m, n), fmt, dev)
        stats = torch.zeros((m, _lib.LN_SLOTS, 2), dtype=torch.float32, device=dev)
        ln.stats_out = _ptr(stats)
    elif epilogue in (EPI_GELU, EPI_STORE_SPLIT):
        out_split = _planes((m, n), fmt if epilogue == EPI_GELU else FMT_BF16, dev)
    else:
        out_f32 = out if out is not None else torch.empty((m, n), dtype=torch.float32, device=dev)
    period = row_table.shape[0] if row_table is not None else 0
    _lib.check(L.ribca_gemm_ln(_ptr(a_split), m * k, _ptr(w_split), n * k, m, n, k, _ptr(bias), _ptr(row_table), period, epilogue,
                               _ptr(out_f32), _ptr(out_split), m * n, PRECISION[precision], int(w_log2_scale), C.byref(ln), _stream()),
               "ribca_gemm_ln")
    if ln_out:
        return out_f32, out_split, stats, int(L.ribca_gemm_ln_slots(n, PRECISION[precision]))
    return out_split if epilogue in (EPI_GELU, EPI_STORE_SPLIT) else out_f32


def _planes(shape, fmt, device):
    return torch.empty((2,) + tuple(shape), dtype=torch.bfloat16 if fmt == FMT_BF16 else torch.int16, device=device)


def layernorm_split(x: torch.Tensor, gamma, beta, eps=1e-6, fmt: int = FMT_BF16) -> torch.Tensor:
    _need_cuda(x, gamma, beta)
    m, d = x.shape
    out = _planes((m, d), fmt, x.device)
    _lib.check(_lib.lib().ribca_layernorm_split(_ptr(x), m, d, _ptr(gamma), _ptr(beta), eps, _ptr(out), m * d, fmt, _stream()),
               "ribca_layernorm_split")
    return out


def attention(qkv: torch.Tensor, cells: int, tokens: int, heads: int, fmt: int = FMT_BF16) -> torch.Tensor:
    _need_cuda(qkv)
    m, d3 = qkv.shape
    d = d3 // 3
    out = _planes((m, d), fmt, qkv.device)
    _lib.check(_lib.lib().ribca_attention(_ptr(qkv), cells, tokens, heads, d // heads, _ptr(out), m * d, fmt, _stream()),
               "ribca_attention")
    return out


def attention_tc(qkv_split: torch.Tensor, cells: int, tokens: int, heads: int, head_dim: int, fmt: int = FMT_BF16) -> torch.Tensor:
    """Tensor-core attention: qkv_split (2, M, 3*heads*hdp) bf16 planes -> (2, M, heads*head_dim) planes in `fmt`."""
    _need_cuda(qkv_split)
    _, m, wq = qkv_split.shape
    d = heads * head_dim
    out = _planes((m, d), fmt, qkv_split.device)
    _lib.check(_lib.lib().ribca_attention_tc(_ptr(qkv_split), m * wq, cells, tokens, heads, head_dim, _ptr(out), m * d,
                                             fmt, _stream()), "ribca_attention_tc")
    return out


# ------------------------------------------------------------------------------------------------
# stage 5
# ------------------------------------------------------------------------------------------------
def merge_votes(probs0, types0, probs1, types1, vote_rank, type_thresh, confidence, want_margin: bool = False):
    """Returns (label uint8 (n,), conf float32 (n,), counts int64 (18,)) and, with want_margin, the decision margin
    float32 (n,) of every cell (include/ribca_b200.h: ribca_merge_votes)."""
    _need_cuda(probs0, probs1)
    n, k0 = probs0.shape
    dev = probs0.device
    label = torch.empty(n, dtype=torch.uint8, device=dev)
    conf = torch.empty(n, dtype=torch.float32, device=dev)
    counts = torch.zeros(18, dtype=torch.int64, device=dev)
    margin = torch.empty(n, dtype=torch.float32, device=dev) if want_margin else None
    if n == 0:
        return (label, conf, counts, margin) if want_margin else (label, conf, counts)
    t0 = (C.c_int * len(types0))(*types0)
    k1 = 0 if probs1 is None else probs1.shape[1]
    t1 = (C.c_int * max(k1, 1))(*(types1 or [0]))
    vr = (C.c_int * 18)(*vote_rank)
    tt = (C.c_float * 18)(*type_thresh)
    _lib.check(_lib.lib().ribca_merge_votes(_ptr(probs0), k0, t0, _ptr(probs1), k1, t1, n, vr, tt, float(confidence),
                                            _ptr(label), _ptr(conf), _ptr(counts), _ptr(margin), _stream()), "ribca_merge_votes")
    return (label, conf, counts, margin) if want_margin else (label, conf, counts)


def paint_cells(mask: torch.Tensor, cells: CellTable, cell_value: torch.Tensor) -> torch.Tensor:
    """Per-pixel map from per-cell uint8 values (n,) or (n, channels<=4): 0 on the background."""
    _need_cuda(mask, cell_value)
    ch = 1 if cell_value.dim() == 1 else cell_value.shape[1]
    h, w = mask.shape
    out = torch.empty((h, w) if cell_value.dim() == 1 else (h, w, ch), dtype=torch.uint8, device=mask.device)
    if cells.n == 0:
        return out.zero_()
    _lib.check(_lib.lib().ribca_paint_cells(_ptr(mask), mask.numel(), _ptr(cells.id_to_index), cells.max_id, _ptr(cell_value), ch,
                                            _ptr(out), _stream()), "ribca_paint_cells")
    return out


# ------------------------------------------------------------------------------------------------
# spatial statistics
# ------------------------------------------------------------------------------------------------
def knn_2d(xy: torch.Tensor, k: int, return_distance: bool = False, points_per_bin: float = 3.0):
    """Exact k nearest neighbours (self included) of the rows of xy (n, 2) float64 on the device:
    sklearn NearestNeighbors(n_neighbors=k).fit(xy).kneighbors(xy) -> indices (n, k) int32 [, distances (n, k)]."""
    _need_cuda(xy)
    if xy.dtype != torch.float64 or xy.dim() != 2 or xy.shape[1] != 2:
        raise TypeError("xy must be an (n, 2) float64 tensor")
    n = xy.shape[0]
    if k > n:
        raise ValueError(f"Expected n_neighbors <= n_samples, but n_samples = {n}, n_neighbors = {k}")      # sklearn's error
    xy = xy.contiguous()
    lo, hi = xy.min(0).values, xy.max(0).values
    x0, y0 = float(lo[0]), float(lo[1])
    ex, ey = max(float(hi[0]) - x0, 1e-9), max(float(hi[1]) - y0, 1e-9)
    cell = max((points_per_bin * ex * ey / n) ** 0.5, max(ex, ey) / 4096.0, 1e-9)
    gx, gy = int(ex / cell) + 1, int(ey / cell) + 1
    bx = ((xy[:, 0] - x0) / cell).to(torch.int64).clamp_(0, gx - 1)
    by = ((xy[:, 1] - y0) / cell).to(torch.int64).clamp_(0, gy - 1)
    key, order = torch.sort(by * gx + bx, stable=True)
    bin_start = torch.searchsorted(key, torch.arange(gx * gy + 1, device=xy.device)).to(torch.int32)
    xy_sorted = xy[order].contiguous()
    order32 = order.to(torch.int32)
    idx = torch.empty((n, k), dtype=torch.int32, device=xy.device)
    d2 = torch.empty((n, k), dtype=torch.float64, device=xy.device) if return_distance else None
    # the kernel recomputes the bin of a query as int((x - x0) * (1 / cell)); any rounding difference against the division
    # above only moves the ring centre by one bin, which the stopping rule (distance to the visited square) tolerates
    _lib.check(_lib.lib().ribca_knn_2d(_ptr(xy_sorted), _ptr(order32), _ptr(bin_start), n, k, x0, y0, cell, gx, gy, _ptr(idx),
                                       _ptr(d2), _stream()), "ribca_knn_2d")
    return (idx, d2.sqrt_()) if return_distance else idx


def neighbor_stats(nbr: torch.Tensor, types: torch.Tensor, n_types: int, levels=None, skip: int = 1, want_matrix: bool = True):
    """-> (type matrix (n_types, n_types) int64 or None, compositions (n, len(levels) * n_types) float64 or None)."""
    _need_cuda(nbr, types)
    n, k = nbr.shape
    types = types.to(torch.int32).contiguous()
    mat = torch.zeros((n_types, n_types), dtype=torch.int64, device=nbr.device) if want_matrix else None
    comp, lv, nl = None, None, 0
    if levels:
        nl = len(levels)
        lv = (C.c_int * nl)(*[int(v) for v in levels])
        comp = torch.empty((n, nl * n_types), dtype=torch.float64, device=nbr.device)
    _lib.check(_lib.lib().ribca_neighbor_stats(_ptr(nbr), _ptr(types), n, k, skip, n_types, _ptr(mat), lv, nl, _ptr(comp), _stream()),
               "ribca_neighbor_stats")
    return mat, comp
