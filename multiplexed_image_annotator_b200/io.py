"""Host-side image / mask decoding (reference cta/preprocess.py:244-250 uses skimage.io.imread, i.e. tifffile).

File decode is outside the timed hot path but dominates the wall-clock once the rest runs on a B200
(SURVEY 8f rank 2), so (multi-page / Big / OME-) TIFF goes through a small reader of its own: `TiffFile` parses the
IFD chain, and uncompressed pages (strips or tiles, either byte order) are read straight into ONE pinned (C, H, W)
buffer with `readinto` - no per-page array, no copy, and the buffer can be handed to the asynchronous, per-channel
upload of `ops.normalize_from_host`.  Compressed pages (LZW / Deflate / PackBits) are decoded page by page with PIL
into the same buffer.  `ome_channel_names` extracts the marker names the Napari widget reads from OME-XML
(reference _widget.py:686-705).  .npy stacks and PNG masks are read with numpy / PIL.
"""
from __future__ import annotations

import os
import struct
import time
import xml.etree.ElementTree as ET

import numpy as np

# TIFF field types -> (struct code, size)
_TIFF_TYPES = {1: ("B", 1), 2: ("c", 1), 3: ("H", 2), 4: ("I", 4), 5: ("II", 8), 6: ("b", 1), 7: ("B", 1), 8: ("h", 2),
               9: ("i", 4), 10: ("ii", 8), 11: ("f", 4), 12: ("d", 8), 13: ("I", 4), 16: ("Q", 8), 17: ("q", 8), 18: ("Q", 8)}
_SAMPLE_DTYPES = {(1, 8): "u1", (1, 16): "u2", (1, 32): "u4", (2, 8): "i1", (2, 16): "i2", (2, 32): "i4", (3, 32): "f4", (3, 64): "f8"}


class TiffPage:
    """One IFD: geometry, sample type and the byte ranges of its strips / tiles."""

    def __init__(self, tags, byteorder):
        g = lambda t, d=None: tags.get(t, d)
        one = lambda v, d: (v[0] if isinstance(v, tuple) else v) if v is not None else d
        self.width, self.height = one(g(256), 0), one(g(257), 0)
        self.samples = one(g(277), 1)
        bits = one(g(258), 1)
        fmt = one(g(339), 1)
        self.compression = one(g(259), 1)
        self.planar = one(g(284), 1)
        self.photometric = one(g(262), 1)
        self.predictor = one(g(317), 1)
        self.description = g(270)
        self.subfile_type = one(g(254), 0)
        key = (fmt if fmt in (1, 2, 3) else 1, bits)
        self.dtype = np.dtype(byteorder + _SAMPLE_DTYPES[key]) if key in _SAMPLE_DTYPES else None
        self.tiled = 322 in tags
        if self.tiled:
            self.tile_w, self.tile_h = one(g(322), 0), one(g(323), 0)
            self.offsets, self.counts = tuple(g(324, ())), tuple(g(325, ()))
        else:
            self.rows_per_strip = min(one(g(278), self.height), self.height) or self.height
            self.offsets, self.counts = tuple(g(273, ())), tuple(g(279, ()))

    @property
    def raw_readable(self) -> bool:
        """Uncompressed, whole-byte samples, one sample per pixel or planar layout: readable without a codec."""
        return (self.compression == 1 and self.dtype is not None and self.predictor == 1 and (self.samples == 1 or self.planar == 2)
                and len(self.offsets) > 0 and len(self.offsets) == len(self.counts))


class TiffFile:
    """IFD index of a classic TIFF or BigTIFF file (no pixel data is read by the constructor)."""

    def __init__(self, path):
        self.path = str(path)
        self.pages = []
        with open(self.path, "rb") as f:
            head = f.read(16)
            if head[:2] == b"II":
                bo = "<"
            elif head[:2] == b"MM":
                bo = ">"
            else:
                raise ValueError(f"{path}: not a TIFF file")
            magic = struct.unpack(bo + "H", head[2:4])[0]
            if magic == 42:
                self.big, off = False, struct.unpack(bo + "I", head[4:8])[0]
            elif magic == 43:
                self.big, off = True, struct.unpack(bo + "Q", head[8:16])[0]
            else:
                raise ValueError(f"{path}: bad TIFF magic {magic}")
            self.byteorder = bo
            seen = set()
            while off and off not in seen:
                seen.add(off)
                tags, off = self._read_ifd(f, off)
                self.pages.append(TiffPage(tags, bo))

    def _read_ifd(self, f, off):
        bo = self.byteorder
        f.seek(off)
        if self.big:
            n = struct.unpack(bo + "Q", f.read(8))[0]
            entry, inline, fmt = 20, 8, bo + "HHQ"
        else:
            n = struct.unpack(bo + "H", f.read(2))[0]
            entry, inline, fmt = 12, 4, bo + "HHI"
        raw = f.read(n * entry + (8 if self.big else 4))
        tags = {}
        for i in range(n):
            e = raw[i * entry:(i + 1) * entry]
            tag, typ, count = struct.unpack(fmt, e[:entry - inline])
            if typ not in _TIFF_TYPES:
                continue
            code, size = _TIFF_TYPES[typ]
            nbytes = size * count
            if nbytes <= inline:
                data = e[entry - inline:entry - inline + nbytes]
            else:
                pos = struct.unpack(bo + ("Q" if self.big else "I"), e[entry - inline:])[0]
                here = f.tell()
                f.seek(pos)
                data = f.read(nbytes)
                f.seek(here)
            if typ == 2:
                tags[tag] = data.split(b"\0", 1)[0].decode("utf-8", "replace")
            elif typ in (5, 10):
                vals = struct.unpack(bo + code[0] * (2 * count), data)
                tags[tag] = tuple(vals[2 * k] / vals[2 * k + 1] if vals[2 * k + 1] else 0.0 for k in range(count))
            else:
                tags[tag] = struct.unpack(bo + code * count, data)
        nxt = struct.unpack(bo + ("Q" if self.big else "I"), raw[n * entry:])[0]
        return tags, nxt

    # ---- pixel data ------------------------------------------------------------------------------------
    def image_pages(self):
        """The full-resolution pages: reduced-resolution (pyramid) and mask sub-files are skipped."""
        full = [p for p in self.pages if not (p.subfile_type & 1) and p.width and p.height]
        if not full:
            return []
        w, h = full[0].width, full[0].height
        return [p for p in full if (p.width, p.height) == (w, h)]

    def read_page_into(self, page: TiffPage, out: np.ndarray, f=None) -> None:
        """Uncompressed page -> out (samples, H, W) or (H, W) native-endian array; strips / tiles via readinto."""
        assert page.raw_readable
        close = f is None
        f = f or open(self.path, "rb")
        try:
            planes = out if out.ndim == 3 else out[None]
            native = page.dtype.newbyteorder("=")
            swap = page.dtype.byteorder not in ("=", "|") and page.dtype != native
            item = page.dtype.itemsize
            if not page.tiled:
                strips_per_plane = (page.height + page.rows_per_strip - 1) // page.rows_per_strip
                for s, (off, cnt) in enumerate(zip(page.offsets, page.counts)):
                    pl, k = divmod(s, strips_per_plane)
                    r0 = k * page.rows_per_strip
                    r1 = min(r0 + page.rows_per_strip, page.height)
                    dst = planes[pl, r0:r1]                          # contiguous rows of one plane
                    want = (r1 - r0) * page.width * item
                    f.seek(off)
                    got = f.readinto(memoryview(dst.reshape(-1).view(np.uint8))[:min(cnt, want)])
                    if got < min(cnt, want):
                        raise IOError(f"{self.path}: truncated strip {s}")
            else:
                tx = (page.width + page.tile_w - 1) // page.tile_w
                ty = (page.height + page.tile_h - 1) // page.tile_h
                buf = np.empty((page.tile_h, page.tile_w), dtype=native)
                for t, (off, cnt) in enumerate(zip(page.offsets, page.counts)):
                    pl, k = divmod(t, tx * ty)
                    y0, x0 = (k // tx) * page.tile_h, (k % tx) * page.tile_w
                    f.seek(off)
                    f.readinto(memoryview(buf.reshape(-1).view(np.uint8))[:min(cnt, buf.nbytes)])
                    h, w = min(page.tile_h, page.height - y0), min(page.tile_w, page.width - x0)
                    planes[pl, y0:y0 + h, x0:x0 + w] = buf[:h, :w]
            if swap:
                planes.byteswap(inplace=True)
        finally:
            if close:
                f.close()


def _alloc(shape, dtype, pin):
    """(numpy view, owner): a pinned torch buffer when `pin` and CUDA is present, else plain numpy."""
    if pin:
        try:
            import torch
            if torch.cuda.is_available():
                t = torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
                return t.numpy(), t
        except Exception:
            pass
    a = np.empty(shape, dtype=dtype)
    return a, a


def read_tiff_stack(path, pin: bool = True) -> np.ndarray:
    """(C, H, W) stack of a multi-page / planar TIFF, decoded into one (pinned) buffer.  Uncompressed pages are read with the
    reader above; compressed ones page by page through PIL.  Chunky multi-sample pages (RGB-like) fall back to PIL too."""
    out = None
    for _, out in iter_tiff_planes(path, pin):
        pass
    return out


def iter_tiff_planes(path, pin: bool = True):
    """Generator over the decode of read_tiff_stack: yields (k, stack) right after planes [..k] of the ONE (C, H, W) buffer are
    filled, so a consumer can ship plane k to the device while plane k + 1 is still being read from disk."""
    tf = TiffFile(path)
    pages = tf.image_pages()
    if not pages:
        raise ValueError(f"{path}: no image pages")
    p0 = pages[0]
    if p0.dtype is None or any((p.dtype, p.samples) != (p0.dtype, p0.samples) for p in pages):
        raise ValueError(f"{path}: pages of mixed or unsupported sample types")
    if p0.samples > 1 and p0.planar != 2:
        raise ValueError(f"{path}: chunky multi-sample pages")
    c = len(pages) * p0.samples
    native = p0.dtype.newbyteorder("=")
    out, _owner = _alloc((c, p0.height, p0.width), native, pin)
    pil = None
    with open(tf.path, "rb") as f:
        for k, page in enumerate(pages):
            dst = out[k * p0.samples:(k + 1) * p0.samples]
            if page.raw_readable:
                tf.read_page_into(page, dst if p0.samples > 1 else dst[0], f)
            else:
                if pil is None:
                    from PIL import Image
                    pil = Image.open(tf.path)
                pil.seek(tf.pages.index(page))
                dst[...] = np.asarray(pil).reshape(dst.shape)
            for j in range(p0.samples):
                yield k * p0.samples + j, out


def iter_npy_planes(path, pin: bool = True, threads: int = 4):
    """The .npy counterpart of iter_tiff_planes: a C-ordered (C, H, W) / (H, W) array of a native numeric dtype is read plane by
    plane (positional reads straight into ONE (pinned) buffer, up to `threads` planes in flight: the page-cache copy of one
    thread is ~3 GB/s); yields (k, stack) in order, after plane k is complete.  Anything else (Fortran order, object arrays,
    other ranks) raises ValueError and the caller falls back to np.load."""
    from concurrent.futures import ThreadPoolExecutor
    with open(path, "rb") as f:
        version = np.lib.format.read_magic(f)
        shape, fortran, dtype = (np.lib.format.read_array_header_1_0 if version == (1, 0) else np.lib.format.read_array_header_2_0)(f)
        if fortran or dtype.hasobject or len(shape) not in (2, 3) or not dtype.isnative or np.dtype(dtype).name not in (
                "uint8", "uint16", "int32", "float32"):
            raise ValueError(f"{path}: not a C-ordered uint8 / uint16 / int32 / float32 image stack")
        c, h, w = shape if len(shape) == 3 else (1,) + tuple(shape)
        out, _owner = _alloc((c, h, w), dtype, pin)
        base, plane_bytes, fd = f.tell(), h * w * np.dtype(dtype).itemsize, f.fileno()

        def read_plane(k):
            dst = memoryview(out[k]).cast("B")
            got = 0
            while got < plane_bytes:
                n = os.preadv(fd, [dst[got:]], base + k * plane_bytes + got)
                if not n:
                    raise ValueError(f"{path}: truncated file")
                got += n
            return k

        with ThreadPoolExecutor(max_workers=max(1, min(threads, c))) as pool:
            ahead = max(1, min(threads, c))
            futs = [pool.submit(read_plane, k) for k in range(min(ahead, c))]
            for k in range(c):
                futs[k].result()
                if k + ahead < c:
                    futs.append(pool.submit(read_plane, k + ahead))
                yield k, out


def is_npy(path) -> bool:
    return str(path).lower().endswith(".npy")


def is_tiff(path) -> bool:
    return str(path).lower().endswith((".tif", ".tiff", ".qptiff"))


def ome_channel_names(path):
    """Marker names from the OME-XML of a TIFF's first ImageDescription (<Channel Name="...">, any OME schema version), or
    None when the file carries no OME metadata (reference _widget.py:686-705 reads the same attribute through tifffile)."""
    tf = TiffFile(path)
    desc = next((p.description for p in tf.pages if isinstance(p.description, str) and "<OME" in p.description), None)
    if not desc:
        return None
    try:
        root = ET.fromstring(desc[desc.index("<OME"):])          # any XML declaration / BOM before the root element is dropped
    except ET.ParseError:
        return None
    names = [el.attrib["Name"] for el in root.iter() if el.tag.rsplit("}", 1)[-1] == "Channel" and "Name" in el.attrib]
    return names or None


def _read_any(path: str) -> np.ndarray:
    path = str(path)
    if path.endswith(".npy"):
        return np.load(path)
    if path.lower().endswith((".tif", ".tiff", ".qptiff")):
        try:
            a = read_tiff_stack(path)            # own reader: one pinned (C, H, W) buffer, no codec for uncompressed pages
            return a[0] if a.shape[0] == 1 else a                 # single page -> 2-D, as imread returns it
        except (ValueError, KeyError, struct.error):
            pass                                 # chunky RGB, exotic sample types ...: let PIL / OpenCV try
        try:
            import cv2
        except ImportError:
            cv2 = None                           # no OpenCV: PIL below
        if cv2 is not None:
            ok, pages = cv2.imreadmulti(path, flags=cv2.IMREAD_UNCHANGED)
            if ok and len(pages):
                # OpenCV decodes colour pages as BGR(A); the reference reads through tifffile (RGB) and read_mask keeps
                # channel 0 = red, so put the channels back in file order
                pages = [pg[..., [2, 1, 0] + list(range(3, pg.shape[-1]))] if pg.ndim == 3 and pg.shape[-1] >= 3 else pg for pg in pages]
                return np.stack(pages, 0) if len(pages) > 1 else pages[0]
    from PIL import Image
    im = Image.open(path)
    frames = []
    try:
        while True:
            frames.append(np.array(im))
            im.seek(im.tell() + 1)
    except EOFError:
        pass
    return np.stack(frames, 0) if len(frames) > 1 else frames[0]


def read_image(path: str) -> np.ndarray:
    """(C, H, W) stack in its stored dtype (uint8 / uint16 / int32 / float32)."""
    a = _read_any(path)
    if a.ndim == 2:
        a = a[None]
    if a.dtype == np.float64:
        a = a.astype(np.float32)
    elif a.dtype not in (np.uint8, np.uint16, np.int32, np.float32):
        a = a.astype(np.float32)                 # the reference casts everything to float32 anyway
    return np.ascontiguousarray(a)


def read_mask(path: str) -> np.ndarray:
    """2-D int32 label mask; a 3-D mask keeps its first channel (preprocess.py:247-250)."""
    m = _read_any(path)
    if m.ndim == 3:
        m = m[:, :, 0]
    return np.ascontiguousarray(m.astype(np.int32))


# ------------------------------------------------------------------------------------------------
# run log: file name and line format of the reference (`<main_dir>/results/log.txt`, a creation line, then
# one message per line; reference cta/logger.py).  Line-buffered: the reference's callers never close it.
# ------------------------------------------------------------------------------------------------
class RunLog:
    FILE = "results/log.txt"

    def __init__(self, main_dir):
        from .parallel import world
        os.makedirs(os.path.join(main_dir, "results"), exist_ok=True)
        self.log_file_path = os.path.join(main_dir, self.FILE)
        # several ranks may share one main_dir (torchrun, one Annotator per rank): rank 0 owns the file
        self.log_file = open(self.log_file_path if world()[0] == 0 else os.devnull, "w", buffering=1)
        self.log(f"Log file created at {time.ctime()}")

    def log(self, message):
        if not self.log_file.closed:
            print(str(message), file=self.log_file)

    def log_all_hyperparameters(self, hyperparameters):
        self.log("Hyperparameters:")
        for name, value in hyperparameters.items():
            self.log(f"{name}: {value}")

    def close(self):
        self.log_file.close()

    def __del__(self):
        try:
            self.log_file.close()
        except Exception:
            pass
