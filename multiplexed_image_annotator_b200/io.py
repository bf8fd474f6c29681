"""Host-side image / mask decoding (reference cta/preprocess.py:244-250 uses skimage.io.imread).

File decode is outside the timed hot path; this reads what the offline image can read:
.npy stacks, PNG, and (multi-page) TIFF through PIL / OpenCV.  Returned arrays are C-contiguous so
they can be pinned and copied to the device in one transfer.
"""
from __future__ import annotations

import os
import time

import numpy as np


def _read_any(path: str) -> np.ndarray:
    path = str(path)
    if path.endswith(".npy"):
        return np.load(path)
    if path.lower().endswith((".tif", ".tiff", ".qptiff")):
        try:
            import tifffile                      # not in the offline image; used when present
            return tifffile.imread(path)
        except ImportError:
            pass
        import cv2
        ok, pages = cv2.imreadmulti(path, flags=cv2.IMREAD_UNCHANGED)
        if ok and len(pages):
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
        os.makedirs(os.path.join(main_dir, "results"), exist_ok=True)
        self.log_file_path = os.path.join(main_dir, self.FILE)
        self.log_file = open(self.log_file_path, "w", buffering=1)
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
