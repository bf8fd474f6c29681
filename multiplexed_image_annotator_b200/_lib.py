"""Build and load libribca_b200.so (the C-ABI CUDA library, include/ribca_b200.h).

`build()` compiles csrc/*.cu for sm_100a with plain nvcc into an in-tree shared object (so it
travels to the GPU box with the repository snapshot); `lib()` loads it through ctypes and declares
every entry point.  There is no fallback: if the library is missing or a call fails, a
RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import shutil
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(_ROOT, "include")
LIB_PATH = os.environ.get("RIBCA_LIB") or os.path.join(_HERE, "libribca_b200.so")      # RIBCA_LIB: A/B-test another build

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I" + INCLUDE]

_lock = threading.Lock()
_lib = None


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libribca_b200.so")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if os.environ.get("RIBCA_LIB"):
        return False
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a and link libribca_b200.so in-tree."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    objdir = os.path.join(_HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        procs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for cmd, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + out)
        if verbose and out.strip():
            print(out)
        objs.append(obj)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: " + " ".join(link) + "\n" + r.stdout)
    _write_build_info(nvcc)
    return LIB_PATH


BUILD_INFO_PATH = LIB_PATH + ".build.json"        # travels with the .so (git-ignored, not gpurun-ignored)


def _sources_digest() -> str:
    import hashlib
    h = hashlib.sha256()
    for path in sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(INCLUDE, "*.h"))):
        h.update(os.path.basename(path).encode())
        h.update(open(path, "rb").read())
    return h.hexdigest()


def _write_build_info(nvcc: str) -> None:
    import json
    import platform
    import time
    ver = subprocess.run([nvcc, "--version"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout.strip().splitlines()
    info = {"built_at_utc": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "host": platform.node(), "nvcc": ver[-1] if ver else "",
            "flags": NVCC_FLAGS[:-1], "sources_sha256": _sources_digest(), "gpu_visible_at_build": _gpu_visible()}
    with open(BUILD_INFO_PATH, "w") as f:
        json.dump(info, f, indent=1)


def _gpu_visible() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def build_info() -> dict:
    """Where and from what the loaded library was built (`build()` writes it next to the .so): `sources_match` tells whether
    the sources in the tree are the ones it was compiled from."""
    import json
    try:
        info = json.load(open(BUILD_INFO_PATH))
    except Exception:
        return {"built_at_utc": None, "sources_match": None}
    info["sources_match"] = info.get("sources_sha256") == _sources_digest()
    return info


class LnFold(C.Structure):
    _fields_ = [("stats_in", C.c_void_p), ("c1", C.c_void_p), ("slots_in", C.c_int), ("eps", C.c_float), ("stats_out", C.c_void_p)]


LN_SLOTS = 8


class BlockDesc(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in ("ln1_g", "ln1_b", "ln2_g", "ln2_b", "qkv_b", "proj_b", "fc1_b", "fc2_b",
                                             "qkv_w", "proj_w", "fc1_w", "fc2_w", "qkv_c1", "qkv_c2", "fc1_c1", "fc1_c2")]


class VitDesc(C.Structure):
    _fields_ = [("dim", C.c_int), ("heads", C.c_int), ("depth", C.c_int), ("in_chans", C.c_int),
                ("classes", C.c_int), ("tokens", C.c_int), ("plane_format", C.c_int), ("w_log2_scale", C.c_int),
                ("ln_folded", C.c_int), ("reserved", C.c_int), ("split_plane", C.c_longlong),
                ("embed_w", C.c_longlong), ("embed_table", C.c_longlong), ("norm_g", C.c_longlong),
                ("norm_b", C.c_longlong), ("head_w", C.c_longlong), ("head_b", C.c_longlong),
                ("blocks", BlockDesc * 16)]


class MaeDesc(C.Structure):
    _fields_ = [("channels", C.c_int), ("enc_dim", C.c_int), ("enc_heads", C.c_int), ("enc_depth", C.c_int),
                ("dec_dim", C.c_int), ("dec_heads", C.c_int), ("dec_depth", C.c_int),
                ("plane_format", C.c_int), ("w_log2_scale", C.c_int), ("split_plane", C.c_longlong), ("embed_w", C.c_longlong), ("embed_bias", C.c_longlong),
                ("cls_token", C.c_longlong), ("pos_embed", C.c_longlong), ("norm_g", C.c_longlong),
                ("norm_b", C.c_longlong), ("dec_embed_w", C.c_longlong), ("dec_embed_b", C.c_longlong),
                ("mask_token", C.c_longlong), ("dec_pos_embed", C.c_longlong), ("dec_norm_g", C.c_longlong),
                ("dec_norm_b", C.c_longlong), ("pred_w", C.c_longlong), ("pred_b", C.c_longlong),
                ("enc_blocks", BlockDesc * 16), ("dec_blocks", BlockDesc * 16)]


_P, _I, _LL, _F, _D, _SZ = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_size_t

# every symbol include/ribca_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "ribca_last_error": (C.c_char_p, []),
    "ribca_version": (_I, []),
    "ribca_launch_count": (_LL, []),
    "ribca_set_interleave": (_I, [_I]),
    "ribca_profile_begin": (_I, []),
    "ribca_profile_end": (_I, [C.POINTER(_D), C.POINTER(_LL), C.POINTER(_D), _I]),
    "ribca_normalize_workspace_bytes": (_SZ, [_I, _I, _I]),
    "ribca_normalize": (_I, [_P, _I, _I, _I, _I, C.POINTER(_D), _I, C.POINTER(_D), _I, _LL, _LL, _F, _P, _P, _P, _SZ, _P]),
    "ribca_channel_min": (_I, [_P, _I, _LL, _P, _P]),
    "ribca_mask_minmax": (_I, [_P, _LL, _P, _P]),
    "ribca_cell_stats": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "ribca_compact_workspace_bytes": (_SZ, [_I]),
    "ribca_compact_cells": (_I, [_P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "ribca_cell_pixels": (_I, [_P, _I, _I, _P, _P, _P, _I, _P, _P, _P]),
    "ribca_build_patches": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _I, _I, _I, C.POINTER(_I), C.POINTER(_I),
                                 C.POINTER(_P), C.POINTER(_D), _P, _P, _P]),
    "ribca_build_patches_resized": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _I, _I, _I, C.POINTER(_I), C.POINTER(_I),
                                         C.POINTER(_P), C.POINTER(_D), _I, C.POINTER(_I), C.POINTER(_D), _I, _P, _P, _P]),
    "ribca_gemm_splitbf16": (_I, [_P, _LL, _P, _LL, _I, _I, _I, _P, _P, _I, _I, _P, _P, _LL, _I, _I, _P]),
    "ribca_gemm_ln": (_I, [_P, _LL, _P, _LL, _I, _I, _I, _P, _P, _I, _I, _P, _P, _LL, _I, _I, C.POINTER(LnFold), _P]),
    "ribca_gemm_ln_slots": (_I, [_I, _I]),
    "ribca_split_bf16": (_I, [_P, _LL, _P, _P, _P]),
    "ribca_split_planes": (_I, [_P, _LL, _I, _I, _I, _P, _P, _P]),
    "ribca_layernorm_split": (_I, [_P, _I, _I, _P, _P, _F, _P, _LL, _I, _P]),
    "ribca_attention": (_I, [_P, _I, _I, _I, _I, _P, _LL, _I, _P]),
    "ribca_attention_tc": (_I, [_P, _LL, _I, _I, _I, _I, _P, _LL, _I, _P]),
    "ribca_vit_workspace_bytes": (_SZ, [C.POINTER(VitDesc), _I]),
    "ribca_vit_forward": (_I, [C.POINTER(VitDesc), _P, _P, _P, _I, _P, _P, _P, _SZ, _I, _P]),
    "ribca_mae_workspace_bytes": (_SZ, [C.POINTER(MaeDesc), _I]),
    "ribca_mae_impute": (_I, [C.POINTER(MaeDesc), _P, _P, _P, _I, C.POINTER(_I), _I, _P, _SZ, _I, _P]),
    "ribca_knn_2d": (_I, [_P, _P, _P, _I, _I, _D, _D, _D, _I, _I, _P, _P, _P]),
    "ribca_neighbor_stats": (_I, [_P, _P, _I, _I, _I, _I, _P, C.POINTER(_I), _I, _P, _P]),
    "ribca_paint_cells": (_I, [_P, _LL, _P, _I, _P, _I, _P, _P]),
    "ribca_merge_votes": (_I, [_P, _I, C.POINTER(_I), _P, _I, C.POINTER(_I), _I, C.POINTER(_I), C.POINTER(_F), _F,
                               _P, _P, _P, _P, _P]),
}


def lib() -> C.CDLL:
    """The loaded library with typed entry points.  Raises if it is not built (no fallback)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(needs nvcc).  There is no CPU fallback.")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)      # AttributeError here = header / library mismatch
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().ribca_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libribca_b200 {what} failed (code {rc}): {msg}")
