"""CPU-only tests: the C-ABI library builds, loads and exports every declared symbol; host-side plans
match numpy / scipy; the reference-facing host classes behave like the reference (golden fixtures);
the multi-rank partition + gather logic works over gloo with world_size 2."""
import contextlib
import io
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from multiplexed_image_annotator_b200 import _lib, ops, parallel, synth, weights
from multiplexed_image_annotator_b200.cell_type_annotation.markerParse import MarkerParser
from multiplexed_image_annotator_b200.cell_type_annotation import model as bmodel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_header_symbol():
    _lib.build()
    header = open(os.path.join(ROOT, "include", "ribca_b200.h")).read()
    declared = set(re.findall(r"\b(ribca_[a-z0-9_]+)\s*\(", header))
    declared -= {"ribca_stream_t"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.lib()                         # raises AttributeError if a symbol is missing
    assert lib.ribca_version() >= 100
    nm = subprocess.run(["nm", "-D", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (ribca_[a-z0-9_]+)", nm))
    assert declared <= exported
    info = _lib.build_info()                 # the build record that travels with the .so: which sources, when, where
    assert info["sources_match"] and info["nvcc"] and info["built_at_utc"]


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    _lib.build()
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "gemm_tcgen05_kernel", _lib.LIB_PATH], capture_output=True, text=True).stdout
    if "UTCHMMA" not in sass:               # -fun needs the mangled name on some toolkits: fall back to everything
        sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "UTCQMMA" in sass                # the e4m3 pass of the default f16f8 operand format (kind::f8f6f4)


def test_ops_refuse_cpu_tensors():
    with pytest.raises(RuntimeError):
        ops.normalize(torch.zeros((1, 8, 8), dtype=torch.uint16))
    with pytest.raises(RuntimeError):
        ops.cell_stats(torch.zeros((8, 8), dtype=torch.int32))


def test_percentile_plan_matches_numpy():
    rng = np.random.default_rng(0)

    def lerp32(a, b, g):
        a, b, g = np.float32(a), np.float32(b), np.float32(g)
        d = np.float32(b - a)
        r = np.float32(a + np.float32(d * g))
        if g >= 0.5:
            r = np.float32(b - np.float32(d * np.float32(np.float32(1) - g)))
        return r

    for n in (1, 2, 3, 10, 1000, 19650, 123457, 17_000_001):
        x = np.maximum(rng.normal(size=n).astype(np.float32) * 50, 0)
        xs = np.sort(x)
        for amax in (99.8, 100, 95.0, 50, 0, 99.99):
            k_lo, k_hi, g = ops.percentile_plan(n, amax)
            assert np.percentile(x, amax) == lerp32(xs[k_lo], xs[k_hi], g), (n, amax)


def test_gaussian_taps_match_scipy():
    from scipy.ndimage._filters import _gaussian_kernel1d
    for s in (20, 0.3, 0.4, 1, 2, 3):
        w, r = ops.gaussian_half_kernel(s)
        full = _gaussian_kernel1d(s, 0, r)[::-1]
        assert r == int(4.0 * float(s) + 0.5) and np.array_equal(full[r:], w)


def test_resize_plan_matches_scipy_zoom():
    import scipy.ndimage as ndi
    for edge in range(8, 81):
        idx, w, r = ops.resize_plan(edge)
        a = np.arange(edge, dtype=np.float64)
        assert np.array_equal(ndi.zoom(a, 40 / edge, order=0, mode="mirror", grid_mode=True), a[idx]), edge
        assert r == (int(4.0 * ((edge / 40 - 1) / 2) + 0.5) if edge > 40 else -1)


def test_marker_parser_matches_reference(golden_dir, tmp_path):
    cases = json.load(open(os.path.join(golden_dir, "markers.json")))
    for name, case in cases.items():
        f = tmp_path / f"{name}.txt"
        f.write_text("\n".join(case["markers"]) + "\n")
        p = MarkerParser(strict=case["strict"], logger=None)
        with contextlib.redirect_stdout(io.StringIO()):
            if "raises" in case:
                with pytest.raises(TypeError):
                    p.parse(str(f))
                continue
            p.parse(str(f))
        assert p.indices == case["indices"], name
        assert [p.immune_base, p.immune_extended, p.immune_full, p.struct, p.nerve] == case["flags"]
        assert [str(m) for m in p.markers] == case["parsed_markers"] and p.n_markers == case["n_markers"]


def test_merge_branch_matches_reference_elif_chain(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "merge.json")))
    for case in cases:
        if "raises" in case:
            with pytest.raises(KeyError):
                bmodel.merge_branch(case["panels"])
        else:
            used = bmodel.merge_branch(case["panels"])
            assert 1 <= len(used) <= 2 and set(used) <= set(case["panels"])
    with pytest.raises(ValueError):
        bmodel.merge_branch([])


def test_random_weights_fit_the_reference_architectures():
    from oracle import ribca_oracle as orc
    for panel in weights.VIT_SPECS:
        orc.make_vit(panel).load_state_dict(weights.random_vit_state(panel, seed=0), strict=True)
    orc.make_mae("immune_base").load_state_dict(weights.random_mae_state("immune_base", seed=0), strict=True)
    tab = weights.sincos_table(512, (2, 5))
    np.testing.assert_allclose(tab[0].numpy(), orc.sincos_2d(512, (2, 5)), atol=1e-6)


def test_synthetic_scene_generator():
    mask = synth.synth_mask(180, 200, seed=2)
    assert mask.dtype == torch.int32 and mask.shape == (180, 200)
    ids = torch.unique(mask)
    assert ids[0] == 0 and len(ids) > 80
    img = synth.to_uint16(synth.synth_image(mask, 4, seed=2))
    assert img.shape == (4, 180, 200) and img.dtype == np.uint16 and img.max() > 500
    assert torch.equal(mask, synth.synth_mask(180, 200, seed=2))


def test_shard_range_is_a_balanced_partition():
    for n in (0, 1, 7, 100, 51843):
        for w in (1, 2, 3, 8):
            parts = [parallel.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from multiplexed_image_annotator_b200 import parallel
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, world = parallel.world()
n = 11
lo, hi = parallel.shard_range(n, rank, world)
label = torch.arange(lo, hi, dtype=torch.uint8)
conf = torch.arange(lo, hi, dtype=torch.float32) * 0.5
full_l = parallel.all_gather_rows(label, n, lo, hi)
full_c = parallel.all_gather_rows(conf, n, lo, hi)
counts = parallel.all_reduce_sum(torch.tensor([hi - lo, 1], dtype=torch.int64))
assert full_l.tolist() == list(range(n)), full_l
assert full_c.tolist() == [0.5 * i for i in range(n)]
assert counts.tolist() == [n, 2]
mat = parallel.all_gather_rows(torch.full((hi - lo, 3), float(rank)), n, lo, hi)
assert mat.shape == (n, 3) and mat[:6].eq(0).all() and mat[6:].eq(1).all()
# batch of 5 "images" sharded round-robin: every rank ends with every image's results
res = []
for i in range(5):
    if parallel.image_owner(i, world) == rank:
        n_i = 3 + i
        res.append((torch.full((n_i,), i, dtype=torch.uint8), torch.arange(n_i, dtype=torch.float32) + i, torch.arange(18) * (i + 1)))
    else:
        res.append(None)
full = parallel.share_image_results(res, torch.device("cpu"))
for i, (l, c, k) in enumerate(full):
    assert l.tolist() == [i] * (3 + i) and c.tolist() == [float(j + i) for j in range(3 + i)] and k.tolist() == [(i + 1) * j for j in range(18)]
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gather_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ok {r}" in out, out


def test_f16f8_weight_scale_keeps_every_plane_finite():
    """t of the f16f8 weight packing: max|w| * 2^t <= 128 < 448 (e4m3 copies) and max|w| * 2^(t+8) <= 32768 < 65504 (fp16 plane)."""
    for m in (1e-6, 3e-4, 0.02, 0.1, 0.99, 1.0, 1.01, 7.5, 128.0, 300.0, 5e4):
        t = ops.weight_log2_scale(m)
        if -20 < t < 30:                                   # inside the clamp the bound is tight within a factor of two
            assert 64.0 < m * 2.0 ** t <= 128.0, (m, t)
        assert m * 2.0 ** (t + 8) <= 32768.0 or t == -20
    assert ops.weight_log2_scale(0.1) == 10 and ops.weight_log2_scale(1.0) == 7
    assert ops.weight_log2_scale(0.0) == 0 and ops.weight_log2_scale(float("inf")) == 0 and ops.weight_log2_scale(float("nan")) == 0
    assert ops.plane_format("f16f8") == ops.FMT_F16F8 and ops.plane_format("bf16x3") == ops.FMT_BF16 == ops.plane_format("simt")


def test_precision_emulation_of_the_operand_formats():
    """The error model behind the default precision, on the CPU: a.w from (fp16 a_hi, e4m3 pairs) operands is within
    2^-14 of sum|a||w| of the exact product, bf16 {hi, lo} within 2^-15, and a single fp16 pass is ~50x worse."""
    g = torch.Generator().manual_seed(0)
    a = torch.randn((64, 576), generator=g, dtype=torch.float64)
    w = torch.randn((96, 576), generator=g, dtype=torch.float64) * 0.05
    exact, bound = a @ w.T, a.abs() @ w.abs().T
    e4 = lambda x: x.clamp(-448, 448).float().to(torch.float8_e4m3fn).double()
    t = ops.weight_log2_scale(float(w.abs().max()))
    ah, wh = a.half().double(), w.half().double()
    f16f8 = (ah @ (wh * 2.0 ** (t + 8)).T + e4((a - ah) * 256) @ e4(wh * 2.0 ** t).T + e4(ah) @ e4((w - wh) * 2.0 ** (t + 8)).T) * 2.0 ** -(t + 8)
    bh, vh = a.bfloat16().double(), w.bfloat16().double()
    bl, vl = (a - bh).bfloat16().double(), (w - vh).bfloat16().double()
    bf16x3 = bl @ vh.T + bh @ vl.T + bh @ vh.T
    r8, r3, r1 = (((x - exact).abs() / bound).max().item() for x in (f16f8, bf16x3, ah @ wh.T))
    assert r8 < 2.0 ** -14 and r3 < 2.0 ** -15 and r1 > 20 * r8, (r8, r3, r1)


# ------------------------------------------------------------------------------------------------
# the TIFF reader of io.py: classic / Big, both byte orders, strips / tiles, codec fallback, OME names
# ------------------------------------------------------------------------------------------------
def _write_tiff(path, pages, big=False, bo="<", tile=None, description=None):
    """Minimal uncompressed TIFF / BigTIFF writer for the tests: one IFD per 2-D page, strips of 7 rows or `tile` x `tile` tiles."""
    import struct
    off_fmt, cnt_fmt, ent = ("Q", "Q", "HHQQ") if big else ("I", "H", "HHII")
    osz = 8 if big else 4
    blob = bytearray(b"II" if bo == "<" else b"MM")
    blob += struct.pack(bo + "H", 43 if big else 42)
    blob += struct.pack(bo + "HHQ", 8, 0, 0) if big else struct.pack(bo + "I", 0)
    first_ptr = len(blob) - osz
    prev_ptr = first_ptr
    for arr in pages:
        a = np.ascontiguousarray(arr.astype(arr.dtype.newbyteorder(bo)))
        h, w = a.shape
        chunks = []
        if tile:
            for y in range(0, h, tile):
                for x in range(0, w, tile):
                    t = np.zeros((tile, tile), a.dtype)
                    t[:min(tile, h - y), :min(tile, w - x)] = a[y:y + tile, x:x + tile]
                    chunks.append(t.tobytes())
        else:
            chunks = [a[y:y + 7].tobytes() for y in range(0, h, 7)]
        offs = []
        for ch in chunks:
            offs.append(len(blob)); blob += ch
        fmt = {"u": 1, "i": 2, "f": 3}[a.dtype.kind]
        tags = [(256, 4, [w]), (257, 4, [h]), (258, 3, [a.dtype.itemsize * 8]), (259, 3, [1]), (262, 3, [1]), (277, 3, [1]), (339, 3, [fmt])]
        tags += [(322, 4, [tile]), (323, 4, [tile]), (324, 16 if big else 4, offs), (325, 16 if big else 4, [len(c) for c in chunks])] if tile else \
                [(278, 4, [7]), (273, 16 if big else 4, offs), (279, 16 if big else 4, [len(c) for c in chunks])]
        if description is not None:
            tags.append((270, 2, description.encode() + b"\0"))
        tags.sort()
        # out-of-line values first
        vals = {}
        for tag, typ, v in tags:
            code = {2: "c", 3: "H", 4: "I", 16: "Q"}[typ]
            data = bytes(v) if typ == 2 else struct.pack(bo + code * len(v), *v)
            if len(data) > osz:
                if len(blob) % 2: blob += b"\0"
                vals[tag] = (len(blob), None); blob += data
            else:
                vals[tag] = (None, data.ljust(osz, b"\0"))
        if len(blob) % 2: blob += b"\0"
        ifd = len(blob)
        blob[prev_ptr:prev_ptr + osz] = struct.pack(bo + off_fmt, ifd)
        blob += struct.pack(bo + cnt_fmt, len(tags))
        for tag, typ, v in tags:
            pos, inline = vals[tag]
            blob += struct.pack(bo + "HH", tag, typ) + struct.pack(bo + off_fmt, len(v))
            blob += struct.pack(bo + off_fmt, pos) if pos is not None else inline
        prev_ptr = len(blob)
        blob += struct.pack(bo + off_fmt, 0)
    open(path, "wb").write(bytes(blob))


@pytest.mark.parametrize("big,bo,tile,dtype", [(False, "<", None, np.uint16), (False, ">", None, np.uint16), (True, "<", 16, np.float32),
                                               (True, ">", 16, np.int32), (False, "<", 32, np.uint8)])
def test_tiff_reader_uncompressed_layouts(tmp_path, big, bo, tile, dtype):
    from multiplexed_image_annotator_b200 import io as bio
    rng = np.random.default_rng(3)
    stack = (rng.random((4, 37, 53)) * 60000).astype(dtype)
    path = str(tmp_path / "s.tif")
    _write_tiff(path, list(stack), big=big, bo=bo, tile=tile)
    tf = bio.TiffFile(path)
    assert tf.big == big and len(tf.pages) == 4 and all(p.raw_readable and p.tiled == bool(tile) for p in tf.pages)
    got = bio.read_tiff_stack(path, pin=False)
    assert got.dtype == np.dtype(dtype) and np.array_equal(got, stack)
    assert np.array_equal(bio.read_image(path), stack if dtype != np.float64 else stack.astype(np.float32))


def test_tiff_reader_matches_pil_and_falls_back_for_codecs(tmp_path):
    from PIL import Image, features
    from multiplexed_image_annotator_b200 import io as bio
    rng = np.random.default_rng(4)
    stack = (rng.random((5, 40, 64)) * 65535).astype(np.uint16)
    ims = [Image.fromarray(p) for p in stack]
    raw = str(tmp_path / "raw.tif")
    ims[0].save(raw, save_all=True, append_images=ims[1:])
    assert all(p.raw_readable for p in bio.TiffFile(raw).pages)
    assert np.array_equal(bio.read_tiff_stack(raw, pin=False), stack)
    if features.check("libtiff"):
        for comp in ("tiff_lzw", "tiff_adobe_deflate"):
            f = str(tmp_path / f"{comp}.tif")
            ims[0].save(f, save_all=True, append_images=ims[1:], compression=comp)
            assert not bio.TiffFile(f).pages[0].raw_readable
            assert np.array_equal(bio.read_tiff_stack(f, pin=False), stack)          # decoded page by page through PIL
    # a single-page TIFF mask comes back 2-D (imread semantics), an RGB-like PNG mask keeps its first channel
    Image.fromarray(stack[0]).save(str(tmp_path / "m.tif"))
    assert bio.read_mask(str(tmp_path / "m.tif")).shape == (40, 64)
    Image.fromarray(np.stack([stack[0] >> 8] * 3, -1).astype(np.uint8)).save(str(tmp_path / "m.png"))
    assert np.array_equal(bio.read_mask(str(tmp_path / "m.png")), (stack[0] >> 8).astype(np.int32))


def test_ome_channel_names(tmp_path):
    from multiplexed_image_annotator_b200 import io as bio
    xml = ('<?xml version="1.0" encoding="UTF-8"?><OME xmlns="http://www.openmicroscopy.org/Schemas/OME/2016-06"><Image ID="Image:0">'
           '<Pixels ID="Pixels:0" SizeC="3"><Channel ID="Channel:0:0" Name="DAPI"/><Channel ID="Channel:0:1" Name="CD45"/>'
           '<Channel ID="Channel:0:2"/><Channel ID="Channel:0:3" Name="PanCK"/></Pixels></Image></OME>')
    path = str(tmp_path / "o.ome.tif")
    _write_tiff(path, [np.zeros((8, 8), np.uint16)] * 3, description=xml)
    assert bio.ome_channel_names(path) == ["DAPI", "CD45", "PanCK"]
    _write_tiff(path, [np.zeros((8, 8), np.uint16)], description="not xml")
    assert bio.ome_channel_names(path) is None


def test_tiff_reader_property_random_layouts(tmp_path):
    """Random page counts / shapes / sample types / byte orders / strip vs tile layouts / classic vs Big: write with the test
    writer, read back bit-exactly; and a page the reader cannot take raw (here: a corrupted offset table) raises instead of
    returning garbage."""
    from hypothesis import given, settings, strategies as st
    from multiplexed_image_annotator_b200 import io as bio
    path = str(tmp_path / "h.tif")

    @settings(max_examples=25, deadline=None)
    @given(st.integers(1, 4), st.integers(1, 40), st.integers(1, 50), st.sampled_from(["u1", "u2", "i2", "u4", "i4", "f4"]),
           st.booleans(), st.sampled_from(["<", ">"]), st.sampled_from([None, 16, 32]), st.integers(0, 2 ** 31))
    def check(pages, h, w, dt, big, bo, tile, seed):
        rng = np.random.default_rng(seed)
        stack = (rng.random((pages, h, w)) * 200).astype(np.dtype(dt))
        _write_tiff(path, list(stack), big=big, bo=bo, tile=tile)
        got = bio.read_tiff_stack(path, pin=False)
        assert got.dtype == np.dtype(dt) and got.shape == stack.shape and np.array_equal(got, stack)

    check()
    # truncated file: the strip read must fail loudly
    _write_tiff(path, [np.arange(200, dtype=np.uint16).reshape(10, 20)])
    blob = open(path, "rb").read()
    open(path, "wb").write(blob[:8] + blob[8:60])          # header + part of the first strip, IFD gone
    with pytest.raises((ValueError, IOError, Exception)):
        bio.read_tiff_stack(path, pin=False)


@pytest.mark.parametrize("dtype,shape", [("uint16", (3, 17, 9)), ("float32", (5, 4)), ("uint8", (2, 3, 3)), ("int32", (4, 8, 8))])
def test_npy_plane_reader_streams_the_same_array(tmp_path, dtype, shape):
    """io.iter_npy_planes (the .npy counterpart of iter_tiff_planes: readinto one buffer, a plane at a time) returns np.load's
    array, yields every plane index once in order, and refuses layouts it cannot stream (the caller falls back to np.load)."""
    from multiplexed_image_annotator_b200 import io
    a = (np.random.default_rng(3).random(shape) * 200).astype(dtype)
    path = str(tmp_path / "stack.npy")
    np.save(path, a)
    seen, out = [], None
    for k, out in io.iter_npy_planes(path, pin=False):
        seen.append(k)
        assert np.array_equal(out[k], a[k] if a.ndim == 3 else a)        # plane k is complete when it is announced
    assert seen == list(range(shape[0] if len(shape) == 3 else 1))
    assert np.array_equal(out.reshape(a.shape), np.load(path))
    for bad in (np.asfortranarray(np.zeros((2, 3, 4), np.float32)), np.zeros((2, 3, 4), np.float64), np.zeros((2, 2, 2, 2), np.uint8)):
        np.save(path, bad)
        with pytest.raises(ValueError):
            list(io.iter_npy_planes(path, pin=False))
    with open(path, "wb") as f:                                            # truncated payload
        np.lib.format.write_array_header_1_0(f, {"descr": "<u2", "fortran_order": False, "shape": (2, 4, 4)})
        f.write(b"\0" * 40)
    with pytest.raises(ValueError):
        list(io.iter_npy_planes(path, pin=False))
