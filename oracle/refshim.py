"""TEST INFRASTRUCTURE - import the UNMODIFIED reference .py files in the build container.

Only `tests/golden/make_golden.py` (run by hand in the build container, where /root/reference is
mounted) uses this.  Nothing that runs on the GPU box may import it: /root/reference does not
exist there.  The reference's own modules are imported as they are; only the third-party packages
that are missing from this image are replaced by the stand-ins of `oracle/standins.py`
(scikit-image, timm) or by inert mocks (matplotlib, seaborn, umap, tifffile: imported by the
reference at module import time but never reached on the hot path).
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("RIBCA_REFERENCE_ROOT", "/root/reference")
_PKG = "src.multiplexed_image_annotator.cell_type_annotation"


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_standins() -> None:
    from oracle import standins

    if "skimage" not in sys.modules:
        import numpy as np
        from PIL import Image

        def imread(path):
            path = str(path)
            if path.endswith(".npy"):
                return np.load(path)
            return np.array(Image.open(path))

        sk = _module("skimage")
        sk.io = _module("skimage.io", imread=imread)
        sk.morphology = _module("skimage.morphology", dilation=standins.dilation, disk=standins.disk)
        sk.filters = _module("skimage.filters", gaussian=standins.gaussian)
        sk.transform = _module("skimage.transform", resize=standins.resize)
    if "timm" not in sys.modules:
        tm = _module("timm")
        tm.models = _module("timm.models")
        tm.models.vision_transformer = _module(
            "timm.models.vision_transformer", VisionTransformer=standins.VisionTransformer,
            PatchEmbed=standins.PatchEmbed, Block=standins.Block)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors",
                 "seaborn", "umap", "tifffile"):
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)
    plt = sys.modules["matplotlib.pyplot"]
    if isinstance(plt, mock.MagicMock):          # `fig, ax = plt.subplots(...)` in reference utils.py:118
        plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())
        sys.modules["matplotlib"].pyplot = plt


def load_reference():
    """Return the reference's cell_type_annotation modules as a namespace.

    The package `src.multiplexed_image_annotator` is pre-registered empty so its __init__ (which
    pulls in napari/magicgui) is skipped; the sub-package modules are imported as-is.
    """
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}; refshim is build-container only")
    _install_standins()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name, sub in (("src", "src"), ("src.multiplexed_image_annotator", "src/multiplexed_image_annotator"),
                      (_PKG, "src/multiplexed_image_annotator/cell_type_annotation")):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(REFERENCE_ROOT, sub)]
            sys.modules[name] = pkg
    ns = types.SimpleNamespace()
    for mod in ("utils", "markerParse", "markerImputer", "preprocess", "logger", "model"):
        setattr(ns, mod, importlib.import_module(f"{_PKG}.{mod}"))
    return ns
