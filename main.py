"""Command-line driver with the reference's flags and defaults (reference main.py:56-157):

    python main.py --marker-list-path markers.txt --batch-id run1 \
        (--image-path img.tif --mask-path mask.png | --batch-csv images.csv) [--device cuda] ...

`run` / `batch_run` keep the reference's signatures and the fixed post-processing sequence of
main.py:19-28; the hot path underneath is libribca_b200.so.
"""
import argparse
import os

import pandas as pd

from multiplexed_image_annotator_b200.cell_type_annotation.gui_api import _intensity_dict, _pipeline
from multiplexed_image_annotator_b200.cell_type_annotation.model import Annotator
from multiplexed_image_annotator_b200.parallel import barrier, is_writer


def run(marker_list_path, image_path, mask_path, device, main_dir, batch_id, bs, strict, infer, min_cells, n_regions,
        normalize, blur, amax, confidence, cell_size, cell_type_confidence, n_jobs):
    path_ = os.path.join(main_dir, "images.csv")
    if is_writer():                               # ranks sharing main_dir (torchrun): rank 0 owns the file
        pd.DataFrame([[image_path, mask_path]]).to_csv(path_, index=False, header=["image_path", "mask_path"])
    barrier()
    annotator = Annotator(marker_list_path, path_, device, main_dir, batch_id, strict, infer, min_cells, normalize, blur,
                          amax, confidence, cell_size, cell_type_confidence, n_jobs=n_jobs)
    _pipeline(annotator, bs, n_regions, export_before_regions=True, from_script=True)
    return _intensity_dict(annotator)


def batch_run(marker_list_path, image_path, device, main_dir, batch_id, bs, strict, infer, min_cells, n_regions,
              normalize, blur, amax, confidence, cell_size, cell_type_confidence, n_jobs=0):
    annotator = Annotator(marker_list_path, image_path, device, main_dir, batch_id, strict, infer, min_cells, normalize,
                          blur, amax, confidence, cell_size, cell_type_confidence, n_jobs=n_jobs)
    _pipeline(annotator, bs, n_regions, export_before_regions=True, from_script=True)


# flag, type (None = store_true), default, required  -- names and defaults of reference main.py:60-104
_OPTIONS = (
    ("--marker-list-path", str, None, True), ("--device", str, "cuda", False), ("--main-dir", str, "./", False),
    ("--batch-id", str, None, True), ("--strict", None, False, False), ("--infer", None, True, False),
    ("--min-cells", int, -1, False), ("--n-regions", int, 3, False), ("--normalize", None, True, False),
    ("--blur", float, 0.3, False), ("--amax", float, 99.8, False), ("--confidence", float, 0.3, False),
    ("--cell-type-confidence", float, None, False), ("--bs", int, 128, False), ("--cell-size", int, 30, False),
    ("--n_jobs", int, 0, False),
)


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="RIBCA cell-type annotation on B200 (reference-compatible flags)")
    for flag, kind, default, required in _OPTIONS:
        if kind is None:                                   # `store_true` with default True cannot be switched off,
            ap.add_argument(flag, action="store_true", default=default)       # exactly as in the reference CLI
        else:
            ap.add_argument(flag, type=kind, default=default, required=required)
    src = ap.add_mutually_exclusive_group(required=True)
    src.add_argument("--image-path", type=str, help="single image (needs --mask-path)")
    src.add_argument("--batch-csv", type=str, help="CSV with columns image_path,mask_path")
    ap.add_argument("--mask-path", type=str)
    args = ap.parse_args(argv)
    if args.image_path and not args.mask_path:
        ap.error("--mask-path is required when using --image-path")
    return args


if __name__ == "__main__":
    a = parse_args()
    common = dict(marker_list_path=a.marker_list_path, device=a.device, main_dir=a.main_dir, batch_id=a.batch_id, bs=a.bs,
                  strict=a.strict, infer=a.infer, min_cells=a.min_cells, n_regions=a.n_regions, normalize=a.normalize,
                  blur=a.blur, amax=a.amax, confidence=a.confidence, cell_size=a.cell_size,
                  cell_type_confidence=a.cell_type_confidence, n_jobs=a.n_jobs)
    if a.batch_csv:
        batch_run(image_path=a.batch_csv, **common)
    else:
        run(image_path=a.image_path, mask_path=a.mask_path, **common)
