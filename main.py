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


def run(marker_list_path, image_path, mask_path, device, main_dir, batch_id, bs, strict, infer, min_cells, n_regions,
        normalize, blur, amax, confidence, cell_size, cell_type_confidence, n_jobs):
    path_ = os.path.join(main_dir, "images.csv")
    pd.DataFrame([[image_path, mask_path]]).to_csv(path_, index=False, header=["image_path", "mask_path"])
    annotator = Annotator(marker_list_path, path_, device, main_dir, batch_id, strict, infer, min_cells, normalize, blur,
                          amax, confidence, cell_size, cell_type_confidence, n_jobs=n_jobs)
    _pipeline(annotator, bs, n_regions, export_before_regions=True, from_script=True)
    return _intensity_dict(annotator)


def batch_run(marker_list_path, image_path, device, main_dir, batch_id, bs, strict, infer, min_cells, n_regions,
              normalize, blur, amax, confidence, cell_size, cell_type_confidence, n_jobs=0):
    annotator = Annotator(marker_list_path, image_path, device, main_dir, batch_id, strict, infer, min_cells, normalize,
                          blur, amax, confidence, cell_size, cell_type_confidence, n_jobs=n_jobs)
    _pipeline(annotator, bs, n_regions, export_before_regions=True, from_script=True)


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description='Process images with markers')
    ap.add_argument('--marker-list-path', type=str, required=True, help='Path to the markers text file')
    ap.add_argument('--device', type=str, default='cuda', help='Device to run on (cuda)')
    ap.add_argument('--main-dir', type=str, default='./', help='Main directory path')
    ap.add_argument('--batch-id', type=str, required=True, help='Batch identifier')
    ap.add_argument('--strict', action='store_true', help='Enable strict mode')
    ap.add_argument('--infer', action='store_true', default=True, help='Enable inference')
    ap.add_argument('--min-cells', type=int, default=-1, help='Minimum number of cells')
    ap.add_argument('--n-regions', type=int, default=3, help='Number of regions')
    ap.add_argument('--normalize', action='store_true', default=True, help='Enable normalization')
    ap.add_argument('--blur', type=float, default=0.3, help='Blur factor')
    ap.add_argument('--amax', type=float, default=99.8, help='Maximum amplitude')
    ap.add_argument('--confidence', type=float, default=0.3, help='Confidence threshold')
    ap.add_argument('--cell-type-confidence', type=float, default=None, help='Cell type confidence threshold')
    ap.add_argument('--bs', type=int, default=128, help='Batch size')
    ap.add_argument('--cell-size', type=int, default=30, help='Cell size')
    ap.add_argument('--n_jobs', type=int, default=0, help='Cell size')
    group = ap.add_mutually_exclusive_group(required=True)
    group.add_argument('--image-path', type=str, help='Path to single image file')
    group.add_argument('--batch-csv', type=str, help='Path to CSV file for batch processing')
    ap.add_argument('--mask-path', type=str, help='Path to mask file (required for single image mode)')
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
