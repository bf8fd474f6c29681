"""Entry points the Napari plugin calls (reference cta/gui_api.py:13-114): same function names,
argument order and JSON keys, driving the B200 `Annotator`."""
import json
import os

import numpy as np
import pandas as pd

from ..parallel import barrier, is_writer
from .model import Annotator


def _applied(annotator):
    p = annotator.channel_parser
    return p.immune_base or p.immune_extended or p.immune_full or p.struct or p.nerve


def _pipeline(annotator, bs, n_regions, export_before_regions, from_script):
    if not _applied(annotator):
        raise ValueError("No panels are applied. Please check the marker list.")
    annotator.preprocess()
    annotator.predict(bs)
    annotator.generate_heatmap(integrate=True)
    if export_before_regions:
        annotator.export_annotations()
    if n_regions > 0:
        annotator.tissue_region_analysis(n_regions)
    annotator.neighborhood_analysis(integrate=True, normalize=True)
    if not export_before_regions:
        annotator.export_annotations()
    annotator.colorize(from_script=from_script) if from_script else annotator.colorize()
    annotator.cell_type_composition()
    annotator.clear_tmp()


def _intensity_dict(annotator):
    inten = annotator.preprocessor.intensity_full[0]
    out = {i + 1: inten[i] for i in range(len(inten))}
    out[0] = np.zeros_like(inten[0])
    return out, annotator.get_cell_type_names()


def gui_run(marker_list_path, image_path, mask_path, device, main_dir, batch_id, bs, strict, infer, min_cells, n_regions,
            normalize, blur, amax, confidence, cell_size, cell_type_confidence, n_jobs=0):
    path_ = os.path.join(main_dir, "images.csv")
    if is_writer():                               # ranks sharing main_dir (torchrun): rank 0 owns the file
        pd.DataFrame([[image_path, mask_path]]).to_csv(path_, index=False, header=["image_path", "mask_path"])
    barrier()
    annotator = Annotator(marker_list_path, path_, device, main_dir, batch_id, strict, infer, min_cells, normalize, blur,
                          amax, confidence, cell_size, cell_type_confidence, n_jobs=n_jobs)
    _pipeline(annotator, bs, n_regions, export_before_regions=False, from_script=False)
    barrier()
    if is_writer():
        os.remove(path_)
    return _intensity_dict(annotator)


def gui_batch_run(marker_list_path, image_path, device, main_dir, batch_id, bs, strict, infer, min_cells, n_regions,
                  normalize, blur, amax, confidence, cell_size, cell_type_confidence, n_jobs=0):
    annotator = Annotator(marker_list_path, image_path, device, main_dir, batch_id, strict, infer, min_cells, normalize,
                          blur, amax, confidence, cell_size, cell_type_confidence, n_jobs=n_jobs)
    _pipeline(annotator, bs, n_regions, export_before_regions=False, from_script=False)


def _load(path):
    with open(path) as f:
        return json.load(f)


def gui_api(working_addr):
    hp = _load(f"{working_addr}/hyperparams.json")
    return gui_run(hp.get('marker_file'), hp.get('image_file'), hp.get('mask_file'), hp.get('device'), hp.get('main_dir'),
                   "single_run", hp.get('batch_size'), hp.get('strict'), hp.get('infer'), hp.get('min_cells'),
                   hp.get('n_regions'), hp.get('normalize'), hp.get('blur'), hp.get('upper_limit'), hp.get('confidence'),
                   hp.get('cell_size'), hp.get('cell_type_confidence'))


def batch_process(working_dir):
    hp = _load(f"{working_dir}/hyperparams_batch.json")
    gui_batch_run(hp.get('marker_file'), hp.get('csv_file'), hp.get('device'), hp.get('main_dir'), hp.get('batch_id'),
                  hp.get('batch_size'), hp.get('strict'), hp.get('infer'), hp.get('min_cells'), hp.get('n_regions'),
                  hp.get('normalize'), hp.get('blur'), hp.get('upper_limit'), hp.get('confidence'), hp.get('cell_size'),
                  hp.get('cell_type_confidence'))
    with open(f"{working_dir}/output.txt", "w") as file:
        file.write("Batch process completed")
