"""`Annotator`: the reference's orchestrator (cta/model.py:90-920) with the same constructor order,
methods and public attributes, sequencing the B200 hot path:

    preprocess()  -> ImageProcessor.transform()            stages 1-3 (device resident)
    predict(bs)   -> ViT ensemble forward + softmax          stage 4  (ribca_vit_forward)
                     merge_by_voting + counts                stage 5  (ribca_merge_votes)
                     one all-gather of labels / confidences when several ranks share an image

Reporting methods (heat map, spatial statistics, colourised masks, pie charts) are outside the hot
path (out of scope / "next"); they are kept as callable no-ops that log a line so the
reference's driver scripts run unchanged, except `export_annotations` whose CSV numbers are in scope.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import exact, ops
from ..engine import VitEngine
from ..parallel import all_gather_rows, all_reduce_sum, barrier, broadcast_object, is_writer, world
from ..weights import MODEL_DIR, VIT_SPECS, load_checkpoint
from .logger import Logger
from .markerParse import MarkerParser
from .preprocess import ImageProcessor
from .utils import VOTE_ORDER, get_colors, get_void_vote  # noqa: F401

ALL_TYPES = ["B cell", "CD4 T cell", "CD8 T cell", "Dendritic cell", "Regulatory T cell", "Granulocyte cell",
             "Mast cell", "M1 macrophage cell", "M2 macrophage cell", "Natural killer cell", "Plasma cell",
             "Endothelial cell", "Epithelial cell", "Stroma cell", "Smooth muscle", "Proliferating/tumor cell",
             "Nerve cell", "Others"]                                         # reference model.py:97-99
OTHERS = ALL_TYPES.index("Others")
_VOTE_RANK = [VOTE_ORDER.index(t) if t in VOTE_ORDER else 0 for t in ALL_TYPES]
_STATE_OVERRIDES = {}


def register_state(panel: str, state_dict) -> None:
    """Provide classifier weights in memory instead of `models/<panel>.pth` (tests, benchmarks)."""
    _STATE_OVERRIDES[panel] = state_dict


def merge_branch(panels):
    """The reference's elif chain (model.py:483-636) -> ordered list of the panels that vote.
    full + structure + nerve raises KeyError('Others') exactly as the reference does (quirk Q1);
    with an immune model and structure present, nerve predictions are ignored (Q2)."""
    immune = next((p for p in ("immune_full", "immune_extended", "immune_base") if p in panels), None)
    has_s, has_n = "structure" in panels, "nerve_cell" in panels
    if immune == "immune_full" and has_s and has_n:
        raise KeyError("Others")
    if immune and has_s:
        return [immune, "structure"]
    if has_s and has_n:
        return ["structure", "nerve_cell"]
    if immune and has_n:
        return [immune, "nerve_cell"]
    if immune:
        return [immune]
    if has_s:
        return ["structure"]
    if has_n:
        return ["nerve_cell"]
    raise ValueError("No predictions to merge")


def merge_on_device(probs: dict, confidence, cell_type_confidence=None, want_margin: bool = False):
    """merge_by_voting for one image on the device: probs maps panel -> (n, classes) float32 CUDA
    tensor.  Returns (label uint8 index into ALL_TYPES, conf float32 with -1 for re-labelled cells,
    counts int64[18]) and, with want_margin, every cell's decision margin (exact.refine_labels)."""
    used = merge_branch(probs.keys())
    ctc = cell_type_confidence or {}
    thresh = [float(ctc.get(t, -1)) for t in ALL_TYPES]
    types = [[ALL_TYPES.index(c) for c in VIT_SPECS[p].classes] for p in used]
    p1 = probs[used[1]] if len(used) > 1 else None
    t1 = types[1] if len(used) > 1 else None
    return ops.merge_votes(probs[used[0]].contiguous(), types[0], None if p1 is None else p1.contiguous(), t1,
                           _VOTE_RANK, thresh, confidence, want_margin=want_margin)


class _AnnotationRows:
    """`annotations_all[i]` of the reference (model.py:464-478): one dict per cell with the full
    pixel lists, built on access instead of up front."""

    def __init__(self, annotator, i):
        self._a, self._i = annotator, i

    def __len__(self):
        return len(self._a.annotations[self._i])

    def __getitem__(self, j):
        if isinstance(j, slice):
            return [self[k] for k in range(*j.indices(len(self)))]
        a, i = self._a, self._i
        pos = a.preprocessor.cell_pos_dict[i]
        key = pos._mask.dtype.type(pos.ids[j])
        row, col = pos[key]
        return {"Cell ID": key, "Cell type": int(np.where(a.cell_types == a.annotations[i][j])[0][0]),
                "Confidence": a.confidence[i][j], "Row": row, "Column": col}

    def __iter__(self):
        return (self[j] for j in range(len(self)))

    def spatial_arrays(self):
        """(xy float64 (n, 2) = [mean column, mean row], type index (n,), ids): what spatial_methods reads from the
        dicts, without materialising a pixel list (np.mean of a list == integer sum / count in float64)."""
        a, i = self._a, self._i
        pos = a.preprocessor.cell_pos_dict[i]
        cent = pos.sums.astype(np.float64) / pos.count.astype(np.float64)[:, None]
        lookup = {str(t): k for k, t in enumerate(a.cell_types)}
        types = np.array([lookup[t] for t in a.annotations[i]], dtype=np.int64)
        return np.ascontiguousarray(cent[:, ::-1]), types, [pos._mask.dtype.type(v) for v in pos.ids]


class Annotator(object):
    """Annotator class to predict cell types and tissue structures using the provided models."""

    def __init__(self, marker_list_path, image_path, device, main_dir='./', batch_id='', strict=True, infer=True,
                 min_cells=-1, normalize=True, blur=False, amax=1, confidence=0.25, cell_size=30,
                 cell_type_confidence=None, n_jobs=0):
        self.device = device
        self.cell_types = list(ALL_TYPES)
        self.batch_id = batch_id
        self.logger = Logger(main_dir)
        self.logger.log_all_hyperparameters({
            "Batch name": batch_id, "Strictly match panel(s)": strict, "Normalize image(s)": normalize,
            "Image blurring kernel size": blur, "Percentile of intensity to upper clip": amax,
            "Confidence threshold": confidence, "Estimated cell size (in pixels)": cell_size})
        self.logger.log("")
        self.logger.log("Start parsing the marker list.")
        self.channel_parser = MarkerParser(strict=strict, logger=self.logger)
        self.channel_parser.parse(marker_list_path)
        self.preprocessor = ImageProcessor(image_path, self.channel_parser, main_dir, device, batch_id, infer, normalize,
                                           blur, amax, cell_size, self.logger, n_jobs=n_jobs)
        self._loaded = False
        self.n_jobs = n_jobs
        self._n_images = 0
        self.min_cells = min_cells
        self.annotations, self.confidence = [], []
        self.probs = []                 # per image: {panel: (N, classes) float32 numpy} softmax outputs
        self.labels_index = []          # per image: uint8 numpy, index into ALL_TYPES
        self.type_counts = []           # per image: int64[18]
        self.confidence_thresh = confidence
        self.extra_cell_types = self.min_cells > 0
        self.n_regions = 0
        self.temp_dir = os.path.join(main_dir, "tmp")
        self.result_dir = os.path.join(main_dir, "results")
        os.makedirs(self.result_dir, exist_ok=True)
        self.cell_type_confidence = ({t: -1 for t in ALL_TYPES} if cell_type_confidence is None else cell_type_confidence)
        self.models = {}
        self.precision = os.environ.get("RIBCA_PRECISION", ops.DEFAULT_PRECISION)
        self.exact_labels = exact.LEVELS         # levels of margin-guarded re-evaluation (RIBCA_EXACT_LABELS, default 2)
        self.refine_stats = []                   # per image: exact.RefineStats of this rank's cells

    # ---- stage 1-3 ---------------------------------------------------------------------------------
    def preprocess(self):
        self.preprocessor.transform()
        self._n_images = self.preprocessor._n_images

    def clear(self):
        self.annotations, self.confidence, self.probs, self.labels_index, self.type_counts = [], [], [], [], []

    # ---- stage 4 -----------------------------------------------------------------------------------
    def load_models(self):
        """reference model.py:188-239: CWD-relative checkpoints, one classifier per panel.  Only the
        panels `predict` consumes are loaded."""
        names = {"immune_base": "Immune base", "immune_extended": "Immune extended", "immune_full": "Immune full",
                 "structure": "Tissue structure", "nerve_cell": "Nerve cell"}
        for panel in self.preprocessor.predicted_panels():
            spec = VIT_SPECS[panel]
            path = os.path.join(MODEL_DIR, spec.ckpt)
            if panel in _STATE_OVERRIDES:
                state = _STATE_OVERRIDES[panel]
            elif os.path.exists(path):
                state = load_checkpoint(path)
            else:
                print(f"{names[panel]} model not found")
                self.logger.log(f"{names[panel]} model not found")
                continue
            self.models[panel] = VitEngine(spec, state, device=self.preprocessor.device, precision=self.precision)
        self._loaded = True

    def _predict_cell_types(self, image_idx, model, tensor_name, celltype_dict=None, batch_size=128):
        """model.py:351-426 for this rank's cells: softmax probabilities (n, classes) on the device."""
        pre = self.preprocessor
        cached = pre.patches[image_idx]
        if cached is not None:
            return model.forward(cached[tensor_name])
        parts = [model.forward(batch[tensor_name]) for _, _, batch, _ in pre.patch_chunks(image_idx, [tensor_name])]
        return torch.cat(parts) if parts else torch.empty((0, len(model.spec.classes)), device=pre.device)

    def predict(self, batch_size=32):
        self.logger.log("\nStart predicting cell types and tissue structures.")
        if not self._loaded:
            self.load_models()
        pre = self.preprocessor
        panels = pre.predicted_panels()
        for p in panels:
            if p not in self.models:
                raise AttributeError(f"model for panel {p} is not loaded")        # the reference fails the same way
        for i in range(self._n_images):
            n_total = pre.cells[i].n
            lo, hi = pre.cell_range[i]
            if pre.patches[i] is not None or len(panels) == 1:
                local = {p: self._predict_cell_types(i, self.models[p], p, None, batch_size) for p in panels}
            else:                                # streamed: build every panel's patches once per chunk
                parts = {p: [] for p in panels}
                for _, _, batch, _ in pre.patch_chunks(i, panels):
                    for p in panels:
                        parts[p].append(self.models[p].forward(batch[p]))
                local = {p: torch.cat(v) if v else torch.empty((0, len(VIT_SPECS[p].classes)), device=pre.device)
                         for p, v in parts.items()}
            # stage 5 on this rank's cells (+ the margin-guarded re-evaluation that makes the labels independent of the
            # fast path's rounding, exact.py), then the single gather
            label, conf, counts, _, stats = exact.refine_labels(
                local, lambda pr, want_margin=True: merge_on_device(pr, self.confidence_thresh, self.cell_type_confidence,
                                                                    want_margin=want_margin),
                lambda sel, prec, i=i: {p: self.models[p].forward(t, precision=prec)
                                        for p, t in pre.patches_of_cells(i, sel, panels, prec).items()},
                levels=self.exact_labels)
            self.refine_stats.append(stats)
            pre.release(i)          # a stack that was re-read because it is outside the residency budget goes again
            label = all_gather_rows(label, n_total, lo, hi)
            conf = all_gather_rows(conf, n_total, lo, hi)
            counts = all_reduce_sum(counts)
            probs = {p: all_gather_rows(v, n_total, lo, hi).cpu().numpy() for p, v in local.items()}
            self.probs.append(probs)
            lab = label.cpu().numpy()
            cf = conf.cpu().numpy()
            self.labels_index.append(lab)
            self.type_counts.append(counts.cpu().numpy())
            self.annotations.append([ALL_TYPES[k] for k in lab.tolist()])
            self.confidence.append([-1 if (k == OTHERS and c == -1.0) else c for k, c in zip(lab.tolist(), cf)])
        self.logger.log("Finished predicting cell types and tissue structures.")
        if self.extra_cell_types:
            self.logger.log("min_cells > 0: extra cell-type discovery (UMAP + HDBSCAN) is outside the B200 hot path; skipped.")
        self.cell_types = self._get_unique_cell_types()
        self.cell_types = np.delete(self.cell_types, np.where(self.cell_types == "Others"))
        self.cell_types = np.append(self.cell_types, "Others")
        self.colors = get_colors(len(self.cell_types))
        self.annotations_all = [_AnnotationRows(self, i) for i in range(len(self.annotations))]

    def merge_by_voting(self):
        """Kept for API parity: `predict` already merged on the device (ribca_merge_votes)."""
        if not self.annotations:
            raise ValueError("No predictions to merge")

    def _get_unique_cell_types(self):
        present = sorted({ALL_TYPES[k] for lab in self.labels_index for k in np.unique(lab).tolist()})
        return np.sort(np.array(present))

    def get_cell_type_names(self):
        txt = ""
        for i in range(len(self.cell_types)):
            txt += f"{i + 1}: {self.cell_types[i]}"
            txt += "\n" if i % 3 == 2 else "  "
        return txt

    def prediction_dicts(self, image_idx, panel):
        """The reference's per-cell {type name: np.float32} dicts (model.py:412-414), on demand."""
        names = VIT_SPECS[panel].classes
        return [{names[k]: row[k] for k in range(len(names))} for row in self.probs[image_idx][panel]]

    # ---- result assembly (in scope: the numbers of the CSV and the counts) -----------------------------
    def export_annotations(self):
        """reference model.py:768-795, same file name, header, rounding and formatting."""
        if len(self.annotations) == 0:
            raise ValueError("No annotations to export")
        for i in range(len(self.annotations) if is_writer() else 0):      # every rank holds the same results: rank 0 writes
            f = os.path.join(self.result_dir, f"{self.batch_id}_annotation_{i}.csv")
            pos = self.preprocessor.cell_pos_dict[i]
            cent = pos.sums.astype(np.float64) / pos.count.astype(np.float64)[:, None]     # = np.mean of the lists
            with open(f, "w") as file:
                file.write("Cell Index,Cell Type,Confidence,Row,Column,Tissue Region\n")
                for j, key in enumerate(pos.ids.tolist()):
                    conf = round(self.confidence[i][j], 3)
                    row = round(cent[j, 0], 2)
                    col = round(cent[j, 1], 2)
                    region = "Region " + str(self.tissue_regions[i][key]) if hasattr(self, 'tissue_regions') else None
                    file.write(f"{key},{self.annotations[i][j]},{conf},{row},{col},{region}\n")
            self.logger.log(f"Exported annotations for image {i} to {f}")
        barrier()

    def cell_type_composition(self, reduction=True, integrate=False):
        """reference model.py:861-912: the per-type counts / fractions (the pie chart itself is
        presentation and is not drawn).  Returns a list of {type: value} per image (or one dict)."""
        if len(self.annotations) == 0:
            raise ValueError("No annotations to analyze")
        per_image = []
        for counts in self.type_counts:
            d = {str(t): int(counts[ALL_TYPES.index(str(t))]) for t in self.cell_types}
            per_image.append(d)
        if integrate:
            tot = {k: sum(d[k] for d in per_image) for k in per_image[0]}
            n = sum(tot.values())
            return {k: v / n for k, v in tot.items()} if reduction else tot
        if reduction:
            per_image = [{k: v / max(sum(d.values()), 1) for k, v in d.items()} for d in per_image]
        return per_image

    # ---- reporting outside the hot path: callable, logged, not drawn -----------------------------------
    def _skipped(self, what):
        self.logger.log(f"{what}: reporting step outside the B200 hot path, not produced by this build.")

    def generate_heatmap(self, integrate=False):
        """reference model.py:700-741: mean marker intensity per predicted cell type.  The matrices are kept in
        `self.heatmaps` ([(cell type names, (types, markers) float64)], one per image or one integrated); the PNG is drawn
        only when matplotlib + seaborn are importable (they are presentation, absent from the offline image)."""
        if len(self.annotations) == 0:
            raise ValueError("No annotations to generate heatmap")
        groups = [list(range(len(self.annotations)))] if integrate else [[i] for i in range(len(self.annotations))]
        self.heatmaps = []
        for g, images in enumerate(groups):
            names = np.concatenate([np.asarray(self.annotations[i]) for i in images])
            inten = np.concatenate([np.asarray(self.preprocessor.intensity_full[i]) for i in images], axis=0)
            assert len(inten) == len(names)
            celltypes = np.unique(names)
            colormap = np.stack([np.mean(inten[names == t], axis=0) for t in celltypes]) if len(celltypes) else np.zeros((0, inten.shape[1]))
            self.heatmaps.append((celltypes, colormap))
            f = os.path.join(self.result_dir, f"{self.batch_id}_Integrated_heatmap.png" if integrate else f"{self.batch_id}_heatmap_{g}.png")
            try:
                if not is_writer():
                    continue
                import matplotlib.pyplot as plt
                import seaborn as sns
            except ImportError:
                self.logger.log(f"generate_heatmap: matplotlib / seaborn not installed, {os.path.basename(f)} not drawn (matrix kept in .heatmaps).")
                continue
            plt.figure(figsize=(max(colormap.shape[1] // 4, 1), max(colormap.shape[0] // 4, 1)))
            sns.heatmap(colormap, cmap='vlag', xticklabels=self.channel_parser.markers, yticklabels=celltypes, linewidth=.5)
            plt.tight_layout()
            plt.savefig(f)
            plt.close()

    def neighborhood_analysis(self, n_neighbors=25, integrate=True, normalize=True):
        """reference model.py:798-800: the neighbourhood matrix CSV(s), from the GPU k-NN (the heat-map PNG is not drawn)."""
        from .spatial_methods import neighborhood_analysis
        out = neighborhood_analysis(self.annotations_all, n_neighbors=n_neighbors, cell_types=self.cell_types, integrate=integrate,
                                    normalize=normalize, result_dir=self.result_dir if is_writer() else None, batch_id=self.batch_id,
                                    device=self.preprocessor.device)
        barrier()
        return out

    def tissue_region_analysis(self, n, method="kmeans"):
        """reference model.py:802-804."""
        from .spatial_methods import tissue_region_partition
        self.n_regions = n
        # the reference's KMeans is unseeded: rank 0 clusters, every rank gets ITS labels
        regions = tissue_region_partition(self.annotations_all, n, self.n_jobs, method=method,
                                          device=self.preprocessor.device) if is_writer() else None
        self.tissue_regions = broadcast_object(regions)

    def colorize(self, from_script=False):
        """reference model.py:806-858: colourised label map, confidence map and (GUI runs) the uint8 label
        map `output_img.png` the Napari widget loads; painted on the device from the per-cell results,
        plus the tissue-region maps when tissue_region_analysis ran (n_regions > 0)."""
        from PIL import Image
        from .utils import number_to_rgb
        pre = self.preprocessor
        if len(pre.masks) == 0:
            raise ValueError("No masks to colorize")
        if len(self.annotations) == 0:
            raise ValueError("No annotations to colorize")
        dev = pre.device
        for i in range(len(pre.masks) if is_writer() else 0):
            type_of_all = np.array([int(np.where(self.cell_types == t)[0][0]) if t in self.cell_types else 0 for t in ALL_TYPES])
            idx = torch.from_numpy(type_of_all[self.labels_index[i]].astype(np.uint8)).to(dev)           # index into cell_types
            rgb = torch.tensor(self.colors, dtype=torch.uint8, device=dev)[idx.long()]
            conf = [number_to_rgb(c) if c > 0 else [192, 192, 192] for c in self.confidence[i]]
            conf = torch.tensor(conf, dtype=torch.uint8, device=dev).reshape(-1, 3)
            m, cells = pre.mask_dev(i), pre.cells[i]
            Image.fromarray(ops.paint_cells(m, cells, rgb.contiguous()).cpu().numpy()).save(
                os.path.join(self.result_dir, f"{self.batch_id}_colorized_annotation_{i}.png"))
            Image.fromarray(ops.paint_cells(m, cells, conf.contiguous()).cpu().numpy()).save(
                os.path.join(self.result_dir, f"{self.batch_id}_confidence_{i}.png"))
            if not from_script:
                gui_dir = "./src/multiplexed_image_annotator/cell_type_annotation/_working_dir_temp"
                if os.path.isdir(gui_dir):
                    Image.fromarray(ops.paint_cells(m, cells, (idx + 1).contiguous()).cpu().numpy()).save(os.path.join(gui_dir, "output_img.png"))
            if self.n_regions > 0 and hasattr(self, "tissue_regions"):
                reg = np.array([int(self.tissue_regions[i][k]) for k in pre.cell_pos_dict[i].ids.tolist()], dtype=np.int64)
                reg_d = torch.from_numpy(reg.astype(np.uint8)).to(dev)
                tcol = torch.tensor(get_colors(self.n_regions + 1), dtype=torch.uint8, device=dev)[reg_d.long()]
                Image.fromarray(ops.paint_cells(m, cells, tcol.contiguous()).cpu().numpy()).save(
                    os.path.join(self.result_dir, f"{self.batch_id}_tissue_region_{i}.png"))
                if not from_script and os.path.isdir("./src/multiplexed_image_annotator/cell_type_annotation/_working_dir_temp"):
                    Image.fromarray(ops.paint_cells(m, cells, (reg_d + 1).contiguous()).cpu().numpy()).save(
                        "./src/multiplexed_image_annotator/cell_type_annotation/_working_dir_temp/output_img_2.png")
            pre.release(i)
        barrier()

    def umap_visualization(self):
        self._skipped("umap_visualization")

    def clear_tmp(self):
        barrier()                                   # no rank is still working in main_dir
        if is_writer() and os.path.isdir(self.temp_dir):
            for f in os.listdir(self.temp_dir):
                try:
                    os.remove(os.path.join(self.temp_dir, f))
                except FileNotFoundError:
                    pass
            try:
                os.rmdir(self.temp_dir)
            except (FileNotFoundError, OSError):
                pass
        self.logger.log("Temporary files cleared")
        barrier()
