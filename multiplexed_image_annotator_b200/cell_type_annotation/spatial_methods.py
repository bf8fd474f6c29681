"""Spatial statistics on the labelled cells, behind the reference's function names
(cta/spatial_methods.py:13-198: `neighborhood_analysis`, `tissue_region_partition`).

What changed under the interface: the reference rebuilds every centroid with np.mean over the pixel lists and asks
a scikit-learn ball tree for the neighbours of one cell at a time from Python; here the centroids come from the
device cell table (exact integer sums / counts, the same float64 values) and ribca_knn_2d + ribca_neighbor_stats
produce the neighbour lists, the type-by-type neighbourhood matrix and the multi-scale neighbour compositions on the
GPU.  The CSV files keep the reference's names and formatting; the heat-map PNGs are presentation and are not drawn.
PCA + KMeans of the tissue regions stay scikit-learn calls on the host, as in the reference (KMeans is unseeded there,
so region labels are only defined up to the clustering's randomness).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import ops

REGION_LEVELS = [10, 20, 30, 50, 75, 100, 150, 200]          # spatial_methods.py:150
REGION_KNN = 201                                              # spatial_methods.py:153


def _per_image(annotation_all, i):
    """(xy float64 (n, 2) [x = mean column, y = mean row], types int (n,), ids) of image i.  `annotation_all[i]` is
    either the Annotator's lazy rows (fast path: the cell table) or the reference's list of dicts."""
    rows = annotation_all[i]
    fast = getattr(rows, "spatial_arrays", None)
    if fast is not None:
        return fast()
    xy = np.array([[np.mean(r["Column"]), np.mean(r["Row"])] for r in rows], dtype=np.float64).reshape(-1, 2)
    return xy, np.array([r["Cell type"] for r in rows]).astype(int), [r["Cell ID"] for r in rows]


def neighborhood_matrix(xy, types, n_types, n_neighbors, device="cuda"):
    """counts[a][b] = number of (cell of type a, one of its n_neighbors - 1 nearest other cells of type b) pairs
    (spatial_methods.py:35-40)."""
    xy_d = torch.as_tensor(xy, dtype=torch.float64, device=device)
    nbr = ops.knn_2d(xy_d, n_neighbors)
    mat, _ = ops.neighbor_stats(nbr, torch.as_tensor(types, device=device), n_types, skip=1)
    return mat.cpu().numpy().astype(np.float64)


def _write_matrix_csv(path, neighborhood, cell_types):
    with open(path, "w") as file:                              # spatial_methods.py:62-73: same text, cell by cell
        file.write("cell_type,")
        for name in cell_types:
            file.write(f"{name},")
        file.write("\n")
        for a in range(len(cell_types)):
            file.write(f"{cell_types[a]},")
            for b in range(len(cell_types)):
                file.write(f"{neighborhood[a][b]:.3f},")
            file.write("\n")


def neighborhood_analysis(annotation_all, n_neighbors=10, cell_types=None, integrate=False, normalize=True, batch_id=None,
                          result_dir=None, device="cuda"):
    n_types = len(cell_types)
    mats = []
    for i in range(len(annotation_all)):
        xy, types, _ = _per_image(annotation_all, i)
        mats.append(neighborhood_matrix(xy, types, n_types, n_neighbors, device))
    if integrate:
        mats = [np.sum(mats, axis=0)] if mats else [np.zeros((n_types, n_types))]
    out = []
    for i, m in enumerate(mats):
        if normalize:
            s = m.sum(1, keepdims=True)
            m = np.divide(m, s, out=m.copy(), where=s > 0)
        name = f"{batch_id}_integrated_neighborhood.csv" if integrate else f"{batch_id}_neighborhood_{i}.csv"
        if result_dir is not None:
            _write_matrix_csv(os.path.join(result_dir, name), m, cell_types)
        out.append(m)
    return out[0] if integrate else out


def neighbor_compositions(xy, types, device="cuda"):
    """(n, 8 * n_celltypes) float64 feature matrix of tissue_region_partition (spatial_methods.py:150-178)."""
    n_celltypes = int(np.max(types)) + 1
    xy_d = torch.as_tensor(xy, dtype=torch.float64, device=device)
    nbr = ops.knn_2d(xy_d, REGION_KNN)
    _, comp = ops.neighbor_stats(nbr, torch.as_tensor(types, device=device), n_celltypes, levels=REGION_LEVELS, skip=1,
                                 want_matrix=False)
    return comp.cpu().numpy()


def tissue_region_partition(annotation_all, n_clusters=3, n_jobs=0, method="kmeans", device="cuda"):
    from sklearn.cluster import HDBSCAN, KMeans, SpectralClustering
    from sklearn.decomposition import PCA
    tissue_labels = []
    for i in range(len(annotation_all)):
        xy, types, ids = _per_image(annotation_all, i)
        compositions = neighbor_compositions(xy, types, device)
        n_jobs = n_jobs if n_jobs is not None and n_jobs > 0 else None
        compositions = PCA(n_components=0.99).fit_transform(compositions)
        if method == "kmeans":
            clusterer = KMeans(n_clusters=n_clusters)
        elif method == "hdbscan":
            clusterer = HDBSCAN(n_clusters=n_clusters)          # (the reference passes n_clusters here too and would fail)
        elif method == "spectral":
            clusterer = SpectralClustering(n_clusters=n_clusters, n_jobs=n_jobs)
        else:
            raise UnboundLocalError(f"unknown tissue-region method {method!r}")
        cluster_labels = clusterer.fit_predict(compositions)
        tissue_labels.append({id_: cluster_labels[j] for j, id_ in enumerate(ids)})
    return tissue_labels
