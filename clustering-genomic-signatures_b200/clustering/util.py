"""Drop-in for ``src/clustering/util.pyx``: the two matrix hand-off layouts of the reference.

    calculate_distances_within_vlmcs(vlmcs, d) -> float32 [N*N, 3] rows (i, j, d(vlmcs[i], vlmcs[j]))
                                                  row-major in (i, j)          (src/clustering/util.pyx:6-21)
    index_distances(vlmcs, distances)          -> float32 [N, N]               (src/clustering/util.pyx:23-33)

The reference fills the first with N^2 Python calls of ``d.distance``; here ``d`` is one of the
device-backed classes of ``..distance`` and the whole matrix is a single device job.
"""
import numpy as np

FLOATTYPE = np.float32


def calculate_distances_within_vlmcs(vlmcs, d):
    if not hasattr(d, "distance_matrix"):
        raise TypeError("d must be a device-backed distance (clustering_genomic_signatures_b200.distance.*); "
                        "this path has no per-pair CPU loop")
    vlmcs = list(vlmcs)
    n = len(vlmcs)
    D32 = np.ascontiguousarray(d.distance_matrix(vlmcs), dtype=FLOATTYPE)
    out = np.empty((n * n, 3), dtype=FLOATTYPE)
    idx = np.arange(n, dtype=FLOATTYPE)       # float32 indices exactly like the reference (:15-18)
    out[:, 0] = np.repeat(idx, n)
    out[:, 1] = np.tile(idx, n)
    out[:, 2] = D32.ravel()
    return out


def index_distances(vlmcs, distances):
    n = len(vlmcs)
    out = np.ndarray([n, n], dtype=FLOATTYPE)
    rows = distances[:, 0].astype(np.int64)
    cols = distances[:, 1].astype(np.int64)
    out[rows, cols] = distances[:, 2]
    return out
