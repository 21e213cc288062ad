"""Taxonomy retrieval metrics of a distance function.

Mirror of ``test_distance_function_`` (``src/test_distance_function.py:50-97``) and of
``update_metrics`` / ``normalise_metrics`` / ``procent_of_taxonomy_in_top`` /
``average_distance_to_taxonomy`` (``src/util/distance_metrics.py:1-60``).  The reference evaluates
``d.distance`` N times per model, sorts the N results and scans them in Python; here the whole
matrix is one device job (``d.distance_matrix``) and the per-row ranking work is one kernel
(``vgs_retrieval_metrics``): no sort, the stable rank of every same-taxon entry is counted directly.

``procent_*_in_top`` values are exact rationals and therefore bit-identical to the reference; the
distance averages are fp64 sums in a fixed device order (the reference adds in sorted order), equal
to ~1e-15 relative.
"""
import numpy as np

from . import native
from .vlmc import _default_device

TAXONOMIES = ("genus", "family")


def _codes(values):
    table = {}
    return np.array([table.setdefault(v, len(table)) for v in values], dtype=np.int32)


def per_model_metrics(d, vlmcs, metadata, device=None, taxonomies=TAXONOMIES):
    """Arrays over ``vlmcs``: {"procent_<tax>_in_top", "distance_to_<tax>"} for every taxonomy and
    "average_distance" -- the per-model values ``update_metrics`` computes (distance_metrics.py:2-17)."""
    import torch
    vlmcs = list(vlmcs)
    device = _default_device() if device is None else device
    d.distance_matrix(vlmcs)
    D = getattr(d, "_D64", None)
    if D is None:
        D = d._D32
    labels = np.stack([_codes([metadata[v.name][t] for v in vlmcs]) for t in taxonomies])
    h = native.Handle(device)
    try:
        Dd = torch.from_numpy(np.ascontiguousarray(D)).to(torch.device("cuda", device))
        in_top, dist_to, avg = h.retrieval_metrics(Dd, labels)
    finally:
        h.close()
    out = {"average_distance": avg}
    for k, t in enumerate(taxonomies):
        out["procent_%s_in_top" % t] = in_top[k]
        out["distance_to_%s" % t] = dist_to[k]
    return out


def distance_function_metrics(d, vlmcs, metadata, device=None):
    """The ``metrics`` dict ``test_distance_function_`` returns (minus the wall-clock ``global_time``)."""
    vlmcs = list(vlmcs)
    per = per_model_metrics(d, vlmcs, metadata, device)
    metrics = {
        "distance_name": d.__class__.__name__,
        "average_procent_of_genus_in_top": 0.0,
        "average_procent_of_family_in_top": 0.0,
        "total_average_distance_to_genus": 0.0,
        "total_average_distance_to_family": 0.0,
        "total_average_distance": 0.0,
    }
    for i in range(len(vlmcs)):  # update_metrics, in list order (the += order is the reference's)
        metrics["procent_genus_in_top"] = float(per["procent_genus_in_top"][i])
        metrics["procent_family_in_top"] = float(per["procent_family_in_top"][i])
        metrics["distance_to_genus"] = float(per["distance_to_genus"][i])
        metrics["distance_to_family"] = float(per["distance_to_family"][i])
        metrics["average_distance"] = float(per["average_distance"][i])
        metrics["average_procent_of_genus_in_top"] += metrics["procent_genus_in_top"]
        metrics["average_procent_of_family_in_top"] += metrics["procent_family_in_top"]
        metrics["total_average_distance_to_genus"] += metrics["distance_to_genus"]
        metrics["total_average_distance_to_family"] += metrics["distance_to_family"]
        metrics["total_average_distance"] += metrics["average_distance"]
    n = len(vlmcs)  # normalise_metrics (distance_metrics.py:29-40)
    metrics["average_procent_of_genus_in_top"] /= n
    metrics["average_procent_of_family_in_top"] /= n
    metrics["total_average_distance"] /= n
    if metrics["total_average_distance"] == 0:
        metrics["total_average_distance_to_genus"] = 0
        metrics["total_average_distance_to_family"] = 0
    else:
        metrics["total_average_distance_to_genus"] /= (n * metrics["total_average_distance"])
        metrics["total_average_distance_to_family"] /= (n * metrics["total_average_distance"])
    return metrics
