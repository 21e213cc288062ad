"""Cluster quality metrics on the device-resident matrix.

Mirror of ``ClusteringMetrics`` (``src/clustering/clustering_metrics.pyx``): same method names and
return shapes.  The reference walks Python sets of VLMC objects and calls ``list.index`` per distance
(O(N^3) for the silhouette); here the O(N^2) passes over the float32 ``[N, N]`` matrix
(silhouette, within-cluster standard deviation) are device kernels behind ``libvgs.so`` and the
label-only counts (taxonomy purity, sensitivity / precision, cluster sizes) are contingency-table
arithmetic on the labels.

The reference accumulates the silhouette sums in numpy float32 in *set iteration order* (id-hash
order of the VLMC objects, not reproducible between runs); the kernels accumulate in fp64 in a
fixed order and round to float32 where the reference returns a C float, so values agree to float32
rounding of the sums (tests: 2e-6 absolute on s_i).
"""
import numpy as np

from .. import native
from ..vlmc import _default_device
from .util import FLOATTYPE


def _codes(values):
    """Dense int32 codes for arbitrary hashable labels (first appearance order)."""
    table = {}
    return np.array([table.setdefault(v, len(table)) for v in values], dtype=np.int32)


class ClusteringMetrics(object):
    """``ClusteringMetrics(G, distance_mean, indexed_distances, vlmcs, metadata, merge_distances)`` in the
    reference (clustering_metrics.pyx:21-28).  ``G`` may be a networkx graph whose connected components are
    the clusters (as in the reference) or an integer label array of length N."""

    def __init__(self, G, distance_mean, indexed_distances, vlmcs, metadata, merge_distances, device=None):
        self.G = G
        self.distance_mean = distance_mean
        self.indexed_distances = np.ascontiguousarray(indexed_distances, dtype=FLOATTYPE)
        self.vlmcs = list(vlmcs)
        self.metadata = metadata
        self.merge_distances = list(merge_distances) if merge_distances is not None else []
        self.device = _default_device() if device is None else device
        n = len(self.vlmcs)
        if hasattr(G, "nodes"):
            import networkx as nx
            row = {id(v): i for i, v in enumerate(self.vlmcs)}
            labels = np.zeros(n, dtype=np.int32)
            for c, comp in enumerate(nx.connected_components(G)):
                for v in comp:
                    labels[row[id(v)]] = c
            self.labels = labels
        else:
            self.labels = np.ascontiguousarray(G, dtype=np.int32)
            assert self.labels.shape == (n,)
        self._sil = None

    # -- device passes ------------------------------------------------------------------------
    def _device_matrix(self):
        import torch
        return torch.from_numpy(self.indexed_distances).to(torch.device("cuda", self.device))

    def silhouette_values(self):
        """(a, b, s) arrays in ``vlmcs`` order."""
        if self._sil is None:
            h = native.Handle(self.device)
            try:
                self._sil = h.silhouette(self._device_matrix(), self.labels)
            finally:
                h.close()
        return self._sil

    def silhouette_metric(self):  # clustering_metrics.pyx:35-62: dict name -> s_i
        s = self.silhouette_values()[2]
        return {v.name: float(s[i]) for i, v in enumerate(self.vlmcs)}

    def average_silhouette(self):  # :29-33
        s = self.silhouette_values()[2]
        return float(sum(float(x) for x in s) / len(s))

    def average_distance_std(self, distance_function=None):  # :168-185
        """The reference re-evaluates ``distance_function.distance`` for every ordered pair inside each cluster;
        here the distances come from the function's device-computed fp64 matrix when it has one for these models,
        else from ``indexed_distances``."""
        import torch
        D = None
        vl = getattr(distance_function, "_vlmcs", None)
        if distance_function is not None and getattr(distance_function, "_D64", None) is not None \
                and vl is not None and len(vl) == len(self.vlmcs) and all(a is b for a, b in zip(vl, self.vlmcs)):
            # the function's batch is exactly these models in this order: its fp64 matrix is what its distance() returns
            D = torch.from_numpy(np.ascontiguousarray(distance_function._D64)).to(torch.device("cuda", self.device))
        if D is None:
            D = self._device_matrix()
        h = native.Handle(self.device)
        try:
            stds = h.cluster_distance_std(D, self.labels)
        finally:
            h.close()
        return float(np.float32(sum(float(x) for x in stds) / len(stds)))  # cpdef float in the reference

    # -- label-only metrics -------------------------------------------------------------------
    def _taxon_codes(self, key):
        return _codes([self.metadata[v.name][key] for v in self.vlmcs])

    def _contingency(self, key):
        tax = self._taxon_codes(key)
        lab = _codes(self.labels.tolist())
        table = np.zeros((int(lab.max()) + 1, int(tax.max()) + 1), dtype=np.int64)
        np.add.at(table, (lab, tax), 1)
        return table

    def percent_same_taxonomy(self, connected_component, taxonomy):  # :95-101
        tax = [self.metadata[v.name][taxonomy] for v in connected_component]
        counts = {}
        for t in tax:
            counts[t] = counts.get(t, 0) + 1
        size = len(tax)
        return sum(counts[t] for t in tax) / (size ** 2)

    def average_percent_same_taxonomy(self, taxonomy):  # :86-93
        table = self._contingency(taxonomy)
        sizes = table.sum(axis=1)
        average = 0.0
        for c in range(table.shape[0]):
            average += (float((table[c] ** 2).sum()) / float(sizes[c] ** 2)) * float(sizes[c])
        return average / len(self.vlmcs)

    def sensitivity_specificity(self, meta_key):  # :113-120 (returns sensitivity, precision like the reference)
        table = self._contingency(meta_key)
        sizes = table.sum(axis=1)
        tp = int((table * (table - 1)).sum())                 # ordered pairs, same cluster, same taxon
        fp = int((sizes * (sizes - 1)).sum()) - tp
        per_taxon = table.sum(axis=0)
        same_taxon_pairs = int((per_taxon * (per_taxon - 1)).sum())
        fn = same_taxon_pairs - tp                            # ordered pairs, different clusters, same taxon
        sensitivity = tp / (tp + fn)
        precision = tp / (tp + fp)
        return sensitivity, precision

    def cluster_size_metrics(self):  # :151-154
        sizes = np.bincount(_codes(self.labels.tolist()))
        return sizes.mean(), np.median(sizes), sizes.min(), sizes.max()

    def get_merge_distances(self):
        return self.merge_distances

    def get_latest_merge_distance(self):
        # cpdef float in the reference (:161-165)
        return float(np.float32(self.merge_distances[-1])) if len(self.merge_distances) > 0 else -1
