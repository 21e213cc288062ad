"""Scalable average-link clustering with the reference's exact results (row f1).

Mirror of ``AverageLinkClustering(vlmcs, d, metadata).cluster(k)``
(``src/clustering/average_link_clustering.pyx``, ``src/clustering/graph_based_clustering.pyx:16-53``):
the constructor builds the distance matrix with ``d`` (one device job), ``cluster(k)`` merges until
``k`` clusters remain.  The reference keeps a heap and a dict per cluster (about 0.4 kB per pair:
~160 GB at N = 20 000); here the merge loop runs on the device over one fp64 matrix and reproduces
the reference's merge order, tie-breaking and mixed float32/float64 arithmetic bit for bit.
"""
import numpy as np

from .. import native
from ..vlmc import _default_device
from .util import FLOATTYPE


class ClusterResult(object):
    def __init__(self, labels, merge_distances, vlmcs):
        self.labels = labels
        self.merge_distances = list(merge_distances)
        n_clusters = int(labels.max()) + 1 if len(labels) else 0
        self.clusters = [[i for i in range(len(labels)) if labels[i] == c] for c in range(n_clusters)]
        self.vlmcs = vlmcs

    def get_merge_distances(self):  # name used by the reference's ClusteringMetrics (:158-159)
        return self.merge_distances

    def named_clusters(self):
        return [[self.vlmcs[i].name for i in comp] for comp in self.clusters]


class AverageLinkClustering(object):
    def __init__(self, vlmcs, d, metadata=None, file_name=None, device=None):
        self.vlmcs = list(vlmcs)
        self.d = d
        self.metadata = metadata
        self.device = _default_device() if device is None else device
        self.indexed_distances = np.ascontiguousarray(d.distance_matrix(self.vlmcs), dtype=FLOATTYPE)
        if file_name is not None:  # the reference always writes 'cluster_distances.npy' into the CWD (:19, :111)
            from .util import calculate_distances_within_vlmcs
            np.save(file_name, calculate_distances_within_vlmcs(self.vlmcs, d))

    @classmethod
    def from_matrix(cls, D32, vlmcs=None, device=None):
        self = cls.__new__(cls)
        self.indexed_distances = np.ascontiguousarray(D32, dtype=FLOATTYPE)
        self.vlmcs = list(vlmcs) if vlmcs is not None else list(range(self.indexed_distances.shape[0]))
        self.d, self.metadata = None, None
        self.device = _default_device() if device is None else device
        return self

    def cluster(self, clusters):
        import torch
        h = native.Handle(self.device)
        try:
            D = torch.from_numpy(self.indexed_distances).to(torch.device("cuda", self.device))
            labels, merges = h.average_link(D, int(clusters))
        finally:
            h.close()
        return ClusterResult(labels, merges, self.vlmcs)
