"""Matrix assembly for the reference's clustering code (src/clustering/util.pyx) and a scalable
average-link consumer (src/clustering/average_link_clustering.pyx)."""
from .util import calculate_distances_within_vlmcs, index_distances  # noqa: F401
from .average_link import AverageLinkClustering, ClusterResult  # noqa: F401,E402
