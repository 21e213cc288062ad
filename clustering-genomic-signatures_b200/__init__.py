"""B200-native all-pairs VLMC distance path (drop-in for the reference's src/vlmc + src/distance
+ src/clustering/util hot path).  See DESIGN.md.

Submodules mirror the reference's module names:
  .vlmc        VLMC                       (src/vlmc/vlmc.pyx)
  .distance    NegativeLogLikelihood, FrobeniusNorm, Projection, EstimateVLMC   (src/distance/*.pyx)
  .clustering  util.calculate_distances_within_vlmcs / index_distances, AverageLinkClustering
  .parse_trees_to_json                    (src/parse_trees_to_json.py)
Device plumbing: .native (ctypes binding of libvgs.so), .engine (handles, sharding).
"""
__version__ = "0.1.0"
