#!/usr/bin/env python
"""Golden vectors for the matrix consumers of row f4 -- TEST INFRASTRUCTURE ONLY.

Executes the reference's own ``test_distance_function_`` / ``distance_metrics`` (plain Python,
imported from /root/reference/src) and its compiled ``AverageLinkClustering`` + ``ClusteringMetrics``
(oracle/_ref) on a small synthetic set with genus / family labels, and commits inputs + outputs as
``tests/golden/metrics_golden.npz``.  Run in the build container:

    python oracle/build_ref.py && python oracle/make_goldens_metrics.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from clustering_genomic_signatures_b200 import synthetic  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    ns = ref_loader.load()
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.lines", "mysql", "mysql.connector"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.lines"].Line2D = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].lines = sys.modules["matplotlib.lines"]
    sys.modules["mysql"].connector = sys.modules["mysql.connector"]
    sys.path.insert(0, os.path.join(REF, "src"))
    for k in ("NaiveParameterSampling", "StationaryDistribution", "FixedLengthSequenceKLDivergence", "PSTMatching"):
        setattr(sys.modules["distance"], k, object)
    import test_distance_function as tdf
    from util import distance_metrics as dm

    n, n_fam = 40, 4
    fm, fam = synthetic.make_family_models(n, 31, 5, 60, n_fam, noise=0.35)
    vlmcs = [ns.VLMC(fm.tree(m), fm.names[m], {}) for m in range(n)]
    # genus = a finer split of the family; two models share a name-independent species
    genus = np.array([int(f) * 10 + (i % 3) for i, f in enumerate(fam)], dtype=np.int32)
    # a few cross-family "mislabelled" genera so that precision / sensitivity are not trivially 1
    genus[5] = genus[25]
    genus[17] = genus[33]
    metadata = {v.name: {"species": v.name, "family": "fam%d" % fam[i], "genus": "gen%d" % genus[i]} for i, v in enumerate(vlmcs)}
    out = {"ctx_offsets": fm.ctx_offsets, "ctx_len": fm.ctx_len, "ctx_code": fm.ctx_code, "prob": fm.prob,
           "family": np.asarray(fam, dtype=np.int32), "genus": genus}

    for dname, d in (("frobenius", ns.FrobeniusNorm()), ("union", ns.FrobeniusNorm(True))):
        D = np.array([[d.distance(a, b) for b in vlmcs] for a in vlmcs], dtype=np.float64)
        out[dname + "_D64"] = D
        # per-model values, exactly as update_metrics computes them
        per = {k: [] for k in ("procent_genus_in_top", "procent_family_in_top", "distance_to_genus", "distance_to_family",
                                "average_distance")}
        for v in vlmcs:
            sorted_results, _ = tdf.calculate_distances(d, v, vlmcs)
            per["procent_genus_in_top"].append(dm.procent_of_taxonomy_in_top(v, vlmcs, sorted_results, metadata, "genus"))
            per["procent_family_in_top"].append(dm.procent_of_taxonomy_in_top(v, vlmcs, sorted_results, metadata, "family"))
            per["distance_to_genus"].append(dm.average_distance_to_taxonomy(v, sorted_results, metadata, "genus"))
            per["distance_to_family"].append(dm.average_distance_to_taxonomy(v, sorted_results, metadata, "family"))
            per["average_distance"].append(sum([x for x, _ in sorted_results]) / len(sorted_results))
        for k, vals in per.items():
            out["%s_%s" % (dname, k)] = np.array(vals, dtype=np.float64)
        m = tdf.test_distance_function_(d, vlmcs, vlmcs, metadata, None, do_print_metrics=False)
        for k, val in m.items():
            if k not in ("distance_name", "global_time"):
                out["%s_tdf_%s" % (dname, k)] = np.float64(val)

    # a matrix with exact ties and duplicates: rounds the Frobenius distances to one decimal
    Dt = np.round(out["frobenius_D64"], 1)
    out["ties_D64"] = Dt

    class Fixed(object):  # a distance object the reference's drivers accept (duck typed)
        def __init__(self, M, vl):
            self.M, self.row = M, {id(v): i for i, v in enumerate(vl)}

        def distance(self, a, b):
            return float(self.M[self.row[id(a)], self.row[id(b)]])

    dt = Fixed(Dt, vlmcs)
    for tax in ("genus", "family"):
        vals = []
        for v in vlmcs:
            sorted_results, _ = tdf.calculate_distances(dt, v, vlmcs)
            vals.append(dm.procent_of_taxonomy_in_top(v, vlmcs, sorted_results, metadata, tax))
        out["ties_procent_%s_in_top" % tax] = np.array(vals)

    # clustering metrics from the compiled reference
    import networkx as nx
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        d = ns.FrobeniusNorm()
        cl = ns.AverageLinkClustering(vlmcs, d, metadata)
        for k in (7, 4, 1):  # the reference continues merging from its previous state: decreasing k only
            met = cl.cluster(k)
            labels = np.zeros(n, dtype=np.int32)
            comps = sorted(sorted(vlmcs.index(v) for v in comp) for comp in nx.connected_components(met.G))
            for ci, comp in enumerate(comps):
                labels[comp] = ci
            pre = "k%d_" % k
            out[pre + "labels"] = labels
            if k > 1:
                sil = met.silhouette_metric()
                out[pre + "silhouette"] = np.array([sil[v.name] for v in vlmcs], dtype=np.float64)
                out[pre + "average_silhouette"] = np.float64(met.average_silhouette())
            for tax in ("genus", "family"):
                out[pre + "same_" + tax] = np.float64(met.average_percent_same_taxonomy(tax))
                if k > 1:
                    out[pre + "sens_prec_" + tax] = np.array(met.sensitivity_specificity(tax), dtype=np.float64)
            out[pre + "size_metrics"] = np.array([float(x) for x in met.cluster_size_metrics()])
            out[pre + "distance_std"] = np.float64(met.average_distance_std(d))
            out[pre + "merge_distances"] = np.array(met.get_merge_distances(), dtype=np.float64)
        out["indexed_f32"] = np.array([[np.float32(d.distance(a, b)) for b in vlmcs] for a in vlmcs], dtype=np.float32)
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(GOLD, "metrics_golden.npz"), **out)
    print("metrics_golden.npz", os.path.getsize(os.path.join(GOLD, "metrics_golden.npz")))


if __name__ == "__main__":
    main()
