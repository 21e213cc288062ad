#!/usr/bin/env python
"""Generate ``tests/golden/*`` by executing the compiled reference -- TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference and ``oracle/_ref`` from ``build_ref.py``):

    python oracle/build_ref.py && python oracle/make_goldens.py

The reference pins no numeric answers itself, so parity is pinned by these
vectors: inputs (models, sampled sequences) and the outputs of the reference's own
``VLMC.get_context`` / ``log_likelihood*`` / ``generate_sequence`` and of its
``NegativeLogLikelihood`` / ``FrobeniusNorm`` / ``Projection`` / ``EstimateVLMC`` /
``calculate_distances_within_vlmcs`` / ``AverageLinkClustering`` on them.  ``/root/reference``
does not exist on the GPU box, so the vectors are committed; this script is how they were made.
"""
import json
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from clustering_genomic_signatures_b200 import synthetic  # noqa: E402
from clustering_genomic_signatures_b200.flat import FlatModels, pack_sequences, SYMBOL_CODE  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def seq_codes(s):
    return np.frombuffer(s.encode("ascii"), dtype=np.uint8).copy()


def str_to_codes(s):
    lut = np.zeros(256, dtype=np.uint8)
    for ch, c in SYMBOL_CODE.items():
        lut[ord(ch)] = c
    return lut[seq_codes(s)]


def ref_vlmcs_from_flat(ns, fm):
    return [ns.VLMC(fm.tree(m), fm.names[m], {}) for m in range(fm.n_models)]


def full_matrix(d, vlmcs):
    n = len(vlmcs)
    D = np.zeros((n, n), dtype=np.float64)
    for i in range(n):
        for j in range(n):
            D[i, j] = d.distance(vlmcs[i], vlmcs[j])
    return D


def lookups(vlmcs, rng, per_model, max_len):
    rows = []
    for m, v in enumerate(vlmcs):
        keys = list(v.tree.keys())
        for _ in range(per_model):
            ln = int(rng.integers(0, max_len + 1))
            hist = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, size=ln))
            if rng.random() < 0.3 and keys:  # bias towards histories ending in a real context
                hist = hist + keys[int(rng.integers(len(keys)))]
            try:
                ctx = v.get_context(hist)
                cid = keys.index(ctx)
                allc = [keys.index(c) for c in v.get_all_contexts(hist)]
            except RuntimeError:
                cid, allc = -1, []
            rows.append({"model": m, "history": hist, "ctx_id": cid, "all_ctx_ids": allc})
    return rows


def clusters_of(metrics_G, vlmcs):
    import networkx as nx
    comps = [sorted(vlmcs.index(v) for v in comp) for comp in nx.connected_components(metrics_G)]
    return sorted(comps)


def fixture_goldens(ns):
    vlmcs = ref_loader.load_fixture_vlmcs(os.path.join(REF, "original_test_trees"))
    models = [{"name": v.name, "contexts": list(v.tree.keys()),
               "probs": [[v.tree[c][ch] for ch in "ACGT"] for c in v.tree]} for v in vlmcs]
    out = {"models": models}
    out["orders"] = [int(v.order) for v in vlmcs]

    out["frobenius_intersection"] = full_matrix(ns.FrobeniusNorm(), vlmcs).tolist()
    out["frobenius_union"] = full_matrix(ns.FrobeniusNorm(True), vlmcs).tolist()
    p = ns.Projection()
    p.set_vlmcs(vlmcs)
    out["projection"] = [[float(p.distance(a, b)) for b in vlmcs] for a in vlmcs]
    out["projection_dimension"] = int(p.dimension)

    # NLL, seed 1234, L = 1000, full loop in list order
    random.seed(1234)
    for v in vlmcs:
        v.reset_sequence()
    d = ns.NegativeLogLikelihood(1000)
    triples = ns.calculate_distances_within_vlmcs(vlmcs, d)
    seqs = [v.sequence for v in vlmcs]
    out["nll_seed"] = 1234
    out["nll_length"] = 1000
    out["nll_sequences"] = seqs
    out["nll_triples_f32"] = [[float(x) for x in row] for row in triples]
    out["nll_matrix_f64"] = full_matrix(d, vlmcs).tolist()
    out["nll_scores"] = [[v.log_likelihood_ignore_initial_bias(s) for v in vlmcs] for s in seqs]
    out["loglik_full_scores"] = [[v.log_likelihood(s) for v in vlmcs] for s in seqs]

    rng = np.random.Generator(np.random.PCG64(77))
    out["lookups"] = lookups(vlmcs, rng, 120, 12)

    # generate_sequence_from-style sampler check: the MT19937 doubles each sequence consumed
    random.seed(99)
    u = [random.random() for _ in range(5 * 700)]
    random.seed(99)
    for v in vlmcs:
        v.reset_sequence()
    out["sampler_seed"] = 99
    out["sampler_sequences"] = [v.generate_sequence(200, 500) for v in vlmcs]
    out["sampler_uniforms_head"] = u[:8]

    # test_distance_function_ with example_test.py-style metadata (src/example_test.py:39-44)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.lines", "mysql", "mysql.connector"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib.lines"].Line2D = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].lines = sys.modules["matplotlib.lines"]
    sys.modules["mysql"].connector = sys.modules["mysql.connector"]
    sys.path.insert(0, os.path.join(REF, "src"))
    sys.modules["distance"].NaiveParameterSampling = object
    sys.modules["distance"].StationaryDistribution = object
    sys.modules["distance"].FixedLengthSequenceKLDivergence = object
    sys.modules["distance"].PSTMatching = object
    import test_distance_function as tdf  # the reference's own driver (plain Python)
    metadata = {v.name: {n: v.name for n in ["species", "family", "genus"]} for v in vlmcs}
    m = tdf.test_distance_function_(ns.FrobeniusNorm(), vlmcs, vlmcs, metadata, None, do_print_metrics=False)
    out["tdf_metrics_frobenius"] = {k: (float(v) if not isinstance(v, str) else v) for k, v in m.items() if k != "global_time"}
    m = tdf.test_distance_function_(d, vlmcs, vlmcs, metadata, None, do_print_metrics=False)
    out["tdf_metrics_nll"] = {k: (float(v) if not isinstance(v, str) else v) for k, v in m.items() if k != "global_time"}

    # clustering on Frobenius (intersection)
    cwd = os.getcwd()
    os.chdir("/tmp")  # the ctor writes cluster_distances.npy into CWD (graph_based_clustering.pyx:111)
    try:
        for cls_name in ("AverageLinkClustering", "MSTClustering"):
            cl = getattr(ns, cls_name)(vlmcs, ns.FrobeniusNorm(), metadata)
            res = {}
            for k in (3, 2):
                met = cl.cluster(k)
                res[str(k)] = {"clusters": clusters_of(met.G, vlmcs),
                               "merge_distances": [float(x) for x in met.get_merge_distances()],
                               "average_silhouette": float(met.average_silhouette())}
            out["clustering_" + cls_name] = res
    finally:
        os.chdir(cwd)

    with open(os.path.join(GOLD, "fixtures_golden.json"), "w") as fh:
        json.dump(out, fh)

    # Estimate: 100 000-symbol sequences are bulky -> separate compressed npz
    random.seed(3)
    for v in vlmcs:
        v.reset_sequence()
    e = ns.EstimateVLMC()
    with ref_loader.patched_power_divergence():
        E = np.zeros((5, 5))
        # first use as *right* model samples its sequence: iterate right-major so the sampling order is list order
        for j in range(5):
            for i in range(5):
                E[i, j] = e.distance(vlmcs[i], vlmcs[j])
    words, offs, lens = pack_sequences([str_to_codes(v.sequence) for v in vlmcs])
    np.savez_compressed(os.path.join(GOLD, "fixtures_estimate.npz"), estimate=E, seq_words=words,
                        seq_word_offsets=offs, seq_lens=lens, seed=3)

    # .tree ingest: the first 60 lines of each real .tree and what the reference parser makes of them
    import parse_trees_to_json as ptj
    trees = {}
    for f in sorted(os.listdir(os.path.join(REF, "original_test_trees"))):
        if f.endswith(".tree"):
            with open(os.path.join(REF, "original_test_trees", f)) as fh:
                head = "".join(fh.readlines()[:60])
            tmp = "/tmp/_vgs_head.tree"
            with open(tmp, "w") as fh:
                fh.write(head)
            trees[f] = {"text": head, "json": json.loads(ptj._parse_file(tmp, False))}
    with open(os.path.join(GOLD, "tree_parse_golden.json"), "w") as fh:
        json.dump(trees, fh)
    return vlmcs


def synthetic_goldens(ns):
    """Small synthetic sets: shallow (depth 6, dense tier), deep (depth 9, hashed tier)."""
    sets = {
        "shallow": synthetic.make_family_models(12, 11, 6, 120, 3),
        "deep": synthetic.make_family_models(8, 12, 9, 300, 2),
        "mixed": None,
    }
    out = {}
    for name, val in sets.items():
        if val is None:
            continue
        fm, fam = val
        vlmcs = ref_vlmcs_from_flat(ns, fm)
        L = 600
        random.seed(21)
        d = ns.NegativeLogLikelihood(L)
        D = full_matrix(d, vlmcs)
        seqs = [v.sequence for v in vlmcs]
        M = np.array([[v.log_likelihood_ignore_initial_bias(s) for v in vlmcs] for s in seqs])
        M0 = np.array([[v.log_likelihood(s) for v in vlmcs] for s in seqs])
        p = ns.Projection()
        p.set_vlmcs(vlmcs)
        P = np.array([[float(p.distance(a, b)) for b in vlmcs] for a in vlmcs])
        words, offs, lens = pack_sequences([str_to_codes(s) for s in seqs])
        out.update({
            name + "_ctx_offsets": fm.ctx_offsets, name + "_ctx_len": fm.ctx_len, name + "_ctx_code": fm.ctx_code,
            name + "_prob": fm.prob, name + "_family": fam,
            name + "_nll_length": L, name + "_nll": D, name + "_scores": M, name + "_scores_full": M0,
            name + "_seq_words": words, name + "_seq_word_offsets": offs, name + "_seq_lens": lens,
            name + "_frobenius_intersection": full_matrix(ns.FrobeniusNorm(), vlmcs),
            name + "_frobenius_union": full_matrix(ns.FrobeniusNorm(True), vlmcs),
            name + "_projection": P,
        })
        # Estimate on the first 4 models (the reference hard-codes L = 100 000, estimate.pyx:27-28)
        sub = vlmcs[:4]
        random.seed(22)
        for v in sub:
            v.reset_sequence()
        e = ns.EstimateVLMC()
        E = np.zeros((4, 4))
        with ref_loader.patched_power_divergence():
            for j in range(4):
                for i in range(4):
                    E[i, j] = e.distance(sub[i], sub[j])
        ew, eo, el = pack_sequences([str_to_codes(v.sequence) for v in sub])
        out.update({name + "_estimate": E, name + "_est_seq_words": ew, name + "_est_seq_word_offsets": eo,
                    name + "_est_seq_lens": el})
        # average-link assignments from the reference on its own NLL / Frobenius matrices
        metadata = {v.name: {n: v.name for n in ["species", "family", "genus"]} for v in vlmcs}
        cwd = os.getcwd()
        os.chdir("/tmp")
        try:
            import networkx as nx
            for v, s in zip(vlmcs, seqs):
                v.sequence = s
            for dn, dd in (("frob", ns.FrobeniusNorm()), ("nll", d)):
                cl = ns.AverageLinkClustering(vlmcs, dd, metadata)
                k = int(fam.max()) + 1
                met = cl.cluster(k)
                labels = np.zeros(len(vlmcs), dtype=np.int64)
                comps = sorted(sorted(vlmcs.index(v) for v in comp) for comp in nx.connected_components(met.G))
                for ci, comp in enumerate(comps):
                    labels[comp] = ci
                out[name + "_avglink_%s_labels" % dn] = labels
                out[name + "_avglink_%s_merges" % dn] = np.array(met.get_merge_distances(), dtype=np.float64)
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(GOLD, "synthetic_small.npz"), **out)


if __name__ == "__main__":
    if not ref_loader.available():
        import build_ref
        assert build_ref.build(REF), "cannot build oracle/_ref"
    ns = ref_loader.load()
    os.makedirs(GOLD, exist_ok=True)
    fixture_goldens(ns)
    synthetic_goldens(ns)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))
