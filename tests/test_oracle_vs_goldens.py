"""Pins the oracle (oracle/vlmc_oracle.py and oracle/vgs_oracle.c) against the golden vectors that
were produced by executing the compiled reference (oracle/make_goldens.py).  CPU only."""
import random

import numpy as np
import pytest

import c_oracle
import golden_io
import vlmc_oracle as O
from clustering_genomic_signatures_b200.flat import encode_history


@pytest.fixture(scope="module")
def g():
    return golden_io.fixtures()


@pytest.fixture(scope="module")
def co(g):
    return c_oracle.COracle(golden_io.fixture_flat(g))


def test_orders(g, co):
    assert [O.order_of(t) for t in g["trees"]] == g["orders"] == [co.order(m) for m in range(5)]


def test_lookup_indices_bit_exact(g, co):
    for r in g["lookups"]:
        t, m = g["trees"][r["model"]], r["model"]
        keys = list(t)
        try:
            cid = keys.index(O.get_context(t, g["orders"][m], r["history"]))
            allc = [keys.index(c) for c in O.get_all_contexts(t, g["orders"][m], r["history"])]
        except RuntimeError:
            cid, allc = -1, []
        assert cid == r["ctx_id"] and allc == r["all_ctx_ids"]
        ln, code = encode_history(r["history"])
        assert co.lookup(m, code, ln) == r["ctx_id"]
        mask = co.lookup_all(m, code, ln)
        got = sorted((keys.index(r["history"][len(r["history"]) - l:] if l else "") for l in range(16) if mask >> l & 1),
                     key=lambda i: -len(keys[i]))
        assert got == r["all_ctx_ids"]


def test_scores_and_nll_bit_exact(g, co):
    M_ref = np.array(g["nll_scores"])
    M = O.score_matrix(g["trees"], g["nll_sequences"])
    assert np.array_equal(M, M_ref)
    codes = np.stack([golden_io.str_codes(s) for s in g["nll_sequences"]])
    Mc = co.score_matrix(codes)
    assert np.array_equal(Mc, M_ref)
    assert np.array_equal(co.score_matrix(codes, skip=0), np.array(g["loglik_full_scores"]))
    D_ref = np.array(g["nll_matrix_f64"])
    assert np.array_equal(O.nll_from_scores(M, g["nll_length"]), D_ref)
    assert np.array_equal(co.nll_from_scores(Mc, g["nll_length"]), D_ref)
    # the [N^2, 3] float32 hand-off layout of src/clustering/util.pyx:6-21
    tri = np.array(g["nll_triples_f32"], dtype=np.float32)
    assert np.array_equal(O.distances_within(D_ref), tri)
    assert np.array_equal(O.index_distances(5, tri), D_ref.astype(np.float32))


def test_full_loglik_python(g):
    M0 = np.array([[O.log_likelihood(t, o, s, 0) for t, o in zip(g["trees"], g["orders"])] for s in g["nll_sequences"][:2]])
    assert np.array_equal(M0, np.array(g["loglik_full_scores"])[:2])


@pytest.mark.parametrize("mode,key", [("intersection", "frobenius_intersection"), ("union", "frobenius_union"),
                                      ("projection", "projection")])
def test_pair_distances(g, co, mode, key):
    ref = np.array(g[key])
    if mode == "projection":
        py, dim = O.projection_matrix(g["trees"])
        assert dim == g["projection_dimension"]
    else:
        py = O.frobenius_matrix(g["trees"], mode == "union")
    c = co.pair_matrix(mode)
    # the reference computes in float32 (BLAS sdot): it agrees with the exact-sum restatement to ~1e-7 rel
    assert np.allclose(py, ref, rtol=3e-7, atol=0)
    # the two restatements agree with each other to fp64 rounding
    assert np.allclose(c, py, rtol=1e-13, atol=0)
    assert np.all(np.diag(c) == 0)


def test_sampler_stream_exact(g, co):
    random.seed(g["sampler_seed"])
    seqs = [O.generate_sequence(t, o, 200, 500, random) for t, o in zip(g["trees"], g["orders"])]
    assert seqs == g["sampler_sequences"]
    random.seed(g["sampler_seed"])
    st = np.array(random.getstate()[1], dtype=np.uint32)
    u = c_oracle.mt_uniforms(st.copy(), 8)
    assert u.tolist() == g["sampler_uniforms_head"]
    for m in range(5):
        assert golden_io.codes_str(co.sample(m, st, 200)) == g["sampler_sequences"][m]


def test_sampler_from_a_context_exact(g):
    """generate_sequence_from with non-empty contexts (shorter / longer than the order, a symbol outside ACGT):
    goldens from the compiled reference (oracle/make_goldens_sampler_from.py)."""
    for c in golden_io.sampler_from_cases():
        m = c["model"]
        random.seed(c["seed"])
        assert O.generate_sequence_from(g["trees"][m], g["orders"][m], c["length"], c["context"], random) == c["sequence"]
        assert random.random() == c["next_random"]


def test_likelihood_and_context_distribution_exact(g):
    for c in golden_io.sampler_from_cases("likelihood"):
        m = c["model"]
        assert O.likelihood(g["trees"][m], g["orders"][m], c["sequence"]) == c["likelihood"]
    for c in golden_io.sampler_from_cases("context_distribution"):
        m = c["model"]
        random.seed(c["seed"])
        seq = O.generate_sequence(g["trees"][m], g["orders"][m], c["length"], 500, random)
        assert seq == c["sequence"]
        d = O.context_distribution(g["trees"][m], g["orders"][m], seq, c["length"])
        assert list(d.keys()) == c["contexts"] and list(d.values()) == c["values"]


def test_estimate(g, co):
    seqs, E = golden_io.fixture_estimate()
    Ec = co.estimate_matrix(seqs)
    assert np.allclose(Ec, E, rtol=1e-12, atol=0)
    # python restatement on two (left, right) pairs incl. the non-suffix-closed model
    for i, j in ((0, 1), (2, 2), (4, 3)):
        assert O.estimate_distance(g["trees"][i], golden_io.codes_str(seqs[j])) == pytest.approx(E[i, j], rel=1e-13)


def test_numpy_pairwise_sum_restated():
    rng = np.random.default_rng(5)
    for n in (0, 1, 7, 8, 9, 127, 128, 129, 1000, 4097):
        a = rng.random(n) * rng.choice([1e-6, 1.0, 1e3], size=n)
        assert c_oracle.np_sum(a) == float(np.sum(a))


def test_average_link_restated(g):
    ref = g["clustering_AverageLinkClustering"]
    D32 = np.array(g["frobenius_intersection"]).astype(np.float32)
    for k in ("3", "2"):
        labels, merges = O.average_link(D32, int(k))
        comps = sorted(sorted(np.nonzero(labels == c)[0].tolist()) for c in set(labels.tolist()))
        assert comps == ref[k]["clusters"]
        assert merges == ref[k]["merge_distances"][: len(merges)]


def test_c_average_link_equals_pinned_restatement(g):
    """oracle/vgs_oracle.c:vo_average_link (O(n^2), used at large n) == vlmc_oracle.average_link (pinned above)
    including forced ties and a non-symmetric matrix."""
    rng = np.random.default_rng(0)
    for n, k, q, sym in ((12, 3, 0, True), (40, 5, 0, True), (60, 4, 16, True), (50, 1, 4, True), (30, 3, 0, False)):
        X = rng.random((n, 6))
        D = np.sqrt(((X[:, None] - X[None]) ** 2).sum(-1)).astype(np.float32)
        if q:
            D = (np.round(D * q) / q).astype(np.float32)
        if not sym:
            D = rng.random((n, n)).astype(np.float32)
        np.fill_diagonal(D, 0)
        l1, m1 = O.average_link(D, k)
        l2, m2 = c_oracle.average_link(D, k)
        assert np.array_equal(l1, l2) and np.array_equal(np.array(m1), m2)
    D32 = np.array(g["frobenius_intersection"]).astype(np.float32)
    labels, merges = c_oracle.average_link(D32, 3)
    assert merges.tolist() == g["clustering_AverageLinkClustering"]["3"]["merge_distances"]


@pytest.mark.parametrize("name", ["shallow", "deep"])
def test_synthetic_small(name):
    fm, seqs, est, z = golden_io.synthetic_small(name)
    co = c_oracle.COracle(fm)
    assert np.array_equal(co.score_matrix(seqs), z["scores"])
    assert np.array_equal(co.score_matrix(seqs, skip=0), z["scores_full"])
    assert np.array_equal(co.nll_matrix(seqs), z["nll"])
    assert np.allclose(co.pair_matrix("intersection"), z["frobenius_intersection"], rtol=3e-7)
    assert np.allclose(co.pair_matrix("union"), z["frobenius_union"], rtol=3e-7)
    assert np.allclose(co.pair_matrix("projection"), z["projection"], rtol=3e-7)
    sub = c_oracle.COracle(fm.subset(range(4)))
    assert np.allclose(sub.estimate_matrix(est), z["estimate"], rtol=1e-12)
    trees = [fm.tree(m) for m in range(fm.n_models)]
    for dn, D in (("frob", z["frobenius_intersection"]), ("nll", z["nll"])):
        labels, merges = O.average_link(D.astype(np.float32), int(z["family"].max()) + 1)
        assert np.array_equal(labels, z["avglink_%s_labels" % dn])
        assert np.array_equal(np.array(merges), z["avglink_%s_merges" % dn])
    # python oracle on a corner of the set
    assert np.array_equal(O.score_matrix(trees[:3], [golden_io.codes_str(s) for s in seqs[:2]]), z["scores"][:2, :3])


def test_synthetic_generator_is_deterministic():
    from clustering_genomic_signatures_b200 import synthetic
    fm, _, _, z = golden_io.synthetic_small("shallow")
    fm2, fam = synthetic.make_family_models(12, 11, 6, 120, 3)
    assert np.array_equal(fm2.ctx_code, fm.ctx_code) and np.array_equal(fm2.prob, fm.prob)
    assert np.array_equal(fam, z["family"])


def test_tree_parser_restated():
    import json, os
    with open(os.path.join(golden_io.GOLD, "tree_parse_golden.json")) as fh:
        gold = json.load(fh)
    for name, item in gold.items():
        tree, occ = O.parse_tree_text(item["text"])
        assert list(tree.keys()) == list(item["json"]["tree"].keys())
        assert tree == item["json"]["tree"]
        assert occ == item["json"]["occurrence_probability"]
