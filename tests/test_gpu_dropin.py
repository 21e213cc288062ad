"""The drop-in boundary on a B200: the host-side mirror classes (VLMC, NegativeLogLikelihood,
FrobeniusNorm, Projection, EstimateVLMC, clustering.util) behave like the reference's, the
reference's own consumer accepts them, and the multi-rank tile path gives the same matrix bits."""
import random

import numpy as np
import pytest

import c_oracle
import golden_io
import vlmc_oracle as O
from clustering_genomic_signatures_b200 import native, synthetic
from clustering_genomic_signatures_b200.clustering import util
from clustering_genomic_signatures_b200.distance import EstimateVLMC, FrobeniusNorm, NegativeLogLikelihood, Projection
from clustering_genomic_signatures_b200.engine import MatrixJob
from clustering_genomic_signatures_b200.flat import pack_sequences
from clustering_genomic_signatures_b200.vlmc import VLMC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    return golden_io.fixtures()


def fixture_vlmcs(g):
    return [VLMC(t, n, {}) for t, n in zip(g["trees"], g["names"])]


def test_vlmc_lookup_methods(g):
    vl = fixture_vlmcs(g)
    for r in g["lookups"][::7]:
        v = vl[r["model"]]
        keys = g["models"][r["model"]]["contexts"]
        if r["ctx_id"] < 0:
            with pytest.raises(RuntimeError, match="get_context"):
                v.get_context(r["history"])
            continue
        assert v.get_context(r["history"]) == keys[r["ctx_id"]]
        assert v.get_all_contexts(r["history"]) == [keys[i] for i in r["all_ctx_ids"]]
    # SURVEY App. A.2 examples on the non-suffix-closed fixture
    v = vl[2]
    for hist, want in (("CA", "A"), ("TTACG", "ACG"), ("GACGTACGT", "ACGTACGT"), ("ACGTACGTA", "ACGTA"), ("CGT", "")):
        assert v.get_context(hist) == want
    with pytest.raises(KeyError):
        v.get_context("ACNT")


def test_vlmc_scoring_methods_bit_exact(g):
    vl = fixture_vlmcs(g)
    for s in (0, 3):
        for m in range(5):
            assert vl[m].log_likelihood_ignore_initial_bias(g["nll_sequences"][s]) == g["nll_scores"][s][m]
            assert vl[m].log_likelihood(g["nll_sequences"][s]) == g["loglik_full_scores"][s][m]
    assert vl[0].log_likelihood("") == 0.0
    assert vl[1].log_likelihood_ignore_initial_bias("ACG") == 0.0  # shorter than the order: nothing scored


def test_generate_sequence_consumes_the_python_stream(g):
    vl = fixture_vlmcs(g)
    random.seed(g["sampler_seed"])
    got = [v.generate_sequence(200, 500) for v in vl]
    assert got == g["sampler_sequences"]
    after = random.random()
    random.seed(g["sampler_seed"])
    for _ in range(5 * 700):
        random.random()
    assert after == random.random()
    assert vl[0].generate_sequence(200, 500) is vl[0].sequence  # cached while the length matches (vlmc.pyx:157-158)
    vl[0].reset_sequence()
    assert vl[0].sequence == ""


def test_generate_sequence_from_a_context(g):
    """VLMC.generate_sequence_from with non-empty contexts (src/vlmc/vlmc.pyx:165-172) against the compiled
    reference's output: same symbols, same position of CPython's MT19937 stream afterwards, cache untouched."""
    vl = fixture_vlmcs(g)
    for c in golden_io.sampler_from_cases():
        v = vl[c["model"]]
        v.sequence = "kept"
        random.seed(c["seed"])
        assert v.generate_sequence_from(c["length"], c["context"]) == c["sequence"]
        assert random.random() == c["next_random"]
        assert v.sequence == "kept"
    random.seed(g["sampler_seed"])
    assert vl[0].generate_sequence_from(700, "")[-200:] == g["sampler_sequences"][0]
    assert vl[0].generate_sequence_from(0, "ACG") == ""


def test_likelihood_and_context_distribution(g):
    """VLMC.likelihood (src/vlmc/vlmc.pyx:103-118) and estimated_context_distribution (:182-206): bit-equal to the
    compiled reference's values, dict in the same order."""
    vl = fixture_vlmcs(g)
    for c in golden_io.sampler_from_cases("likelihood"):
        assert vl[c["model"]].likelihood(c["sequence"]) == c["likelihood"]
    with pytest.raises(KeyError):
        vl[0].likelihood("ACGNT")
    for c in golden_io.sampler_from_cases("context_distribution"):
        v = vl[c["model"]]
        random.seed(c["seed"])
        v.reset_sequence()
        d = v.estimated_context_distribution(c["length"])
        assert v.sequence == c["sequence"]
        assert list(d.keys()) == c["contexts"] and list(d.values()) == c["values"]


def test_nll_class_reproduces_reference_run(g):
    """random.seed(1234); N^2 loop of d.distance in list order == the golden run of the reference."""
    vl = fixture_vlmcs(g)
    random.seed(g["nll_seed"])
    d = NegativeLogLikelihood(g["nll_length"], mode="sequential")   # the reference's own summation order: bit-exact
    tri = util.calculate_distances_within_vlmcs(vl, d)
    assert [v.sequence for v in vl] == g["nll_sequences"]
    assert np.array_equal(tri, np.array(g["nll_triples_f32"], dtype=np.float32))
    want = np.array(g["nll_matrix_f64"])
    for i in range(5):
        for j in range(5):
            assert d.distance(vl[i], vl[j]) == want[i, j]
    assert np.array_equal(util.index_distances(vl, tri), want.astype(np.float32))
    # an un-batched pair goes to the device as a 2-model set and gives the same value
    d2 = NegativeLogLikelihood(g["nll_length"], mode="sequential")
    assert d2.distance(vl[0], vl[1]) == want[0, 1]
    # the default mode (exact int8 tensor-core contraction where the set allows it, else the warp scan): same sequences,
    # distances within 1e-9, the same float32 triples up to last-place rounding
    d3 = NegativeLogLikelihood(g["nll_length"])
    assert d3.mode == "kmer"
    tri3 = util.calculate_distances_within_vlmcs(vl, d3)
    got = np.array([[d3.distance(a, b) for b in vl] for a in vl])
    off = ~np.eye(5, dtype=bool)
    assert np.max(np.abs(got[off] - want[off]) / np.abs(want[off])) < 1e-9 and np.all(np.diag(got) == 0)
    assert np.allclose(tri3, tri, rtol=2e-7, atol=0)


def test_frobenius_projection_classes(g):
    vl = fixture_vlmcs(g)
    for use_union, key in ((False, "frobenius_intersection"), (True, "frobenius_union")):
        d = FrobeniusNorm(use_union)
        d.set_vlmcs(vl)
        got = np.array([[d.distance(a, b) for b in vl] for a in vl])
        assert np.allclose(got, np.array(g[key]), rtol=3e-7, atol=0) and np.all(np.diag(got) == 0)
        assert FrobeniusNorm(use_union).distance(vl[0], vl[1]) == got[0, 1]
        if not use_union:
            # test_distance_function_'s retrieval metric (src/util/distance_metrics.py:14,29-33) = mean distance
            assert float(np.mean(got)) == pytest.approx(g["tdf_metrics_frobenius"]["total_average_distance"], rel=1e-6)
    p = Projection()
    p.set_vlmcs(vl)
    assert p.dimension == g["projection_dimension"]
    got = np.array([[p.distance(a, b) for b in vl] for a in vl])
    assert isinstance(p.distance(vl[0], vl[1]), np.float32)
    assert np.allclose(got, np.array(g["projection"]), rtol=3e-7, atol=1e-6)


def test_estimate_class(g):
    vl = fixture_vlmcs(g)
    seqs, E = golden_io.fixture_estimate()
    for v, s in zip(vl, seqs):
        v.sequence = golden_io.codes_str(s)
    d = EstimateVLMC()
    d.set_vlmcs(vl)
    got = np.array([[d.distance(a, b) for b in vl] for a in vl])
    assert np.allclose(got, E, rtol=1e-9, atol=0)
    # sampling path: seed 3, right-major first use -> the golden sequences themselves
    vl2 = fixture_vlmcs(g)
    random.seed(3)
    d2 = EstimateVLMC()
    assert d2.distance(vl2[0], vl2[0]) == pytest.approx(E[0, 0], rel=1e-9)
    assert vl2[0].sequence == golden_io.codes_str(seqs[0])


def test_reference_consumer_accepts_our_distance(g):
    """The compiled reference's AverageLinkClustering (oracle/_ref) driven by OUR distance object:
    identical cluster assignments and merge distances to the all-reference golden run."""
    import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not built")
    import os
    ns = ref_loader.load()
    vl = fixture_vlmcs(g)
    d = FrobeniusNorm()
    d.set_vlmcs(vl)
    metadata = {v.name: {k: v.name for k in ("species", "family", "genus")} for v in vl}
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        import networkx as nx
        cl = ns.AverageLinkClustering(vl, d, metadata)
        for k in ("3", "2"):
            met = cl.cluster(int(k))
            comps = sorted(sorted(vl.index(v) for v in comp) for comp in nx.connected_components(met.G))
            assert comps == g["clustering_AverageLinkClustering"][k]["clusters"]
            assert np.allclose(met.get_merge_distances(), g["clustering_AverageLinkClustering"][k]["merge_distances"], rtol=1e-6)
    finally:
        os.chdir(cwd)


@pytest.mark.parametrize("name", ["shallow", "deep"])
def test_cluster_assignments_identical(name):
    """Average link (the oracle's restatement of the reference's algorithm, pinned on its goldens) on OUR
    float32 matrices gives the reference's labels and merge distances."""
    fm, seqs, est, z = golden_io.synthetic_small(name)
    h = native.Handle(0)
    h.load_models(fm)
    h.load_sequences_packed(*pack_sequences(seqs))
    k = int(z["family"].max()) + 1
    D32, _ = h.nll_matrix_h(int(z["nll_length"]), "sequential")
    labels, merges = O.average_link(D32, k)
    assert np.array_equal(labels, z["avglink_nll_labels"]) and np.array_equal(np.array(merges), z["avglink_nll_merges"])
    D32, _ = h.pair_matrix_h("frobenius_intersection")
    labels, merges = O.average_link(D32, k)
    assert np.array_equal(labels, z["avglink_frob_labels"])
    assert np.allclose(merges, z["avglink_frob_merges"], rtol=1e-6)


@pytest.mark.parametrize("kind", ["nll", "frobenius_intersection", "frobenius_union", "projection", "estimate"])
def test_tiling_invariance_1_2_4_8_ranks(kind):
    """north_star (c): the matrix is bit-identical however the tile space is dealt to ranks (the other
    ranks' tiles are computed on this GPU one after the other and assembled by the same kernel)."""
    import torch
    n, L = (45, 300) if kind != "estimate" else (14, 3000)
    fm, fam = synthetic.make_family_models(n, 31, 7 if kind != "estimate" else 5, 150, 4)
    seqs = synthetic.sample_sequences(fm, L, 5)
    h = native.Handle(0)
    h.load_models(fm)
    h.load_sequences_packed(*pack_sequences(seqs))
    base = MatrixJob(h, kind, n, length=L, tile=8).run().clone()
    torch.cuda.synchronize()
    for world in (2, 4, 8):
        for tile in (8, 16):
            out = MatrixJob(h, kind, n, length=L, tile=tile).run_emulated(world)
            assert torch.equal(out, base), (world, tile)
    if kind == "nll":  # the contraction mode shards at 128-granularity
        n2 = 300
        fm2, _ = synthetic.make_family_models(n2, 32, 5, 100, 4)
        seqs2 = synthetic.sample_sequences(fm2, 400, 6)
        h2 = native.Handle(0)
        h2.load_models(fm2)
        h2.load_sequences_packed(*pack_sequences(seqs2))
        base2 = MatrixJob(h2, "nll", n2, length=400, mode="kmer", tile=128).run().clone()
        for world in (2, 4):
            assert torch.equal(MatrixJob(h2, "nll", n2, length=400, mode="kmer", tile=128).run_emulated(world), base2)
        assert np.allclose(base2.cpu().numpy(), c_oracle.COracle(fm2).nll_matrix(seqs2).astype(np.float32), rtol=2e-7, atol=0)
    co = c_oracle.COracle(fm)
    if kind == "nll":
        want = co.nll_matrix(seqs)
    elif kind == "estimate":
        want = co.estimate_matrix(seqs)
    else:
        want = co.pair_matrix({"frobenius_intersection": "intersection", "frobenius_union": "union", "projection": "projection"}[kind])
    assert np.allclose(base.cpu().numpy(), want.astype(np.float32), rtol=2e-7, atol=0)


def test_average_link_on_device_matches_reference_semantics(g):
    """Row f1: the device merge loop gives the reference's labels and merge distances bit for bit --
    golden fixtures, forced ties, asymmetric input, and n = 1500 against the O(n^2) C oracle."""
    import torch
    from clustering_genomic_signatures_b200.clustering import AverageLinkClustering
    vl = fixture_vlmcs(g)
    cl = AverageLinkClustering(vl, FrobeniusNorm())
    for k in ("3", "2"):
        res = cl.cluster(int(k))
        assert sorted(res.clusters) == g["clustering_AverageLinkClustering"][k]["clusters"]
        assert np.allclose(res.merge_distances, g["clustering_AverageLinkClustering"][k]["merge_distances"], rtol=1e-6)
    h = native.Handle(0)
    rng = np.random.default_rng(0)
    for n, k, q, sym in ((12, 3, 0, True), (60, 4, 16, True), (90, 7, 8, True), (50, 1, 4, True), (30, 3, 0, False),
                         (1500, 20, 0, True), (700, 5, 64, True)):
        X = rng.random((n, 6))
        D = np.sqrt(((X[:, None] - X[None]) ** 2).sum(-1)).astype(np.float32)
        if q:
            D = (np.round(D * q) / q).astype(np.float32)  # many exactly equal distances
        if not sym:
            D = rng.random((n, n)).astype(np.float32)
        np.fill_diagonal(D, 0)
        want_labels, want_merges = c_oracle.average_link(D, k)
        labels, merges = h.average_link(torch.from_numpy(D).cuda(), k)
        assert np.array_equal(labels, want_labels), (n, k, q)
        assert np.array_equal(merges, want_merges), (n, k, q)
        if n <= 90:
            py_labels, py_merges = O.average_link(D, k)
            assert np.array_equal(labels, py_labels) and np.array_equal(merges, np.array(py_merges))


def test_full_size_properties():
    """BASELINE cfg3 size (1000 models x 10 kbp): size-independent properties -- symmetry, exact zero
    diagonal, warp mode within 1e-9 of sequential, float32 matrix == rounded fp64, a 40-model corner
    bit-equal to the C oracle."""
    import torch
    fm, fam = synthetic.config2(1000, 2)
    seqs = synthetic.sample_sequences(fm, 10000, 3)
    h = native.Handle(0)
    h.load_models(fm)
    h.load_sequences_packed(*pack_sequences(seqs))
    D32, D64 = h.nll_matrix_h(10000, "sequential")
    assert np.array_equal(D64, D64.T) and np.all(np.diag(D64) == 0) and np.array_equal(D32, D64.astype(np.float32))
    D32w, D64w = h.nll_matrix_h(10000, "warp")
    off = ~np.eye(1000, dtype=bool)
    assert np.max(np.abs(D64w[off] - D64[off]) / np.abs(D64[off])) < 1e-9
    D32k, D64k = h.nll_matrix_h(10000, "kmer")
    assert np.max(np.abs(D64k[off] - D64[off]) / np.abs(D64[off])) < 1e-9 and np.all(np.diag(D64k) == 0)
    print("float32 entries that differ from the bit-exact mode: warp %d, kmer %d of %d" % (
        int((D32w != D32).sum()), int((D32k != D32).sum()), D32.size))
    k = 40
    co = c_oracle.COracle(fm.subset(range(k)))
    M = h.nll_scores_h(native.SKIP_ORDER, "sequential")
    assert np.array_equal(M[:k, :k], co.score_matrix(seqs[:k]))
    F32, F64 = h.pair_matrix_h("frobenius_intersection")   # cfg 2 fills 9 % of its universe: the tensor-core path
    assert h.last_pair_kernel() == 1
    assert np.array_equal(F64, F64.T) and np.all(np.diag(F64) == 0)
    want = co.pair_matrix("intersection")
    assert np.all(np.abs(F64[:k, :k] - want) <= 1.3e-7 * want + 1e-9)
    h.set_pair_path("cuda")
    F32c, F64c = h.pair_matrix_h("frobenius_intersection")
    assert h.last_pair_kernel() == 0 and np.array_equal(F64c, F64c.T) and np.all(np.diag(F64c) == 0)
    assert np.allclose(F64c[:k, :k], want, rtol=1e-9, atol=0)
    # downstream: both paths give the same average-link clusters
    import torch
    kf = int(fam.max()) + 1
    la, ma = h.average_link(torch.from_numpy(F32).cuda(), kf)
    lb, mb = h.average_link(torch.from_numpy(F32c).cuda(), kf)
    assert np.array_equal(la, lb) and np.allclose(ma, mb, rtol=1e-6)
    # same-family pairs are closer than cross-family pairs on average (the set clusters)
    same = fam[:, None] == fam[None, :]
    assert F64[same & off].mean() < 0.5 * F64[~same].mean()
