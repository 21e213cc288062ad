"""Parity of the CUDA path (through the C ABI, libvgs.so) against the oracle and the golden vectors
recorded from the compiled reference.  Needs a B200: run with ``pytest -m gpu``.

Bars: lookup indices and the sequential NLL mode bit-exact; warp NLL mode and all fp64-accumulated
distances within 1e-9 relative of the oracle; Frobenius / Projection within 3e-7 relative of the
reference's own float32 BLAS result (its rounding)."""
import random

import numpy as np
import pytest

import c_oracle
import golden_io
from clustering_genomic_signatures_b200 import native, synthetic
from clustering_genomic_signatures_b200.flat import FlatModels, encode_history, pack_sequences, sequences_to_ascii

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def g():
    return golden_io.fixtures()


@pytest.fixture(scope="module")
def fx(g):
    fm = golden_io.fixture_flat(g)
    h = native.Handle(0)
    h.load_models(fm)
    return fm, h


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    denom = np.maximum(np.abs(b), 1e-300)
    e = np.abs(a - b) / denom
    e[(a == b)] = 0
    return float(e.max()) if e.size else 0.0


def test_logp_table_is_host_math_log(fx):
    import math
    fm, h = fx
    lp = h.get_logp()
    want = np.array([[-1000.0 if p == 0 else math.log(p) for p in row] for row in fm.prob])
    assert np.array_equal(lp, want)


def test_model_info(fx, g):
    fm, h = fx
    for m in range(fm.n_models):
        info = h.model_info(m)
        assert info["order"] == g["orders"][m] and info["n_ctx"] == fm.n_ctx(m)
        assert info["tier"] == (1 if g["orders"][m] <= 6 else 2)


def test_lookup_indices_bit_exact(fx, g):
    fm, h = fx
    rows = g["lookups"]
    enc = [encode_history(r["history"]) for r in rows]
    ids, masks = h.lookup([r["model"] for r in rows], [e[1] for e in enc], [e[0] for e in enc], want_all=True)
    assert ids.tolist() == [r["ctx_id"] for r in rows]
    for r, mask in zip(rows, masks):
        keys = g["models"][r["model"]]["contexts"]
        got = sorted((keys.index(r["history"][len(r["history"]) - l:] if l else "") for l in range(16) if int(mask) >> l & 1),
                     key=lambda i: -len(keys[i]))
        assert got == r["all_ctx_ids"]


def test_nll_fixture_scores_bit_exact(fx, g):
    fm, h = fx
    buf, off = sequences_to_ascii(g["nll_sequences"])
    h.load_sequences_ascii(buf, off)
    M = h.nll_scores_h(native.SKIP_ORDER, "sequential")
    assert np.array_equal(M, np.array(g["nll_scores"]))
    M0 = h.nll_scores_h(native.SKIP_NONE, "sequential")
    assert np.array_equal(M0, np.array(g["loglik_full_scores"]))
    Mw = h.nll_scores_h(native.SKIP_ORDER, "warp")
    assert rel_err(Mw, g["nll_scores"]) < 1e-12
    M0w = h.nll_scores_h(native.SKIP_NONE, "warp")
    assert rel_err(M0w, g["loglik_full_scores"]) < 1e-12


def test_nll_fixture_matrix(fx, g):
    fm, h = fx
    buf, off = sequences_to_ascii(g["nll_sequences"])
    h.load_sequences_ascii(buf, off)
    D32, D64 = h.nll_matrix_h(g["nll_length"], "sequential")
    assert np.array_equal(D64, np.array(g["nll_matrix_f64"]))
    tri = np.array(g["nll_triples_f32"], dtype=np.float32)
    assert np.array_equal(D32.ravel(), tri[:, 2])
    D32w, D64w = h.nll_matrix_h(g["nll_length"], "warp")
    assert rel_err(D64w, g["nll_matrix_f64"]) < RTOL
    assert np.all(np.diag(D64w) == 0)


def test_packed_and_ascii_loaders_agree(fx, g):
    fm, h = fx
    codes = [golden_io.str_codes(s) for s in g["nll_sequences"]]
    words, woff, lens = pack_sequences(codes)
    h.load_sequences_packed(words, woff, lens)
    M = h.nll_scores_h(native.SKIP_ORDER, "sequential")
    assert np.array_equal(M, np.array(g["nll_scores"]))
    for i, c in enumerate(codes):
        assert np.array_equal(h.get_sequence(i, len(c)), c)


def test_bad_symbol_raises_keyerror(fx):
    fm, h = fx
    buf, off = sequences_to_ascii(["ACGTNACGT"] * fm.n_models)
    with pytest.raises(KeyError):
        h.load_sequences_ascii(buf, off)


@pytest.mark.parametrize("kind,key", [("frobenius_intersection", "frobenius_intersection"),
                                      ("frobenius_union", "frobenius_union"), ("projection", "projection")])
def test_pair_fixture(fx, g, kind, key):
    fm, h = fx
    co = c_oracle.COracle(fm)
    want = co.pair_matrix({"frobenius_intersection": "intersection", "frobenius_union": "union", "projection": "projection"}[kind])
    D32, D64 = h.pair_matrix_h(kind)
    assert rel_err(D64, want) < RTOL
    assert rel_err(D64, g[key]) < 3e-7  # vs the reference's float32 arithmetic
    assert np.all(np.diag(D64) == 0)
    assert np.array_equal(D32, D64.astype(np.float32))


def test_estimate_fixture(fx):
    fm, h = fx
    seqs, E = golden_io.fixture_estimate()
    words, woff, lens = pack_sequences(seqs)
    h.load_sequences_packed(words, woff, lens)
    D32, D64 = h.estimate_matrix_h()
    assert rel_err(D64, E) < RTOL


def test_sampler_reproduces_reference_stream(fx, g):
    fm, h = fx
    random.seed(g["sampler_seed"])
    st = np.array(random.getstate()[1], dtype=np.uint32)
    h.sample_sequences(200, 500, st)
    for m in range(fm.n_models):
        assert golden_io.codes_str(h.get_sequence(m, 200)) == g["sampler_sequences"][m]
    # the stream position afterwards equals 5 * 700 calls of random.random()
    random.seed(g["sampler_seed"])
    for _ in range(5 * 700):
        random.random()
    assert np.array_equal(st, np.array(random.getstate()[1], dtype=np.uint32))


def test_sampler_zero_weight_row_raises_like_random_choices():
    # random.choices raises ValueError("Total of weights must be greater than zero") on an all-zero row (vlmc.pyx:177)
    tree = {"": {"A": 0.0, "C": 0.0, "G": 0.0, "T": 0.0}}
    fm = FlatModels.from_trees([tree], ["z"])
    h = native.Handle(0)
    h.load_models(fm, skip_nll=True)
    random.seed(5)
    st = np.array(random.getstate()[1], dtype=np.uint32)
    with pytest.raises(ValueError, match="Total of weights"):
        h.sample_sequences(20, 5, st)
    h.close()


def test_missing_root_raises_runtime_error():
    # a tree without "": histories that match nothing make the reference raise RuntimeError (vlmc.pyx:141)
    tree = {"A": {"A": 0.25, "C": 0.25, "G": 0.25, "T": 0.25}, "CA": {"A": 0.1, "C": 0.2, "G": 0.3, "T": 0.4}}
    fm = FlatModels.from_trees([tree, tree], ["a", "b"])
    h = native.Handle(0)
    h.load_models(fm)
    assert h.lookup([0], [encode_history("GG")[1]], [2])[0] == -1
    buf, off = sequences_to_ascii(["ACGTACGTAAAA", "AAAAAAAAAAAA"])
    h.load_sequences_ascii(buf, off)
    with pytest.raises(RuntimeError, match="get_context"):
        h.nll_matrix_h(12, "sequential")
    M = h.nll_scores_h(native.SKIP_ORDER, "sequential")
    assert np.isnan(M[0, 0]) and not np.isnan(M[1, 0])


@pytest.mark.parametrize("name", ["shallow", "deep"])
def test_synthetic_small(name):
    fm, seqs, est, z = golden_io.synthetic_small(name)
    h = native.Handle(0)
    h.load_models(fm)
    words, woff, lens = pack_sequences(seqs)
    h.load_sequences_packed(words, woff, lens)
    assert np.array_equal(h.nll_scores_h(native.SKIP_ORDER, "sequential"), z["scores"])
    assert np.array_equal(h.nll_scores_h(native.SKIP_NONE, "sequential"), z["scores_full"])
    assert rel_err(h.nll_scores_h(native.SKIP_ORDER, "warp"), z["scores"]) < 1e-12
    D32, D64 = h.nll_matrix_h(int(z["nll_length"]), "sequential")
    assert np.array_equal(D64, z["nll"])
    co = c_oracle.COracle(fm)
    for kind, mode in (("frobenius_intersection", "intersection"), ("frobenius_union", "union"), ("projection", "projection")):
        D32, D64 = h.pair_matrix_h(kind)
        assert rel_err(D64, co.pair_matrix(mode)) < RTOL
        assert rel_err(D64, z[kind]) < 3e-7
    sub = fm.subset(range(4))
    h2 = native.Handle(0)
    h2.load_models(sub)
    words, woff, lens = pack_sequences(est)
    h2.load_sequences_packed(words, woff, lens)
    D32, D64 = h2.estimate_matrix_h()
    assert rel_err(D64, z["estimate"]) < RTOL


@pytest.mark.parametrize("n,depth,ctx,L", [(150, 6, 300, 2000), (70, 9, 900, 1500), (40, 12, 2500, 700), (33, 3, 40, 515)])
def test_against_c_oracle_medium(n, depth, ctx, L):
    """Seeded synthetic sets at sizes the C oracle finishes in seconds; ragged tile edges (n not a
    multiple of the tile) and all three NLL table tiers."""
    fm, fam = synthetic.make_family_models(n, 100 + depth, depth, ctx, max(2, n // 10))
    seqs = synthetic.sample_sequences(fm, L, 7)
    co = c_oracle.COracle(fm)
    h = native.Handle(0)
    h.load_models(fm)
    words, woff, lens = pack_sequences(seqs)
    h.load_sequences_packed(words, woff, lens)
    M_or = co.score_matrix(seqs)
    assert np.array_equal(h.nll_scores_h(native.SKIP_ORDER, "sequential"), M_or)
    assert rel_err(h.nll_scores_h(native.SKIP_ORDER, "warp"), M_or) < 1e-12
    assert rel_err(h.nll_scores_h(native.SKIP_ORDER, "kmer"), M_or) < 1e-12  # contraction (depth <= 6) or its fallback
    D_or = co.nll_from_scores(M_or, L)
    D32, D64 = h.nll_matrix_h(L, "sequential")
    assert np.array_equal(D64, D_or)
    assert np.array_equal(D32, D_or.astype(np.float32))
    D32w, D64w = h.nll_matrix_h(L, "warp")
    assert rel_err(D64w, D_or) < RTOL
    D32k, D64k = h.nll_matrix_h(L, "kmer")
    assert rel_err(D64k, D_or) < RTOL and np.all(np.diag(D64k) == 0) and np.array_equal(D64k, D64k.T)
    for kind, mode in (("frobenius_intersection", "intersection"), ("frobenius_union", "union"), ("projection", "projection")):
        h.set_pair_path("cuda")          # fp64-accumulating CUDA-core kernels: the 1e-9 bar
        D32, D64 = h.pair_matrix_h(kind)
        assert h.last_pair_kernel() == 0
        want = co.pair_matrix(mode)
        assert rel_err(D64, want) < RTOL
        assert np.allclose(D64, D64.T, rtol=1e-14, atol=0)
        h.set_pair_path("auto")          # dense enough sets go to the tensor cores: their stated bound
        D32a, D64a = h.pair_matrix_h(kind)
        if h.last_pair_kernel() == 1:
            assert np.all(np.abs(D64a - want) <= 1.3e-7 * np.abs(want) + 1e-9) and np.array_equal(D64a, D64a.T)
        else:
            assert np.array_equal(D64a, D64)
    if depth <= 9:
        k = 12
        sub = fm.subset(range(k))
        hs = native.Handle(0)
        hs.load_models(sub)
        w2, o2, l2 = pack_sequences(seqs[:k])
        hs.load_sequences_packed(w2, o2, l2)
        D32, D64 = hs.estimate_matrix_h()
        assert rel_err(D64, c_oracle.COracle(sub).estimate_matrix(seqs[:k])) < RTOL


def test_kmer_mode_mixed_orders_and_short_sequences():
    """The contraction mode with models of different orders (head terms), ragged block edges and
    sequences shorter than the deepest order."""
    fa, _ = synthetic.make_family_models(70, 41, 6, 150, 5)
    fb, _ = synthetic.make_family_models(40, 42, 3, 30, 4)
    fc, _ = synthetic.make_family_models(25, 43, 1, 5, 2)
    trees = [f.tree(m) for f in (fa, fb, fc) for m in range(f.n_models)]
    order = np.random.default_rng(0).permutation(len(trees))
    fm = FlatModels.from_trees([trees[i] for i in order])
    assert sorted(set(fm.orders().tolist())) == [1, 3, 6]
    for L in (700, 5):
        seqs = synthetic.sample_sequences(fm, L, 8)
        h = native.Handle(0)
        h.load_models(fm)
        h.load_sequences_packed(*pack_sequences(seqs))
        co = c_oracle.COracle(fm)
        M_or = co.score_matrix(seqs)
        assert np.array_equal(h.nll_scores_h(native.SKIP_ORDER, "sequential"), M_or)
        Mk = h.nll_scores_h(native.SKIP_ORDER, "kmer")
        assert rel_err(Mk, M_or) < 1e-12
        D32, D64 = h.nll_matrix_h(L, "kmer")
        assert rel_err(D64, co.nll_from_scores(M_or, L)) < RTOL
