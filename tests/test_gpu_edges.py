"""Edge cases of the hot path on the B200, against the oracle: single model, root-only models, ragged and
empty sequences, the maximum context length, the wide (>= 32767 contexts) table tier, duplicate models,
empty intersections, and the argument errors the reference raises."""
import numpy as np
import pytest

import c_oracle
import vlmc_oracle as O
from clustering_genomic_signatures_b200 import native, synthetic
from clustering_genomic_signatures_b200.flat import FlatModels, pack_sequences, sequences_to_ascii

pytestmark = pytest.mark.gpu

UNIFORM = {"A": 0.25, "C": 0.25, "G": 0.25, "T": 0.25}


def row(a, c, g, t):
    return {"A": a, "C": c, "G": g, "T": t}


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    e = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    e[a == b] = 0
    return float(e.max()) if e.size else 0.0


def codes_of(s):
    return np.array(["ACGT".index(ch) for ch in s], dtype=np.uint8)


def test_single_model_matrices():
    fm = FlatModels.from_trees([{"": row(0.1, 0.2, 0.3, 0.4), "A": row(0.4, 0.3, 0.2, 0.1)}])
    h = native.Handle(0)
    h.load_models(fm)
    h.load_sequences_packed(*pack_sequences([codes_of("ACGTTGCAAC" * 20)]))
    for mode in ("sequential", "warp", "kmer"):
        D32, D64 = h.nll_matrix_h(200, mode)
        assert D32.shape == (1, 1) and D64[0, 0] == 0.0
    for kind in ("frobenius_intersection", "frobenius_union", "projection"):
        assert h.pair_matrix_h(kind)[1][0, 0] == 0.0
    D32, D64 = h.estimate_matrix_h()
    assert D64.shape == (1, 1) and np.isfinite(D64[0, 0])


def test_root_only_models_order_zero():
    trees = [{"": row(0.1, 0.2, 0.3, 0.4)}, {"": row(0.25, 0.25, 0.25, 0.25)}, {"": row(0.7, 0.1, 0.1, 0.1)}]
    fm = FlatModels.from_trees(trees)
    seqs = [codes_of("ACGT" * 50), codes_of("AAAC" * 50), codes_of("TTTTGGGC" * 25)]
    strs = ["ACGT" * 50, "AAAC" * 50, "TTTTGGGC" * 25]
    h = native.Handle(0)
    h.load_models(fm)
    h.load_sequences_packed(*pack_sequences(seqs))
    want = np.array([[O.log_likelihood_ignore_initial_bias(t, 0, s) for t in trees] for s in strs])
    assert np.array_equal(h.nll_scores_h(native.SKIP_ORDER, "sequential"), want)
    assert rel_err(h.nll_scores_h(native.SKIP_ORDER, "warp"), want) < 1e-12
    assert rel_err(h.nll_scores_h(native.SKIP_ORDER, "kmer"), want) < 1e-12
    co = c_oracle.COracle(fm)
    for kind, mode in (("frobenius_intersection", "intersection"), ("frobenius_union", "union"), ("projection", "projection")):
        assert rel_err(h.pair_matrix_h(kind)[1], co.pair_matrix(mode)) < 1e-9


def test_ragged_and_empty_sequences():
    """Sequences of different lengths in one set, including length 0 and lengths below the model order."""
    fm, _ = synthetic.make_family_models(9, 77, 5, 80, 3)
    trees = [fm.tree(m) for m in range(fm.n_models)]
    orders = fm.orders()
    rng = np.random.default_rng(3)
    lengths = [0, 1, 3, 5, 17, 64, 65, 1000, 4097]
    seqs = [rng.integers(0, 4, size=L).astype(np.uint8) for L in lengths]
    strs = ["".join("ACGT"[c] for c in s) for s in seqs]
    h = native.Handle(0)
    h.load_models(fm)
    h.load_sequences_packed(*pack_sequences(seqs))
    for skip, fn in ((native.SKIP_ORDER, O.log_likelihood_ignore_initial_bias), (native.SKIP_NONE, None)):
        if fn is None:
            want = np.array([[O.log_likelihood(t, int(o), s, 0) for t, o in zip(trees, orders)] for s in strs])
        else:
            want = np.array([[fn(t, int(o), s) for t, o in zip(trees, orders)] for s in strs])
        assert np.array_equal(h.nll_scores_h(skip, "sequential"), want)
        assert rel_err(h.nll_scores_h(skip, "warp"), want) < 1e-12
    assert rel_err(h.nll_scores_h(native.SKIP_ORDER, "kmer"), np.array(
        [[O.log_likelihood_ignore_initial_bias(t, int(o), s) for t, o in zip(trees, orders)] for s in strs])) < 1e-12
    # the ASCII loader packs the same ragged set
    h.load_sequences_ascii(*sequences_to_ascii(strs))
    assert np.array_equal(h.nll_scores_h(native.SKIP_NONE, "sequential"),
                          np.array([[O.log_likelihood(t, int(o), s, 0) for t, o in zip(trees, orders)] for s in strs]))


def test_maximum_context_length_and_beyond():
    deep = "ACGTACGTACGTACG"  # 15 symbols: the longest supported context
    tree = {"": row(0.1, 0.2, 0.3, 0.4), "G": row(0.3, 0.3, 0.2, 0.2), deep: row(0.01, 0.01, 0.01, 0.97)}
    other = {"": row(0.25, 0.25, 0.25, 0.25), "CG": row(0.5, 0.1, 0.2, 0.2), deep: row(0.4, 0.3, 0.2, 0.1)}
    fm = FlatModels.from_trees([tree, other])
    assert list(fm.orders()) == [15, 15]
    s = ("TT" + deep) * 30 + "ACGTTTGACCA" * 9
    seqs = [codes_of(s), codes_of(s[::-1])]
    h = native.Handle(0)
    h.load_models(fm)
    h.load_sequences_packed(*pack_sequences(seqs))
    want = np.array([[O.log_likelihood_ignore_initial_bias(t, 15, x) for t in (tree, other)] for x in (s, s[::-1])])
    assert np.array_equal(h.nll_scores_h(native.SKIP_ORDER, "sequential"), want)
    assert rel_err(h.nll_scores_h(native.SKIP_ORDER, "warp"), want) < 1e-12
    co = c_oracle.COracle(fm)
    assert rel_err(h.pair_matrix_h("frobenius_union")[1], co.pair_matrix("union")) < 1e-9
    with pytest.raises((ValueError, NotImplementedError)):
        FlatModels.from_trees([{"": UNIFORM, deep + "A": UNIFORM}])
    bad = FlatModels(np.array([0, 2]), np.array([0, 16], dtype=np.uint8), np.array([0, 0], dtype=np.uint32),
                     np.full((2, 4), 0.25), ["x"])
    with pytest.raises(NotImplementedError):
        native.Handle(0).load_models_raw(1, bad.ctx_offsets.astype(np.int64), bad.ctx_len, bad.ctx_code, bad.prob)


def test_wide_tier_model_with_more_than_32767_contexts():
    """A model with >= 32767 contexts takes the 32-bit first-level map (tier 3)."""
    fm, _ = synthetic.make_family_models(3, 91, 9, 36000, 1, leaf_jitter=0.01)
    assert fm.n_ctx(0) >= 0x7FFF
    h = native.Handle(0)
    h.load_models(fm)
    assert h.model_info(0)["tier"] == 3
    seqs = synthetic.sample_sequences(fm, 3000, 4)
    h.load_sequences_packed(*pack_sequences(seqs))
    co = c_oracle.COracle(fm)
    M_or = co.score_matrix(seqs)
    assert np.array_equal(h.nll_scores_h(native.SKIP_ORDER, "sequential"), M_or)
    assert rel_err(h.nll_scores_h(native.SKIP_ORDER, "warp"), M_or) < 1e-12
    D32, D64 = h.nll_matrix_h(3000, "sequential")
    assert np.array_equal(D64, co.nll_from_scores(M_or, 3000))
    for kind, mode in (("frobenius_intersection", "intersection"), ("frobenius_union", "union")):
        assert rel_err(h.pair_matrix_h(kind)[1], co.pair_matrix(mode)) < 1e-9


def test_duplicate_models_have_zero_distance():
    fm, _ = synthetic.make_family_models(6, 5, 6, 100, 2)
    idx = [0, 3, 0, 5, 3]
    dup = fm.subset(idx)
    seqs = synthetic.sample_sequences(fm, 900, 2)[idx]
    h = native.Handle(0)
    h.load_models(dup)
    h.load_sequences_packed(*pack_sequences(seqs))
    for kind in ("frobenius_intersection", "frobenius_union", "projection"):
        D = h.pair_matrix_h(kind)[1]
        assert D[0, 2] == 0.0 and D[1, 4] == 0.0 and D[0, 1] > 0.0
    for mode in ("sequential", "warp"):
        D = h.nll_matrix_h(900, mode)[1]
        assert D[0, 2] == 0.0 and D[1, 4] == 0.0 and D[2, 0] == 0.0
    # the int8 contraction sums exact integers: duplicates get identical score bits, so the reference's exact 0.0 holds
    D = h.nll_matrix_h(900, "kmer")[1]
    assert h.last_nll_kernel() == 3
    assert D[0, 2] == 0.0 and D[1, 4] == 0.0 and D[2, 0] == 0.0


def test_empty_intersection_is_nan_like_the_reference():
    # frobenius.pyx:36 divides by sqrt(0): numpy gives nan (0/0); the union and Projection stay finite
    a = {"A": row(0.1, 0.2, 0.3, 0.4), "": UNIFORM}
    b = {"C": row(0.4, 0.3, 0.2, 0.1), "G": UNIFORM}
    c = {"T": UNIFORM}
    fm = FlatModels.from_trees([a, b, c])
    h = native.Handle(0)
    h.load_models(fm, skip_nll=True)
    D = h.pair_matrix_h("frobenius_intersection")[1]
    assert np.isnan(D[0, 1]) and np.isnan(D[1, 2]) and np.isnan(D[0, 2]) and D[0, 0] == 0.0
    P = h.pair_matrix_h("projection")[1]
    assert np.all(np.isfinite(P))
    assert rel_err(P, O.projection_matrix([a, b, c])[0]) < 1e-9


def test_argument_errors():
    h = native.Handle(0)
    with pytest.raises(ValueError):
        h.nll_matrix_h(10, "sequential")                       # nothing loaded
    fm = FlatModels.from_trees([{"": UNIFORM}, {"": UNIFORM}])
    h.load_models(fm, skip_nll=True)
    h.load_sequences_packed(*pack_sequences([codes_of("ACGT"), codes_of("ACGT")]))
    with pytest.raises(ValueError):
        h.nll_matrix_h(4, "sequential")                        # loaded without NLL tables
    neg = FlatModels(np.array([0, 1]), np.array([0], dtype=np.uint8), np.array([0], dtype=np.uint32),
                     np.array([[0.5, -0.1, 0.3, 0.3]]), ["neg"])
    with pytest.raises(ValueError, match="math domain"):
        native.Handle(0).load_models_raw(1, neg.ctx_offsets.astype(np.int64), neg.ctx_len, neg.ctx_code, neg.prob)
    h.load_models(fm)
    h.load_sequences_packed(*pack_sequences([codes_of("ACGT")]))
    with pytest.raises(ValueError):
        h.nll_matrix_h(4, "sequential")                        # one sequence per model is required


@pytest.mark.parametrize("L", [30000, 70000])
def test_int8_contraction_with_frequent_kmers(L):
    """Shallow models and long sequences: every k-mer count is far above one byte, so the int8 tensor-core path
    runs entirely on its 'frequent k-mer' term (count >> 8) plus the low bytes; still exact to fixed-point rounding.
    L = 70000 also takes the histogram kernel with 32-bit counters."""
    fm, _ = synthetic.make_family_models(37, 9, 2, 12, 4)
    assert int(fm.orders().max()) == 2
    seqs = synthetic.sample_sequences(fm, L, 5)
    h = native.Handle(0)
    h.load_models(fm)
    h.load_sequences_packed(*pack_sequences(seqs))
    co = c_oracle.COracle(fm)
    M_or = co.score_matrix(seqs)
    Mk = h.nll_scores_h(native.SKIP_ORDER, "kmer")
    assert h.last_nll_kernel() == 3, "expected the int8 tensor-core contraction"
    assert rel_err(Mk, M_or) < 1e-12
    D = h.nll_matrix_h(L, "kmer")[1]
    assert rel_err(D, co.nll_from_scores(M_or, L)) < 1e-9 and np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
    # the fp64 MMA path gives the same scores to rounding
    import os
    import subprocess
    import sys
    code = ("import numpy as np, sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from clustering_genomic_signatures_b200 import native, synthetic\n"
            "from clustering_genomic_signatures_b200.flat import pack_sequences\n"
            "fm, _ = synthetic.make_family_models(37, 9, 2, 12, 4); seqs = synthetic.sample_sequences(fm, int(sys.argv[2]), 5)\n"
            "h = native.Handle(0); h.load_models(fm); h.load_sequences_packed(*pack_sequences(seqs))\n"
            "M = h.nll_scores_h(native.SKIP_ORDER, 'kmer'); assert h.last_nll_kernel() == 1\n"
            "np.save(sys.argv[1], M)\n") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                            os.path.dirname(os.path.abspath(__file__)))
    out = "/tmp/_vgs_dmma_scores.npy"
    env = dict(os.environ, VGS_KMER_GEMM="dmma")
    subprocess.check_call([sys.executable, "-c", code, out, str(L)], env=env)
    assert rel_err(np.load(out), Mk) < 1e-12


def test_too_many_frequent_kmers_voids_the_speculative_pass():
    """vgs_nll_matrix_h launches the int8 contraction without waiting for the once-per-sequence-set check of the k-mer
    counts; when a sequence turns out to have more k-mers above 255 counts than the side list holds (order 3, 256 k-mers,
    all of them counted ~500 times), the pass is void and the call redoes it on the fp64 contraction -- same answer as the
    oracle, first call and second."""
    fm, _ = synthetic.make_family_models(12, 4, 3, 60, 2)
    assert int(fm.orders().max()) == 3
    fm.prob[:] = 0.25 + 0.3 * (fm.prob - 0.25)         # flattened rows: most 4-mers occur ~ L / 256 = 500 times
    fm.prob[:] /= fm.prob.sum(axis=1, keepdims=True)
    L = 130000
    seqs = synthetic.sample_sequences(fm, L, 9)
    co = c_oracle.COracle(fm)
    want = co.nll_from_scores(co.score_matrix(seqs), L)
    h = native.Handle(0)
    h.load_models(fm)
    h.load_sequences_packed(*pack_sequences(seqs))
    for _ in range(2):
        D = h.nll_matrix_h(L, "kmer")[1]
        assert h.last_nll_kernel() == 1, "expected the fp64 contraction after the void int8 pass"
        assert np.allclose(D, want, rtol=1e-9, atol=1e-13) and np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
    h.close()
