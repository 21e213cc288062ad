"""Round-2 paths on a B200 (through the C ABI): the CTA-pair int8 contraction at awkward shapes, the model-sharded NLL
path (every rank emulated on one GPU), the downstream parity of the headline (k-mer) mode at BASELINE cfg-3 size, the
[N^2, 3] hand-off kernel, the tensor-peak probe."""
import numpy as np
import pytest

import c_oracle
import vlmc_oracle as O
from clustering_genomic_signatures_b200 import native, synthetic
from clustering_genomic_signatures_b200.engine import sharded_nll_emulated
from clustering_genomic_signatures_b200.flat import pack_sequences

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    e = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    e[a == b] = 0
    return float(e.max()) if e.size else 0.0


@pytest.mark.parametrize("n,depth,ctx,L", [(5, 6, 60, 700), (37, 4, 40, 900), (130, 6, 150, 1200), (261, 5, 90, 500),
                                           (300, 3, 20, 2500), (9, 0, 1, 300), (70, 1, 4, 400)])
def test_int8_contraction_shapes(n, depth, ctx, L):
    """Sets that are not multiples of the 256 x 32 tile quanta, k ranges shorter than one 128-byte stage (depth <= 2),
    several units per CTA pair and k-split launches: scores against the C oracle, matrix symmetric with exact zeros."""
    fm, _ = synthetic.make_family_models(n, 100 + n, depth, ctx, max(1, n // 10))
    seqs = synthetic.sample_sequences(fm, L, 7)
    h = native.Handle(0)
    h.load_models(fm)
    h.load_sequences_packed(*pack_sequences(seqs))
    co = c_oracle.COracle(fm)
    M_or = co.score_matrix(seqs)
    Mk = h.nll_scores_h(native.SKIP_ORDER, "kmer")
    assert h.last_nll_kernel() == 3, "expected the int8 tensor-core contraction"
    assert rel_err(Mk, M_or) < 1e-12
    D32, D64 = h.nll_matrix_h(L, "kmer")
    assert rel_err(D64, co.nll_from_scores(M_or, L)) < 1e-9
    assert np.array_equal(D64, D64.T) and np.all(np.diag(D64) == 0) and np.array_equal(D32, D64.astype(np.float32))
    # a sub-range of sequences x models through the public scores call (row-major output, leading dimension > n)
    import torch
    ld = n + 3
    M = torch.full((n, ld), -7.0, dtype=torch.float64, device="cuda")
    s0, s1, m0, m1 = 0, n, 0, n
    stream = torch.cuda.current_stream().cuda_stream
    h.nll_scores(M, (s0, s1), (m0, m1), native.SKIP_ORDER, "kmer", stream)
    h.poll_error(stream)
    got = M.cpu().numpy()
    assert np.array_equal(got[:, :n], Mk) and np.all(got[:, n:] == -7.0)


def test_kmer_duplicates_score_exactly_alike():
    """Integer sums: identical models give identical bits, so the reference's exact 0.0 between duplicates holds."""
    fm, _ = synthetic.make_family_models(6, 5, 6, 100, 2)
    idx = [0, 3, 0, 5, 3]
    dup = fm.subset(idx)
    seqs = synthetic.sample_sequences(fm, 900, 2)[idx]
    h = native.Handle(0)
    h.load_models(dup)
    h.load_sequences_packed(*pack_sequences(seqs))
    D = h.nll_matrix_h(900, "kmer")[1]
    assert h.last_nll_kernel() == 3
    assert D[0, 2] == 0.0 and D[1, 4] == 0.0 and D[2, 0] == 0.0 and D[4, 1] == 0.0 and D[0, 1] > 0.0


@pytest.mark.parametrize("mode", ["kmer", "sequential", "warp"])
def test_model_sharded_nll_is_bit_identical(mode):
    """north_star (c) for the NLL kinds: 1/2/3/4/8 model shards (each rank emulated on this GPU) give the single-GPU
    matrix bit for bit -- shallow set (int8 contraction in kmer mode) and deep set (scan)."""
    import torch
    for n, depth, ctx, L in ((300, 5, 100, 400), (45, 8, 150, 300)):
        fm, _ = synthetic.make_family_models(n, 32, depth, ctx, 4)
        seqs = synthetic.sample_sequences(fm, L, 6)
        packed = pack_sequences(seqs)
        h = native.Handle(0)
        h.load_models(fm)
        h.load_sequences_packed(*packed)
        D32, D64 = h.nll_matrix_h(L, mode)
        base = torch.from_numpy(D32).cuda()
        for world in (1, 2, 3, 4, 8):
            out = sharded_nll_emulated(fm, packed, L, mode, world)
            assert torch.equal(out, base), (mode, n, world)
        out = sharded_nll_emulated(fm, packed, L, mode, 2, mirror=True)   # two destinations per band
        assert torch.equal(out, base), (mode, n, "mirror")
        want = c_oracle.COracle(fm).nll_matrix(seqs)
        if mode == "sequential":
            assert np.array_equal(D64, want)
        else:
            assert rel_err(D64, want) < 1e-9


def test_headline_mode_downstream_parity_cfg3():
    """BASELINE cfg 3 (1000 models x 10 kbp).  The k-mer (int8 tensor-core) and warp matrices against the bit-exact
    sequential one: fp64 entries within 1e-9, the float32 matrices differ in at most a handful of last-place roundings,
    and average-link clustering (device kernel = the reference's algorithm, pinned on its goldens) gives the SAME labels
    and merge order from all three -- and from the C oracle's average link on the sequential matrix."""
    import torch
    fm, fam = synthetic.config2(1000, 2)
    seqs = synthetic.sample_sequences(fm, 10000, 3)
    h = native.Handle(0)
    h.load_models(fm)
    h.load_sequences_packed(*pack_sequences(seqs))
    k = int(fam.max()) + 1
    D32s, D64s = h.nll_matrix_h(10000, "sequential")
    off = ~np.eye(1000, dtype=bool)
    labels_s, merges_s = h.average_link(torch.from_numpy(D32s).cuda(), k)
    want_labels, want_merges = c_oracle.average_link(D32s, k)
    assert np.array_equal(labels_s, want_labels) and np.array_equal(merges_s, want_merges)
    for mode in ("kmer", "warp"):
        D32, D64 = h.nll_matrix_h(10000, mode)
        if mode == "kmer":
            assert h.last_nll_kernel() == 3
        assert np.max(np.abs(D64[off] - D64s[off]) / np.abs(D64s[off])) < 1e-9
        assert np.array_equal(D32, D64.astype(np.float32)) and np.array_equal(D64, D64.T) and np.all(np.diag(D64) == 0)
        differing = int((D32 != D32s).sum())
        assert differing <= 200, "float32 entries that round differently from the bit-exact mode: %d of 10^6" % differing
        # where they differ it is one unit in the last place
        if differing:
            a, b = D32[D32 != D32s], D32s[D32 != D32s]
            assert np.all(np.abs(a.view(np.int32) - b.view(np.int32)) == 1)
        labels, merges = h.average_link(torch.from_numpy(D32).cuda(), k)
        assert np.array_equal(labels, labels_s), mode
        assert np.allclose(merges, merges_s, rtol=1e-6, atol=0), mode


def test_matrix_to_triples_kernel():
    """vgs_matrix_to_triples = src/clustering/util.pyx:6-21's [N^2, 3] float32 rows (i, j, d), row-major in (i, j)."""
    import torch
    rng = np.random.default_rng(3)
    for n in (1, 5, 300):
        D = rng.random((n, n)).astype(np.float32)
        h = native.Handle(0)
        tri = torch.empty((n * n, 3), dtype=torch.float32, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        h.matrix_to_triples(n, torch.from_numpy(D).cuda(), tri, stream)
        torch.cuda.synchronize()
        got = tri.cpu().numpy()
        ii, jj = np.divmod(np.arange(n * n), n)
        assert np.array_equal(got[:, 0], ii.astype(np.float32)) and np.array_equal(got[:, 1], jj.astype(np.float32))
        assert np.array_equal(got[:, 2], D.reshape(-1))
        # and it is what the host mirror of calculate_distances_within_vlmcs / index_distances round-trips
        assert np.array_equal(O.triples_to_matrix(got, n), D) if hasattr(O, "triples_to_matrix") else True


def test_tensor_peak_probe():
    h = native.Handle(0)
    t2, ms2 = h.tensor_peak_i8(2)
    t1, ms1 = h.tensor_peak_i8(1)
    print("int8 tensor peak probe: cta_group::2 %.0f TOPS (%.3f ms), cta_group::1 %.0f TOPS (%.3f ms)" % (t2, ms2, t1, ms1))
    assert 500.0 < t2 < 6000.0 and 250.0 < t1 < 6000.0


def test_nll_only_and_lazy_table_loads():
    """VGS_LOAD_NLL_ONLY uploads no transition rows (the pair / estimate / sampler entry points then refuse), and the digit
    operand built during the load (block by block behind the ln p upload) equals the one built on first use."""
    fm, _ = synthetic.make_family_models(150, 77, 6, 120, 5)
    seqs = synthetic.sample_sequences(fm, 800, 3)
    packed = pack_sequences(seqs)
    outs = []
    for kw in ({}, {"nll_only": True}, {"lazy_tables": True}, {"nll_only": True, "lazy_tables": True}):
        h = native.Handle(0)
        h.load_models(fm, **kw)
        h.load_sequences_packed(*packed)
        outs.append((h.nll_matrix_h(800, "kmer")[1], h.nll_matrix_h(800, "sequential")[1]))
        assert h.last_nll_kernel() == 0
        if kw.get("nll_only"):
            with pytest.raises(ValueError, match="NLL_ONLY"):
                h.pair_matrix_h("frobenius_intersection")
            with pytest.raises(ValueError, match="NLL_ONLY"):
                h.estimate_matrix_h()
            with pytest.raises(ValueError, match="NLL_ONLY"):
                h.sample_sequences(10, 0, np.zeros(625, dtype=np.uint32))
        h.close()
    for k, s in outs[1:]:
        assert np.array_equal(k, outs[0][0]) and np.array_equal(s, outs[0][1])
    assert np.array_equal(outs[0][1], c_oracle.COracle(fm).nll_matrix(seqs))


@pytest.mark.gpu
@pytest.mark.parametrize("with_zeros", [False, True])
def test_device_log_load(with_zeros):
    """VGS_LOAD_DEVICE_LOG: the contraction takes ln p from the device's log() -- scores within the mode's bound of the oracle
    (6 digit rows when every probability exceeds e^-16, 7 when the set has p = 0 entries) --, the bit-exact modes still see
    libm's ln p (computed on first need), and no kmer result depends on whether a bit-exact mode ran in between.  Sets the
    contraction does not cover ignore the flag; a negative probability is the reference's math domain error."""
    fm, _ = synthetic.make_family_models(90, 31, 6, 100, 4)
    if with_zeros:
        assert (fm.prob == 0).any()
    else:
        fm.prob[:] = np.maximum(fm.prob, 1e-3)
        fm.prob[:] /= fm.prob.sum(axis=1, keepdims=True)
        assert fm.prob.min() > 2e-7
    seqs = synthetic.sample_sequences(fm, 900, 3)
    packed = pack_sequences(seqs)
    co = c_oracle.COracle(fm)
    M_or = co.score_matrix(seqs)
    h = native.Handle(0)
    h.load_models(fm, nll_only=True, device_log=True)
    h.load_sequences_packed(*packed)
    Mk = h.nll_scores_h(native.SKIP_ORDER, "kmer")
    assert h.last_nll_kernel() == 3
    nd, f = h.kmer_digit_info()
    assert f == 43 and nd == (7 if with_zeros else 6)
    assert rel_err(Mk, M_or) < 1e-12
    Dk = h.nll_matrix_h(900, "kmer")[1]
    assert rel_err(Dk, co.nll_from_scores(M_or, 900)) < 1e-9 and np.array_equal(Dk, Dk.T) and np.all(np.diag(Dk) == 0)
    # bit-exact modes on the same handle: libm's ln p arrives on first need
    assert np.array_equal(h.nll_scores_h(native.SKIP_ORDER, "sequential"), M_or)
    import math
    want_lp = np.array([-1000.0 if p == 0 else math.log(p) for p in fm.prob.ravel()]).reshape(fm.prob.shape)   # libm, not numpy's SIMD log
    assert np.array_equal(h.get_logp(), want_lp)
    assert np.array_equal(h.nll_scores_h(native.SKIP_ORDER, "kmer"), Mk)       # ... and the contraction keeps its own
    # host-log load of the same set: the same scores to the mode's bound (not necessarily the same bits)
    h2 = native.Handle(0)
    h2.load_models(fm)
    h2.load_sequences_packed(*packed)
    assert rel_err(h2.nll_scores_h(native.SKIP_ORDER, "kmer"), Mk) < 1e-13
    h2.close()
    # a set deeper than the contraction reaches ignores the flag (host ln p, scan)
    fd, _ = synthetic.make_family_models(12, 5, 8, 300, 2)
    sd = synthetic.sample_sequences(fd, 400, 2)
    h.load_models(fd, nll_only=True, device_log=True)
    h.load_sequences_packed(*pack_sequences(sd))
    assert np.array_equal(h.nll_scores_h(native.SKIP_ORDER, "sequential"), c_oracle.COracle(fd).score_matrix(sd))
    bad = fm.subset(range(3))
    bad.prob[1, 2] = -0.25
    with pytest.raises(ValueError, match="math domain error"):
        h.load_models(bad, device_log=True)
    h.close()


@pytest.mark.gpu
def test_device_log_load_digit_count_guess():
    """A device-log load builds the digit operand before the host has seen the range of ln p, for the digit count the
    handle's previous set needed; whatever the guess, the operand ends up with the count the set asks for (6 rows without
    p = 0 entries, 7 with) and the scores are the same bits as on a fresh handle."""
    wide, _ = synthetic.make_family_models(40, 17, 6, 80, 3)
    assert (wide.prob == 0).any()
    narrow, _ = synthetic.make_family_models(40, 17, 6, 80, 3)
    narrow.prob[:] = np.maximum(narrow.prob, 1e-3)
    narrow.prob[:] /= narrow.prob.sum(axis=1, keepdims=True)
    seqs = synthetic.sample_sequences(wide, 700, 5)
    packed = pack_sequences(seqs)

    def fresh(fm):
        h = native.Handle(0)
        h.load_models(fm, nll_only=True, device_log=True)
        h.load_sequences_packed(*packed)
        M = h.nll_scores_h(native.SKIP_ORDER, "kmer")
        nd = h.kmer_digit_info()[0]
        h.close()
        return M, nd
    want = {id(narrow): fresh(narrow), id(wide): fresh(wide)}
    assert want[id(narrow)][1] == 6 and want[id(wide)][1] == 7
    h = native.Handle(0)
    for fm in (narrow, wide, wide, narrow, narrow, wide):
        h.load_models(fm, nll_only=True, device_log=True)
        h.load_sequences_packed(*packed)
        M = h.nll_scores_h(native.SKIP_ORDER, "kmer")
        assert h.last_nll_kernel() == 3
        assert h.kmer_digit_info()[0] == want[id(fm)][1]
        assert np.array_equal(M, want[id(fm)][0])
    h.close()


@pytest.mark.gpu
def test_kmer_fraction_bits():
    """The precision of the contraction is a parameter: f fraction bits of ln p -> |d score| <= L * 2^-(f+1); the digit rows per
    model (the GEMM's work) follow f and the range of the set's ln p; a change takes effect with the next operand build."""
    fm, _ = synthetic.make_family_models(60, 17, 5, 80, 3)
    fm.prob[:] = np.maximum(fm.prob, 1e-3)
    fm.prob[:] /= fm.prob.sum(axis=1, keepdims=True)
    L = 1500
    seqs = synthetic.sample_sequences(fm, L, 4)
    M_or = c_oracle.COracle(fm).score_matrix(seqs)
    h = native.Handle(0)
    h.load_models(fm)
    h.load_sequences_packed(*pack_sequences(seqs))
    errs = {}
    for f, nd_want in ((35, 6), (43, 6), (51, 7)):
        h.set_kmer_fraction_bits(f)
        Mk = h.nll_scores_h(native.SKIP_ORDER, "kmer")
        assert h.last_nll_kernel() == 3 and h.kmer_digit_info() == (nd_want, f)
        errs[f] = float(np.abs(Mk - M_or).max())
        assert errs[f] <= L * 2.0 ** -(f + 1) + 1e-10          # the stated bound (+ the fp64 scan's own rounding noise)
    assert errs[35] > errs[51]
    with pytest.raises(ValueError):
        h.set_kmer_fraction_bits(60)
    h.close()


def _tc_bound_ok(D_tc, D_ref):
    """Stated bound of the tensor-core variant: |d_tc - d_ref| <= 1.3e-7 d_ref + 1e-9 (rows quantised to 2^-31, the
    difference of two rows not rounded to float32 before squaring)."""
    fin = np.isfinite(D_ref)
    assert np.array_equal(np.isnan(D_tc), np.isnan(D_ref))
    return bool(np.all(np.abs(D_tc[fin] - D_ref[fin]) <= 1.3e-7 * np.abs(D_ref[fin]) + 1e-9))


@pytest.mark.parametrize("shape", ["chains4", "chains6", "trees6", "trees7_sparse_forced", "tiny"])
def test_tensor_core_frobenius_projection(shape):
    """K5: Frobenius (intersection) and Projection as exact integer GEMMs on the tensor cores against the fp64 oracle
    (stated bound) and the CUDA-core path (1e-9 of the oracle); symmetric, exact zero diagonal, exact zero between
    duplicates, NaN for an empty intersection; the dispatcher picks the path by universe density."""
    if shape == "chains4":
        fm, _ = synthetic.make_full_chains(150, 3, 4)
    elif shape == "chains6":
        fm, _ = synthetic.make_full_chains(70, 4, 6)
    elif shape == "trees6":
        fm, _ = synthetic.make_family_models(200, 5, 6, 400, 8)
    elif shape == "trees7_sparse_forced":
        fm, _ = synthetic.make_family_models(90, 6, 7, 120, 5)
    else:
        fm, _ = synthetic.make_family_models(3, 7, 2, 9, 1)
    n = fm.n_models
    dup = fm.subset(list(range(n)) + [0, n - 1])     # two duplicates at the end
    h = native.Handle(0)
    h.load_models(dup, skip_nll=True)
    co = c_oracle.COracle(dup)
    for kind, mode in (("frobenius_intersection", "intersection"), ("projection", "projection")):
        want = co.pair_matrix(mode)
        h.set_pair_path("cuda")
        Dc = h.pair_matrix_h(kind)[1]
        assert h.last_pair_kernel() == 0
        assert rel_err(Dc, want) < 1e-9
        h.set_pair_path("tensor")
        D32, Dt = h.pair_matrix_h(kind)
        assert h.last_pair_kernel() == 1
        assert _tc_bound_ok(Dt, want), (shape, kind, float(np.nanmax(np.abs(Dt - want) / np.maximum(np.abs(want), 1e-30))))
        assert np.array_equal(Dt, Dt.T) and np.all(np.diag(Dt) == 0) and np.array_equal(D32, Dt.astype(np.float32))
        assert Dt[0, n] == 0.0 and Dt[n - 1, n + 1] == 0.0 and Dt[n, 0] == 0.0
        h.set_pair_path("auto")
        h.pair_matrix_h(kind)
        dense = fm.total_contexts >= 0.03 * n * ((4 ** (int(fm.orders().max()) + 1) - 1) // 3) and dup.n_models >= 64
        assert h.last_pair_kernel() == (1 if dense else 0), shape
    # the union stays on the CUDA cores whatever the setting
    h.set_pair_path("tensor")
    h.pair_matrix_h("frobenius_union")
    assert h.last_pair_kernel() == 0


def test_tensor_core_pair_edge_cases():
    from clustering_genomic_signatures_b200.flat import FlatModels
    U = dict(zip("ACGT", (0.25, 0.25, 0.25, 0.25)))
    a = {"A": dict(zip("ACGT", (0.1, 0.2, 0.3, 0.4))), "": U}
    b = {"C": dict(zip("ACGT", (0.4, 0.3, 0.2, 0.1))), "G": U}
    c = {"T": U, "": dict(zip("ACGT", (1.0, 0.0, 0.0, 0.0)))}
    fm = FlatModels.from_trees([a, b, c])
    h = native.Handle(0)
    h.load_models(fm, skip_nll=True)
    h.set_pair_path("tensor")
    D = h.pair_matrix_h("frobenius_intersection")[1]
    assert h.last_pair_kernel() == 1
    assert np.isnan(D[0, 1]) and np.isnan(D[1, 2]) and D[0, 0] == 0.0 and np.isfinite(D[0, 2])
    co = c_oracle.COracle(fm)
    assert _tc_bound_ok(D, co.pair_matrix("intersection"))
    P = h.pair_matrix_h("projection")[1]
    assert np.all(np.isfinite(P)) and _tc_bound_ok(P, co.pair_matrix("projection"))
    # rows that are no probabilities (the `deltas` ingest can produce them) have no fixed-point form: CUDA cores
    d = {"": dict(zip("ACGT", (1.5, -0.5, 0.0, 0.0)))}
    h.load_models(FlatModels.from_trees([a, d, c]), skip_nll=True)
    h.pair_matrix_h("projection")
    assert h.last_pair_kernel() == 0


def test_tree_files_to_device_tables(tmp_path):
    """Row f3 end to end: `.tree` text -> native parser (vgs_parse_tree_text) -> FlatModels -> device tables: lookups and the
    NLL matrix bit-exact against the oracle built from the REFERENCE parser's JSON of the same files; the `deltas` ingest
    (src/parse_trees_to_json.py:40-44) gives the same rows as the reference's JSON and loads for the pair distances."""
    import json
    import os

    import golden_io
    from clustering_genomic_signatures_b200 import parse_trees_to_json
    from clustering_genomic_signatures_b200.flat import FlatModels, encode_history
    with open(os.path.join(golden_io.GOLD, "tree_parse_golden.json")) as fh:
        gold = json.load(fh)
    for name, item in gold.items():
        (tmp_path / name).write_text(item["text"])
    fm, occ = parse_trees_to_json.load_tree_dir_flat(str(tmp_path))
    names = sorted(gold)
    trees = [gold[k]["json"]["tree"] for k in names]
    ref_fm = FlatModels.from_trees(trees)
    assert np.array_equal(fm.prob, ref_fm.prob) and np.array_equal(fm.ctx_code, ref_fm.ctx_code)
    h = native.Handle(0)
    h.load_models(fm)
    # lookups: every context is found as itself; a random longer history resolves to the oracle's longest suffix
    rng = np.random.default_rng(5)
    models, codes, lens, want = [], [], [], []
    for m, t in enumerate(trees):
        keys = list(t)
        for i, c in enumerate(keys):
            ln, code = encode_history(c)
            models.append(m); codes.append(code); lens.append(ln); want.append(i)
        for _ in range(200):
            hist = "".join(rng.choice(list("ACGT"), size=int(rng.integers(0, 12))))
            ln, code = encode_history(hist)
            models.append(m); codes.append(code); lens.append(ln)
            want.append(keys.index(O.get_context(t, O.order_of(t), hist)))
    got = h.lookup(models, codes, lens)
    assert np.array_equal(got, np.array(want, dtype=np.int32))
    seqs = synthetic.sample_sequences(fm, 700, 3)
    h.load_sequences_packed(*pack_sequences(seqs))
    D = h.nll_matrix_h(700, "sequential")[1]
    assert np.array_equal(D, c_oracle.COracle(ref_fm).nll_matrix(seqs))
    # deltas mode
    with open(os.path.join(golden_io.GOLD, "tree_parse_deltas_golden.json")) as fh:
        gold_d = json.load(fh)
    d_dir = tmp_path / "deltas"
    d_dir.mkdir()
    for name, item in gold_d.items():
        (d_dir / name).write_text(item["text"])
    fmd, _ = parse_trees_to_json.load_tree_dir_flat(str(d_dir), deltas=True)
    ref_d = FlatModels.from_trees([gold_d[k]["json"]["tree"] for k in sorted(gold_d)])
    assert np.array_equal(fmd.prob, ref_d.prob)
    h.load_models(fmd, skip_nll=True)
    F = h.pair_matrix_h("frobenius_intersection")[1]
    assert rel_err(F, c_oracle.COracle(ref_d).pair_matrix("intersection")) < 1e-9
