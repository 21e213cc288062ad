"""CPU checks of the reformulations the kernels rely on (numpy / Python integers, no GPU):
the exact fixed-point form of the k-mer contraction (vgs_kmer_i8.cu), the one-direction form of Projection
(vgs_pairs.cu) and the sort-free stable rank of the retrieval metrics (vgs_metrics.cu)."""
from fractions import Fraction

import numpy as np

import vlmc_oracle as O

import pytest


def digits_of(q, nd):
    """nd balanced base-256 digits of a Python int, as k_i8_digits computes them (a remainder flags 'does not fit')."""
    out = []
    for _ in range(nd):
        d = ((q + 128) & 255) - 128
        out.append(d)
        q = (q - d) >> 8
    return out, q == 0


def fold(S, scale):
    """i8_fold: two int64 halves, exact power-of-two scaling."""
    lo = S[0] + S[1] * 256 + S[2] * 65536 + S[3] * 16777216
    hi = 0
    for p in range(len(S) - 1, 3, -1):
        hi = hi * 256 + S[p]
    assert abs(lo) < 2 ** 63 and abs(hi) < 2 ** 63
    v = float(np.float64(hi)) * 4294967296.0 + float(np.float64(lo))   # fma(hi, 2^32, lo): both products exact here
    return v / scale


@pytest.mark.parametrize("frac_bits,nd,with_zero_p", [(43, 6, False), (43, 7, True), (51, 7, False), (51, 8, True)])
def test_fixed_point_contraction_is_exact(frac_bits, nd, with_zero_p):
    """round(ln p * 2^f) in nd digits: six hold |ln p| < 16 at the default f = 43, seven the -1000 of p = 0 as well."""
    SCALE = 2 ** frac_bits
    rng = np.random.Generator(np.random.PCG64(1))
    K = 4096
    p = rng.dirichlet([0.3, 0.3, 0.3, 0.3], size=K // 4).ravel()
    T = np.log(np.maximum(p, 2e-7))
    if with_zero_p:
        T = np.where(rng.random(K) < 0.02, -1000.0, T)                             # includes the p == 0 marker
    C = rng.integers(0, 4, size=K)
    C[rng.integers(0, K, 7)] += rng.integers(300, 40000, 7)                        # frequent k-mers: above one byte
    q = [int(np.rint(t * SCALE)) for t in T]
    dg = [digits_of(x, nd) for x in q]
    assert all(ok for _, ok in dg)
    if with_zero_p:
        assert not all(digits_of(x, nd - 1)[1] for x in q)                         # ... and one digit fewer would not do
    D = np.array([d for d, _ in dg])                                               # [K, nd]
    assert D.min() >= -128 and D.max() <= 127
    assert all(sum(int(D[k, j]) * 256 ** j for j in range(nd)) == q[k] for k in range(0, K, 97))
    # tensor-core part: low bytes, int32 sums; epilogue part: 256 * (count >> 8) for the frequent k-mers
    lo, hi = C & 255, C >> 8
    S_lo = [int(np.dot(lo.astype(np.int64), D[:, j].astype(np.int64))) for j in range(nd)]
    assert max(abs(x) for x in S_lo) < 2 ** 31
    S_hi = [int(sum(256 * int(hi[k]) * int(D[k, j]) for k in np.nonzero(hi)[0])) for j in range(nd)]
    got = fold([a + b for a, b in zip(S_lo, S_hi)], SCALE)
    exact = Fraction(sum(int(C[k]) * q[k] for k in range(K)), SCALE)                # exact sum of the quantised terms
    assert abs(Fraction(got) - exact) <= Fraction(3, 2 ** 53) * abs(exact)          # the fold's fp64 roundings
    # against the real-number sum: only the 2^-(f+1) quantisation of each ln p
    real = sum(Fraction(int(C[k])) * Fraction(float(T[k])) for k in range(K))
    assert abs(exact - real) <= Fraction(int(C.sum()), 2 ** (frac_bits + 1))
    # and an fp64 dot product in any order is no closer than the stated bound
    assert abs(got - float(np.dot(C.astype(np.float64), T))) <= max(1e-12, 2.0 ** -(frac_bits + 1) / 1.0) * abs(got)


def test_byte_parallel_digit_split_equals_the_digit_loop():
    """k_i8_digits splits all digits of a value at once: q + 0x80..80 (nd bytes) has the balanced digits, plus 128, as its plain
    bytes; what is left above nd bytes tells whether q fits."""
    rng = np.random.Generator(np.random.PCG64(7))
    for nd in (6, 7, 8):
        bias = 0x8080808080808080 >> (8 * (8 - nd))
        for e in range(1, 61):
            for q in [int(x) for x in rng.integers(-2 ** e, 2 ** e, size=200)] + [2 ** e - 1, -2 ** e, 2 ** (e - 1), -2 ** (e - 1) - 1]:
                want, fits = digits_of(q, nd)
                b = q + bias
                left = (b >> (8 * nd)) if nd < 8 else 0
                dg = (b & (2 ** 64 - 1)) ^ bias
                got = [((dg >> (8 * p)) & 255) - (256 if (dg >> (8 * p)) & 128 else 0) for p in range(nd)]
                assert (left == 0) == fits
                if fits:
                    assert got == want


def test_projection_one_direction_identity():
    rng = np.random.Generator(np.random.PCG64(2))
    ctx = ["", "A", "C", "G", "T", "AA", "CA", "GT", "TT", "ACG", "TTA"]
    def model(keys):
        return {c: dict(zip("ACGT", rng.dirichlet([2, 2, 2, 2]))) for c in keys}
    x = model(ctx[:9])
    y = model([ctx[i] for i in (0, 1, 2, 4, 5, 8, 9, 10)])
    P, dim = O.projection_matrix([x, y])
    def rows32(t):
        return {c: np.array([r[ch] for ch in "ACGT"], dtype=np.float32) for c, r in t.items()}
    rx, ry = rows32(x), rows32(y)
    shared = [c for c in rx if c in ry]
    S = sum(float(np.sum((rx[c] - ry[c]).astype(np.float64) ** 2)) for c in shared)   # float32 subtraction, fp64 squares
    n_x = sum(float(np.sum(r.astype(np.float64) ** 2)) for r in rx.values())
    n_y = sum(float(np.sum(r.astype(np.float64) ** 2)) for r in ry.values())
    A_x = sum(float(np.sum(rx[c].astype(np.float64) ** 2)) for c in shared)
    A_y = sum(float(np.sum(ry[c].astype(np.float64) ** 2)) for c in shared)
    d = np.sqrt(S + (n_x - A_x) + (n_y - A_y))
    assert abs(d - P[0, 1]) <= 3e-7 * P[0, 1]          # the reference's float32 norm
    # every unshared context contributes a whole probability row: its squares sum to at least 1/4
    assert (n_x - A_x) >= 0.25 * (len(rx) - len(shared)) - 1e-12
    assert (n_y - A_y) >= 0.25 * (len(ry) - len(shared)) - 1e-12


def test_stable_rank_without_sorting():
    rng = np.random.Generator(np.random.PCG64(3))
    row = np.round(rng.random(300), 2)                 # many ties
    order = sorted(range(len(row)), key=lambda j: row[j])          # Python's stable sort, as the reference
    rank_sorted = {j: r for r, j in enumerate(order)}
    for j in range(0, len(row), 7):
        rank = int(np.sum(row < row[j]) + np.sum(row[:j] == row[j]))
        assert rank == rank_sorted[j]
