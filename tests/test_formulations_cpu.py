"""CPU checks of the reformulations the kernels rely on (numpy / Python integers, no GPU):
the exact fixed-point form of the k-mer contraction (vgs_kmer_i8.cu), the one-direction form of Projection
(vgs_pairs.cu) and the sort-free stable rank of the retrieval metrics (vgs_metrics.cu)."""
from fractions import Fraction

import numpy as np

import vlmc_oracle as O

SCALE = 2 ** 51


def digits_of(q):
    """Eight balanced base-256 digits of a Python int, as k_i8_digits computes them."""
    out = []
    for _ in range(8):
        d = ((q + 128) & 255) - 128
        out.append(d)
        q = (q - d) >> 8
    assert q == 0
    return out


def fold(S):
    """i8_fold: two int64 halves, one fp64 rounding each, exact power-of-two scaling."""
    lo = S[0] + S[1] * 256 + S[2] * 65536 + S[3] * 16777216
    hi = S[4] + S[5] * 256 + S[6] * 65536 + S[7] * 16777216
    assert abs(lo) < 2 ** 63 and abs(hi) < 2 ** 63
    v = float(np.float64(hi)) * 4294967296.0 + float(np.float64(lo))   # fma(hi, 2^32, lo): both products exact here
    return v / SCALE


def test_fixed_point_contraction_is_exact():
    rng = np.random.Generator(np.random.PCG64(1))
    K = 4096
    p = rng.dirichlet([0.3, 0.3, 0.3, 0.3], size=K // 4).ravel()
    T = np.where(rng.random(K) < 0.02, -1000.0, np.log(np.maximum(p, 1e-300)))   # includes the p == 0 marker
    C = rng.integers(0, 4, size=K)
    C[rng.integers(0, K, 7)] += rng.integers(300, 40000, 7)                        # frequent k-mers: above one byte
    q = [int(np.rint(t * SCALE)) for t in T]
    assert max(abs(x) for x in q) < 2 ** 61
    D = np.array([digits_of(x) for x in q])                                        # [K, 8]
    assert D.min() >= -128 and D.max() <= 127
    assert all(sum(int(D[k, j]) * 256 ** j for j in range(8)) == q[k] for k in range(0, K, 97))
    # tensor-core part: low bytes, int32 sums; epilogue part: 256 * (count >> 8) for the frequent k-mers
    lo, hi = C & 255, C >> 8
    S_lo = [int(np.dot(lo.astype(np.int64), D[:, j].astype(np.int64))) for j in range(8)]
    assert max(abs(x) for x in S_lo) < 2 ** 31
    S_hi = [int(sum(256 * int(hi[k]) * int(D[k, j]) for k in np.nonzero(hi)[0])) for j in range(8)]
    got = fold(S_lo) + fold(S_hi)
    exact = Fraction(sum(int(C[k]) * q[k] for k in range(K)), SCALE)                # exact sum of the quantised terms
    assert abs(Fraction(got) - exact) <= Fraction(2, 2 ** 52) * abs(exact)          # two fp64 roundings
    # against the real-number sum: only the 2^-52 quantisation of each ln p
    real = sum(Fraction(int(C[k])) * Fraction(float(T[k])) for k in range(K))
    assert abs(exact - real) <= Fraction(int(C.sum()), 2 ** 52)
    # and an fp64 dot product in any order is no closer
    assert abs(got - float(np.dot(C.astype(np.float64), T))) <= 1e-12 * abs(got)


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
