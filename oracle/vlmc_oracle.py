"""CPU restatement of the reference's all-pairs VLMC distance path -- TEST INFRASTRUCTURE ONLY.

This file is the *oracle*: a plain Python / numpy restatement of the algorithms of
Kalior/clustering-genomic-signatures for the one hot path this repository accelerates.  It is
never imported by the product package (``clustering-genomic-signatures_b200/``); only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg use it, and
only as the checker.

Parity status: PINNED.  The reference holds no known-answer tests of its own,
so the oracle is pinned by executing the reference itself: ``oracle/build_ref.py`` compiles
the unmodified reference sources into ``oracle/_ref`` and ``oracle/make_goldens.py`` records
its outputs under ``tests/golden/``; ``tests/test_oracle_vs_goldens.py`` checks every function
below against those vectors (NLL and lookups bit-exact, Estimate to 1e-12, Frobenius/Projection
to the reference's own float32 rounding).

Every function cites the reference file:line (relative to /root/reference) it restates.
Trees are ``dict[str, dict[str, float]]`` exactly as in the reference (``src/vlmc/vlmc.pyx:9``);
iteration order of the dict is the canonical context index.
"""
import bisect
import itertools
import math
import re

import numpy as np

ALPHABET = ["A", "C", "G", "T"]  # src/vlmc/vlmc.pyx:21


# ----------------------------------------------------------------------------- model (a1-a5)
def order_of(tree):
    """src/vlmc/vlmc.pyx:179-180 -- order = max context length."""
    return max(len(k) for k in tree.keys())


def strip_parameters_from_name(signature_name):
    """src/vlmc/vlmc.pyx:62-70 -- 'pst_vlmc_name_23_23' -> 'vlmc_name'."""
    parts = signature_name.split("_")
    kept = [p for p in parts[1:len(parts) - 2] if p != ""]
    return "_".join(kept)


def get_context(tree, order, sequence):
    """src/vlmc/vlmc.pyx:131-141 -- longest suffix (length <= order) that is a key.

    Raises RuntimeError when not even "" matches, like the reference.
    """
    top = min(order, len(sequence))
    for ln in range(top, -1, -1):
        cand = sequence[len(sequence) - ln:] if ln > 0 else ""
        if cand in tree:
            return cand
    raise RuntimeError("get_context vlmc.pyx")


def get_all_contexts(tree, order, sequence):
    """src/vlmc/vlmc.pyx:143-154 -- every suffix (longest first) that is a key."""
    top = min(order, len(sequence))
    out = []
    for ln in range(top, -1, -1):
        cand = sequence[len(sequence) - ln:] if ln > 0 else ""
        if cand in tree:
            out.append(cand)
    return out


def probability(tree, order, character, history):
    """src/vlmc/vlmc.pyx:122-128."""
    return tree[get_context(tree, order, history)][character]


# ----------------------------------------------------------------------------- log-likelihood (a6)
def log_likelihood(tree, order, sequence, skip):
    """src/vlmc/vlmc.pyx:87-101 -- sum_{t>=skip} f(p_t), f(p)=-1000 if p==0 else ln p; fp64, in order."""
    acc = 0.0
    for t in range(skip, len(sequence)):
        hist = sequence[max(0, t - order):t] if order > 0 else sequence[:t]
        p = probability(tree, order, sequence[t], hist)
        if p == 0:
            acc -= 1000
        else:
            acc += math.log(p)
    return acc


def log_likelihood_ignore_initial_bias(tree, order, sequence):
    """src/vlmc/vlmc.pyx:79-82."""
    return log_likelihood(tree, order, sequence, order)


def score_matrix(trees, sequences):
    """M[s][m] = log_likelihood_ignore_initial_bias of sequence s under model m."""
    orders = [order_of(t) for t in trees]
    M = np.zeros((len(sequences), len(trees)), dtype=np.float64)
    for s, seq in enumerate(sequences):
        for m, tree in enumerate(trees):
            M[s, m] = log_likelihood(tree, orders[m], seq, orders[m])
    return M


# ----------------------------------------------------------------------------- NLL distance (a7)
def nll_from_scores(M, length):
    """src/distance/negloglikelihood.pyx:17-26 in exactly the reference's operation order."""
    n = M.shape[0]
    D = np.zeros((n, n), dtype=np.float64)
    for i in range(n):
        for j in range(n):
            d_lr = (M[i, i] - M[i, j]) / length
            d_rl = (M[j, j] - M[j, i]) / length
            D[i, j] = (d_lr + d_rl) / 2
    return D


def nll_matrix(trees, sequences, length):
    return nll_from_scores(score_matrix(trees, sequences), length)


# ----------------------------------------------------------------------------- sampler (a12)
def generate_sequence(tree, order, sequence_length, pre_sample_length, rng):
    """src/vlmc/vlmc.pyx:156-177 with random.choices spelled out.

    ``random.choices(pop, weights)`` (CPython Lib/random.py) is
    ``pop[bisect(cum_weights, random() * total, 0, n - 1)]`` with ``cum_weights =
    list(accumulate(weights))`` and ``total = cum_weights[-1] + 0.0``: one ``random()`` per symbol.
    ``rng`` is a ``random.Random`` (or the ``random`` module) so the MT19937 stream is shared.
    """
    total_length = sequence_length + pre_sample_length
    out = []
    hist = ""
    for _ in range(total_length):
        row = tree[get_context(tree, order, hist)]
        cum = list(itertools.accumulate(row[c] for c in ALPHABET))
        total = cum[-1] + 0.0
        out.append(ALPHABET[bisect.bisect(cum, rng.random() * total, 0, 3)])
        hist = (hist + out[-1])[-order:] if order > 0 else ""
    s = "".join(out)
    return s[len(s) - sequence_length:]


def generate_sequence_from(tree, order, sequence_length, context, rng):
    """src/vlmc/vlmc.pyx:165-172: ``sequence_length`` symbols continuing ``context`` (which is not returned); every
    step hands the last ``order`` symbols of context + output to the longest-suffix lookup (``[-0:]`` is the whole
    string when order = 0, and then only "" can match)."""
    generated = context
    for _ in range(sequence_length):
        hist = generated[-order:] if order > 0 else generated
        row = tree[get_context(tree, order, hist)]
        cum = list(itertools.accumulate(row[c] for c in ALPHABET))
        total = cum[-1] + 0.0
        generated += ALPHABET[bisect.bisect(cum, rng.random() * total, 0, 3)]
    return generated[len(context):]


def likelihood(tree, order, sequence):
    """src/vlmc/vlmc.pyx:103-118: left-to-right product, 0.0 from the first impossible symbol on."""
    value = 1.0
    for t, ch in enumerate(sequence):
        hist = sequence[:t][-order:] if order > 0 else sequence[:t]
        prob = tree[get_context(tree, order, hist)][ch]
        if prob == 0:
            return 0.0
        value *= prob
    return value


def context_distribution(tree, order, sequence, sequence_length):
    """src/vlmc/vlmc.pyx:182-206 for an already sampled ``sequence``: the context of every ``order``-symbol window,
    counted, in order of first occurrence, divided by the requested length."""
    counts = {}
    for i in range(len(sequence) - order):
        c = get_context(tree, order, sequence[i:i + order])
        counts[c] = counts.get(c, 0) + 1
    return {c: n / sequence_length for c, n in counts.items()}


# ----------------------------------------------------------------------------- Frobenius (a8)
def frobenius_terms(tree_l, order_l, tree_r, order_r, use_union):
    """Context set and the two row lists of src/distance/frobenius.pyx:21-53 (sorted for determinism)."""
    if use_union:
        ctxs = set(tree_l.keys()) | set(tree_r.keys())
    else:
        ctxs = set(c for c in tree_l if c in tree_r)
    ctxs = sorted(ctxs, key=lambda c: (len(c), c))

    def rows(tree, order):
        out = np.empty((len(ctxs), 4), dtype=np.float32)
        for i, c in enumerate(ctxs):
            src = tree[c] if c in tree else tree[get_context(tree, order, c)]
            for j, ch in enumerate(ALPHABET):
                out[i, j] = src[ch]
        return out

    return ctxs, rows(tree_l, order_l), rows(tree_r, order_r)


def frobenius(tree_l, tree_r, use_union=False):
    """src/distance/frobenius.pyx:21-36.  float32 subtraction like the reference, then an *exact*
    fp64 sum of squares (the reference uses a float32 BLAS dot whose order is implementation
    defined: it agrees with this value to ~1e-7 rel)."""
    ctxs, a, b = frobenius_terms(tree_l, order_of(tree_l), tree_r, order_of(tree_r), use_union)
    diff = (a - b).astype(np.float64)  # float32 subtraction first
    ssq = math.fsum((diff * diff).ravel().tolist())
    n = len(ctxs)
    if n == 0:
        return float("nan")  # reference: 0/0 -> nan with a numpy warning (frobenius.pyx:36)
    return math.sqrt(ssq) / math.sqrt(n)


def frobenius_matrix(trees, use_union=False):
    n = len(trees)
    D = np.zeros((n, n), dtype=np.float64)
    for i in range(n):
        for j in range(n):
            D[i, j] = frobenius(trees[i], trees[j], use_union)
    return D


# ----------------------------------------------------------------------------- Projection (a9)
def projection_matrix(trees):
    """src/distance/projection.pyx:7-36 -- Euclidean norm over the union universe, zeros for
    absent contexts, float32 entries; exact fp64 sum of squares of the float32 differences."""
    universe = sorted(set(itertools.chain.from_iterable(t.keys() for t in trees)), key=lambda c: (len(c), c))
    index = {c: i for i, c in enumerate(universe)}
    vecs = np.zeros((len(trees), len(universe), 4), dtype=np.float32)
    for m, t in enumerate(trees):
        for c, row in t.items():
            for j, ch in enumerate(ALPHABET):
                vecs[m, index[c], j] = row[ch]
    n = len(trees)
    D = np.zeros((n, n), dtype=np.float64)
    for i in range(n):
        for j in range(n):
            diff = (vecs[i] - vecs[j]).astype(np.float64)
            D[i, j] = math.sqrt(math.fsum((diff * diff).ravel().tolist()))
    return D, 4 * len(universe)


# ----------------------------------------------------------------------------- Estimate (a10)
def count_events(tree, order, sequence):
    """src/distance/estimate.pyx:53-71 restated as substring counts:
    counts[c][x] = #{t : sequence[t-|c|:t] == c and sequence[t] == x}."""
    counts = {c: [0, 0, 0, 0] for c in tree.keys()}
    sym = {"A": 0, "C": 1, "G": 2, "T": 3}
    by_len = {}
    for c in tree.keys():
        by_len.setdefault(len(c), set()).add(c)
    for t, ch in enumerate(sequence):
        x = sym[ch]
        for ln, group in by_len.items():
            if ln <= t and ln <= order:
                c = sequence[t - ln:t]
                if c in group:
                    counts[c][x] += 1
    return counts


def add_pseudo_counts(counts):
    """src/distance/estimate.pyx:73-78."""
    for c, v in counts.items():
        if sum(v) != 0 and any(k == 0 for k in v):
            counts[c] = [k + 1 for k in v]
    return counts


def estimate_statistic(tree, counts):
    """src/distance/estimate.pyx:40-51 and :80-108 -- Pearson sum over visited contexts, over the
    symbols with p>0 except the last such symbol; observed = count/total, expected = p."""
    obs, exp = [], []
    for c in tree.keys():
        total = sum(counts[c])
        if total > 0:
            pos = [(j, tree[c][ch]) for j, ch in enumerate(ALPHABET) if tree[c][ch] > 0]
            for j, p in pos[:-1]:
                obs.append(counts[c][j] / total)
                exp.append(p)
    obs = np.array(obs, dtype=np.float64)
    exp = np.array(exp, dtype=np.float64)
    terms = (obs - exp) ** 2 / exp  # scipy.stats.power_divergence, lambda_ = 1 ("pearson")
    return float(np.sum(terms, axis=0)) if len(terms) else 0.0


def estimate_distance(tree_left, right_sequence):
    """src/distance/estimate.pyx:19-38 -- asymmetric: left's contexts counted along right's sequence."""
    order = order_of(tree_left)
    counts = add_pseudo_counts(count_events(tree_left, order, right_sequence))
    return estimate_statistic(tree_left, counts)


def estimate_matrix(trees, sequences):
    n = len(trees)
    D = np.zeros((n, n), dtype=np.float64)
    for i in range(n):
        for j in range(n):
            D[i, j] = estimate_distance(trees[i], sequences[j])
    return D


# ----------------------------------------------------------------------------- matrix assembly (a11)
def distances_within(D):
    """src/clustering/util.pyx:6-21 -- float32 [N*N, 3] rows (i, j, d), row-major in (i, j)."""
    n = D.shape[0]
    out = np.zeros((n * n, 3), dtype=np.float32)
    k = 0
    for i in range(n):
        for j in range(n):
            out[k, 0] = i
            out[k, 1] = j
            out[k, 2] = D[i, j]
            k += 1
    return out


def index_distances(n, triples):
    """src/clustering/util.pyx:23-33 -- float32 [N, N]."""
    out = np.zeros((n, n), dtype=np.float32)
    for row in triples:
        out[int(row[0]), int(row[1])] = row[2]
    return out


# ----------------------------------------------------------------------------- .tree ingest (f3)
_NUM = re.compile(r"-?[0-9]+\.?[0-9]*")
_STR = re.compile(r"[A-Za-z#]+")


def parse_tree_text(text, deltas=False):
    """src/parse_trees_to_json.py:18-55 -- 'Node:' lines -> (tree, occurrence_probability)."""
    tree, occ = {}, {}
    total_occurrences = -1
    i = 0
    for line in text.splitlines():
        if line[0:5] != "Node:":
            continue
        numbers = _NUM.findall(line)
        total = sum(int(numbers[k]) for k in (6, 7, 8, 9))
        if deltas:
            row = {ch: float(numbers[14 + j]) for j, ch in enumerate(ALPHABET)}
        else:
            row = {ch: int(numbers[6 + j]) / total for j, ch in enumerate(ALPHABET)}
        key = _STR.findall(line)[1]
        if key == "#":
            key = ""
        tree[key] = row
        if i == 0:
            total_occurrences = total
        occ[key] = total / (total_occurrences - max(len(key) - 1, 0))
        i += 1
    return tree, occ


# ----------------------------------------------------------------------------- average link (f1)
def average_link(D32, k):
    """UPGMA restating src/clustering/average_link_clustering.pyx:21-146 on the float32 [N,N]
    matrix: repeatedly merge the pair of clusters with the smallest average distance (strict '<',
    first found wins), Lance-Williams size-weighted update (:119-122), until k clusters remain.
    Returns (labels, merge_distances).  Arithmetic follows the reference's mixed precision:
    original entries are np.float32, merged entries are Python floats."""
    n = D32.shape[0]
    keys = [(i,) for i in range(n)]  # dict insertion order of cluster_heaps
    dist = {a: {b: D32[a[0], b[0]] for b in keys if b != a} for a in keys}
    members = {a: [a[0]] for a in keys}
    merges = []
    while len(keys) > k:
        best, bl, br = np.inf, None, None
        for a in keys:
            if not dist[a]:
                continue
            # heap top == smallest (distance, key) tuple
            d, b = min(((v, kk) for kk, v in dist[a].items()), key=lambda t: (t[0], t[1]))
            if d < best:
                best, bl, br = d, a, b
        if bl is None:
            break
        merges.append(float(best))
        nl, nr = len(members[bl]), len(members[br])
        new_key = tuple(sorted(members[bl] + members[br]))
        new_d = {}
        for a in keys:
            if a in (bl, br):
                continue
            ld = dist[a].pop(bl)
            rd = dist[a].pop(br)
            v = nl * ld + nr * rd
            v /= nl + nr
            v = float(v)  # cdef double return (:119)
            dist[a][new_key] = v
        for a in dist[bl]:
            if a in dist[br] and a not in (bl, br):
                v = nl * dist[bl][a] + nr * dist[br][a]
                v /= nl + nr
                new_d[a] = float(v)
        members[new_key] = members[bl] + members[br]
        for dead in (bl, br):
            dist.pop(dead)
            members.pop(dead)
            keys.remove(dead)
        dist[new_key] = new_d
        keys.append(new_key)
    labels = np.zeros(n, dtype=np.int64)
    for ci, a in enumerate(sorted(keys, key=lambda kk: min(members[kk]))):
        for i in members[a]:
            labels[i] = ci
    return labels, merges


# ----------------------------------------------------------------------------- metrics (f4)
def retrieval_metrics(D, labels):
    """Per-model retrieval metrics restating src/test_distance_function.py:96-104 (``calculate_distances``:
    distances to every model, ``sorted`` on the distance alone, stable) and src/util/distance_metrics.py:43-60
    (``procent_of_taxonomy_in_top``, ``average_distance_to_taxonomy``) plus the row mean of :11.
    D: [N, N] array of the values ``d.distance`` returns (Python floats); labels: one taxon label per model.
    Returns (in_top[N], dist_to[N], avg[N]) as Python-float lists, summed in the reference's (sorted) order."""
    n = len(labels)
    in_top, dist_to, avg = [], [], []
    for i in range(n):
        sorted_results = sorted(zip([float(x) for x in D[i]], range(n)), key=lambda t: t[0])
        same = [j for j in range(n) if labels[j] == labels[i]]
        k = len(same)
        hits = 0
        for r in range(k):
            _, j = sorted_results[r]
            if labels[j] == labels[i]:
                hits += 1
        in_top.append(hits / k)
        st = [abs(dist) for dist, j in sorted_results if labels[j] == labels[i]]
        dist_to.append(sum(st) / len(st))
        avg.append(sum([dist for dist, _ in sorted_results]) / len(sorted_results))
    return in_top, dist_to, avg


def silhouette(D32, clusters):
    """src/clustering/clustering_metrics.pyx:35-84 on the float32 [N, N] matrix.  Sums accumulate in numpy
    float32 like the reference (here in ascending member order; the reference uses set order) and the two
    averages pass through a C float.  Returns (a[N], b[N], s[N]); s is NaN where the reference would fail
    (single cluster: inf/inf; a == b == 0: 0/0)."""
    D32 = np.asarray(D32, dtype=np.float32)
    n = D32.shape[0]
    comps = {}
    for i, c in enumerate(clusters):
        comps.setdefault(c, []).append(i)
    comps = list(comps.values())
    a = np.zeros(n)
    b = np.zeros(n)
    for comp in comps:
        for v1 in comp:
            total = 0
            for v2 in comp:
                if v1 == v2:
                    continue
                total += D32[v1, v2]
            others = len(comp) - 1
            a[v1] = float(np.float32(0 if others == 0 else total / others))
            best = np.inf
            for other in comps:
                if other is comp:
                    continue
                total = 0
                for v2 in other:
                    total += D32[v1, v2]
                average = float(np.float32(total / len(other)))
                if average < best:
                    best = average
            b[v1] = best
    with np.errstate(invalid="ignore", divide="ignore"):
        s = (b - a) / np.where(a > b, a, b)
    return a, b, s


def cluster_distance_std(D, clusters):
    """src/clustering/clustering_metrics.pyx:168-185: numpy.std of the distances between distinct members of each
    cluster (0 for a singleton), clusters in ascending label order."""
    comps = {}
    for i, c in enumerate(clusters):
        comps.setdefault(c, []).append(i)
    out = []
    for c in sorted(comps):
        comp = comps[c]
        vals = [float(D[x, y]) for x in comp for y in comp if x != y]
        out.append(0.0 if len(vals) < 1 else float(np.std(vals)))
    return out
