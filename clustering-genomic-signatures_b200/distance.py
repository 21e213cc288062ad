"""Drop-in distance classes: same names, constructors and ``distance(left, right)`` method as the
reference's ``src/distance`` package, computed on the B200 through ``libvgs.so``.

    NegativeLogLikelihood(sequence_length)      src/distance/negloglikelihood.pyx:13-26
    FrobeniusNorm(use_union=False)              src/distance/frobenius.pyx:14-53
    Projection()  (+ set_vlmcs)                 src/distance/projection.pyx:7-36
    EstimateVLMC(d=None)                        src/distance/estimate.pyx:16-108

The reference evaluates ``d.distance`` once per ordered pair inside Python loops
(``src/clustering/util.pyx:12-14``, ``src/test_distance_function.py:102``).  Here the whole N x N
matrix is one device job: ``set_vlmcs(vlmcs)`` (the reference's own batch hook, honoured by
``test_distance_function_`` for Projection, ``src/test_distance_function.py:63-64``) or
``distance_matrix(vlmcs)`` runs it, and ``distance(left, right)`` then indexes the result.  A
``distance`` call on models that were not batched computes that pair on the device as a 2-model set.
"""
import numpy as np

from . import native
from .flat import sequences_to_ascii
from .vlmc import _default_device, sample_missing_sequences, vlmcs_to_flat


class _DeviceDistance(object):
    """Shared batching / caching logic.  Subclasses implement
    ``_compute(handle, vlmcs, pair_only=False) -> (D32, D64)``."""

    needs_nll_tables = False   # only NegativeLogLikelihood needs ln p and the NLL tables on the device

    def __init__(self):
        self._index = None     # id(vlmc) -> row
        self._vlmcs = None     # the batched objects themselves (strong references: an id() is only unique while its object lives)
        self._D64 = None
        self._D32 = None
        self.device = None
        self._handle = None

    def _get_handle(self):
        """One device handle per distance object, created on first use (its device buffers are grow-only, so an
        N^2 loop of un-batched distance() calls re-uses the same allocations)."""
        dev = _default_device() if self.device is None else self.device
        if self._handle is None or self._handle.device != dev:
            if self._handle is not None:
                self._handle.close()
            self._handle = native.Handle(dev)
        return self._handle

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None

    # -- batch hook (same name as Projection.set_vlmcs in the reference)
    def set_vlmcs(self, vlmcs):
        self.distance_matrix(vlmcs)

    def distance_matrix(self, vlmcs):
        """float32 [N, N] matrix of ``distance(vlmcs[i], vlmcs[j])`` (src/clustering/util.pyx:23-33 layout)."""
        vlmcs = list(vlmcs)
        h = self._get_handle()
        h.load_models(vlmcs_to_flat(vlmcs), skip_nll=not self.needs_nll_tables, nll_only=self._nll_only(vlmcs),
                      device_log=self._device_log())
        D32, D64 = self._compute(h, vlmcs)
        self._D32, self._D64 = D32, D64
        self._vlmcs = vlmcs
        self._index = {id(v): i for i, v in enumerate(vlmcs)}
        self._batch_done(vlmcs)
        return D32

    def _batch_done(self, vlmcs):
        pass

    def _nll_only(self, vlmcs):
        """True when the device needs nothing but the log-likelihood tables for these models (NegativeLogLikelihood with
        every sequence already sampled): the transition rows are then not uploaded at all."""
        return False

    def _device_log(self):
        """True when ln p may be taken on the device (VGS_LOAD_DEVICE_LOG): only the k-mer contraction reads it."""
        return False

    def _row(self, v):
        """Row of ``v`` in the batched matrix, or None.  A hit needs the very same object: the reference reads the
        two objects' trees on every call and never resolves a model by name (names are stripped of their training
        parameters, so different models share one, src/vlmc/vlmc.pyx:62-70, src/test_increasing_parameter_distance.py)."""
        if self._index is None:
            return None
        r = self._index.get(id(v))
        if r is None or self._vlmcs[r] is not v:
            return None
        return r

    def _value(self, i, j):
        return float(self._D64[i, j])

    def distance(self, left_vlmc, right_vlmc):
        i, j = self._row(left_vlmc), self._row(right_vlmc)
        if i is not None and j is not None and self._still_valid(left_vlmc, right_vlmc):
            return self._value(i, j)
        # not batched: this pair alone, still on the device
        pair = [left_vlmc, right_vlmc]
        h = self._get_handle()
        h.load_models(vlmcs_to_flat(pair), skip_nll=not self.needs_nll_tables, nll_only=self._nll_only(pair),
                      device_log=self._device_log())
        D32, D64 = self._compute(h, pair, pair_only=True)
        return self._scalar(D32, D64)

    def _scalar(self, D32, D64):
        return float(D64[0, 1])

    def _still_valid(self, left, right):
        return True


class NegativeLogLikelihood(_DeviceDistance):
    """D(l, r) = (CE(l, r) + CE(r, l)) / 2 with CE(x, y) = (ln Pr(S_x|x) - ln Pr(S_x|y)) / L on a sequence
    S_x of L symbols sampled from x (cached on the model), src/distance/negloglikelihood.pyx:17-26."""

    length_of_pregenerated_sequence = 500  # src/distance/negloglikelihood.pyx:12
    needs_nll_tables = True

    def __init__(self, sequence_length, mode="kmer"):
        super().__init__()
        self.generated_sequence_length = sequence_length
        # "kmer" (default): the exact int8 tensor-core contraction bench.py times -- within 1e-9 of the reference's
        # fp64 sum (ln p in 2^-43 fixed point: |d distance| <= 1.1e-13), identical clusters downstream; sets deeper than 6 take the "warp" scan.
        # "sequential": one lane per pair, the reference's own summation order, bit-identical fp64.
        self.mode = mode
        self._seq_ids = None

    def _compute(self, h, vlmcs, pair_only=False):
        L = self.generated_sequence_length
        sample_missing_sequences(vlmcs, L, self.length_of_pregenerated_sequence, h.device)
        buf, off = sequences_to_ascii([v.sequence for v in vlmcs])
        h.load_sequences_ascii(buf, off)
        return h.nll_matrix_h(L, self.mode)

    def _batch_done(self, vlmcs):
        self._seq_ids = {id(v): v.sequence for v in vlmcs}

    def _nll_only(self, vlmcs):
        # scoring reads ln p only; missing sequences are sampled on a handle of their own (vlmc.sample_missing_sequences)
        return True

    def _device_log(self):
        # the contraction quantises ln p to 2^-43: the device's log() (<= 1 ulp from libm's) is 64 times finer than that.  The
        # library ignores the flag for sets the contraction does not cover; "sequential" / "warp" always take libm's ln p
        return self.mode == "kmer"

    def _still_valid(self, left, right):
        # the reference re-samples when a model's cached sequence was reset or has another length
        return self._seq_ids.get(id(left)) is left.sequence and self._seq_ids.get(id(right)) is right.sequence


class FrobeniusNorm(_DeviceDistance):
    """||P_l - P_r||_F / sqrt(#C) over the shared (default) or united context set, src/distance/frobenius.pyx:21-36."""

    def __init__(self, use_union=False, path="auto"):
        super().__init__()
        self.use_union = use_union
        # "auto": sets that fill >= 3 % of their universe (depth <= 7) go to the exact integer tensor-core GEMMs (within
        # 1.3e-7 d + 1e-9 of the fp64 result), the rest to the CUDA-core kernels (1e-9); "cuda" / "tensor" force one
        self.path = path

    def _compute(self, h, vlmcs, pair_only=False):
        h.set_pair_path(self.path)
        return h.pair_matrix_h("frobenius_union" if self.use_union else "frobenius_intersection")


class Projection(_DeviceDistance):
    """Euclidean distance between the models embedded over the union of all context sets,
    src/distance/projection.pyx:7-36.  ``distance`` returns ``numpy.float32`` like the reference."""

    def __init__(self, path="auto"):
        super().__init__()
        self.vlmcs = None
        self.dimension = 0
        self.context_transition_to_array_index = None
        self.path = path   # see FrobeniusNorm

    def set_vlmcs(self, vlmcs):
        self.vlmcs = list(vlmcs)
        universe = {}
        for v in self.vlmcs:
            for c in v.tree.keys():
                universe.setdefault(c, len(universe))
        self.context_transition_to_array_index = {c: {ch: 4 * i + k for k, ch in enumerate("ACGT")} for c, i in universe.items()}
        self.dimension = 4 * len(universe)
        self.distance_matrix(self.vlmcs)

    def _compute(self, h, vlmcs, pair_only=False):
        h.set_pair_path(self.path)
        return h.pair_matrix_h("projection")

    def _value(self, i, j):
        return np.float32(self._D64[i, j])

    def _scalar(self, D32, D64):
        return np.float32(D64[0, 1])


class EstimateVLMC(_DeviceDistance):
    """Pearson statistic between the left model and its re-estimation from a 100 000-symbol sequence of
    the right model; asymmetric, src/distance/estimate.pyx:19-38.  ``d`` is accepted and ignored exactly
    as in the reference (its use is commented out there, :36)."""

    pre_sample_length = 500      # src/distance/estimate.pyx:27
    sequence_length = 100000     # src/distance/estimate.pyx:28

    def __init__(self, d=None):
        super().__init__()
        self.d = d
        self._seq_ids = None

    def _compute(self, h, vlmcs, pair_only=False):
        if pair_only:
            # distance(left, right) samples only the RIGHT model's sequence (estimate.pyx:30); the left
            # slot just needs a sequence of some kind for the (unused) other direction
            sample_missing_sequences([vlmcs[1]], self.sequence_length, self.pre_sample_length, h.device)
            seqs = [vlmcs[1].sequence, vlmcs[1].sequence]
        else:
            sample_missing_sequences(vlmcs, self.sequence_length, self.pre_sample_length, h.device)
            seqs = [v.sequence for v in vlmcs]
        buf, off = sequences_to_ascii(seqs)
        h.load_sequences_ascii(buf, off)
        return h.estimate_matrix_h()

    def _batch_done(self, vlmcs):
        self._seq_ids = {id(v): v.sequence for v in vlmcs}

    def _still_valid(self, left, right):
        return self._seq_ids.get(id(right)) is right.sequence
