"""Host-side mirror of the reference's ``VLMC`` class (``src/vlmc/vlmc.pyx``) for the device path.

Same constructor, attributes and method names as the reference so that the reference's drivers
(``test_distance_function_``, ``AverageLinkClustering`` ...) accept these objects unchanged:

    VLMC(tree, name, occurrence_probability)      .tree .name .order .sequence .alphabet
    VLMC.from_json / from_json_dir / strip_parameters_from_name / to_json
    get_context / get_all_contexts                (src/vlmc/vlmc.pyx:131-154)   -> device lookup kernel
    log_likelihood / log_likelihood_ignore_initial_bias   (:79-101)             -> device scoring kernel
    generate_sequence / reset_sequence            (:156-177, :208-209)          -> device sampler, CPython MT stream

Nothing here computes on the CPU: every method that the reference implements with dict lookups
goes through ``libvgs.so`` and raises if the library or a CUDA device is missing.
"""
import json
import os
import random

import numpy as np

from . import native
from .flat import ALPHABET, FlatModels, decode_context, encode_history, sequences_to_ascii

_CODE_TO_CHAR = np.frombuffer(b"ACGT", dtype=np.uint8)
_CHAR_TO_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def _default_device():
    try:
        import torch
        return torch.cuda.current_device() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def take_mt_state():
    """CPython's global MT19937 state as uint32[625] (random.getstate()[1])."""
    version, state, gauss = random.getstate()
    return np.array(state, dtype=np.uint32), (version, gauss)


def put_mt_state(arr, tag):
    version, gauss = tag
    random.setstate((version, tuple(int(x) for x in arr), gauss))


class VLMC(object):
    def __init__(self, tree, name, occurrence_probability):
        self.tree = tree
        self.name = name
        self.order = max(len(k) for k in tree.keys())  # src/vlmc/vlmc.pyx:179-180
        self.sequence = ""
        self.alphabet = list(ALPHABET)
        self.occurrence_probabilites = occurrence_probability
        self._handle = None
        self._keys = None

    # identity is by name, as in the reference (src/vlmc/vlmc.pyx:24-31)
    def __str__(self):
        return self.name

    def __hash__(self):
        return hash(self.name)

    def __eq__(self, other):
        return other is not None and self.name == getattr(other, "name", None)

    # ---- construction (src/vlmc/vlmc.pyx:33-70)
    @classmethod
    def from_json(cls, s, name=""):
        info = json.loads(s)
        if "tree" in info and "occurrence_probability" in info:
            return cls(info["tree"], name, info["occurrence_probability"])
        return cls(info, name, {})  # old format: the json is the bare tree

    @classmethod
    def from_json_dir(cls, directory):
        out = []
        for file in [f for f in os.listdir(directory) if f.endswith(".json")]:
            stem, _ = os.path.splitext(file)
            with open(os.path.join(directory, file)) as fh:
                d = json.load(fh)
            out.append(cls(d["tree"], cls.strip_parameters_from_name(stem), d["occurrence_probability"]))
        return out

    @classmethod
    def strip_parameters_from_name(cls, signature_name):
        parts = signature_name.split("_")
        return "_".join(p for p in parts[1:len(parts) - 2] if p != "")

    def to_json(self):
        return json.dumps({"tree": self.tree, "occurrence_probability": self.occurrence_probabilites})

    def occurrence_probability(self, state):
        if len(self.occurrence_probabilites) == 0:
            raise RuntimeError("No information about occurrence probabilites")
        return self.occurrence_probabilites[state]

    # ---- device plumbing
    def flat(self):
        return FlatModels.from_trees([self.tree], [self.name])

    def _device(self):
        if self._handle is None:
            h = native.Handle(_default_device())
            h.load_models(self.flat())
            self._handle = h
            self._keys = list(self.tree.keys())
        return self._handle

    # ---- lookups (src/vlmc/vlmc.pyx:131-154)
    def _lookup(self, sequence, want_all):
        for ch in sequence:
            if ch not in ("A", "C", "G", "T"):
                raise KeyError(ch)
        ln, code = encode_history(sequence)
        return self._device().lookup([0], [code], [ln], want_all=want_all)

    def get_context(self, sequence):
        ids = self._lookup(sequence, False)
        if ids[0] < 0:
            raise RuntimeError("get_context vlmc.pyx")
        return self._keys[int(ids[0])]

    def get_all_contexts(self, sequence):
        ids, masks = self._lookup(sequence, True)
        out = []
        for ln in range(min(len(sequence), self.order), -1, -1):
            if int(masks[0]) >> ln & 1:
                out.append(sequence[len(sequence) - ln:] if ln else "")
        return out

    # ---- scoring (src/vlmc/vlmc.pyx:79-101)
    def _score(self, sequence, skip_mode):
        if len(sequence) == 0:
            return 0.0
        h = self._device()
        try:
            buf, off = sequences_to_ascii([sequence])
        except UnicodeEncodeError:
            raise KeyError(sequence)
        h.load_sequences_ascii(buf, off)
        M = h.nll_scores_h(skip_mode, "sequential")
        if np.isnan(M[0, 0]):
            raise RuntimeError("get_context vlmc.pyx")
        return float(M[0, 0])

    def log_likelihood_ignore_initial_bias(self, sequence):
        return self._score(sequence, native.SKIP_ORDER)

    def log_likelihood(self, sequence):
        return self._score(sequence, native.SKIP_NONE)

    def _history_ids(self, sequence, first, last, exact_length=False):
        """ctx_id of the longest-suffix context of the history before every position t in [first, last): the last
        ``order`` symbols before t (``exact_length``: the window sequence[t:t + order] instead).  One device lookup."""
        codes = np.array([_CHAR_TO_CODE[ch] for ch in sequence], dtype=np.uint32)
        t = np.arange(first, last, dtype=np.int64)
        end = t + self.order if exact_length else t
        lens = np.minimum(end, self.order).astype(np.int64)
        hist = np.zeros(len(t), dtype=np.uint32)
        for k in range(self.order, 0, -1):   # oldest symbol first: code = (code << 2) | symbol
            use = lens >= k
            sym = codes[np.where(use, end - k, 0)]
            hist = np.where(use, (hist << np.uint32(2)) | sym, hist).astype(np.uint32)
        ids = self._device().lookup(np.zeros(len(t), dtype=np.int32), hist, lens.astype(np.uint8))
        if (ids < 0).any():
            raise RuntimeError("get_context vlmc.pyx")
        return ids

    def likelihood(self, sequence):
        """Product of the transition probabilities along ``sequence``, 0.0 from the first impossible symbol on
        (src/vlmc/vlmc.pyx:103-118).  The context lookups run on the device; the product is taken left to right."""
        if len(sequence) == 0:
            return 1.0
        for ch in sequence:
            if ch not in _CHAR_TO_CODE:
                break
        else:
            ch = None
        stop = sequence.index(ch) if ch is not None else len(sequence)
        ids = self._history_ids(sequence[:stop], 0, stop) if stop else []
        likelihood = 1.0
        for t in range(stop):
            prob = self.tree[self._keys[int(ids[t])]][sequence[t]]
            if prob == 0:
                return 0.0
            likelihood *= prob
        if ch is not None:
            raise KeyError(ch)
        return likelihood

    def estimated_context_distribution(self, sequence_length):
        """Relative frequency of the context every ``order``-symbol window of a sampled sequence falls into, keyed in
        order of first occurrence (src/vlmc/vlmc.pyx:182-206)."""
        sequence = self.generate_sequence(sequence_length, 500)
        n = len(sequence) - self.order
        if n <= 0:
            return {}
        ids = self._history_ids(sequence, 0, n, exact_length=True)
        uniq, first, counts = np.unique(ids, return_index=True, return_counts=True)
        out = {}
        for k in np.argsort(first, kind="stable"):
            out[self._keys[int(uniq[k])]] = int(counts[k]) / sequence_length
        return out

    # ---- sampling (src/vlmc/vlmc.pyx:156-177)
    def generate_sequence(self, sequence_length, pre_sample_length):
        if len(self.sequence) == sequence_length:
            return self.sequence
        h = self._device()
        state, tag = take_mt_state()
        h.sample_sequences(sequence_length, pre_sample_length, state)
        put_mt_state(state, tag)  # the global stream advanced by length + pre doubles, like the reference
        self.sequence = _CODE_TO_CHAR[h.get_sequence(0, sequence_length)].tobytes().decode("ascii")
        return self.sequence

    def generate_sequence_from(self, sequence_length, context):
        """``sequence_length`` symbols continuing ``context`` (src/vlmc/vlmc.pyx:165-172); neither the cached
        ``.sequence`` nor the context is part of the result.  Symbols outside ACGT can never be part of a matching
        suffix (get_context, :131-141), so the chain effectively starts after the last one."""
        if sequence_length <= 0:
            return ""
        usable = context
        for pos in range(len(context) - 1, -1, -1):
            if context[pos] not in _CHAR_TO_CODE:
                usable = context[pos + 1:]
                break
        codes = np.array([_CHAR_TO_CODE[ch] for ch in usable[-15:]], dtype=np.uint8)
        h = self._device()
        state, tag = take_mt_state()
        h.sample_sequences(sequence_length, 0, state, context_codes=codes)
        put_mt_state(state, tag)  # one random() per symbol, like the reference's loop
        return _CODE_TO_CHAR[h.get_sequence(0, sequence_length)].tobytes().decode("ascii")

    def reset_sequence(self):
        self.sequence = ""


def vlmcs_to_flat(vlmcs):
    return FlatModels.from_trees([v.tree for v in vlmcs], [v.name for v in vlmcs])


def sample_missing_sequences(vlmcs, sequence_length, pre_sample_length, device=None):
    """Give every VLMC whose cached ``.sequence`` has another length a freshly sampled one, consuming
    CPython's global MT19937 stream in list order exactly as the reference's N^2 loop does on first use
    (``src/distance/negloglikelihood.pyx:22-24``, ``src/vlmc/vlmc.pyx:156-163``)."""
    todo = [v for v in vlmcs if len(v.sequence) != sequence_length]
    seen, uniq = set(), []
    for v in todo:  # equal names alias one object in the reference's dict/index based code
        if id(v) not in seen:
            seen.add(id(v))
            uniq.append(v)
    if not uniq:
        return
    h = native.Handle(_default_device() if device is None else device)
    h.load_models(vlmcs_to_flat(uniq), skip_nll=True)   # the sampler reads the transition rows, not their logarithms
    state, tag = take_mt_state()
    h.sample_sequences(sequence_length, pre_sample_length, state)
    put_mt_state(state, tag)
    for i, v in enumerate(uniq):
        v.sequence = _CODE_TO_CHAR[h.get_sequence(i, sequence_length)].tobytes().decode("ascii")
    h.close()
