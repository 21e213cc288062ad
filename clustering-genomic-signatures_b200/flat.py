"""Flat (structure-of-arrays) form of a set of VLMC context tables.

The reference keeps a VLMC as ``dict[str ctx] -> dict['A','C','G','T'] -> float``
(``src/vlmc/vlmc.pyx:9``).  The device path wants every model of a set in one contiguous,
pointer-free layout.  ``FlatModels`` is that layout on the host; ``native.Handle.load_models``
hands the arrays to the C-ABI library, which derives the device-side tables (hash, dense
longest-suffix maps, fp32 rows, host-``log`` tables) from them.

Encoding: alphabet order A,C,G,T -> 0..3 (``src/vlmc/vlmc.pyx:21``).  A
context string reads oldest -> newest; its *code* packs two bits per symbol with the newest
symbol in the two lowest bits, so a rolling history ``h = (h << 2) | x`` yields the code of the
length-k suffix as ``h & (4**k - 1)``.  ``ctx_id`` of a context is its position in the tree
dict's iteration order -- the only canonical index the reference has.
"""
from dataclasses import dataclass

import numpy as np

ALPHABET = ["A", "C", "G", "T"]
SYMBOL_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
MAX_ORDER = 15  # key = (1 << 2*len) | code must fit in 32 bits


def encode_context(ctx):
    """'' -> (0, 0); 'CAA' -> (3, 0b010000).  KeyError on a non-ACGT symbol like the reference
    (``src/vlmc/vlmc.pyx:125``)."""
    code = 0
    for ch in ctx:
        code = (code << 2) | SYMBOL_CODE[ch]
    return len(ctx), code


def decode_context(length, code):
    return "".join(ALPHABET[(code >> (2 * (length - 1 - k))) & 3] for k in range(length))


@dataclass
class FlatModels:
    """All context tables of a model set, concatenated in model order then dict order."""

    ctx_offsets: np.ndarray  # int64  [n_models + 1]
    ctx_len: np.ndarray      # uint8  [total]
    ctx_code: np.ndarray     # uint32 [total]
    prob: np.ndarray         # float64 [total, 4]  (bit-identical to the dict values)
    names: list

    @property
    def n_models(self):
        return len(self.ctx_offsets) - 1

    @property
    def total_contexts(self):
        return int(self.ctx_offsets[-1])

    def n_ctx(self, m):
        return int(self.ctx_offsets[m + 1] - self.ctx_offsets[m])

    def orders(self):
        out = np.zeros(self.n_models, dtype=np.int32)
        for m in range(self.n_models):
            a, b = int(self.ctx_offsets[m]), int(self.ctx_offsets[m + 1])
            out[m] = int(self.ctx_len[a:b].max()) if b > a else 0
        return out

    def validate(self):
        if self.ctx_len.size and int(self.ctx_len.max()) > MAX_ORDER:
            raise ValueError("context longer than %d symbols is not supported" % MAX_ORDER)
        if np.any(np.diff(self.ctx_offsets) <= 0):
            raise ValueError("every VLMC needs at least one context")
        # negative probabilities are not rejected here: the reference only fails on them when it takes their
        # logarithm (ValueError from math.log, src/vlmc/vlmc.pyx:99) -- vgs_load_models does the same when it
        # derives ln p for the NLL path, and Frobenius / Projection / Estimate accept them like the reference
        return self

    @classmethod
    def from_trees(cls, trees, names=None):
        """trees: iterable of ``dict[str, dict[str, float]]`` in the reference's format."""
        trees = list(trees)
        offsets = np.zeros(len(trees) + 1, dtype=np.int64)
        for m, t in enumerate(trees):
            offsets[m + 1] = offsets[m] + len(t)
        total = int(offsets[-1])
        ctx_len = np.zeros(total, dtype=np.uint8)
        ctx_code = np.zeros(total, dtype=np.uint32)
        prob = np.zeros((total, 4), dtype=np.float64)
        k = 0
        for t in trees:
            for ctx, row in t.items():
                if len(ctx) > MAX_ORDER:
                    raise ValueError("context longer than %d symbols is not supported" % MAX_ORDER)
                ln, code = encode_context(ctx)
                ctx_len[k] = ln
                ctx_code[k] = code
                prob[k, 0] = row["A"]
                prob[k, 1] = row["C"]
                prob[k, 2] = row["G"]
                prob[k, 3] = row["T"]
                k += 1
        if names is None:
            names = ["m%d" % i for i in range(len(trees))]
        return cls(offsets, ctx_len, ctx_code, prob, list(names)).validate()

    def tree(self, m):
        """Back to the reference's dict-of-dicts (same iteration order, same doubles)."""
        a, b = int(self.ctx_offsets[m]), int(self.ctx_offsets[m + 1])
        out = {}
        for k in range(a, b):
            out[decode_context(int(self.ctx_len[k]), int(self.ctx_code[k]))] = {
                "A": float(self.prob[k, 0]), "C": float(self.prob[k, 1]),
                "G": float(self.prob[k, 2]), "T": float(self.prob[k, 3])}
        return out

    def subset(self, idx):
        idx = list(idx)
        parts_len, parts_code, parts_prob = [], [], []
        offsets = np.zeros(len(idx) + 1, dtype=np.int64)
        for k, m in enumerate(idx):
            a, b = int(self.ctx_offsets[m]), int(self.ctx_offsets[m + 1])
            parts_len.append(self.ctx_len[a:b])
            parts_code.append(self.ctx_code[a:b])
            parts_prob.append(self.prob[a:b])
            offsets[k + 1] = offsets[k] + (b - a)
        return FlatModels(offsets, np.concatenate(parts_len), np.concatenate(parts_code),
                          np.concatenate(parts_prob), [self.names[m] for m in idx])


def encode_history(hist):
    """History string -> (length, code) using at most the last MAX_ORDER symbols."""
    return encode_context(hist[-MAX_ORDER:] if len(hist) > MAX_ORDER else hist)


def sequences_to_ascii(sequences):
    """list[str] -> (uint8 concatenation, int64 offsets[n+1]) for ``vgs_load_sequences_ascii``."""
    offsets = np.zeros(len(sequences) + 1, dtype=np.int64)
    for i, s in enumerate(sequences):
        offsets[i + 1] = offsets[i] + len(s)
    buf = np.frombuffer("".join(sequences).encode("ascii"), dtype=np.uint8)
    return buf, offsets


def pack_sequences(codes, lengths=None):
    """Pack symbol codes (uint8 in 0..3) 16 per uint32, FIRST symbol in the HIGHEST two bits, so
    that the 2(k)-bit field holding a k-mer is contiguous with its oldest symbol on top and a
    single funnel shift extracts it.  Each sequence starts on a 16-byte boundary (128-bit loads).

    codes: 2-D uint8 array [n_seq, L] (equal lengths) or list of 1-D arrays.
    Returns (words uint32, word_offsets int64 [n+1], lengths int64 [n]).
    """
    if isinstance(codes, np.ndarray) and codes.ndim == 2:
        seqs = [codes[i] for i in range(codes.shape[0])]
    else:
        seqs = list(codes)
    n = len(seqs)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    words_per = ((lens + 15) // 16 + 3) // 4 * 4
    offs = np.zeros(n + 1, dtype=np.int64)
    offs[1:] = np.cumsum(words_per)
    words = np.zeros(int(offs[-1]), dtype=np.uint32)
    shifts = (30 - 2 * np.arange(16)).astype(np.uint32)
    for i, s in enumerate(seqs):
        L = int(lens[i])
        if L == 0:
            continue
        pad = (-L) % 16
        a = np.concatenate([np.asarray(s, dtype=np.uint32), np.zeros(pad, dtype=np.uint32)]).reshape(-1, 16)
        w = np.bitwise_or.reduce(a << shifts, axis=1).astype(np.uint32)
        words[int(offs[i]):int(offs[i]) + len(w)] = w
    return words, offs, lens
