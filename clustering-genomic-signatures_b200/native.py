"""ctypes binding of ``libvgs.so`` (the C-ABI CUDA library, ``include/vgs.h``).

This is the only module that touches the shared library.  There is no CPU fallback: if the
library has not been built, or no CUDA device is present, the calls raise.
"""
import ctypes
import os

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libvgs.so")

PAIR_KINDS = {"frobenius_intersection": 0, "frobenius_union": 1, "projection": 2}
NLL_MODES = {"sequential": 0, "warp": 1, "kmer": 2}
SKIP_ORDER, SKIP_NONE = 0, 1

_lib = None

c_i64, c_i32, c_int, c_vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_int, ctypes.c_void_p
_P = ctypes.POINTER

# (name, restype, argtypes) -- one entry per declaration in include/vgs.h
SIGNATURES = [
    ("vgs_version", c_int, []),
    ("vgs_last_error", ctypes.c_char_p, []),
    ("vgs_create", c_int, [c_int, _P(c_vp)]),
    ("vgs_destroy", c_int, [c_vp]),
    ("vgs_device_info", c_int, [c_vp, _P(c_int), _P(c_int), _P(c_int), _P(c_i64)]),
    ("vgs_launch_count", c_i64, [c_vp]),
    ("vgs_poll_error", c_int, [c_vp, c_vp]),
    ("vgs_profile_enable", c_int, [c_vp, c_int]),
    ("vgs_profile_read", c_int, [c_vp, c_int, _P(ctypes.c_double), _P(c_i64)]),
    ("vgs_load_models", c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    ("vgs_load_models_ex", c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_int]),
    ("vgs_model_count", c_int, [c_vp, _P(c_i64), _P(c_i64)]),
    ("vgs_model_info", c_int, [c_vp, c_i64, _P(c_int), _P(c_i64), _P(c_int)]),
    ("vgs_get_logp", c_int, [c_vp, c_vp]),
    ("vgs_parse_tree_text", c_int, [ctypes.c_char_p, c_i64, c_int, c_i64, c_vp, c_vp, c_vp, c_vp, _P(c_i64)]),
    ("vgs_lookup", c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    ("vgs_load_sequences_ascii", c_int, [c_vp, c_i64, c_vp, c_vp]),
    ("vgs_load_sequences_packed", c_int, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    ("vgs_sequence_count", c_int, [c_vp, _P(c_i64), _P(c_i64)]),
    ("vgs_get_sequence", c_int, [c_vp, c_i64, c_vp, c_i64]),
    ("vgs_sample_sequences", c_int, [c_vp, c_i64, c_i64, c_vp]),
    ("vgs_sample_sequences_from", c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_i64]),
    ("vgs_nll_scores", c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_vp, c_i64, c_vp]),
    ("vgs_nll_scores_h", c_int, [c_vp, c_int, c_int, c_vp]),
    ("vgs_nll_matrix", c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp]),
    ("vgs_nll_matrix_h", c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    ("vgs_pair_matrix", c_int, [c_vp, c_int, c_vp, c_vp, c_vp]),
    ("vgs_pair_matrix_h", c_int, [c_vp, c_int, c_vp, c_vp]),
    ("vgs_last_pair_kernel", c_int, [c_vp]),
    ("vgs_set_pair_path", c_int, [c_vp, c_int]),
    ("vgs_pair_kernel_choice", c_int, [c_vp, c_int]),
    ("vgs_estimate_matrix", c_int, [c_vp, c_vp, c_vp, c_vp]),
    ("vgs_estimate_matrix_h", c_int, [c_vp, c_vp, c_vp]),
    ("vgs_tile_plan", c_int, [c_i64, c_i64, c_int, c_int, _P(c_i64), _P(c_i64)]),
    ("vgs_nll_tiles", c_int, [c_vp, c_i64, c_int, c_i64, c_int, c_int, c_vp, c_vp]),
    ("vgs_pair_tiles", c_int, [c_vp, c_int, c_i64, c_int, c_int, c_vp, c_vp]),
    ("vgs_estimate_tiles", c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp]),
    ("vgs_assemble", c_int, [c_vp, c_i64, c_i64, c_int, c_int, c_vp, c_vp, c_vp]),
    ("vgs_nll_scores_band", c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_int, c_vp]),
    ("vgs_nll_combine", c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp]),
    ("vgs_nll_sharded_step", c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_int, c_vp, c_int, c_int, ctypes.c_uint32, c_i64, c_vp, c_vp]),
    ("vgs_peer_alloc", c_int, [c_vp, c_i64, _P(c_vp), c_vp]),
    ("vgs_peer_open", c_int, [c_vp, c_vp, _P(c_vp)]),
    ("vgs_peer_close", c_int, [c_vp, c_vp]),
    ("vgs_peer_free", c_int, [c_vp, c_vp]),
    ("vgs_peer_barrier", c_int, [c_vp, c_vp, c_int, c_int, ctypes.c_uint32, c_vp]),
    ("vgs_tensor_peak_i8", c_int, [c_vp, c_int, _P(ctypes.c_double), _P(ctypes.c_double)]),
    ("vgs_tensor_rate_i8", c_int, [c_vp, c_int, c_int, _P(ctypes.c_double), _P(ctypes.c_double)]),
    ("vgs_g8_plan", c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, _P(c_int), _P(c_int)]),
    ("vgs_g8_plan_gran", c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_int, _P(c_int), _P(c_int)]),
    ("vgs_set_kmer_fraction_bits", c_int, [c_vp, c_int]),
    ("vgs_kmer_digit_info", c_int, [c_vp, _P(c_int), _P(c_int)]),
    ("vgs_matrix_to_triples", c_int, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    ("vgs_last_nll_kernel", c_int, [c_vp]),
    ("vgs_average_link", c_int, [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, _P(c_i64)]),
    ("vgs_retrieval_metrics", c_int, [c_vp, c_i64, c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    ("vgs_silhouette", c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    ("vgs_cluster_distance_std", c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, _P(c_i64)]),
]


def lib():
    """Load libvgs.so (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libvgs.so is missing (%s): build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C clustering-genomic-signatures_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, res, args in SIGNATURES:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class VgsError(RuntimeError):
    pass


def _raise(status):
    msg = lib().vgs_last_error().decode("utf-8", "replace")
    if status == 1:
        raise ValueError(msg)
    if status == 3:
        raise RuntimeError(msg)  # "get_context vlmc.pyx", same text as the reference
    if status == 4:
        raise KeyError(msg)
    if status == 5:
        raise NotImplementedError(msg)
    if status == 6:
        raise MemoryError(msg)
    raise VgsError(msg)


def _check(status):
    if status != 0:
        _raise(status)


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(c_vp)


def _dev_ptr(t):
    """torch CUDA tensor (or None / raw int) -> void*."""
    if t is None:
        return None
    if isinstance(t, int):
        return c_vp(t)
    return c_vp(t.data_ptr())


def tile_plan(n, tile, world, symmetric=True):
    nt, slots = c_i64(), c_i64()
    _check(lib().vgs_tile_plan(n, tile, world, int(bool(symmetric)), ctypes.byref(nt), ctypes.byref(slots)))
    return nt.value, slots.value


def parse_tree_text(text, deltas=False):
    """Text (bytes or str) of one ``.tree`` file -> (ctx_len uint8 [n], ctx_code uint32 [n], prob float64 [n, 4],
    occurrence float64 [n]) in file order, parsed by the library (no device needed)."""
    if isinstance(text, str):
        text = text.encode("utf-8")
    n = c_i64()
    _check(lib().vgs_parse_tree_text(text, len(text), int(bool(deltas)), 0, None, None, None, None, ctypes.byref(n)))
    ln = np.zeros(n.value, dtype=np.uint8)
    code = np.zeros(n.value, dtype=np.uint32)
    prob = np.zeros((n.value, 4), dtype=np.float64)
    occ = np.zeros(n.value, dtype=np.float64)
    _check(lib().vgs_parse_tree_text(text, len(text), int(bool(deltas)), n.value, _np_ptr(ln), _np_ptr(code), _np_ptr(prob),
                                     _np_ptr(occ), ctypes.byref(n)))
    return ln, code, prob, occ


def g8_plan(rects, kblocks, n_cta_pairs, max_split=4, gran=32):
    """Unit list of the int8 GEMM for output rectangles [(a_blk, b_row0, b_rows), ...] (host only); ``gran`` = granule of the
    row ranges and unit widths (32; 96 / 224 for the NLL contraction with 6 / 7 digit rows per model).
    Returns (units int32 [n, 6] = a_blk, b_row0, n, kb0, kbn, partial; split)."""
    rects = np.ascontiguousarray(rects, dtype=np.int32).reshape(-1, 3)
    cap = 64 + 16 * max_split * int(sum(-(-int(r[2]) // 32) for r in rects))
    out = np.zeros((cap, 6), dtype=np.int32)
    n, split = c_int(), c_int()
    _check(lib().vgs_g8_plan_gran(_np_ptr(rects), len(rects), kblocks, n_cta_pairs, max_split, gran, _np_ptr(out), cap, ctypes.byref(n),
                                  ctypes.byref(split)))
    return out[: n.value].copy(), split.value


class Handle:
    """One device context: model tables + sequences resident in HBM, kernels on demand."""

    def __init__(self, device=0):
        self._h = c_vp()
        self.device = int(device)
        _check(lib().vgs_create(self.device, ctypes.byref(self._h)))
        self._keep = []
        self.n_models = 0
        self.total_contexts = 0
        self.n_seq = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().vgs_destroy(self._h)
            self._h = c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- info
    def device_info(self):
        sm, ma, mi, fr = c_int(), c_int(), c_int(), c_i64()
        _check(lib().vgs_device_info(self._h, ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi), ctypes.byref(fr)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "free_bytes": fr.value}

    def launch_count(self):
        return int(lib().vgs_launch_count(self._h))

    def profile_enable(self, on=True):
        _check(lib().vgs_profile_enable(self._h, int(bool(on))))

    def profile_read(self, kernel_id):
        """(total device ms, launches) of one kernel kind since the last read."""
        ms, n = ctypes.c_double(), c_i64()
        _check(lib().vgs_profile_read(self._h, kernel_id, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def last_nll_kernel(self):
        """0 scan, 1 fp64 MMA contraction, 2 fp64 FMA contraction, 3 exact int8 tensor-core contraction."""
        return int(lib().vgs_last_nll_kernel(self._h))

    def set_kmer_fraction_bits(self, bits):
        """Fraction bits f of the fixed-point ln p in the k-mer contraction (35..51, default 43): |d distance| <= 2^-f."""
        _check(lib().vgs_set_kmer_fraction_bits(self._h, int(bits)))

    def kmer_digit_info(self):
        """(digit rows per model of the int8 operand as built -- 0 when not built --, fraction bits)."""
        nd, f = c_int(), c_int()
        _check(lib().vgs_kmer_digit_info(self._h, ctypes.byref(nd), ctypes.byref(f)))
        return nd.value, f.value

    def last_pair_kernel(self):
        """0 CUDA-core pair kernels, 1 tensor-core (exact integer GEMM) path."""
        return int(lib().vgs_last_pair_kernel(self._h))

    def pair_kernel_choice(self, kind):
        """Which kernels pair_matrix would use for this kind on the loaded set: 0 CUDA cores, 1 tensor cores."""
        return int(lib().vgs_pair_kernel_choice(self._h, PAIR_KINDS[kind]))

    def set_pair_path(self, path):
        """-1 by universe density (default), 0 CUDA-core kernels, 1 tensor-core path wherever it applies."""
        _check(lib().vgs_set_pair_path(self._h, {"auto": -1, "cuda": 0, "tensor": 1}.get(path, path)))

    def poll_error(self, stream):
        """Synchronise `stream` and raise what the kernels flagged (RuntimeError / KeyError)."""
        _check(lib().vgs_poll_error(self._h, c_vp(stream)))

    # ---- models
    @staticmethod
    def _load_flags(skip_nll, nll_only, lazy_tables, device_log=False):
        return (1 if skip_nll else 0) | (2 if nll_only else 0) | (4 if lazy_tables else 0) | (8 if device_log else 0)

    def load_models(self, fm, logp=None, skip_nll=False, nll_only=False, lazy_tables=False, device_log=False):
        """``device_log``: VGS_LOAD_DEVICE_LOG -- ln p for the k-mer contraction is taken on the device (no host arithmetic in the
        load); the bit-exact modes still get libm's ln p on first use."""
        off = np.ascontiguousarray(fm.ctx_offsets, dtype=np.int64)
        ln = np.ascontiguousarray(fm.ctx_len, dtype=np.uint8)
        code = np.ascontiguousarray(fm.ctx_code, dtype=np.uint32)
        prob = np.ascontiguousarray(fm.prob, dtype=np.float64)
        lp = None if logp is None else np.ascontiguousarray(logp, dtype=np.float64)
        _check(lib().vgs_load_models_ex(self._h, fm.n_models, _np_ptr(off), _np_ptr(ln), _np_ptr(code), _np_ptr(prob), _np_ptr(lp),
                                        self._load_flags(skip_nll, nll_only, lazy_tables, device_log)))
        self.n_models = fm.n_models
        self.total_contexts = fm.total_contexts

    def load_models_raw(self, n_models, off, ln, code, prob, skip_nll=False, nll_only=False, lazy_tables=False, device_log=False):
        """Same as load_models but from already-contiguous (e.g. pinned) arrays; no conversions."""
        _check(lib().vgs_load_models_ex(self._h, n_models, _np_ptr(off), _np_ptr(ln), _np_ptr(code), _np_ptr(prob), None,
                                        self._load_flags(skip_nll, nll_only, lazy_tables, device_log)))
        self.n_models = n_models
        self.total_contexts = int(off[-1])

    def model_info(self, m):
        order, n_ctx, tier = c_int(), c_i64(), c_int()
        _check(lib().vgs_model_info(self._h, m, ctypes.byref(order), ctypes.byref(n_ctx), ctypes.byref(tier)))
        return {"order": order.value, "n_ctx": n_ctx.value, "tier": tier.value}

    def get_logp(self):
        out = np.zeros((self.total_contexts, 4), dtype=np.float64)
        _check(lib().vgs_get_logp(self._h, _np_ptr(out)))
        return out

    def lookup(self, models, codes, lengths, want_all=False):
        models = np.ascontiguousarray(models, dtype=np.int32)
        codes = np.ascontiguousarray(codes, dtype=np.uint32)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint8)
        n = len(models)
        ids = np.zeros(n, dtype=np.int32)
        masks = np.zeros(n, dtype=np.uint32) if want_all else None
        _check(lib().vgs_lookup(self._h, n, _np_ptr(models), _np_ptr(codes), _np_ptr(lengths), _np_ptr(ids), _np_ptr(masks)))
        return (ids, masks) if want_all else ids

    # ---- sequences
    def load_sequences_ascii(self, buf, offsets):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        _check(lib().vgs_load_sequences_ascii(self._h, len(offsets) - 1, _np_ptr(buf), _np_ptr(offsets)))
        self.n_seq = len(offsets) - 1

    def load_sequences_packed(self, words, word_offsets, lengths):
        words = np.ascontiguousarray(words, dtype=np.uint32)
        word_offsets = np.ascontiguousarray(word_offsets, dtype=np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        _check(lib().vgs_load_sequences_packed(self._h, len(lengths), _np_ptr(words), _np_ptr(word_offsets), _np_ptr(lengths)))
        self.n_seq = len(lengths)

    def get_sequence(self, s, length):
        out = np.zeros(length, dtype=np.uint8)
        _check(lib().vgs_get_sequence(self._h, s, _np_ptr(out), length))
        return out

    def sample_sequences(self, length, pre_sample, mt_state, context_codes=None):
        """mt_state: uint32[625] view of random.getstate()[1]; advanced in place.  ``context_codes`` (uint8 symbol
        codes, oldest first): the history every chain starts from (generate_sequence_from, vlmc.pyx:165-172)."""
        assert mt_state.dtype == np.uint32 and mt_state.size == 625 and mt_state.flags["C_CONTIGUOUS"]
        if context_codes is None or len(context_codes) == 0:
            _check(lib().vgs_sample_sequences(self._h, length, pre_sample, _np_ptr(mt_state)))
        else:
            ctx = np.ascontiguousarray(context_codes, dtype=np.uint8)
            _check(lib().vgs_sample_sequences_from(self._h, length, pre_sample, _np_ptr(mt_state), _np_ptr(ctx), ctx.size))
        self.n_seq = self.n_models

    # ---- host-buffer entry points (copies inside)
    def nll_scores_h(self, skip_mode=SKIP_ORDER, mode="sequential"):
        M = np.zeros((self.n_seq, self.n_models), dtype=np.float64)
        _check(lib().vgs_nll_scores_h(self._h, skip_mode, NLL_MODES[mode], _np_ptr(M)))
        return M

    def nll_matrix_h(self, length, mode="sequential", want_f64=True, out32=None):
        n = self.n_models
        D32 = out32 if out32 is not None else np.zeros((n, n), dtype=np.float32)
        D64 = np.zeros((n, n), dtype=np.float64) if want_f64 else None
        _check(lib().vgs_nll_matrix_h(self._h, length, NLL_MODES[mode], _np_ptr(D32), _np_ptr(D64)))
        return D32, D64

    def pair_matrix_h(self, kind, want_f64=True, out32=None):
        n = self.n_models
        D32 = out32 if out32 is not None else np.zeros((n, n), dtype=np.float32)
        D64 = np.zeros((n, n), dtype=np.float64) if want_f64 else None
        _check(lib().vgs_pair_matrix_h(self._h, PAIR_KINDS[kind], _np_ptr(D32), _np_ptr(D64)))
        return D32, D64

    def estimate_matrix_h(self, want_f64=True):
        n = self.n_models
        D32 = np.zeros((n, n), dtype=np.float32)
        D64 = np.zeros((n, n), dtype=np.float64) if want_f64 else None
        _check(lib().vgs_estimate_matrix_h(self._h, _np_ptr(D32), _np_ptr(D64)))
        return D32, D64

    # ---- device-pointer entry points (torch tensors; enqueue on `stream`)
    def nll_scores(self, M, s_range, m_range, skip_mode, mode, stream):
        _check(lib().vgs_nll_scores(self._h, s_range[0], s_range[1], m_range[0], m_range[1], skip_mode, NLL_MODES[mode],
                                    _dev_ptr(M), M.shape[1], c_vp(stream)))

    def nll_matrix(self, length, mode, D32, D64, stream):
        _check(lib().vgs_nll_matrix(self._h, length, NLL_MODES[mode], _dev_ptr(D32), _dev_ptr(D64), c_vp(stream)))

    def pair_matrix(self, kind, D32, D64, stream):
        _check(lib().vgs_pair_matrix(self._h, PAIR_KINDS[kind], _dev_ptr(D32), _dev_ptr(D64), c_vp(stream)))

    def estimate_matrix(self, D32, D64, stream):
        _check(lib().vgs_estimate_matrix(self._h, _dev_ptr(D32), _dev_ptr(D64), c_vp(stream)))

    def nll_tiles(self, length, mode, tile, rank, world, payload, stream):
        _check(lib().vgs_nll_tiles(self._h, length, NLL_MODES[mode], tile, rank, world, _dev_ptr(payload), c_vp(stream)))

    def pair_tiles(self, kind, tile, rank, world, payload, stream):
        _check(lib().vgs_pair_tiles(self._h, PAIR_KINDS[kind], tile, rank, world, _dev_ptr(payload), c_vp(stream)))

    def estimate_tiles(self, tile, rank, world, payload, stream):
        _check(lib().vgs_estimate_tiles(self._h, tile, rank, world, _dev_ptr(payload), c_vp(stream)))

    def assemble(self, n, tile, world, symmetric, gathered, D32, stream):
        _check(lib().vgs_assemble(self._h, n, tile, world, int(bool(symmetric)), _dev_ptr(gathered), _dev_ptr(D32), c_vp(stream)))

    def average_link(self, D32, k):
        """D32: torch CUDA float32 [n, n].  Returns (labels int32[n], merge distances float64[n - k])."""
        n = int(D32.shape[0])
        labels = np.zeros(n, dtype=np.int32)
        merges = np.zeros(max(n - k, 1), dtype=np.float64)
        m = c_i64()
        _check(lib().vgs_average_link(self._h, n, _dev_ptr(D32), k, _np_ptr(labels), _np_ptr(merges), ctypes.byref(m)))
        return labels, merges[: m.value]

    def retrieval_metrics(self, D, labels):
        """D: torch CUDA [n, n] float32 or float64; labels int32 [n_tax, n].
        Returns (in_top [n_tax, n], dist_to [n_tax, n], avg [n]) float64."""
        import torch
        n = int(D.shape[0])
        labels = np.ascontiguousarray(labels, dtype=np.int32).reshape(-1, n)
        t = labels.shape[0]
        in_top = np.zeros((t, n)); dist_to = np.zeros((t, n)); avg = np.zeros(n)
        d32 = _dev_ptr(D) if D.dtype == torch.float32 else None
        d64 = _dev_ptr(D) if D.dtype == torch.float64 else None
        assert d32 is not None or d64 is not None, "matrix must be float32 or float64"
        _check(lib().vgs_retrieval_metrics(self._h, n, d32, d64, t, _np_ptr(labels), _np_ptr(in_top), _np_ptr(dist_to), _np_ptr(avg)))
        return in_top, dist_to, avg

    def silhouette(self, D32, clusters):
        """D32: torch CUDA float32 [n, n]; clusters int labels [n].  Returns (a, b, s) float64 [n]."""
        n = int(D32.shape[0])
        clusters = np.ascontiguousarray(clusters, dtype=np.int32)
        a = np.zeros(n); b = np.zeros(n); s = np.zeros(n)
        _check(lib().vgs_silhouette(self._h, n, _dev_ptr(D32), _np_ptr(clusters), _np_ptr(a), _np_ptr(b), _np_ptr(s)))
        return a, b, s

    def cluster_distance_std(self, D, clusters):
        """Per-cluster numpy.std of the distances between distinct members (ascending label order)."""
        import torch
        n = int(D.shape[0])
        clusters = np.ascontiguousarray(clusters, dtype=np.int32)
        out = np.zeros(n)
        k = c_i64()
        d32 = _dev_ptr(D) if D.dtype == torch.float32 else None
        d64 = _dev_ptr(D) if D.dtype == torch.float64 else None
        _check(lib().vgs_cluster_distance_std(self._h, n, d32, d64, _np_ptr(clusters), n, _np_ptr(out), ctypes.byref(k)))
        return out[: k.value]

    # ---- model-sharded NLL / peer memory (multi-GPU)
    def nll_scores_band(self, mode, m_col0, ld, dst_ptrs, stream):
        """dst_ptrs: device addresses (ints) of the transposed score matrices to write, this GPU's first."""
        arr = (c_vp * len(dst_ptrs))(*[c_vp(int(p)) for p in dst_ptrs])
        _check(lib().vgs_nll_scores_band(self._h, NLL_MODES[mode], m_col0, ld, arr, len(dst_ptrs), c_vp(stream)))

    def nll_sharded_step(self, mode, m_col0, n_total, dst_ptrs, flag_ptrs, rank, world, epoch, length, D32, stream):
        """band -> (peer barrier) -> combine in one library call; dst_ptrs / flag_ptrs: device addresses (ints)."""
        dsts = (c_vp * len(dst_ptrs))(*[c_vp(int(p)) for p in dst_ptrs])
        flags = (c_vp * len(flag_ptrs))(*[c_vp(int(p)) for p in flag_ptrs]) if flag_ptrs else None
        _check(lib().vgs_nll_sharded_step(self._h, NLL_MODES[mode], m_col0, n_total, dsts, len(dst_ptrs), flags, rank, world,
                                          epoch & 0xFFFFFFFF, length, _dev_ptr(D32), c_vp(stream)))

    def nll_combine(self, n, Mt, ld, length, D32, D64, stream):
        _check(lib().vgs_nll_combine(self._h, n, _dev_ptr(Mt), ld, length, _dev_ptr(D32), _dev_ptr(D64), c_vp(stream)))

    def peer_alloc(self, nbytes):
        """(device address, 64-byte IPC handle) of a zeroed buffer other ranks can map."""
        ptr = c_vp()
        ipc = ctypes.create_string_buffer(64)
        _check(lib().vgs_peer_alloc(self._h, nbytes, ctypes.byref(ptr), ipc))
        return int(ptr.value), bytes(ipc.raw)

    def peer_open(self, ipc):
        ptr = c_vp()
        _check(lib().vgs_peer_open(self._h, ctypes.create_string_buffer(ipc, 64), ctypes.byref(ptr)))
        return int(ptr.value)

    def peer_close(self, ptr):
        _check(lib().vgs_peer_close(self._h, c_vp(ptr)))

    def peer_free(self, ptr):
        _check(lib().vgs_peer_free(self._h, c_vp(ptr)))

    def peer_barrier(self, flag_ptrs, rank, epoch, stream):
        arr = (c_vp * len(flag_ptrs))(*[c_vp(int(p)) for p in flag_ptrs])
        _check(lib().vgs_peer_barrier(self._h, arr, rank, len(flag_ptrs), epoch & 0xFFFFFFFF, c_vp(stream)))

    def tensor_peak_i8(self, cta_group=2):
        """(TOPS, ms) of back-to-back tcgen05.mma kind::i8 from resident shared memory on every SM."""
        tops, ms = ctypes.c_double(), ctypes.c_double()
        _check(lib().vgs_tensor_peak_i8(self._h, cta_group, ctypes.byref(tops), ctypes.byref(ms)))
        return tops.value, ms.value

    def tensor_rate_i8(self, n_cols, cta_group=2):
        """(TOPS, ms) of the same probe at instruction width N = n_cols (multiple of 16, <= 256)."""
        tops, ms = ctypes.c_double(), ctypes.c_double()
        _check(lib().vgs_tensor_rate_i8(self._h, cta_group, int(n_cols), ctypes.byref(tops), ctypes.byref(ms)))
        return tops.value, ms.value

    def matrix_to_triples(self, n, D32, triples, stream):
        _check(lib().vgs_matrix_to_triples(self._h, n, _dev_ptr(D32), _dev_ptr(triples), c_vp(stream)))
