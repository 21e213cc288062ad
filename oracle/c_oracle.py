"""ctypes binding of oracle/_build/libvgs_oracle.so -- TEST INFRASTRUCTURE ONLY (see vgs_oracle.c)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libvgs_oracle.so")

_lib = None


def build(force=False):
    src = os.path.join(HERE, "vgs_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE] + (["-B"] if force else []))
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        vp, i64, i32, u32, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_uint32, ctypes.c_double
        L.vo_create.restype = vp
        L.vo_create.argtypes = [i64, vp, vp, vp, vp]
        L.vo_destroy.argtypes = [vp]
        L.vo_order.argtypes = [vp, i64]
        L.vo_lookup.restype = i32
        L.vo_lookup.argtypes = [vp, i64, u32, ctypes.c_int]
        L.vo_lookup_all.restype = u32
        L.vo_lookup_all.argtypes = [vp, i64, u32, ctypes.c_int]
        L.vo_loglik.restype = dbl
        L.vo_loglik.argtypes = [vp, i64, vp, i64, i64]
        L.vo_score_matrix.argtypes = [vp, vp, i64, i64, i64, vp]
        L.vo_nll_from_scores.argtypes = [vp, i64, i64, vp]
        L.vo_pair_distance.restype = dbl
        L.vo_pair_distance.argtypes = [vp, i64, i64, ctypes.c_int]
        L.vo_pair_matrix.argtypes = [vp, ctypes.c_int, vp]
        L.vo_np_sum.restype = dbl
        L.vo_np_sum.argtypes = [vp, i64]
        L.vo_estimate.restype = dbl
        L.vo_estimate.argtypes = [vp, i64, vp, i64]
        L.vo_estimate_matrix.argtypes = [vp, vp, i64, i64, vp]
        L.vo_mt_uniforms.argtypes = [vp, vp, i64]
        L.vo_sample.restype = ctypes.c_int
        L.vo_sample.argtypes = [vp, i64, vp, i64, i64, vp]
        L.vo_average_link.restype = i64
        L.vo_average_link.argtypes = [vp, i64, i64, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class COracle:
    """The C oracle over one FlatModels set (arrays are kept alive by this object)."""

    MODE = {"intersection": 0, "union": 1, "projection": 2}

    def __init__(self, fm):
        self.off = np.ascontiguousarray(fm.ctx_offsets, dtype=np.int64)
        self.len = np.ascontiguousarray(fm.ctx_len, dtype=np.uint8)
        self.code = np.ascontiguousarray(fm.ctx_code, dtype=np.uint32)
        self.prob = np.ascontiguousarray(fm.prob, dtype=np.float64)
        self.n = fm.n_models
        self.h = lib().vo_create(self.n, _p(self.off), _p(self.len), _p(self.code), _p(self.prob))

    def __del__(self):
        if getattr(self, "h", None):
            lib().vo_destroy(self.h)
            self.h = None

    def order(self, m):
        return lib().vo_order(self.h, m)

    def lookup(self, m, code, length):
        return lib().vo_lookup(self.h, m, int(code), int(length))

    def lookup_all(self, m, code, length):
        return lib().vo_lookup_all(self.h, m, int(code), int(length))

    def score_matrix(self, codes, skip=-1):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        M = np.zeros((codes.shape[0], self.n), dtype=np.float64)
        lib().vo_score_matrix(self.h, _p(codes), codes.shape[0], codes.shape[1], skip, _p(M))
        return M

    @staticmethod
    def nll_from_scores(M, length):
        M = np.ascontiguousarray(M, dtype=np.float64)
        D = np.zeros_like(M)
        lib().vo_nll_from_scores(_p(M), M.shape[0], length, _p(D))
        return D

    def nll_matrix(self, codes):
        return self.nll_from_scores(self.score_matrix(codes), codes.shape[1])

    def pair_matrix(self, mode):
        D = np.zeros((self.n, self.n), dtype=np.float64)
        lib().vo_pair_matrix(self.h, self.MODE[mode], _p(D))
        return D

    def pair_distance(self, i, j, mode):
        return lib().vo_pair_distance(self.h, i, j, self.MODE[mode])

    def estimate(self, left, codes):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        return lib().vo_estimate(self.h, left, _p(codes), len(codes))

    def estimate_matrix(self, codes):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        D = np.zeros((self.n, codes.shape[0]), dtype=np.float64)
        lib().vo_estimate_matrix(self.h, _p(codes), codes.shape[0], codes.shape[1], _p(D))
        return D

    def sample(self, m, state, length, pre=500):
        """state: uint32[625] (random.getstate()[1]); advanced in place."""
        out = np.zeros(length, dtype=np.uint8)
        rc = lib().vo_sample(self.h, m, _p(state), length, pre, _p(out))
        if rc != 0:
            raise RuntimeError("get_context vlmc.pyx")
        return out


def mt_uniforms(state, n):
    out = np.zeros(n, dtype=np.float64)
    lib().vo_mt_uniforms(_p(state), _p(out), n)
    return out


def np_sum(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib().vo_np_sum(_p(a), len(a))


def average_link(D32, k):
    """Exact reference semantics (mixed precision, tie-breaking); see vo_average_link."""
    D32 = np.ascontiguousarray(D32, dtype=np.float32)
    n = D32.shape[0]
    labels = np.zeros(n, dtype=np.int64)
    merges = np.zeros(max(n - k, 1), dtype=np.float64)
    m = lib().vo_average_link(_p(D32), n, k, _p(labels), _p(merges))
    return labels, merges[:m]
