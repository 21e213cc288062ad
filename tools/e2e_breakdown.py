"""Where does the host-buffer (e2e) path spend its time?  Developer tool."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from clustering_genomic_signatures_b200 import native
from clustering_genomic_signatures_b200.flat import pack_sequences

fm, fam, seqs = bench.make_inputs("nll", 1000, 10000, 6, 500)
h = native.Handle(0)
def pinned(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t, t.numpy()
keep = []
arrs = {}
for name, arr in (("off", fm.ctx_offsets.astype(np.int64)), ("len", fm.ctx_len), ("code", fm.ctx_code), ("prob", fm.prob)):
    t, a = pinned(arr); keep.append(t); arrs[name] = a
words, woff, lens = pack_sequences(seqs)
for name, arr in (("words", words), ("woff", woff), ("lens", lens)):
    t, a = pinned(arr); keep.append(t); arrs[name] = a
out_t = torch.empty((1000, 1000), dtype=torch.float32).pin_memory(); out = out_t.numpy()
def T(f, n=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print("load_models            %.2f ms" % T(lambda: h.load_models_raw(1000, arrs["off"], arrs["len"], arrs["code"], arrs["prob"])))
print("load_sequences_packed  %.2f ms" % T(lambda: h.load_sequences_packed(arrs["words"], arrs["woff"], arrs["lens"])))
print("nll_matrix_h (no f64)  %.2f ms" % T(lambda: h.nll_matrix_h(10000, "sequential", want_f64=False, out32=out)))
print("pair_matrix_h          %.2f ms" % T(lambda: h.pair_matrix_h("frobenius_intersection", want_f64=False, out32=out)))
print("nll_matrix_h kmer (no f64) %.2f ms" % T(lambda: h.nll_matrix_h(10000, "kmer", want_f64=False, out32=out)))
import os as _os
print("host cores:", _os.cpu_count())
print("load_models skip_nll   %.2f ms" % T(lambda: h.load_models_raw(1000, arrs["off"], arrs["len"], arrs["code"], arrs["prob"], skip_nll=True)))
import ctypes
a = torch.empty(16 << 20, dtype=torch.uint8).pin_memory(); d = torch.empty(16 << 20, dtype=torch.uint8, device="cuda")
print("H2D 16 MiB pinned      %.2f ms" % T(lambda: d.copy_(a, non_blocking=True)))
