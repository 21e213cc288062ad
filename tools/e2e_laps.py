"""Developer tool: wall-clock phases of the host-buffer NLL path (vgs_load_models laps via VGS_LOAD_TIMING=1)."""
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
keep, arrs = [], {}
for name, arr in (("off", fm.ctx_offsets.astype(np.int64)), ("len", fm.ctx_len), ("code", fm.ctx_code), ("prob", fm.prob)):
    t, a = pinned(arr); keep.append(t); arrs[name] = a
words, woff, lens = pack_sequences(seqs)
for name, arr in (("words", words), ("woff", woff), ("lens", lens)):
    t, a = pinned(arr); keep.append(t); arrs[name] = a
out_t = torch.empty((1000, 1000), dtype=torch.float32).pin_memory(); out = out_t.numpy()
for kw in ({"nll_only": True}, {"nll_only": True, "device_log": True}, {"nll_only": True, "lazy_tables": True}, {}):
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h.load_models_raw(1000, arrs["off"], arrs["len"], arrs["code"], arrs["prob"], **kw)
        t1 = time.perf_counter()
        h.load_sequences_packed(arrs["words"], arrs["woff"], arrs["lens"])
        t2 = time.perf_counter()
        h.nll_matrix_h(10000, "kmer", want_f64=False, out32=out)
        t3 = time.perf_counter()
        if rep:
            print("%-44s load %.2f  seqs %.2f  matrix_h %.2f  total %.2f ms" % (kw, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3), flush=True)
