"""Developer tool: a few device-log loads of the cfg-3 set (target for ncu captures of the load's kernels)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from clustering_genomic_signatures_b200 import native
fm, fam, seqs = bench.make_inputs("nll", 1000, 10000, 6, 500)
h = native.Handle(0)
def pinned(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t, t.numpy()
keep, arrs = [], {}
for name, arr in (("off", fm.ctx_offsets.astype(np.int64)), ("len", fm.ctx_len), ("code", fm.ctx_code), ("prob", fm.prob)):
    t, a = pinned(arr); keep.append(t); arrs[name] = a
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    h.load_models_raw(1000, arrs["off"], arrs["len"], arrs["code"], arrs["prob"], nll_only=True, device_log=True)
torch.cuda.synchronize()
print("ok")
