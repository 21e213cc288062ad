"""Developer tool: both Frobenius / Projection paths (CUDA cores / tensor cores) on one set; prints ms per matrix.
usage: python tools/pair_tune.py chains|trees N DEPTH [CTX]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clustering_genomic_signatures_b200 import native, synthetic  # noqa: E402

shape, n, depth = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ctx = int(sys.argv[4]) if len(sys.argv) > 4 else 500
fm, _ = synthetic.make_full_chains(n, 5, depth) if shape == "chains" else synthetic.make_family_models(n, 2, depth, ctx, max(1, n // 50))
h = native.Handle(0)
h.load_models(fm, skip_nll=True)
D = torch.empty((n, n), dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for kind in ("frobenius_intersection", "projection"):
    res = {}
    for path in ("cuda", "tensor"):
        h.set_pair_path(path)
        for _ in range(2):
            h.pair_matrix(kind, D, None, stream)
        torch.cuda.synchronize()
        evs = []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); h.pair_matrix(kind, D, None, stream); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        res[path] = (sum(a.elapsed_time(b) for a, b in evs) / 5, h.last_pair_kernel(), D.double().cpu().numpy().copy())
    diff = np.nanmax(np.abs(res["cuda"][2] - res["tensor"][2]) / np.maximum(np.abs(res["cuda"][2]), 1e-30))
    print("%s n=%d depth=%d contexts/model=%.0f | %-22s cuda-core %.3f ms, tensor-core %.3f ms (kernel %d), max rel diff %.2e" % (
        shape, n, depth, fm.total_contexts / n, kind, res["cuda"][0], res["tensor"][0], res["tensor"][1], diff), flush=True)
