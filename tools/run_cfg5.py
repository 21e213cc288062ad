#!/usr/bin/env python
"""BASELINE configs[4]: context-depth / parameter-count sweep (depth 3..12), mirroring
src/plot_metrics_vs_models_varying_nbr_of_parameters.py: full order-D Markov chains (:169-196, parameters =
3 * 4^D) for D <= 7 and pruned VLMCs with a growing number of contexts capped by depth (:40, :126-143).
Frobenius (intersection) all-pairs matrix per point; one JSON line per point."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from clustering_genomic_signatures_b200 import native, synthetic
    h = native.Handle(0)
    stream = torch.cuda.current_stream().cuda_stream
    points = []
    for depth in range(3, 8):
        n = 1000 if depth <= 6 else 400
        points.append(("full_chain", depth, None, n))
    for depth, ctx in ((3, 40), (4, 120), (5, 400), (6, 1300), (7, 4300), (8, 14000), (9, 46000), (10, 150000), (12, 333000)):
        n = int(max(24, min(1000, 12_000_000 // ctx)))
        points.append(("vlmc", depth, ctx, n))
    for arm, depth, ctx, n in points:
        t0 = time.time()
        if arm == "full_chain":
            fm, fam = synthetic.make_full_chains(n, 50 + depth, depth)
        else:
            fm, fam = synthetic.make_family_models(n, 60 + depth, depth, ctx, max(2, n // 20), leaf_jitter=0.02)
        gen = time.time() - t0
        h.load_models(fm, skip_nll=True)
        D32 = torch.empty((n, n), dtype=torch.float32, device="cuda")
        h.pair_matrix("frobenius_intersection", D32, None, stream)
        h.poll_error(stream)
        evs = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); h.pair_matrix("frobenius_intersection", D32, None, stream); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs) / 3
        pairs = n * (n - 1) // 2
        n_ctx = np.diff(fm.ctx_offsets).astype(np.float64)
        alg = 20.0 * (n - 1) * n_ctx.sum() + 4.0 * pairs
        Dh = D32.cpu().numpy()
        same = fam[:, None] == fam[None, :]
        off = ~np.eye(n, dtype=bool)
        print(json.dumps({"arm": arm, "depth": depth, "n_models": n, "mean_contexts": float(n_ctx.mean()),
                          "parameters_per_model": float(3 * n_ctx.mean()), "ms": ms, "pair_distances_per_s": pairs / (ms / 1e3),
                          "algorithmic_GBps": alg / (ms / 1e3) / 1e9, "gen_s": gen,
                          "within_over_between": float(Dh[same & off].mean() / Dh[~same].mean()),
                          "symmetric": bool(np.array_equal(Dh, Dh.T))}), flush=True)


if __name__ == "__main__":
    main()
