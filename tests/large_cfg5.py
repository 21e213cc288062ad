#!/usr/bin/env python
"""BASELINE configs[4]: context-depth / parameter-count sweep (depth 3..12), mirroring
src/plot_metrics_vs_models_varying_nbr_of_parameters.py: full order-D Markov chains (:169-196, parameters = 3 * 4^D) for
D <= 7 and pruned VLMCs with a growing number of contexts capped by depth (:40, :126-143).

Per point, one JSON line: the Frobenius (intersection) all-pairs matrix on the CUDA-core kernels and -- where it applies --
on the tensor cores, which one the dispatcher picks, the NLL matrix (L = 1000; k-mer contraction up to depth 6, scan
beyond), and parity of a corner of each matrix against the C oracle (test infrastructure: this driver lives in tests/).

    python tests/large_cfg5.py > profiles/r2_cfg5_sweep_1gpu.jsonl
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def timed(fn, reps=3):
    import torch
    fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / reps


def main():
    import torch
    import c_oracle
    from clustering_genomic_signatures_b200 import native, synthetic
    from clustering_genomic_signatures_b200.flat import pack_sequences
    h = native.Handle(0)
    stream = torch.cuda.current_stream().cuda_stream
    L = 1000
    points = []
    for depth in range(3, 8):
        points.append(("full_chain", depth, None, 1000 if depth <= 6 else 400))
    for depth, ctx in ((3, 40), (4, 120), (5, 400), (6, 1300), (7, 4300), (8, 14000), (9, 46000), (10, 150000), (12, 333000)):
        points.append(("vlmc", depth, ctx, int(max(24, min(1000, 12_000_000 // ctx)))))
    for arm, depth, ctx, n in points:
        t0 = time.time()
        if arm == "full_chain":
            fm, fam = synthetic.make_full_chains(n, 50 + depth, depth)
        else:
            fm, fam = synthetic.make_family_models(n, 60 + depth, depth, ctx, max(2, n // 20), leaf_jitter=0.02)
        seqs = synthetic.sample_sequences(fm, L, 9)
        gen = time.time() - t0
        h.load_models(fm)
        h.load_sequences_packed(*pack_sequences(seqs))
        D32 = torch.empty((n, n), dtype=torch.float32, device="cuda")
        D64 = torch.empty((n, n), dtype=torch.float64, device="cuda")
        pairs = n * (n - 1) // 2
        n_ctx = np.diff(fm.ctx_offsets).astype(np.float64)
        k = min(n, 16)
        co = c_oracle.COracle(fm.subset(range(k)))
        want_f = co.pair_matrix("intersection")
        rec = {"arm": arm, "depth": depth, "n_models": n, "mean_contexts": float(n_ctx.mean()), "parameters_per_model": float(3 * n_ctx.mean()),
               "universe_fill": float(n_ctx.mean() / ((4 ** (depth + 1) - 1) // 3)), "pairs": pairs, "gen_s": round(gen, 1)}
        # Frobenius, both paths
        for path in ("cuda", "tensor"):
            h.set_pair_path(path)
            ms = timed(lambda: h.pair_matrix("frobenius_intersection", D32, D64, stream))
            h.poll_error(stream)
            used = h.last_pair_kernel()
            if path == "tensor" and used == 0:
                rec["frobenius_tensor"] = None   # deeper than 7: no dense universe
                continue
            got = D64.cpu().numpy()
            err = np.abs(got[:k, :k] - want_f)
            ok = bool(np.all(err <= (1.3e-7 * np.abs(want_f) + 1e-9 if used else 1e-9 * np.abs(want_f))))
            rec["frobenius_" + path] = {"ms": ms, "pair_distances_per_s": pairs / (ms / 1e3), "corner_within_bound_of_c_oracle": ok,
                                        "symmetric": bool(np.array_equal(got, got.T)), "zero_diagonal": bool(np.all(np.diag(got) == 0))}
        h.set_pair_path("auto")
        h.pair_matrix("frobenius_intersection", D32, None, stream)
        h.poll_error(stream)
        rec["dispatcher_picks"] = "tensor" if h.last_pair_kernel() else "cuda"
        Dh = D32.cpu().numpy()
        same = fam[:, None] == fam[None, :]
        off = ~np.eye(n, dtype=bool)
        rec["within_over_between_family"] = float(Dh[same & off].mean() / Dh[~same].mean())
        # NLL
        ms = timed(lambda: h.nll_matrix(L, "kmer", D32, D64, stream))
        h.poll_error(stream)
        got = D64.cpu().numpy()
        want_n = co.nll_from_scores(c_oracle.COracle(fm.subset(range(k))).score_matrix(seqs[:k]), L)
        offk = ~np.eye(k, dtype=bool)
        rel = float(np.max(np.abs(got[:k, :k][offk] - want_n[offk]) / np.abs(want_n[offk]))) if k > 1 else 0.0
        rec["nll"] = {"ms": ms, "pair_distances_per_s": pairs / (ms / 1e3), "kernel": {0: "scan", 1: "dmma", 2: "fma", 3: "int8 tensor-core contraction"}[h.last_nll_kernel()],
                      "corner_max_rel_err_vs_c_oracle": rel, "within_1e-9": bool(rel < 1e-9), "symmetric": bool(np.array_equal(got, got.T)),
                      "zero_diagonal": bool(np.all(np.diag(got) == 0))}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
