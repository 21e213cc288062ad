#!/usr/bin/env python
"""BASELINE configs[3]: N virus-genome-scale VLMCs (depth <= 9, ~2000 contexts), full N x N matrix
tile-sharded over the GPUs of one box + average-link clustering to k clusters.

    python tools/run_cfg4.py --n-models 4000                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_cfg4.py

Prints one JSON object (rank 0) with the timings of every stage, pair-distances/s and the SURVEY §8d
roofline fractions; used to fill profiles/r1_cfg4_*.json.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-models", type=int, default=20000)
    ap.add_argument("--depth", type=int, default=9)
    ap.add_argument("--ctx", type=int, default=2000)
    ap.add_argument("--seq-len", type=int, default=1000)
    ap.add_argument("--k", type=int, default=200)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--skip-nll", action="store_true")
    ap.add_argument("--host-sequences", action="store_true",
                    help="sample the sequences with the numpy sampler on the host (slow at N = 20000) instead of vgs_sample_sequences")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from clustering_genomic_signatures_b200 import native, synthetic
    from clustering_genomic_signatures_b200.engine import MatrixJob
    from clustering_genomic_signatures_b200.flat import pack_sequences

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, L = args.n_models, args.seq_len
    out = {"config": {"n_models": n, "depth": args.depth, "target_ctx": args.ctx, "seq_len": L, "k": args.k, "n_gpus": world}}

    t0 = time.time()
    fm, fam = synthetic.make_family_models(n, 4, args.depth, args.ctx, max(1, n // 100))
    out["gen_models_s"] = time.time() - t0
    out["config"]["total_contexts"] = fm.total_contexts
    seqs = None
    if not args.skip_nll and args.host_sequences:
        t0 = time.time()
        seqs = synthetic.sample_sequences(fm, L, 5)
        out["gen_sequences_s"] = time.time() - t0

    h = native.Handle(local)
    t0 = time.time()
    h.load_models(fm)
    torch.cuda.synchronize()
    out["load_models_s"] = time.time() - t0
    if seqs is not None:
        h.load_sequences_packed(*pack_sequences(seqs))
    elif not args.skip_nll:
        # the device sampler (row f2): every rank draws the same MT19937 stream, so the sequence sets are identical
        import random
        t0 = time.time()
        st = np.array(random.Random(5).getstate()[1], dtype=np.uint32)
        h.sample_sequences(L, 500, st)
        torch.cuda.synchronize()
        out["sample_sequences_on_device_s"] = time.time() - t0
        seqs = True
    free_after = h.device_info()["free_bytes"]
    out["hbm_free_after_load_gb"] = free_after / 1e9

    peak = 6536.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    n_ctx = np.diff(fm.ctx_offsets).astype(np.float64)
    orders = fm.orders().astype(np.float64)
    pairs = n * (n - 1) // 2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def run(kind, mode="sequential"):
        job = MatrixJob(h, kind, n, length=L, mode=mode)
        job.run(); job.check()  # warm-up (builds lazily built tables)
        barrier()
        evs = []
        for _ in range(args.steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); job.run(); b.record()
            evs.append((a, b))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        job.check()
        return job, ms

    # ---- Frobenius (intersection): the primary full-matrix distance of this config
    job, ms = run("frobenius_intersection")
    alg = 20.0 * (n - 1) * n_ctx.sum() + 4.0 * pairs
    out["frobenius"] = {"ms": ms, "pair_distances_per_s": pairs / (ms / 1e3), "tile": job.tile,
                        "algorithmic_GBps_per_gpu": alg / world / (ms / 1e3) / 1e9,
                        "roofline_frac_of_measured_hbm": alg / world / (ms / 1e3) / 1e9 / peak}
    D32 = job.D32
    # ---- NLL
    if seqs is not None:
        T_m = 32.0 * n_ctx + 2.0 * 4.0 ** orders
        algn = n * (np.ceil(L / 4.0) + np.minimum(T_m, L * 10.0) + 8.0).sum()
        for mode in ("sequential", "warp"):
            jobn, msn = run("nll", mode)
            out["nll" if mode == "sequential" else "nll_warp"] = {
                "ms": msn, "pair_distances_per_s": pairs / (msn / 1e3), "tile": jobn.tile, "mode": mode,
                "algorithmic_GBps_per_gpu": algn / world / (msn / 1e3) / 1e9,
                "roofline_frac_of_measured_hbm": algn / world / (msn / 1e3) / 1e9 / peak}
    # ---- average link on the Frobenius matrix (rank 0; the matrix is resident on every rank)
    if rank == 0:
        torch.cuda.synchronize()
        t0 = time.time()
        labels, merges = h.average_link(D32, args.k)
        out["average_link"] = {"s": time.time() - t0, "merges": int(len(merges)), "k": args.k}
        # how well do the clusters recover the generating families?
        from collections import Counter
        purity = sum(Counter(fam[labels == c].tolist()).most_common(1)[0][1] for c in range(int(labels.max()) + 1)) / n
        out["average_link"]["family_purity"] = purity
        Dh = D32[:64, :64].cpu().numpy()
        out["matrix_checks"] = {"symmetric": bool(np.array_equal(Dh, Dh.T)), "zero_diagonal": bool(np.all(np.diag(Dh) == 0))}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
