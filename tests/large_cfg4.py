#!/usr/bin/env python
"""BASELINE configs[3]: N = 20 000 virus-genome-scale VLMCs (depth <= 9, ~2000 contexts): the full N x N matrix over the GPUs of
one box (Frobenius: tables replicated, tile space sharded; NLL: models sharded, sequences replicated) + average-link clustering
to k = 200.  Test infrastructure (lives in tests/: it checks against the oracle).

    # timing on all GPUs of the box; records the exact bit checksum of every matrix
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/large_cfg4.py > profiles/r2_cfg4_n20000_8gpu.json
    # verification on one GPU: the same matrices (same checksums), >= 10^4 entries of each against the C oracle, the
    # 20 000 average-link labels and merge distances against the C oracle's average link
    python tests/large_cfg4.py --verify > profiles/r2_cfg4_n20000_verify_1gpu.json
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-models", type=int, default=20000)
    ap.add_argument("--depth", type=int, default=9)
    ap.add_argument("--ctx", type=int, default=2000)
    ap.add_argument("--seq-len", type=int, default=1000)
    ap.add_argument("--k", type=int, default=200)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--verify", action="store_true", help="compare with the C oracle (one GPU)")
    ap.add_argument("--sample-models", type=int, default=104, help="--verify: models of the sampled sub-matrices (104^2 > 10^4 entries)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from clustering_genomic_signatures_b200 import native, synthetic
    from clustering_genomic_signatures_b200.engine import MatrixJob, ShardedNllJob
    from clustering_genomic_signatures_b200.flat import FlatModels, pack_sequences

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, L = args.n_models, args.seq_len
    out = {"config": {"n_models": n, "depth": args.depth, "target_ctx": args.ctx, "seq_len": L, "k": args.k, "n_gpus": world}}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def bit_checksum(t):
        return int(t.contiguous().view(torch.int32).to(torch.int64).sum().item())

    # ---- inputs: rank 0 generates, the others read them from shared memory
    shm = "/dev/shm/vgs_cfg4_%d_%d_%d.npz" % (n, args.depth, args.ctx)
    t0 = time.time()
    if rank == 0:
        fm, fam = synthetic.make_family_models(n, 4, args.depth, args.ctx, max(1, n // 100))
        if world > 1:
            np.savez(shm, off=fm.ctx_offsets, ln=fm.ctx_len, code=fm.ctx_code, prob=fm.prob, fam=fam)
    barrier()
    if rank != 0:
        z = np.load(shm)
        fm = FlatModels(z["off"], z["ln"], z["code"], z["prob"], ["m%d" % i for i in range(n)])
        fam = z["fam"]
    barrier()
    if rank == 0 and world > 1:
        os.unlink(shm)
    out["gen_models_s"] = round(time.time() - t0, 1)
    out["config"]["total_contexts"] = fm.total_contexts
    pairs = n * (n - 1) // 2

    def timed(job):
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
        return ms

    # ---- Frobenius (intersection): tables replicated, tile space sharded
    hF = native.Handle(local)
    t0 = time.time()
    hF.load_models(fm, skip_nll=True)
    out["load_models_frobenius_s"] = round(time.time() - t0, 2)
    jobF = MatrixJob(hF, "frobenius_intersection", n)
    ms = timed(jobF)
    out["frobenius"] = {"ms": ms, "pair_distances_per_s": pairs / (ms / 1e3), "tile": jobF.tile, "checksum": bit_checksum(jobF.D32),
                        "kernel": "tensor" if hF.last_pair_kernel() else "cuda-core (2000 of 349 525 possible contexts: sparse)"}
    # sequences: the device sampler (row f2) on the full set; every rank draws the same MT19937 stream
    t0 = time.time()
    st = np.array(random.Random(5).getstate()[1], dtype=np.uint32)
    hF.sample_sequences(L, 500, st)
    seqs = np.stack([hF.get_sequence(i, L) for i in range(n)])
    out["sample_sequences_s"] = round(time.time() - t0, 2)
    packed = pack_sequences(seqs)
    # ---- NLL: models sharded, sequences replicated
    m0, m1 = ShardedNllJob.shard_bounds(n, world)[rank]
    hN = native.Handle(local)
    t0 = time.time()
    hN.load_models(fm.subset(range(m0, m1)) if world > 1 else fm, nll_only=True)
    hN.load_sequences_packed(*packed)
    out["load_models_nll_shard_s"] = round(time.time() - t0, 2)
    for mode in ("warp", "sequential"):
        jobN = ShardedNllJob(hN, n, L, mode=mode, rank=rank, world=world)
        ms = timed(jobN)
        out["nll_" + mode] = {"ms": ms, "pair_distances_per_s": pairs / (ms / 1e3), "exchange": jobN.exchange,
                              "checksum": bit_checksum(jobN.D32)}
        if mode == "sequential":
            Dn = jobN.D32.clone()
        jobN.close()
    D32 = jobF.D32
    # ---- average link on the Frobenius matrix (rank 0; the matrix is resident on every rank)
    if rank == 0:
        torch.cuda.synchronize()
        t0 = time.time()
        labels, merges = hF.average_link(D32, args.k)
        out["average_link"] = {"s": round(time.time() - t0, 3), "merges": int(len(merges)), "k": args.k,
                               "labels_checksum": int(np.sum(labels.astype(np.int64) * (np.arange(n) % 1009 + 1)))}
        from collections import Counter
        out["average_link"]["family_purity"] = sum(Counter(fam[labels == c].tolist()).most_common(1)[0][1]
                                                   for c in range(int(labels.max()) + 1)) / n
        if args.verify:
            import c_oracle
            rng = np.random.default_rng(11)
            S = np.sort(rng.choice(n, size=args.sample_models, replace=False))
            sub = fm.subset(S.tolist())
            co = c_oracle.COracle(sub)
            idx = torch.from_numpy(S).cuda()
            got_f = D32[idx][:, idx].double().cpu().numpy()
            want_f = co.pair_matrix("intersection")
            w32 = want_f.astype(np.float32).astype(np.float64)
            got_n = Dn[idx][:, idx].cpu().numpy()
            want_n = co.nll_matrix(seqs[S]).astype(np.float32)
            out["verify"] = {
                "sampled_entries": int(got_f.size),
                "frobenius_max_rel_err_vs_c_oracle_f32": float(np.max(np.abs(got_f - w32) / np.maximum(np.abs(w32), 1e-30))),
                "frobenius_entries_equal_to_rounded_oracle": int((got_f == w32).sum()),
                "nll_sequential_bit_equal_to_c_oracle": bool(np.array_equal(got_n, want_n)),
            }
            t0 = time.time()
            Dh = D32.cpu().numpy()
            want_labels, want_merges = c_oracle.average_link(Dh, args.k)
            out["verify"]["c_oracle_average_link_s"] = round(time.time() - t0, 1)
            out["verify"]["labels_identical_to_c_oracle"] = bool(np.array_equal(labels, want_labels))
            out["verify"]["merge_distances_identical_to_c_oracle"] = bool(np.array_equal(merges, want_merges))
            out["verify"]["matrix_symmetric"] = bool(np.array_equal(Dh, Dh.T))
            out["verify"]["matrix_zero_diagonal"] = bool(np.all(np.diag(Dh) == 0))
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
