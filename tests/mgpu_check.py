"""Multi-GPU check, launched with torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_check.py

Every rank computes its tiles, the payloads are all-gathered over NCCL and assembled; the result must be
bit-identical on every rank to the single-GPU matrix computed locally by the same rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clustering_genomic_signatures_b200 import native, synthetic  # noqa: E402
from clustering_genomic_signatures_b200.engine import MatrixJob, ShardedNllJob  # noqa: E402
from clustering_genomic_signatures_b200.flat import pack_sequences  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n, L = 300, 2000
    fm, fam = synthetic.make_family_models(n, 9, 6, 200, 6)
    seqs = synthetic.sample_sequences(fm, L, 4)
    h = native.Handle(local)
    h.load_models(fm)
    h.load_sequences_packed(*pack_sequences(seqs))
    stream = torch.cuda.current_stream().cuda_stream
    ok = True
    for kind in ("nll", "nll_kmer", "frobenius_intersection", "frobenius_union", "projection", "estimate"):
        if kind == "nll_kmer":   # the int8 tensor-core contraction: 128-granular tiles
            job = MatrixJob(h, "nll", n, length=L, mode="kmer", tile=128)
        else:
            job = MatrixJob(h, kind, n, length=L)
        D = job.run().clone()
        job.check()
        ref = torch.empty_like(D)
        if kind == "nll":
            h.nll_matrix(L, "sequential", ref, None, stream)
        elif kind == "nll_kmer":
            h.nll_matrix(L, "kmer", ref, None, stream)
        elif kind == "estimate":
            h.estimate_matrix(ref, None, stream)
        else:
            h.pair_matrix(kind, ref, None, stream)
        h.poll_error(stream)
        same = bool(torch.equal(D, ref))
        # and identical across ranks
        chk = D.double().sum().reshape(1)
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        same_ranks = all(float(c) == float(allc[0]) for c in allc)
        ok = ok and same and same_ranks
        if rank == 0:
            print("mgpu %-24s world=%d tile=%d bit-identical to 1 GPU: %s, identical on all ranks: %s" % (kind, world, job.tile, same, same_ranks))
    # model-sharded NLL (each rank holds only its models): peer-memory exchange (GEMM epilogue stores into every rank's
    # matrix over NVLink + flag barrier) and NCCL all-gather exchange, k-mer / sequential / warp kernels, several steps
    packed = pack_sequences(seqs)
    m0, m1 = ShardedNllJob.shard_bounds(n, world)[rank]
    hs = native.Handle(local)
    hs.load_models(fm.subset(range(m0, m1)))
    hs.load_sequences_packed(*packed)
    for exchange in ("peer", "nccl"):
        for mode in ("kmer", "sequential", "warp"):
            try:
                job = ShardedNllJob(hs, n, L, mode=mode, exchange=exchange)
            except Exception as exc:
                if rank == 0:
                    print("mgpu sharded nll %-10s exchange=%s: setup failed: %s" % (mode, exchange, str(exc)[:200]))
                ok = False
                continue
            ref = torch.empty((n, n), dtype=torch.float32, device="cuda")
            h.nll_matrix(L, mode, ref, None, stream)
            h.poll_error(stream)
            same = True
            for step in range(3):
                D = job.run()
                job.check()
                same = same and bool(torch.equal(D, ref))
            flag = torch.tensor([1 if same else 0], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = ok and bool(int(flag.item()))
            if rank == 0:
                print("mgpu sharded nll %-10s world=%d exchange=%s (used %s): bit-identical to 1 GPU on every rank, 3 steps: %s"
                      % (mode, world, exchange, job.exchange, bool(int(flag.item()))))
            job.close()
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
