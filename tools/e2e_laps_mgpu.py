"""Developer tool: per-rank wall-clock phases of the multi-GPU end-to-end NLL step (torchrun, one process per GPU)."""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from clustering_genomic_signatures_b200 import native
from clustering_genomic_signatures_b200.engine import ShardedNllJob
from clustering_genomic_signatures_b200.flat import pack_sequences

world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n, L = 1000, 10000
fm_full, fam, seqs = bench.make_inputs("nll", n, L, 6, 500)
m0, m1 = ShardedNllJob.shard_bounds(n, world)[rank]
fm = fm_full.subset(range(m0, m1))
h = native.Handle(lr)
def pinned(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t, t.numpy()
keep, A = [], {}
for name, arr in (("off", fm.ctx_offsets.astype(np.int64)), ("len", fm.ctx_len), ("code", fm.ctx_code), ("prob", fm.prob)):
    t, a = pinned(arr); keep.append(t); A[name] = a
words, woff, lens = pack_sequences(seqs)
for name, arr in (("words", words), ("woff", woff), ("lens", lens)):
    t, a = pinned(arr); keep.append(t); A[name] = a
out_t = torch.empty((n, n), dtype=torch.float32).pin_memory()
def load():
    h.load_models_raw(fm.n_models, A["off"], A["len"], A["code"], A["prob"], nll_only=True, device_log=True)
    h.load_sequences_packed(A["words"], A["woff"], A["lens"])
load()
job = ShardedNllJob(h, n, L, mode="kmer")
for variant in ("all ranks read the matrix", "rank 0 alone reads the matrix", "no barrier before the step"):
    acc = np.zeros(6)
    reps = 12
    for it in range(reps + 2):
        torch.cuda.synchronize()
        if variant != "no barrier before the step":
            dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        h.load_models_raw(fm.n_models, A["off"], A["len"], A["code"], A["prob"], nll_only=True, device_log=True)
        t1 = time.perf_counter()
        h.load_sequences_packed(A["words"], A["woff"], A["lens"])
        t2 = time.perf_counter()
        e0.record(); job.run(); e1.record()
        t3 = time.perf_counter()
        if variant != "rank 0 alone reads the matrix" or rank == 0:
            out_t.copy_(job.D32, non_blocking=True)
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        if it >= 2:
            acc += np.array([t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0, e0.elapsed_time(e1) * 1e-3]) * 1e3
    acc /= reps
    t = torch.tensor(acc, dtype=torch.float64, device=dev)
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
    if rank == 0:
        M = torch.stack(allr).cpu().numpy()
        print("== %s (world %d, exchange %s, cpus %d): ms per rank, [load_models, load_sequences, enqueue step, D2H + sync, total, step on device]" % (variant, world, job.exchange, len(os.sched_getaffinity(0))))
        for r in range(world):
            print("   rank %d: " % r + "  ".join("%.3f" % v for v in M[r]))
        print("   max:    " + "  ".join("%.3f" % v for v in M.max(axis=0)), flush=True)
job.close()
h.close()
dist.destroy_process_group()
