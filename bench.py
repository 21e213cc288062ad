#!/usr/bin/env python
"""bench.py -- VLMC pair-distances/sec on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload nll|frobenius|frobenius_union|projection|estimate]
    python bench.py --impl reference ...        # the reference's own CPU implementation (oracle/_ref)

Default workload: BASELINE configs[2] -- 1,000 synthetic VLMCs (depth <= 6, ~500 contexts, SURVEY.md
§8d recipe, seed 2) x 10 kbp sequences sampled from the models, negative-log-likelihood distance
(10^6 pair scorings, 499,500 unordered pair-distances), the configuration the metric "pair-distances/
sec at 1/2/4/8 B200" is quoted on.  ``--workload frobenius`` is BASELINE configs[1].

A step = one full N x N float32 distance matrix:
  value  tables and packed sequences already resident in HBM -> matrix resident on every GPU
         (N > 1, NLL: models sharded over ranks, score bands stored into every rank's matrix by the GEMM epilogue over NVLink
         peer memory + flag barrier -- or NCCL all-gather --, local combine; other kinds: upper-triangular tile space sharded
         over ranks, NCCL all-gather, local assemble);
  e2e    the same matrix through the C-ABI with HOST buffers: H2D of the flat model tables and packed
         sequences (pinned), ln p (k-mer mode: on the device, VGS_LOAD_DEVICE_LOG; scan modes: libm on the host, streamed),
         table construction, kernels, D2H of the float32 matrix (pinned).  N > 1: every rank holds the whole matrix in HBM
         after the exchange; it crosses to the host ONCE, on rank 0 (the process that hands it to the clustering code, as in
         the reference's single-process pipeline) -- eight ranks each pulling all 4 MB at the same moment measured the host's
         aggregate PCIe ingest (0.33-0.58 ms per rank at 8 GPUs, tools/e2e_laps_mgpu.py), not the path.
Timing: CUDA events on the launching stream, one pair per step, L2 flushed (256 MiB memset) between
steps outside the timed pairs, barrier + synchronize on both sides, max over ranks.
Extra keys: `matrix_checksum` (exact int64 sum of the float32 bit patterns: equal across GPU counts iff the matrices are),
`parity_check` (N > 1: this rank's matrix == the single-GPU matrix of the unsharded set, min over ranks), `scale_probe` (a cfg-4
slice, N = 8000 depth 9: Frobenius with sharded tiles and NLL with sharded models at this GPU count; `--probe-models 0` skips).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, n_models, seq_len, depth, ctx, description)
    "nll": ("nll", 1000, 10000, 6, 500, "cfg3: 1000 VLMCs (depth<=6, ~500 ctx) x 10 kbp, NLL distance"),
    "frobenius": ("frobenius_intersection", 1000, 0, 6, 500, "cfg2: 1000 VLMCs (depth<=6, ~500 ctx), Frobenius intersection"),
    "frobenius_union": ("frobenius_union", 1000, 0, 6, 500, "cfg2: 1000 VLMCs, Frobenius union"),
    "projection": ("projection", 1000, 0, 6, 500, "cfg2: 1000 VLMCs, Projection"),
    "estimate": ("estimate", 1000, 100000, 6, 500, "1000 VLMCs x 100 kbp, EstimateVLMC"),
    "nll_deep": ("nll", 2000, 1000, 9, 2000, "cfg4 slice: 2000 VLMCs (depth<=9, ~2000 ctx) x 1 kbp, NLL"),
    "frobenius_deep": ("frobenius_intersection", 2000, 0, 9, 2000, "cfg4 slice: 2000 VLMCs (depth<=9, ~2000 ctx), Frobenius"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="nll", choices=sorted(WORKLOADS))
    ap.add_argument("--n-models", type=int, default=0)
    ap.add_argument("--seq-len", type=int, default=0)
    ap.add_argument("--mode", default="kmer", choices=["sequential", "warp", "kmer"],
                    help="NLL kernel: kmer = k-mer histogram + exact int8 tcgen05 contraction (ln p in 2^-43 fixed point, "
                         "<= 1e-9 rel.), sequential = bit-exact scan, warp = warp-per-pair scan")
    ap.add_argument("--no-mode-sweep", action="store_true", help="skip the short timings of the other NLL modes")
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="models in the CPU-baseline sample")
    ap.add_argument("--probe-models", type=int, default=8000,
                    help="scale_probe: models of the cfg-4 slice (depth 9, ~2000 contexts) timed at this GPU count; 0 = off")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1, NLL: how the score bands travel (peer = GEMM epilogue stores into peer-mapped memory)")
    ap.add_argument("--breakdown", action="store_true",
                    help="N > 1: also time the phases of a step (tiles / all-gather / assemble) with CUDA events; stderr")
    return ap.parse_args()


def make_inputs(kind, n, L, depth, ctx):
    from clustering_genomic_signatures_b200 import synthetic
    fm, fam = synthetic.make_family_models(n, 2 if depth == 6 else 4, depth, ctx, max(1, n // 50))
    seqs = None
    if kind in ("nll", "estimate"):
        seqs = synthetic.sample_sequences(fm, L, 3)
    return fm, fam, seqs


# ---------------------------------------------------------------------------------------------
def algorithmic_bytes(kind, fm, L, orders):
    """SURVEY.md §8(d) per-unit figures evaluated on the actual set.  Returns (bytes per step of the
    dominant kernel, description)."""
    n = fm.n_models
    n_ctx = np.diff(fm.ctx_offsets).astype(np.float64)
    if kind == "nll":
        T_m = 32.0 * n_ctx + 2.0 * 4.0 ** orders
        per_scoring = np.ceil(L / 4.0) + np.minimum(T_m, L * 10.0) + 8.0
        return float(n * per_scoring.sum()), "N scorings per model x (ceil(L/4) + min(32 n_ctx + 2*4^order, 10 L) + 8) B"
    if kind == "estimate":
        per = np.ceil(L / 4.0) + n_ctx * 36.0 + 8.0
        return float(n * per.sum()), "N^2 x (ceil(L/4) + 36 n_ctx + 8) B"
    # pair kinds: unordered pairs, (n_i + n_j) * 20 + 4
    s = n_ctx.sum()
    return float(20.0 * (n - 1) * s + 4.0 * n * (n - 1) / 2), "P x ((n_ctx_i + n_ctx_j) x 20 + 4) B"


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7 or not (t0 - 0.05 <= t <= t1 + 0.1):
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # timed region shorter than the sampling period: use every sample we have
            for t, line in self.rows:
                parts = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(parts[0])); mx.append(float(parts[1]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def cpu_baseline(kind, fm, seqs, L, sample_n, threads=1):
    """Bounded CPU sample of the same workload.  Returns the cpu_baseline dict.
    Uses the compiled reference (oracle/_ref) when present, else the C oracle port."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    from clustering_genomic_signatures_b200 import synthetic
    n = min(sample_n, fm.n_models)
    if ref_loader.available():
        ns = ref_loader.load()
        vl = [ns.VLMC(fm.tree(m), fm.names[m], {}) for m in range(n)]
        if kind == "nll":
            d = ns.NegativeLogLikelihood(L)
            for v, s in zip(vl, synthetic.codes_to_strings(seqs[:n])):
                v.sequence = s
        elif kind == "estimate":
            d = ns.EstimateVLMC()
            for v, s in zip(vl, synthetic.codes_to_strings(seqs[:n])):
                v.sequence = s
        elif kind == "projection":
            d = ns.Projection()
            d.set_vlmcs(vl)
        else:
            d = ns.FrobeniusNorm(kind == "frobenius_union")
        t0 = time.perf_counter()
        if kind == "estimate":
            with ref_loader.patched_power_divergence():
                ns.calculate_distances_within_vlmcs(vl, d)
        else:
            ns.calculate_distances_within_vlmcs(vl, d)
        dt = time.perf_counter() - t0
        return {"value": n * n / dt, "unit": "pair-distances/s", "cores": 1, "kind": "reference",
                "sample": "compiled reference (oracle/_ref), calculate_distances_within_vlmcs on the first %d models "
                          "= %d distance() calls in %.2f s, single thread" % (n, n * n, dt)}
    import c_oracle
    co = c_oracle.COracle(fm.subset(range(n)))
    t0 = time.perf_counter()
    if kind == "nll":
        co.nll_matrix(seqs[:n])
    elif kind == "estimate":
        co.estimate_matrix(seqs[:n])
    else:
        co.pair_matrix({"frobenius_intersection": "intersection", "frobenius_union": "union", "projection": "projection"}[kind])
    dt = time.perf_counter() - t0
    return {"value": n * n / dt, "unit": "pair-distances/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "C oracle port (oracle/vgs_oracle.c, OpenMP) on the first %d models = %d distances in %.2f s" % (n, n * n, dt)}


def _ref_worker(args):
    kind, trees, names, seq_strs, L, rows = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    ns = ref_loader.load()
    vl = [ns.VLMC(t, nm, {}) for t, nm in zip(trees, names)]
    if kind in ("nll", "estimate"):
        for v, s in zip(vl, seq_strs):
            v.sequence = s
    if kind == "nll":
        d = ns.NegativeLogLikelihood(L)
    elif kind == "estimate":
        d = ns.EstimateVLMC()
    elif kind == "projection":
        d = ns.Projection(); d.set_vlmcs(vl)
    else:
        d = ns.FrobeniusNorm(kind == "frobenius_union")
    out = 0
    ctx = ref_loader.patched_power_divergence() if kind == "estimate" else None
    if ctx:
        ctx.__enter__()
    for i in rows:
        for j in range(len(vl)):
            d.distance(vl[i], vl[j])
            out += 1
    if ctx:
        ctx.__exit__(None, None, None)
    return out


def run_reference(args, kind, n, L, depth, ctx, desc):
    """--impl reference: the reference's own CPU implementation with every host core (row-sharded
    multiprocessing, the only parallelism the reference itself uses), on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import multiprocessing as mp
    import ref_loader
    from clustering_genomic_signatures_b200 import synthetic
    cores = os.cpu_count() or 1
    sample_n = args.cpu_sample or {"nll": max(24, min(96, 2 * cores)), "estimate": max(6, min(32, cores)),
                                   "frobenius_union": 400, "projection": 400}.get(kind, 600)
    sample_n = min(sample_n, n)
    fm, fam, seqs = make_inputs(kind, n, L, depth, ctx)  # the same set as our arm; the sample is its first models
    fm = fm.subset(range(sample_n))
    seq_strs = synthetic.codes_to_strings(seqs[:sample_n]) if seqs is not None else None
    trees = [fm.tree(m) for m in range(sample_n)]
    if not ref_loader.available():
        cb = cpu_baseline(kind, fm, seqs, L, sample_n)
        ms = 1000.0 * sample_n * sample_n / cb["value"]
        line = {"impl": "reference", "metric": "pair-distances/sec", "value": cb["value"], "unit": "pair-distances/s",
                "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": desc, "sample_models": sample_n},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "pair-distances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return
    ref_loader.load()   # in the parent too: the forked workers inherit it, and the process the driver watches shows oracle/_ref
    workers = min(cores, sample_n)
    shards = [list(range(w, sample_n, workers)) for w in range(workers)]
    jobs = [(kind, trees, fm.names, seq_strs, L, rows) for rows in shards]
    times = []
    with mp.get_context("fork").Pool(workers) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            total = sum(pool.map(_ref_worker, jobs))
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
            if sum(times) > 150:  # keep the whole run within a few minutes
                break
    steps = len(times)
    ms = 1000.0 * sum(times) / steps
    value = sample_n * sample_n / (ms / 1000.0)
    cb = {"value": value, "unit": "pair-distances/s", "cores": workers, "kind": "reference",
          "sample": "compiled reference (oracle/_ref): all %d x %d distance() calls of the first %d models per step, "
                    "row-sharded over %d processes" % (sample_n, sample_n, sample_n, workers)}
    line = {"impl": "reference", "metric": "pair-distances/sec", "value": value, "unit": "pair-distances/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64" if kind in ("nll", "estimate") else "f32", "data": "synthetic",
            "config": {"workload": desc, "sample_models": sample_n, "seq_len": L},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "pair-distances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def scale_probe(args, world, rank, local_rank, dev, bit_checksum):
    """BASELINE cfg 4 in small: N models of depth 9 with ~2000 contexts; the full Frobenius (intersection) matrix with the
    tables replicated and the tile space sharded, and the full NLL matrix (L = 1000, scan kernel: the sets are deeper
    than the k-mer contraction reaches) with the models sharded.  Device time per step, max over ranks; the driver's
    1/2/4/8 lines give the scaling, the checksums must agree across them."""
    import torch
    import torch.distributed as dist
    from clustering_genomic_signatures_b200 import native, synthetic
    from clustering_genomic_signatures_b200.engine import MatrixJob, ShardedNllJob
    from clustering_genomic_signatures_b200.flat import pack_sequences
    n, L = args.probe_models, 1000
    t0 = time.time()
    fm, _ = synthetic.make_family_models(n, 4, 9, 2000, max(1, n // 100))
    seqs = synthetic.sample_sequences(fm, L, 5)
    t_gen = time.time() - t0
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, steps=3, warmup=1):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            flush.zero_()
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    out = {"n_models": n, "depth": 9, "total_contexts": fm.total_contexts, "seq_len": L, "pairs": n * (n - 1) // 2,
           "host_generation_s": round(t_gen, 1)}
    # Frobenius intersection: replicated tables, sharded tiles
    hF = native.Handle(local_rank)
    hF.load_models(fm, skip_nll=True)
    jobF = MatrixJob(hF, "frobenius_intersection", n)
    ms = timed(jobF.run)
    jobF.check()
    out["frobenius_ms"] = ms
    out["frobenius_pairs_per_s"] = out["pairs"] / (ms / 1e3)
    out["frobenius_checksum"] = bit_checksum(jobF.D32)
    del jobF
    hF.close()
    # NLL: sharded models, replicated sequences
    m0, m1 = ShardedNllJob.shard_bounds(n, world)[rank]
    hN = native.Handle(local_rank)
    hN.load_models(fm.subset(range(m0, m1)) if world > 1 else fm)
    hN.load_sequences_packed(*pack_sequences(seqs))
    jobN = ShardedNllJob(hN, n, L, mode="warp", exchange=args.exchange, rank=rank, world=world)
    ms = timed(jobN.run)
    jobN.check()
    out["nll_ms"] = ms
    out["nll_pairs_per_s"] = out["pairs"] / (ms / 1e3)
    out["nll_checksum"] = bit_checksum(jobN.D32)
    out["nll_exchange"] = jobN.exchange
    jobN.close()
    hN.close()
    return out


# ---------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    kind, n, L, depth, ctx, desc = WORKLOADS[args.workload]
    n = args.n_models or n
    L = args.seq_len or L
    if args.impl == "reference":
        run_reference(args, kind, n, L, depth, ctx, desc)
        return

    import torch
    import torch.distributed as dist
    from clustering_genomic_signatures_b200 import native
    from clustering_genomic_signatures_b200.flat import pack_sequences

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from clustering_genomic_signatures_b200.engine import ShardedNllJob
    fm_full, fam, seqs = make_inputs(kind, n, L, depth, ctx)
    # N > 1, NLL: the models are sharded over the ranks (each rank loads, logs and tabulates only its own), the
    # sequences are replicated; every other kind replicates the tables and shards the tile space
    sharded = world > 1 and kind == "nll"
    if sharded:
        m0, m1 = ShardedNllJob.shard_bounds(n, world)[rank]
        fm = fm_full.subset(range(m0, m1))
    else:
        fm = fm_full
    h = native.Handle(local_rank)
    sm_count = h.device_info()["sm_count"]

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    keep = []
    host = {}
    for name, arr in (("off", fm.ctx_offsets.astype(np.int64)), ("len", fm.ctx_len.astype(np.uint8)),
                      ("code", fm.ctx_code.astype(np.uint32)), ("prob", fm.prob.astype(np.float64))):
        t, a = pinned(arr); keep.append(t); host[name] = a
    h2d = sum(a.nbytes for a in host.values())
    if seqs is not None:
        words, woff, lens = pack_sequences(seqs)
        for name, arr in (("words", words), ("woff", woff), ("lens", lens)):
            t, a = pinned(arr); keep.append(t); host[name] = a
        h2d += host["words"].nbytes + host["woff"].nbytes + host["lens"].nbytes
    out_t = torch.empty((n, n), dtype=torch.float32).pin_memory()
    out_h = out_t.numpy()
    d2h = out_h.nbytes

    def load_all():
        # NLL: the sequences are given, so nothing on the device reads the transition rows themselves -- the same load the
        # drop-in NegativeLogLikelihood class does when every model already carries its sampled sequence
        h.load_models_raw(fm.n_models, host["off"], host["len"], host["code"], host["prob"], skip_nll=(kind != "nll"),
                          nll_only=(kind == "nll"), device_log=(kind == "nll" and args.mode == "kmer"))
        if seqs is not None:
            h.load_sequences_packed(host["words"], host["woff"], host["lens"])

    load_all()
    stream = torch.cuda.current_stream().cuda_stream
    D32 = torch.empty((n, n), dtype=torch.float32, device=dev)
    symmetric = kind != "estimate"
    tile = args.tile or (32 if n <= 1024 else 64 if n <= 4096 else 128)
    if kind == "nll" and args.mode == "kmer":
        tile = max(128, tile // 128 * 128)
    job = None
    replicated = False
    if sharded:
        job = ShardedNllJob(h, n, L, mode=args.mode, exchange=args.exchange)
        D32 = job.D32
    elif world > 1 and kind in native.PAIR_KINDS and h.pair_kernel_choice(kind) == 1:
        replicated = True   # tensor-core pair path: every rank computes the whole (cheap) matrix, nothing to shard
    elif world > 1:
        nt, slots = native.tile_plan(n, tile, world, symmetric)
        payload = torch.zeros(slots * tile * tile, dtype=torch.float32, device=dev)
        gathered = torch.empty(world * slots * tile * tile, dtype=torch.float32, device=dev)

    def device_step():
        if sharded:
            job.mode = args.mode
            job.run()
        elif world == 1 or replicated:
            if kind == "nll":
                h.nll_matrix(L, args.mode, D32, None, stream)
            elif kind == "estimate":
                h.estimate_matrix(D32, None, stream)
            else:
                h.pair_matrix(kind, D32, None, stream)
        else:
            if kind == "nll":
                h.nll_tiles(L, args.mode, tile, rank, world, payload, stream)
            elif kind == "estimate":
                h.estimate_tiles(tile, rank, world, payload, stream)
            else:
                h.pair_tiles(kind, tile, rank, world, payload, stream)
            dist.all_gather_into_tensor(gathered, payload)
            h.assemble(n, tile, world, symmetric, gathered, D32, stream)

    def e2e_step():
        load_all()
        if world == 1:
            if kind == "nll":
                h.nll_matrix_h(L, args.mode, want_f64=False, out32=out_h)
            elif kind == "estimate":
                device_step(); out_t.copy_(D32, non_blocking=True); torch.cuda.synchronize()
            else:
                h.pair_matrix_h(kind, want_f64=False, out32=out_h)
        else:
            device_step()
            if rank == 0:   # the matrix goes to the host once: rank 0 is the process that consumes it (see the docstring)
                out_t.copy_(D32, non_blocking=True)
            torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup, use_events):
        for _ in range(warmup):
            fn()
        barrier()
        total_ms = 0.0
        t_begin = time.time()
        if use_events:
            # device time: one event pair per step on the launching stream, the L2 flush enqueued between the pairs;
            # the host does not wait between steps (launch latency is not the thing measured here, e2e measures that)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for a, b in evs:
                flush.zero_()
                a.record(); fn(); b.record()
            torch.cuda.synchronize()
            total_ms = sum(a.elapsed_time(b) for a, b in evs)
        else:
            for _ in range(steps):
                flush.zero_()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
                total_ms += (time.perf_counter() - t0) * 1000.0
        barrier()
        t_end = time.time()
        if world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms / steps, t_begin, t_end

    pairs = n * n if kind == "estimate" else n * (n - 1) // 2
    unit = "pair-distances/s"

    # ---- device-resident throughput (value) + per-kernel roofline
    prof_id = {"nll": 0, "estimate": 2}.get(kind, 1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        device_step()
    torch.cuda.synchronize()
    l0 = h.launch_count()
    ms_step, tb, te = timed(device_step, args.steps, 0, True)
    launches = h.launch_count() - l0
    nll_kernel = h.last_nll_kernel() if kind == "nll" else -1   # which score kernel the timed steps used
    pair_tc = kind not in ("nll", "estimate") and h.last_pair_kernel() == 1
    clocks = sampler.stop(tb, te) if rank == 0 else None
    # the dominant kernel's own duration (the roofline's denominator): the same K steps once more, now with the library's
    # event pair around that kernel on the launching stream.  Kept out of the step timing above: an event recorded between
    # two kernels ends the overlap of the second kernel's launch with the first one's tail (programmatic dependent launch)
    h.profile_enable(True)
    h.profile_read(prof_id)
    timed(device_step, args.steps, 0, True)
    k_ms, k_n = h.profile_read(prof_id)
    h.profile_enable(False)
    h.poll_error(stream)

    # ---- end to end through the host-buffer C-ABI
    e2e_ms, _, _ = timed(e2e_step, max(3, args.steps // 4), 1, False)

    if args.breakdown and world > 1 and not sharded and not replicated:
        def ev():
            return torch.cuda.Event(enable_timing=True)
        acc = np.zeros(3)
        reps = 10
        for _ in range(reps):
            flush.zero_()
            e0, e1, e2, e3 = ev(), ev(), ev(), ev()
            barrier()
            e0.record()
            h.nll_tiles(L, args.mode, tile, rank, world, payload, stream) if kind == "nll" else (
                h.estimate_tiles(tile, rank, world, payload, stream) if kind == "estimate" else
                h.pair_tiles(kind, tile, rank, world, payload, stream))
            e1.record()
            dist.all_gather_into_tensor(gathered, payload)
            e2.record()
            h.assemble(n, tile, world, symmetric, gathered, D32, stream)
            e3.record()
            torch.cuda.synchronize()
            acc += np.array([e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)])
        sys.stderr.write("breakdown rank %d: tiles %.3f ms, all-gather (incl. wait for the slowest rank) %.3f ms, assemble %.3f ms\n"
                         % (rank, acc[0] / reps, acc[1] / reps, acc[2] / reps))

    sweep = None
    if kind == "nll" and not args.no_mode_sweep:
        # the other kernels of the same path, 3 timed steps each, for the record (not the headline)
        sweep = {}
        main_mode = args.mode
        for mname in ("sequential", "warp", "kmer"):
            if mname == main_mode:
                sweep[mname] = ms_step
                continue
            args.mode = mname
            if world > 1 and not sharded and mname == "kmer" and tile % 128:
                continue
            try:
                m_ms, _, _ = timed(device_step, 3, 1, True)
                sweep[mname] = m_ms
            except Exception as exc:  # e.g. tile not a multiple of 128 for kmer
                sweep[mname] = "n/a: %s" % str(exc)[:60]
        args.mode = main_mode
    def bit_checksum(t):
        """Exact, order-independent fingerprint of a float32 matrix: the int64 sum of its bit patterns."""
        return int(t.contiguous().view(torch.int32).to(torch.int64).sum().item())

    device_step()
    torch.cuda.synchronize()
    h.poll_error(stream)
    matrix_checksum = bit_checksum(D32)

    # ---- N > 1: is this rank's matrix bit-identical to the one a single GPU computes from the full, unsharded set?
    parity_check = None
    if world > 1:
        h1 = native.Handle(local_rank)
        h1.load_models(fm_full, skip_nll=(kind != "nll"), device_log=(kind == "nll" and args.mode == "kmer"))
        if seqs is not None:
            h1.load_sequences_packed(*pack_sequences(seqs))
        ref = torch.empty((n, n), dtype=torch.float32, device=dev)
        if kind == "nll":
            h1.nll_matrix(L, args.mode, ref, None, stream)
        elif kind == "estimate":
            h1.estimate_matrix(ref, None, stream)
        else:
            h1.pair_matrix(kind, ref, None, stream)
        h1.poll_error(stream)
        same = torch.tensor([1 if torch.equal(ref, D32) else 0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        parity_check = {"bit_identical_to_single_gpu_on_every_rank": bool(int(same.item())),
                        "single_gpu_checksum": bit_checksum(ref), "what": "torch.equal(this rank's [N,N] float32 matrix, "
                        "the matrix one GPU computes from the unsharded set with the same kernel mode), min over ranks"}
        del ref
        h1.close()

    # ---- scale_probe: the same code on a cfg-4 slice that is big enough for 8 GPUs to matter
    probe = None
    if args.probe_models > 0:
        probe = scale_probe(args, world, rank, local_rank, dev, bit_checksum)

    i8_peak = None
    if (nll_kernel == 3 or pair_tc) and rank == 0:
        i8_peak = {"cta_group_2_tops": h.tensor_peak_i8(2)[0], "cta_group_1_tops": h.tensor_peak_i8(1)[0]}

    exchange_used = None
    if job is not None:
        exchange_used = job.exchange
        job.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
    orders = fm_full.orders().astype(np.float64)
    alg_bytes, alg_desc = algorithmic_bytes(kind, fm_full, L, orders)
    alg_bytes_rank = alg_bytes / world  # each rank's launch covers its share of the units
    k_avg_ms = k_ms / max(k_n, 1)
    launches_per_step = max(1, k_n // args.steps)
    achieved = alg_bytes_rank / launches_per_step / (k_avg_ms / 1000.0) / 1e9 if k_avg_ms > 0 else None
    kernel_name = {0: {3: "k_g8_gemm<EpiNll> (int8 tcgen05 contraction)", 1: "k_kmer_dmma", 2: "k_kmer_gemm"}.get(nll_kernel, "k_nll_scores")
                   if kind == "nll" else "", 1: "k_g8_gemm<EpiRaw> (int8 tcgen05, a.a digit GEMM)" if pair_tc else "k_pair_partials",
                   2: "k_est_pairs"}[prof_id]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": (achieved / peak_gbs) if achieved else None, "traffic": None,
                "kernel": kernel_name,
                "kernel_ms_per_launch": k_avg_ms, "kernel_launches_per_step": launches_per_step,
                "kernel_share_of_step": (k_ms / args.steps) / ms_step if ms_step > 0 else None,
                "algorithmic_bytes_per_step": alg_bytes, "algorithmic_bytes_formula": alg_desc, "peak_source": peak_src}
    # ncu dram bytes per launch of the dominant kernel, when a committed profile summary provides it
    traffic_key = args.workload + ("_" + args.mode if kind == "nll" else "")
    if nll_kernel == 3:
        traffic_key = args.workload + "_i8mma"
    if pair_tc:
        traffic_key = args.workload + "_tc"
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            roofline["traffic"] = json.load(fh).get(traffic_key)
    except Exception:
        pass
    Kdim = max(16.0, 4.0 ** (orders.max() + 1)) if kind == "nll" else 0.0
    if nll_kernel == 3 and k_avg_ms > 0:
        # the dominant kernel is an int8 tensor-core GEMM (tcgen05.mma cta_group::2 kind::i8): [N seq] x [8 digit rows per
        # model] x [4^(D+1) k-mers], 2 integer ops per MAC.  Its roof is the tensor pipe.  MEASURED_PEAKS.json has no int8
        # figure, so the peak is measured live by the library's probe (the same instruction back to back from resident
        # shared memory on every SM pair, vgs_tensor_peak_i8); 2 x the measured cuBLAS bf16 burst is reported beside it.
        K8 = max(128.0, 4.0 ** (orders.max() + 1))
        nd, frac_bits = h.kmer_digit_info()   # digit rows per model of the operand as built (6 / 7 / 8), fraction bits of ln p
        ops = 2.0 * n * (float(nd) * n) * K8 / world
        peak_bf16 = float(peaks.get("bf16_tflops", 1590.0))
        a_tops = ops / (k_avg_ms / 1e3) / 1e12
        probe_peak = (i8_peak or {}).get("cta_group_2_tops") or 0.0
        peak_used = probe_peak if probe_peak > 0 else 2.0 * peak_bf16
        roofline.update({"bound": "tensor", "achieved": a_tops, "peak": peak_used, "unit": "TFLOP/s",
                         "frac": a_tops / peak_used,
                         "peak_source": ("int8 tensor-pipe peak measured in this run by vgs_tensor_peak_i8 (tcgen05.mma cta_group::2 "
                                         "kind::i8 M256 N256 K32 from resident shared memory, all SM pairs; of measured)")
                         if probe_peak > 0 else "2 x MEASURED_PEAKS.json bf16_tflops (probe unavailable)",
                         "int8_peak_probe": i8_peak,
                         "frac_of_2x_bf16_burst": a_tops / (2.0 * peak_bf16),
                         "two_x_bf16_burst_tflops": 2.0 * peak_bf16,
                         "ops_per_step": ops * world,
                         "ops_formula": "2 x N_seq x (%d digit rows x N_models) x max(128, 4^(D+1)) int8 ops (real rows only: "
                                        "padding of the 256 x N tiles is not counted)" % nd,
                         "digit_rows_per_model": nd, "ln_p_fraction_bits": frac_bits,
                         "binding_resource": "the SM's shared-memory data pipe, shared by the bulk-copy writes of the operand stages and the "
                                             "MMAs' operand reads (ncu, profiles/r2_ncu_full_k_g8_gemm_nll_cfg3.txt: integer tensor pipe and "
                                             "sm__mem_tensor active 89 % of the working SMs' cycles; 104 KB per 128-byte k-block at 128 B/clk = 812 of "
                                             "the ~915 cycles it takes, MMA floor 784); at cfg 3 the 4 x 16 units occupy 64 of the 74 CTA pairs",
                         "useful_fp64_equivalent_macs_per_s": n * n * K8 / world / (k_avg_ms / 1e3),
                         "note": "every int8 MAC carries one base-256 digit of round(ln p x 2^%d) (%d digits: this set has "
                                 "%s): the useful rate is one fp64-equivalent MAC per %d executed int8 MACs"
                                 % (frac_bits, nd, "p = 0 entries / probabilities below e^-16" if nd > 6 else "no probability below e^-16", nd)})
        for k in ("algorithmic_bytes_per_step", "algorithmic_bytes_formula"):
            roofline.pop(k, None)   # SURVEY 8d's scan-bytes model does not describe a contraction
    elif pair_tc and k_avg_ms > 0:
        # Frobenius (intersection) / Projection on the tensor cores: the dominant kernel is the a.a digit GEMM over the dense
        # universe: 4 digit rows per model on either side, upper triangle of 256-row blocks, K = 4 x universe (padded to 128)
        D_ = int(orders.max())
        U_ = (4 ** (D_ + 1) - 1) // 3
        K8 = (4 * U_ + 127) // 128 * 128
        rows = 4 * n
        g_ld = (rows + 31) // 32 * 32
        macs = sum(256.0 * (g_ld - 256 * a) * K8 for a in range((rows + 255) // 256))
        ops = 2.0 * macs
        a_tops = ops / (k_avg_ms / 1e3) / 1e12
        probe_peak = (i8_peak or {}).get("cta_group_2_tops") or 0.0
        peak_bf16 = float(peaks.get("bf16_tflops", 1590.0))
        peak_used = probe_peak if probe_peak > 0 else 2.0 * peak_bf16
        roofline.update({"bound": "tensor", "achieved": a_tops, "peak": peak_used, "unit": "TFLOP/s", "frac": a_tops / peak_used,
                         "peak_source": "int8 tensor-pipe peak measured in this run by vgs_tensor_peak_i8 (of measured)" if probe_peak > 0
                         else "2 x MEASURED_PEAKS.json bf16_tflops (probe unavailable)",
                         "int8_peak_probe": i8_peak, "frac_of_2x_bf16_burst": a_tops / (2.0 * peak_bf16), "ops_per_launch": ops,
                         "ops_formula": "2 x sum over 256-row blocks a of 256 x (4N - 256 a) x (4 x universe): the executed int8 MACs of the a.a "
                                        "digit GEMM (16 digit pairs per model pair, upper triangle)",
                         "note": "second GEMM (sq.n, 9 x N x N x universe MACs) and the 128-bit recombination follow; the CUDA-core "
                                 "kernel k_pair_partials runs the same matrix in the time profiles/r2_cfg5_sweep_1gpu.jsonl lists"})
        for k in ("algorithmic_bytes_per_step", "algorithmic_bytes_formula"):
            roofline.pop(k, None)
    elif kind != "nll" and k_avg_ms > 0:
        # CUDA-core pair kernels / Estimate: the tables are staged in shared memory (or live in L2) and reused for many pairs,
        # so SURVEY 8d's per-pair bytes are NOT DRAM traffic and their rate is not bounded by HBM (it exceeded the peak in
        # round 1).  The roofline entry is therefore the kernel's measured DRAM traffic (ncu, profiles/traffic.json) against
        # HBM -- a small fraction: the kernel is bound by shared-memory gathers and issue slots (ncu: XU pipe 45-55 %,
        # shared-memory wavefronts 60 %, profiles/r2_ncu_full_pairs_tc_vs_cuda_core.txt) -- and the algorithmic rate is kept
        # beside it, labelled as what it is.
        tr = roofline.get("traffic")
        roofline["algorithmic_rate_GBps_not_a_roofline"] = achieved
        if tr:
            a_dram = tr / (k_avg_ms / 1e3) / 1e9
            roofline.update({"achieved": a_dram, "frac": a_dram / peak_gbs})
        else:
            roofline.update({"achieved": None, "frac": None})
        roofline["binding_resource"] = "shared-memory gathers + issue slots (not HBM): see profiles/r2_ncu_full_pairs_tc_vs_cuda_core.txt"
    elif kind == "nll" and nll_kernel in (1, 2) and k_avg_ms > 0:
        # the contraction is an fp64 MMA kernel: report it against the fp64 pipe as well
        # (64 fp64 FMA lanes per SM: 2 * 64 * SMs * max SM clock)
        flops = 2.0 * n * n * Kdim / world
        peak64 = 2.0 * 64 * sm_count * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12
        roofline["fp64"] = {"achieved_tflops": flops / (k_avg_ms / 1e3) / 1e12, "peak_tflops": peak64,
                            "frac": flops / (k_avg_ms / 1e3) / 1e12 / peak64,
                            "peak_source": "64 fp64 FMA lanes/SM x %d SMs x sm_max_mhz" % sm_count}

    if kind == "nll" and nll_kernel == 0 and roofline.get("frac") and roofline["frac"] > 1.0:
        # a staged scan whose SURVEY 8d bytes "exceed" the HBM peak is not HBM-bound (the table bytes are re-read from shared
        # memory, not DRAM): the entry becomes the measured DRAM traffic, the algorithmic rate stays beside it
        roofline["algorithmic_rate_GBps_not_a_roofline"] = roofline["achieved"]
        tr = roofline.get("traffic")
        a_dram = tr / launches_per_step / (k_avg_ms / 1e3) / 1e9 if tr else None
        roofline.update({"achieved": a_dram, "frac": (a_dram / peak_gbs) if a_dram else None,
                         "binding_resource": "shared-memory gathers (see smem_gather), not HBM"})
    if kind == "nll" and nll_kernel == 0 and k_avg_ms > 0:
        # scan kernel: one random 8-byte shared-memory gather per scored base.  Conflict-free that is 16 gathers per clock and
        # SM (128 B/clk); random 8-byte addresses from a half-warp collide on the 16 bank pairs 3.1 x on average (measured 6.36
        # wavefronts per warp-wide LDS.64 against 2.04 ideal), so the reachable rate is ~5.2 gathers per clock and SM
        gathers = float(n) * n * L / world * (2.0 if orders.max() > 6 else 1.0)   # deep sets: first-level map + row
        per_clk_sm = gathers / (k_avg_ms / 1e3) / (sm_count * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6)
        roofline["smem_gather"] = {"gathers_per_clk_per_sm": per_clk_sm, "conflict_free_peak": 16.0, "random_address_expectation": 5.2,
                                   "frac_of_random_address_expectation": per_clk_sm / 5.2}
    line = {
        "metric": "pair-distances/sec", "value": pairs / (ms_step / 1000.0), "unit": unit, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": ("int8 x int8 -> int32 exact fixed point (ln p quantised to 2^-%d), f64 result" % h.kmer_digit_info()[1] if nll_kernel == 3 else
                  "f64" if kind in ("nll", "estimate") else "f32 rows, f64 accumulate"), "data": "synthetic",
        "config": {"workload": desc, "n_models": n, "seq_len": L, "pairs_per_step": pairs, "total_contexts": fm_full.total_contexts,
                   "nll_mode": args.mode if kind == "nll" else None, "nll_kernel": kernel_name if kind == "nll" else None, "tile": tile if world > 1 else None,
                   "l2": "flushed between steps (256 MiB memset outside the timed events)",
                   "parallelism": ("models sharded over ranks (each rank loads / tabulates its own), sequences replicated; score bands "
                                   + ("stored into every rank's matrix by the GEMM epilogue over NVLink peer memory + flag barrier"
                                      if exchange_used == "peer" else "all-gathered with NCCL") + "; local combine") if sharded else
                                  ("replicas: the tensor-core pair path computes the whole matrix on every rank (N <= 4096, nothing to shard)"
                                   if replicated else "tile-sharded upper triangle + NCCL all-gather" if world > 1 else "single GPU"),
                   "exchange": exchange_used,
                   "sm_count": sm_count},
        "e2e": {"value": pairs / (e2e_ms / 1000.0), "unit": unit, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h),
                "includes": "vgs_load_models (H2D of the flat model arrays -- N > 1: this rank's model shard --, ln p "
                            + ("on the device (VGS_LOAD_DEVICE_LOG: the k-mer contraction quantises it to 2^-43)" if kind == "nll" and args.mode == "kmer"
                               else "on the host (libm, streamed)" if kind == "nll" else "not needed") +
                            "), vgs_load_sequences_packed, device table builds, kernels, exchange, D2H of the float32 matrix"
                            + (" (N > 1: once, on rank 0 -- the consuming process; d2h_bytes_per_step is that one copy; every rank's H2D is "
                               "h2d_bytes_per_step of its own shard)" if world > 1 else "")},
        "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
        "matrix_checksum": matrix_checksum, "parity_check": parity_check, "scale_probe": probe,
    }
    if sweep is not None:
        line["config"]["ms_per_step_by_nll_mode"] = sweep
    if world == 1 and not args.no_cpu_baseline:
        sample_n = args.cpu_sample or {"nll": 24, "estimate": 6, "frobenius_union": 150, "projection": 150}.get(kind, 400)
        line["cpu_baseline"] = cpu_baseline(kind, fm, seqs, L, sample_n)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
