"""Device-resident matrix jobs, single- and multi-GPU (north_star (c)).

One process per GPU (``torch.distributed``, backend ``nccl``).  The model tables (and packed
sequences) are replicated on every GPU; the upper-triangular tile space of the N x N matrix
(the whole tile grid for the asymmetric EstimateVLMC) is dealt round-robin to the ranks
(tile k -> rank k % world, ``include/vgs.h``); each rank computes its tiles into a payload of
``slots * T * T`` float32; one NCCL all-gather over NVLink/NVSwitch replicates all payloads and a
local scatter/mirror kernel (``vgs_assemble``) writes the [N, N] matrix on every rank.  There is no
other collective on the data path: tiles are independent given the replicated tables.

PyTorch is used only for device buffers, the current stream and the collective.
"""
from . import native

KINDS = ("nll", "frobenius_intersection", "frobenius_union", "projection", "estimate")


def default_tile(n, world):
    """Tile edge: enough tiles for balance (>= ~16 per rank), not so small that staging dominates."""
    t = 256
    while t > 16:
        nt = -(-n // t)
        if nt * (nt + 1) // 2 >= 16 * world:
            break
        t //= 2
    return t


class ShardedNllJob(object):
    """NegativeLogLikelihood matrix over several GPUs with the MODELS sharded (one process per GPU).

    Rank r loads only its models (``shard_bounds``) -- so the host ln p, the uploads and the table builds shrink with the
    number of ranks -- and all sequences.  It scores every sequence against its models; the result is a contiguous band
    of the transposed fp64 score matrix ``Mt[model][sequence]``.  Exchange:

    * ``peer`` (default when CUDA IPC works): every rank maps every other rank's ``Mt`` over NVLink and the int8 GEMM's
      epilogue stores its band straight into all of them; a flag barrier in peer memory ends the step.  No payload, no
      collective, no scatter kernel.
    * ``nccl``: one ``all_gather_into_tensor`` of the bands (they are contiguous, so the gather IS the assembly).

    Then every rank turns ``Mt`` into the float32 [N, N] matrix locally (``vgs_nll_combine``).  Per-entry arithmetic does
    not depend on the sharding, so the matrix is bit-identical to the single-GPU one.
    """

    @staticmethod
    def shard_bounds(n, world):
        per = -(-n // world)
        return [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]

    def __init__(self, handle, n_total, length, mode="kmer", group=None, exchange="auto", rank=None, world=None):
        import torch
        import torch.distributed as dist
        self.h, self.n, self.length, self.mode, self.group = handle, n_total, length, mode, group
        live = dist.is_available() and dist.is_initialized()
        self.world = world if world is not None else (dist.get_world_size(group) if live else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if live else 0)
        self.per = -(-n_total // self.world)
        self.m0, self.m1 = self.shard_bounds(n_total, self.world)[self.rank]
        assert handle.n_models == self.m1 - self.m0, "the handle must hold exactly this rank's model shard"
        assert handle.n_seq == n_total, "the handle must hold all sequences"
        self.device = torch.device("cuda", handle.device)
        self.rows = self.per * self.world           # Mt rows, padded so that every rank's band has the same size
        self.D32 = torch.empty((n_total, n_total), dtype=torch.float32, device=self.device)
        self.exchange = "nccl"
        self.epoch = 0
        self._peer = None
        if exchange in ("auto", "peer") and self.world > 1 and live:
            try:
                self._setup_peer()
                self.exchange = "peer"
            except Exception as exc:  # CUDA IPC unavailable (container, MIG, ...): the NCCL exchange needs nothing special
                if exchange == "peer":
                    raise
                self._peer_error = str(exc)
            # every rank must take the same path
            flag = torch.tensor([1 if self.exchange == "peer" else 0], device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0 and self.exchange == "peer":
                self._teardown_peer()
                self.exchange = "nccl"
        if self.exchange == "nccl":
            self.Mt = torch.zeros((self.rows, n_total), dtype=torch.float64, device=self.device)
            self.mt_ptr = self.Mt.data_ptr()

    # -- peer memory
    def _setup_peer(self):
        import torch.distributed as dist
        # two score matrices, used alternately: a fast rank's stores of step k + 1 must not land in the matrix a slow rank
        # is still reading for step k (it cannot get two steps ahead: the barrier of step k + 1 needs everyone's signal)
        nbytes = self.rows * self.n * 8
        self._mt_bytes = nbytes
        mt_ptr, mt_ipc = self.h.peer_alloc(2 * nbytes)
        fl_ptr, fl_ipc = self.h.peer_alloc(256)
        self._peer = {"own": (mt_ptr, fl_ptr), "opened": []}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, (mt_ipc, fl_ipc), group=self.group)
        mts, fls = [], []
        for r, (mi, fi) in enumerate(everyone):
            if r == self.rank:
                mts.append(mt_ptr); fls.append(fl_ptr)
            else:
                a, b = self.h.peer_open(mi), self.h.peer_open(fi)
                self._peer["opened"] += [a, b]
                mts.append(a); fls.append(b)
        self.mt_ptr = mt_ptr
        self._dsts = [mt_ptr] + [p for r, p in enumerate(mts) if r != self.rank]
        self._flags = fls

    def _teardown_peer(self):
        if not self._peer:
            return
        for p in self._peer["opened"]:
            try:
                self.h.peer_close(p)
            except Exception:
                pass
        for p in self._peer["own"]:
            try:
                self.h.peer_free(p)
            except Exception:
                pass
        self._peer = None

    def close(self):
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        if self._peer and dist.is_initialized():
            dist.barrier(group=self.group)   # nobody unmaps while a peer may still store
        self._teardown_peer()

    def run(self):
        """Enqueue one full matrix on the current stream; returns the device tensor [N, N] float32."""
        import torch
        import torch.distributed as dist
        stream = torch.cuda.current_stream().cuda_stream
        mt = self.mt_ptr
        if self.exchange == "peer":
            self.epoch += 1
            off = (self.epoch & 1) * self._mt_bytes
            # peer order for the library: destination 0 = this rank's matrix; the flag arrays stay in rank order
            self.h.nll_sharded_step(self.mode, self.m0, self.n, [p + off for p in self._dsts], self._flags, self.rank, self.world,
                                    self.epoch, self.length, self.D32, stream)
            return self.D32
        else:
            self.h.nll_scores_band(self.mode, self.m0, self.n, [self.mt_ptr], stream)
            if self.world > 1:
                band = self.Mt[self.rank * self.per:(self.rank + 1) * self.per]
                dist.all_gather_into_tensor(self.Mt, band, group=self.group)
        self.h.nll_combine(self.n, mt, self.n, self.length, self.D32, None, stream)
        return self.D32

    def check(self):
        import torch
        self.h.poll_error(torch.cuda.current_stream().cuda_stream)


def sharded_nll_emulated(fm, packed, length, mode, world, device=0, mirror=False):
    """The model-sharded NLL path with every rank emulated on ONE GPU, one after the other: `world` handles, each with
    its model shard and all sequences, write their bands into one transposed score matrix; then the local combine.
    ``mirror`` makes every band go to two destinations (the multi-destination epilogue of the peer-store path, here
    with both matrices on this GPU) and returns the matrix combined from the second one."""
    import torch
    n = fm.n_models
    dev = torch.device("cuda", device)
    stream = torch.cuda.current_stream().cuda_stream
    per = -(-n // world)
    Mt = torch.zeros((per * world, n), dtype=torch.float64, device=dev)
    Mt2 = torch.zeros_like(Mt) if mirror else None
    handles = []
    for r, (m0, m1) in enumerate(ShardedNllJob.shard_bounds(n, world)):
        if m1 <= m0:
            continue
        h = native.Handle(device)
        h.load_models(fm.subset(range(m0, m1)))
        h.load_sequences_packed(*packed)
        h.nll_scores_band(mode, m0, n, [Mt.data_ptr()] + ([Mt2.data_ptr()] if mirror else []), stream)
        h.poll_error(stream)
        handles.append(h)
    D = torch.empty((n, n), dtype=torch.float32, device=dev)
    handles[0].nll_combine(n, Mt2 if mirror else Mt, n, length, D, None, stream)
    handles[0].poll_error(stream)
    for h in handles:
        h.close()
    return D


class MatrixJob(object):
    """All-pairs matrix of one kind on a loaded ``native.Handle``; reusable across steps."""

    def __init__(self, handle, kind, n, length=0, mode="sequential", tile=None, group=None):
        import torch
        import torch.distributed as dist
        assert kind in KINDS
        self.h, self.kind, self.n, self.length, self.mode = handle, kind, n, length, mode
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.symmetric = kind != "estimate"
        self.device = torch.device("cuda", handle.device)
        self.D32 = torch.empty((n, n), dtype=torch.float32, device=self.device)
        self.tile = tile or default_tile(n, self.world)
        # dense-universe Frobenius / Projection sets go to the tensor cores (N <= 4096, a fraction of a millisecond): every
        # rank computes the whole matrix itself -- nothing to shard, and the same bits as on one GPU
        self.replicated = self.world > 1 and kind in native.PAIR_KINDS and handle.pair_kernel_choice(kind) == 1
        if self.world > 1 and not self.replicated:
            n_tiles, slots = native.tile_plan(n, self.tile, self.world, self.symmetric)
            self.n_tiles, self.slots = n_tiles, slots
            self.payload = torch.zeros(slots * self.tile * self.tile, dtype=torch.float32, device=self.device)
            self.gathered = torch.empty(self.world * slots * self.tile * self.tile, dtype=torch.float32, device=self.device)

    def compute_tiles(self, rank, world, payload, stream):
        if self.kind == "nll":
            self.h.nll_tiles(self.length, self.mode, self.tile, rank, world, payload, stream)
        elif self.kind == "estimate":
            self.h.estimate_tiles(self.tile, rank, world, payload, stream)
        else:
            self.h.pair_tiles(self.kind, self.tile, rank, world, payload, stream)

    def run(self):
        """Enqueue one full matrix on the current stream; returns the device tensor [N, N] float32."""
        import torch
        import torch.distributed as dist
        stream = torch.cuda.current_stream().cuda_stream
        if self.world == 1 or self.replicated:
            if self.kind == "nll":
                self.h.nll_matrix(self.length, self.mode, self.D32, None, stream)
            elif self.kind == "estimate":
                self.h.estimate_matrix(self.D32, None, stream)
            else:
                self.h.pair_matrix(self.kind, self.D32, None, stream)
            return self.D32
        self.compute_tiles(self.rank, self.world, self.payload, stream)
        dist.all_gather_into_tensor(self.gathered, self.payload, group=self.group)
        self.h.assemble(self.n, self.tile, self.world, self.symmetric, self.gathered, self.D32, stream)
        return self.D32

    def run_emulated(self, world):
        """Every rank's tiles computed one after the other on THIS GPU and assembled -- the multi-GPU
        data path without a second device (used by the single-GPU invariance test)."""
        import torch
        stream = torch.cuda.current_stream().cuda_stream
        n_tiles, slots = native.tile_plan(self.n, self.tile, world, self.symmetric)
        per = slots * self.tile * self.tile
        gathered = torch.zeros(world * per, dtype=torch.float32, device=self.device)
        for r in range(world):
            self.compute_tiles(r, world, gathered[r * per:(r + 1) * per], stream)
        out = torch.empty((self.n, self.n), dtype=torch.float32, device=self.device)
        self.h.assemble(self.n, self.tile, world, self.symmetric, gathered, out, stream)
        self.h.poll_error(stream)
        return out

    def check(self):
        import torch
        self.h.poll_error(torch.cuda.current_stream().cuda_stream)
