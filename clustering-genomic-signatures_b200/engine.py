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
        if self.world > 1:
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
        if self.world == 1:
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
