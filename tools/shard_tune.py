"""Developer tool: one rank's share of the model-sharded cfg-3 step on ONE GPU (band of WORLD-th of the models + local combine),
for launch lists / timing of the multi-GPU step's kernels without a multi-rank run.  usage: python tools/shard_tune.py WORLD [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clustering_genomic_signatures_b200 import native, synthetic  # noqa: E402
from clustering_genomic_signatures_b200.engine import ShardedNllJob  # noqa: E402
from clustering_genomic_signatures_b200.flat import pack_sequences  # noqa: E402

world = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n, L = 1000, 10000
fm, fam = synthetic.config2(n, 2)
seqs = synthetic.sample_sequences(fm, L, 3)
m0, m1 = ShardedNllJob.shard_bounds(n, world)[0]
h = native.Handle(0)
h.load_models(fm.subset(range(m0, m1)), nll_only=True)
h.load_sequences_packed(*pack_sequences(seqs))
Mt = torch.zeros((n, n), dtype=torch.float64, device="cuda")
D = torch.empty((n, n), dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def step():
    h.nll_sharded_step("kmer", m0, n, [Mt.data_ptr()], None, 0, 1, 0, L, D, stream)
for _ in range(3):
    step()
torch.cuda.synchronize()
evs = []
for _ in range(steps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step(); b.record()
    evs.append((a, b))
torch.cuda.synchronize()
print("world %d share on one GPU: step %.4f ms" % (world, sum(a.elapsed_time(b) for a, b in evs) / steps))
