"""Developer tool: time the cfg-3 k-mer NLL step (and print the GEMM's cycle counters) under tuning knobs.
usage: python tools/i8_tune.py [steps]   (knobs via env: VGS_G8_PROF, VGS_G8_STAGES, VGS_G8_FORCE_N, VGS_I8_SPLITK)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clustering_genomic_signatures_b200 import native, synthetic  # noqa: E402
from clustering_genomic_signatures_b200.flat import pack_sequences  # noqa: E402

n, L = int(os.environ.get("TUNE_N", "1000")), int(os.environ.get("TUNE_L", "10000"))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
fm, fam = synthetic.config2(n, 2)
if os.environ.get("TUNE_NOZERO"):   # a set without p = 0 entries: six digit rows per model instead of seven
    fm.prob[:] = np.maximum(fm.prob, 1e-3)
    fm.prob[:] /= fm.prob.sum(axis=1, keepdims=True)
seqs = synthetic.sample_sequences(fm, L, 3)
h = native.Handle(0)
h.load_models(fm)
h.load_sequences_packed(*pack_sequences(seqs))
D = torch.empty((n, n), dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    h.nll_matrix(L, "kmer", D, None, stream)
torch.cuda.synchronize()
h.profile_enable(True)
h.profile_read(0)
evs = []
for _ in range(steps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); h.nll_matrix(L, "kmer", D, None, stream); b.record()
    evs.append((a, b))
torch.cuda.synchronize()
ms, k = h.profile_read(0)
print("digits per model %d, fraction bits %d | " % h.kmer_digit_info(), end="")
print("knobs %s | step %.4f ms | gemm %.4f ms | checksum %d" % (
    {k2: v for k2, v in os.environ.items() if k2.startswith(("VGS_", "TUNE_"))}, sum(a.elapsed_time(b) for a, b in evs) / steps, ms / max(k, 1),
    int(D.view(torch.int32).to(torch.int64).sum().item())))
