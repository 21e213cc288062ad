"""Small end-to-end pass over the paths compute-sanitizer should see (run under `compute-sanitizer --tool memcheck|racecheck`):
the CTA-pair int8 GEMM (single unit, several units per pair, k-split + fold kernel, frequent k-mers, shallow-model heads),
the model-sharded band path with two destinations, the scan kernels, the pair kernels, Estimate, average link."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clustering_genomic_signatures_b200 import native, synthetic  # noqa: E402
from clustering_genomic_signatures_b200.engine import sharded_nll_emulated  # noqa: E402
from clustering_genomic_signatures_b200.flat import pack_sequences  # noqa: E402

checks = []
for n, depth, ctx, L in ((37, 4, 40, 600), (261, 5, 90, 400), (40, 2, 12, 30000), (9, 0, 1, 200)):
    fm, _ = synthetic.make_family_models(n, 100 + n, depth, ctx, max(1, n // 10))
    seqs = synthetic.sample_sequences(fm, L, 7)
    packed = pack_sequences(seqs)
    h = native.Handle(0)
    h.load_models(fm)
    h.load_sequences_packed(*packed)
    Dk = h.nll_matrix_h(L, "kmer")[1]
    assert h.last_nll_kernel() == 3
    Ds = h.nll_matrix_h(L, "sequential")[1]
    Dw = h.nll_matrix_h(L, "warp")[1]
    off = ~np.eye(n, dtype=bool)
    checks.append(float(np.max(np.abs(Dk[off] - Ds[off]) / np.abs(Ds[off]))))
    checks.append(float(np.max(np.abs(Dw[off] - Ds[off]) / np.abs(Ds[off]))))
    if n <= 261 and L <= 600:
        out = sharded_nll_emulated(fm, packed, L, "kmer", 2, mirror=True)
        assert np.array_equal(out.cpu().numpy(), Dk.astype(np.float32))
        for kind in ("frobenius_intersection", "frobenius_union", "projection"):
            h.pair_matrix_h(kind)
    h.close()
fm, _ = synthetic.make_family_models(12, 3, 5, 60, 3)
seqs = synthetic.sample_sequences(fm, 3000, 4)
h = native.Handle(0)
h.load_models(fm)
h.load_sequences_packed(*pack_sequences(seqs))
h.estimate_matrix_h()
D32 = torch.from_numpy(h.pair_matrix_h("frobenius_intersection")[0]).cuda()
h.average_link(D32, 3)
print("sanitize target ok; max relative differences between modes:", max(checks))
