"""Developer check: k-mer contraction (int8 tensor-core path unless VGS_KMER_GEMM=dmma) vs the bit-exact scan."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clustering_genomic_signatures_b200 import native, synthetic
from clustering_genomic_signatures_b200.flat import pack_sequences

n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 6
L = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
fm, _ = synthetic.make_family_models(n, 100 + depth, depth, 300, max(2, n // 10))
seqs = synthetic.sample_sequences(fm, L, 7)
h = native.Handle(0)
h.load_models(fm)
h.load_sequences_packed(*pack_sequences(seqs))
M_seq = h.nll_scores_h(native.SKIP_ORDER, "sequential")
t0 = time.time()
M_k = h.nll_scores_h(native.SKIP_ORDER, "kmer")
print("kmer call %.3f s" % (time.time() - t0))
err = np.abs(M_k - M_seq) / np.maximum(np.abs(M_seq), 1e-300)
print("n=%d depth=%d L=%d  max rel err %.3e  (mean |M| %.1f)  worst at %s" % (n, depth, L, err.max(), np.abs(M_seq).mean(),
                                                                          np.unravel_index(err.argmax(), err.shape)))
print("M_seq[0,:4]", M_seq[0, :4]); print("M_k  [0,:4]", M_k[0, :4])
D32, D64 = h.nll_matrix_h(L, "kmer")
D32s, D64s = h.nll_matrix_h(L, "sequential")
e2 = np.abs(D64 - D64s) / np.maximum(np.abs(D64s), 1e-300); e2[D64 == D64s] = 0
print("distance max rel err %.3e  diag zero %s  symmetric %s" % (e2.max(), bool(np.all(np.diag(D64) == 0)), bool(np.array_equal(D64, D64.T))))
sys.exit(0 if err.max() < 1e-12 else 1)
