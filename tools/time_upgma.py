"""Average-link timing: persistent cooperative kernel vs the multi-kernel path (VGS_UPGMA_MULTI_KERNEL=1), and
equality of their results.  Developer tool: python tools/time_upgma.py [n] [k]"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n, k):
    import torch
    from clustering_genomic_signatures_b200 import native
    rng = np.random.Generator(np.random.PCG64(11))
    # clustered points in the plane -> a matrix with structure (and float32 ties from the rounding)
    centers = rng.random((max(2, n // 50), 2)) * 10
    pts = centers[rng.integers(0, len(centers), n)] + rng.normal(0, 0.3, (n, 2))
    D = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)).astype(np.float32)
    D = np.round(D, 3)
    h = native.Handle(0)
    Dd = torch.from_numpy(D).cuda()
    h.average_link(Dd, max(k, n - 10))  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    labels, merges = h.average_link(Dd, k)
    dt = time.perf_counter() - t0
    return dt, labels, merges


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else max(1, n // 100)
    if os.environ.get("_VGS_CHILD"):
        dt, labels, merges = run(n, k)
        np.savez(os.environ["_VGS_CHILD"], dt=dt, labels=labels, merges=merges)
        sys.exit(0)
    res = {}
    for name, env in (("persistent", {}), ("multi_kernel", {"VGS_UPGMA_MULTI_KERNEL": "1"})):
        out = "/tmp/_upgma_%s.npz" % name
        e = dict(os.environ); e.update(env); e["_VGS_CHILD"] = out
        subprocess.check_call([sys.executable, os.path.abspath(__file__), str(n), str(k)], env=e)
        res[name] = np.load(out)
        print("%s: n=%d k=%d  %.3f s  (%d merges, %.1f us per merge)" % (name, n, k, float(res[name]["dt"]), len(res[name]["merges"]),
                                                                       1e6 * float(res[name]["dt"]) / max(1, len(res[name]["merges"]))))
    same = np.array_equal(res["persistent"]["labels"], res["multi_kernel"]["labels"]) and \
        np.array_equal(res["persistent"]["merges"], res["multi_kernel"]["merges"])
    print("identical labels and merge distances:", same)
    sys.exit(0 if same else 1)
