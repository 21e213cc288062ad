#!/usr/bin/env python
"""Golden vectors for ``VLMC.generate_sequence_from(length, context)`` with NON-EMPTY contexts
(src/vlmc/vlmc.pyx:165-172) -- TEST INFRASTRUCTURE ONLY.  The compiled reference (``oracle/_ref``) samples from each
of the five fixture models after ``random.seed(k)``; contexts shorter than, equal to and longer than the order, one
that no deep context matches, and one with a symbol outside ACGT; plus ``VLMC.likelihood`` and
``estimated_context_distribution`` on the same models.  Run here (the reference is not on the GPU box);
output: tests/golden/sampler_from_golden.json."""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402

REF = "/root/reference"
CONTEXTS = ["A", "GT", "ACGTACGT", "TTTTTTTTTTTTTTTTTTTT", "CAGNAC", "GATTACAGATTACAGATTACA", "ACGTACGTACG"]


def main():
    ref_loader.load()
    vlmcs = ref_loader.load_fixture_vlmcs(os.path.join(REF, "original_test_trees"))
    cases = []
    for m, v in enumerate(vlmcs):
        for k, ctx in enumerate(CONTEXTS):
            seed = 1000 + 17 * m + k
            random.seed(seed)
            s = v.generate_sequence_from(150, ctx)
            follow = random.random()   # where the stream stands afterwards
            cases.append({"model": m, "context": ctx, "seed": seed, "length": 150, "sequence": s, "next_random": follow})
    # VLMC.likelihood (vlmc.pyx:103-118) and estimated_context_distribution (:182-206) on the same models
    rng = random.Random(5)
    like = []
    for m, v in enumerate(vlmcs):
        seqs = ["", "A", "ACGT", "TTTTTTTTTTTT", "AAAAAAAAAAAAAAAAAAAAAAAA"]
        seqs += ["".join(rng.choice("ACGT") for _ in range(n)) for n in (7, 40, 200)]
        for s in seqs:
            like.append({"model": m, "sequence": s, "likelihood": v.likelihood(s)})
    dist = []
    for m, v in enumerate(vlmcs):
        random.seed(300 + m)
        v.reset_sequence()
        d = v.estimated_context_distribution(400)
        dist.append({"model": m, "seed": 300 + m, "length": 400, "contexts": list(d.keys()), "values": list(d.values()),
                     "sequence": v.sequence})
    with open(os.path.join(ROOT, "tests", "golden", "sampler_from_golden.json"), "w") as fh:
        json.dump({"cases": cases, "likelihood": like, "context_distribution": dist}, fh)
    print("wrote", len(cases), "sampler cases,", len(like), "likelihoods,", len(dist), "context distributions")


if __name__ == "__main__":
    main()
