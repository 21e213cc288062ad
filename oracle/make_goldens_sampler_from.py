#!/usr/bin/env python
"""Golden vectors for ``VLMC.generate_sequence_from(length, context)`` with NON-EMPTY contexts
(src/vlmc/vlmc.pyx:165-172) -- TEST INFRASTRUCTURE ONLY.  The compiled reference (``oracle/_ref``) samples from each
of the five fixture models after ``random.seed(k)``; contexts shorter than, equal to and longer than the order, one
that no deep context matches, and one with a symbol outside ACGT.  Run here (the reference is not on the GPU box);
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
    with open(os.path.join(ROOT, "tests", "golden", "sampler_from_golden.json"), "w") as fh:
        json.dump({"cases": cases}, fh)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
