"""Golden vectors for the `deltas` mode of the .tree ingest (src/parse_trees_to_json.py:40-44), which the shipped
fixtures do not exercise: the excerpts of tests/golden/tree_parse_golden.json get four more numbers per Node line (the
layout the classifier's delta output has: numbers[14..17]) and the REFERENCE's own parser (imported from
/root/reference/src) turns them into JSON.  Run here (the reference is not on the GPU box); output:
tests/golden/tree_parse_deltas_golden.json."""
import importlib.util
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def main(reference="/root/reference"):
    spec = importlib.util.spec_from_file_location("ref_parse_trees", os.path.join(reference, "src", "parse_trees_to_json.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    with open(os.path.join(GOLD, "tree_parse_golden.json")) as fh:
        base = json.load(fh)
    rng = np.random.Generator(np.random.PCG64(11))
    out = {}
    for name, item in base.items():
        lines = []
        for line in item["text"].splitlines():
            if line.startswith("Node:"):
                d = rng.dirichlet([2, 2, 2, 2])
                # mixed spellings: plain decimals, a negative, an integer-looking token, many digits
                toks = ["%.6f" % d[0], "%.17g" % d[1], "-%.3f" % d[2] if rng.random() < 0.1 else "%.9f" % d[2], "%d" % round(d[3] * 3)]
                line = line + "[ " + " ".join(toks) + " ]"
            lines.append(line)
        text = "\n".join(lines) + "\n"
        with tempfile.NamedTemporaryFile("w", suffix=".tree", delete=False) as fh:
            fh.write(text)
            path = fh.name
        out[name] = {"text": text, "json": json.loads(ref._parse_file(path, True))}
        os.unlink(path)
    with open(os.path.join(GOLD, "tree_parse_deltas_golden.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote", len(out), "deltas goldens")


if __name__ == "__main__":
    main(*sys.argv[1:])
