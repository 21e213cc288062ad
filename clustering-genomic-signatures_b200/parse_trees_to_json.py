"""``.tree`` -> ``{"tree": ..., "occurrence_probability": ...}`` ingest, the on-disk format either side of
the loader (reference: ``src/parse_trees_to_json.py:9-55``).

A classifier ``.tree`` file lists one context per ``Node:`` line,

    Node: <id> <context or #> [ prev counts x4 ] <total> [ next counts x4 ][ children x4 ]

and the transition row of a context is ``next_count / sum(next counts)`` (:39, :46-49); ``#`` is the
root context ``""`` (:53-54); the occurrence probability of a context is its total divided by the
root's total minus ``max(len(context) - 1, 0)`` (:29).
"""
import json
import os
import re

_NUMBER = re.compile(r"-?[0-9]+\.?[0-9]*")
_WORD = re.compile(r"[A-Za-z#]+")


def parse_tree_lines(lines, deltas=False):
    tree, occurrence = {}, {}
    root_total = None
    for line in lines:
        if not line.startswith("Node:"):
            continue
        numbers = _NUMBER.findall(line)
        counts = [int(numbers[k]) for k in (6, 7, 8, 9)]
        total = sum(counts)
        if deltas:
            row = dict(zip("ACGT", (float(numbers[k]) for k in (14, 15, 16, 17))))
        else:
            row = dict(zip("ACGT", (c / total for c in counts)))
        key = _WORD.findall(line)[1]
        key = "" if key == "#" else key
        tree[key] = row
        if root_total is None:
            root_total = total
        occurrence[key] = total / (root_total - max(len(key) - 1, 0))
    return {"tree": tree, "occurrence_probability": occurrence}


def parse_file(path, deltas=False):
    with open(path) as fh:
        return parse_tree_lines(fh, deltas)


def parse_trees(directory, deltas=False):
    """Write ``<name>.json`` next to every ``<name>.tree`` in ``directory`` (reference :9-15)."""
    for file in [f for f in os.listdir(directory) if f.endswith(".tree")]:
        stem, _ = os.path.splitext(file)
        with open(os.path.join(directory, stem + ".json"), "w") as fh:
            fh.write(json.dumps(parse_file(os.path.join(directory, file), deltas)))


def load_tree_dir_flat(directory, deltas=False):
    """``.tree`` files of a directory (sorted by name) straight to ``FlatModels`` through the library's native parser
    (``vgs_parse_tree_text``): no Python dicts, no JSON.  Returns (FlatModels, list of occurrence-probability arrays)."""
    import numpy as np

    from . import native
    from .flat import FlatModels
    from .vlmc import VLMC
    lens, codes, probs, occs, names, offsets = [], [], [], [], [], [0]
    for file in sorted(f for f in os.listdir(directory) if f.endswith(".tree")):
        with open(os.path.join(directory, file), "rb") as fh:
            ln, code, prob, occ = native.parse_tree_text(fh.read(), deltas)
        lens.append(ln); codes.append(code); probs.append(prob); occs.append(occ)
        offsets.append(offsets[-1] + len(ln))
        names.append(VLMC.strip_parameters_from_name(os.path.splitext(file)[0]))
    fm = FlatModels(np.array(offsets, dtype=np.int64), np.concatenate(lens), np.concatenate(codes), np.concatenate(probs), names)
    return fm.validate(), occs


def load_tree_dir(directory):
    """``.tree`` files of a directory straight to VLMC objects (no JSON round trip)."""
    from .vlmc import VLMC
    out = []
    for file in sorted(f for f in os.listdir(directory) if f.endswith(".tree")):
        stem, _ = os.path.splitext(file)
        d = parse_file(os.path.join(directory, file))
        out.append(VLMC(d["tree"], VLMC.strip_parameters_from_name(stem), d["occurrence_probability"]))
    return out
