"""Helpers shared by the tests: load tests/golden/* into trees / FlatModels / code arrays."""
import json
import os

import numpy as np

from clustering_genomic_signatures_b200.flat import FlatModels

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_LUT = np.zeros(256, dtype=np.uint8)
for _i, _ch in enumerate("ACGT"):
    _LUT[ord(_ch)] = _i


def fixtures():
    with open(os.path.join(GOLD, "fixtures_golden.json")) as fh:
        g = json.load(fh)
    g["trees"] = [{c: dict(zip("ACGT", p)) for c, p in zip(m["contexts"], m["probs"])} for m in g["models"]]
    g["names"] = [m["name"] for m in g["models"]]
    return g


def fixture_flat(g=None):
    g = g or fixtures()
    return FlatModels.from_trees(g["trees"], g["names"])


def str_codes(s):
    return _LUT[np.frombuffer(s.encode("ascii"), dtype=np.uint8)]


def unpack(words, word_offsets, lens, i):
    """Inverse of flat.pack_sequences for sequence i -> uint8 codes."""
    w = np.asarray(words[int(word_offsets[i]):int(word_offsets[i + 1])], dtype=np.uint32)
    sh = (30 - 2 * np.arange(16)).astype(np.uint32)
    return ((w[:, None] >> sh) & 3).astype(np.uint8).ravel()[: int(lens[i])]


def codes_str(codes):
    return np.frombuffer(b"ACGT", dtype=np.uint8)[codes].tobytes().decode("ascii")


def synthetic_small(name):
    z = np.load(os.path.join(GOLD, "synthetic_small.npz"))
    fm = FlatModels(z[name + "_ctx_offsets"], z[name + "_ctx_len"], z[name + "_ctx_code"], z[name + "_prob"],
                    ["%s%d" % (name, i) for i in range(len(z[name + "_ctx_offsets"]) - 1)])
    n = fm.n_models
    seqs = np.stack([unpack(z[name + "_seq_words"], z[name + "_seq_word_offsets"], z[name + "_seq_lens"], i)
                     for i in range(n)])
    est = np.stack([unpack(z[name + "_est_seq_words"], z[name + "_est_seq_word_offsets"], z[name + "_est_seq_lens"], i)
                    for i in range(4)])
    return fm, seqs, est, {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "_")}


def fixture_estimate():
    z = np.load(os.path.join(GOLD, "fixtures_estimate.npz"))
    seqs = np.stack([unpack(z["seq_words"], z["seq_word_offsets"], z["seq_lens"], i) for i in range(5)])
    return seqs, z["estimate"]


def sampler_from_cases(section="cases"):
    """generate_sequence_from(length, context) outputs of the compiled reference (oracle/make_goldens_sampler_from.py);
    sections "likelihood" and "context_distribution": VLMC.likelihood / estimated_context_distribution."""
    with open(os.path.join(GOLD, "sampler_from_golden.json")) as fh:
        return json.load(fh)[section]
