"""Host-side logic that needs no GPU: flat encoding, packing, tile plan (vs the C library), the
.tree ingest, the C-ABI export list, loud failure without a device, gloo world_size-2 sharding."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import golden_io
import tiling_ref
from clustering_genomic_signatures_b200 import flat, native, parse_trees_to_json, synthetic
from clustering_genomic_signatures_b200.vlmc import VLMC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_context_encoding_round_trip():
    assert flat.encode_context("") == (0, 0)
    assert flat.encode_context("CAA") == (3, 0b010000)
    for ctx in ["A", "T", "ACGTACGT", "TTTTTTTTTTTTTTT", "GATTACA"]:
        assert flat.decode_context(*flat.encode_context(ctx)) == ctx
    with pytest.raises(KeyError):
        flat.encode_context("ACN")
    # rolling history: the length-k suffix code is h & (4**k - 1)
    h = 0
    s = "GATTACAGATTACA"
    for ch in s:
        h = (h << 2) | flat.SYMBOL_CODE[ch]
    for k in range(1, 8):
        assert flat.encode_context(s[-k:])[1] == h & (4 ** k - 1)


def test_flat_models_from_trees_and_back():
    g = golden_io.fixtures()
    fm = golden_io.fixture_flat(g)
    assert fm.n_models == 5 and fm.orders().tolist() == g["orders"]
    for m in range(5):
        assert fm.tree(m) == g["trees"][m] and list(fm.tree(m)) == list(g["trees"][m])
    sub = fm.subset([3, 1])
    assert sub.tree(0) == g["trees"][3] and sub.tree(1) == g["trees"][1]
    with pytest.raises(ValueError):
        flat.FlatModels.from_trees([{"A" * 16: {"A": 1.0, "C": 0.0, "G": 0.0, "T": 0.0}}])
    # negative probabilities load like in the reference, which only fails when it takes their logarithm (vlmc.pyx:99):
    # the NLL load raises that "math domain error" (tests/test_gpu_edges.py::test_argument_errors)
    neg = flat.FlatModels.from_trees([{"": {"A": -0.1, "C": 0.5, "G": 0.3, "T": 0.3}}])
    assert neg.prob[0, 0] == -0.1


def test_pack_sequences_layout():
    rng = np.random.default_rng(0)
    seqs = [rng.integers(0, 4, size=L).astype(np.uint8) for L in (0, 1, 15, 16, 17, 64, 65, 1000)]
    words, offs, lens = flat.pack_sequences(seqs)
    assert np.all(offs % 4 == 0) and lens.tolist() == [len(s) for s in seqs]
    for i, s in enumerate(seqs):
        assert np.array_equal(golden_io.unpack(words, offs, lens, i), s)
    # first symbol in the two highest bits
    w, _, _ = flat.pack_sequences([np.array([3, 0, 1], dtype=np.uint8)])
    assert w[0] == (3 << 30) | (0 << 28) | (1 << 26)
    buf, off = flat.sequences_to_ascii(["ACGT", "", "TT"])
    assert buf.tobytes() == b"ACGTTT" and off.tolist() == [0, 4, 4, 6]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "vgs.h")).read()
    declared = sorted(set(re.findall(r"\b(vgs_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    L = native.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert sorted(s[0] for s in native.SIGNATURES) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (vgs_[a-z0-9_]+)", out))
    assert set(declared) <= exported


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        native.Handle(0)
    v = VLMC({"": {"A": 0.25, "C": 0.25, "G": 0.25, "T": 0.25}}, "x", {})
    with pytest.raises(RuntimeError):
        v.get_context("ACG")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "clustering-genomic-signatures_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "vlmc_oracle" not in text and "c_oracle" not in text and "ref_loader" not in text, f
                assert "oracle/" not in text, f


@pytest.mark.parametrize("n,tile,world,sym", [(5, 2, 1, True), (10, 3, 2, True), (33, 8, 4, True), (1000, 64, 8, True),
                                              (7, 2, 3, False), (20, 5, 8, False)])
def test_tile_plan_matches_c_library_and_round_trips(n, tile, world, sym):
    nt, n_tiles, slots = tiling_ref.plan(n, tile, world, sym)
    assert native.tile_plan(n, tile, world, sym) == (n_tiles, slots)
    rng = np.random.default_rng(n)
    D = rng.random((n, n)).astype(np.float32)
    if sym:
        D = ((D + D.T) / 2).astype(np.float32)
    gathered = np.stack([tiling_ref.payload_of_rank(D, tile, r, world, sym) for r in range(world)])
    assert np.array_equal(tiling_ref.assemble(gathered, n, tile, world, sym), D)
    # every tile is owned exactly once
    owned = [k for r in range(world) for k in range(r, n_tiles, world)]
    assert sorted(owned) == list(range(n_tiles))


def test_tree_ingest_matches_reference_parser():
    with open(os.path.join(golden_io.GOLD, "tree_parse_golden.json")) as fh:
        gold = json.load(fh)
    for name, item in gold.items():
        got = parse_trees_to_json.parse_tree_lines(item["text"].splitlines(True))
        assert list(got["tree"]) == list(item["json"]["tree"])
        assert got == item["json"]
    ref_dir = "/root/reference/original_test_trees"
    if os.path.isdir(ref_dir):  # full files, against the shipped .json (same keys, order and doubles)
        for f in os.listdir(ref_dir):
            if f.endswith(".tree"):
                got = parse_trees_to_json.parse_file(os.path.join(ref_dir, f))
                want = json.load(open(os.path.join(ref_dir, f[:-5] + ".json")))
                assert got["tree"] == want and list(got["tree"]) == list(want)


def _flat_of(tree):
    return flat.FlatModels.from_trees([tree])


def test_native_tree_parser_matches_reference_parser(tmp_path):
    """Row f3: vgs_parse_tree_text (C, host) against the goldens the REFERENCE's parser produced -- plain mode on the
    shipped excerpts, `deltas` mode (src/parse_trees_to_json.py:40-44) on oracle/make_goldens_tree_deltas.py's vectors:
    same context order, bit-identical rows and occurrence probabilities; and the directory loader's FlatModels equal the
    ones built from the reference's JSON."""
    for fname, deltas in (("tree_parse_golden.json", False), ("tree_parse_deltas_golden.json", True)):
        with open(os.path.join(golden_io.GOLD, fname)) as fh:
            gold = json.load(fh)
        for name, item in gold.items():
            ln, code, prob, occ = native.parse_tree_text(item["text"], deltas)
            want = item["json"]
            keys = [flat.decode_context(int(a), int(b)) for a, b in zip(ln, code)]
            assert keys == list(want["tree"])
            assert np.array_equal(prob, np.array([[want["tree"][k][c] for c in "ACGT"] for k in keys]))
            assert np.array_equal(occ, np.array([want["occurrence_probability"][k] for k in keys]))
            py = parse_trees_to_json.parse_tree_lines(item["text"].splitlines(True), deltas)
            assert py == want
            (tmp_path / name).write_text(item["text"])
        fm, occs = parse_trees_to_json.load_tree_dir_flat(str(tmp_path), deltas)
        names = sorted(gold)
        ref = flat.FlatModels.from_trees([gold[n]["json"]["tree"] for n in names])
        assert np.array_equal(fm.ctx_offsets, ref.ctx_offsets) and np.array_equal(fm.ctx_len, ref.ctx_len)
        assert np.array_equal(fm.ctx_code, ref.ctx_code) and np.array_equal(fm.prob, ref.prob)
        assert fm.names == ["AB026117.1", "KF234407.1"]
    ref_dir = "/root/reference/original_test_trees"
    if os.path.isdir(ref_dir):   # the full files against the shipped .json
        fm, _ = parse_trees_to_json.load_tree_dir_flat(ref_dir)
        for m, f in enumerate(sorted(x for x in os.listdir(ref_dir) if x.endswith(".tree"))):
            assert fm.tree(m) == json.load(open(os.path.join(ref_dir, f[:-5] + ".json")))


def test_native_tree_parser_errors_and_dict_semantics():
    node = "Node: %d %s [ 1 2 3 4 ] 10 [ %s ][ -1 -1 -1 -1  ]\n"
    text = "Name: x\n" + node % (0, "#", "4 3 2 1") + node % (1, "A", "1 1 1 1") + node % (2, "A", "0 2 0 2") + "junk Node: 3 C\n"
    ln, code, prob, occ = native.parse_tree_text(text)
    assert ln.tolist() == [0, 1] and np.array_equal(prob, [[0.4, 0.3, 0.2, 0.1], [0.0, 0.5, 0.0, 0.5]])   # the repeated key overwrites in place
    assert np.array_equal(occ, [1.0, 0.4])
    assert parse_trees_to_json.parse_tree_lines(text.splitlines(True))["tree"] == {"": dict(zip("ACGT", (0.4, 0.3, 0.2, 0.1))),
                                                                                    "A": dict(zip("ACGT", (0.0, 0.5, 0.0, 0.5)))}
    with pytest.raises(ValueError):
        native.parse_tree_text(node % (0, "#", "0 0 0 0"))               # ZeroDivisionError in the reference
    with pytest.raises(ValueError):
        native.parse_tree_text(node % (0, "#", "4 3 2 1"), deltas=True)  # IndexError: no numbers[14..17]
    with pytest.raises(ValueError):
        native.parse_tree_text("Node: 0 # [ 1 2 3 4 ] 10 [ 1.5 1 1 1 ][ -1 -1 -1 -1 ]\n")   # int('1.5'): ValueError
    with pytest.raises(KeyError):
        native.parse_tree_text(node % (0, "#", "4 3 2 1") + node % (1, "AN", "1 1 1 1"))
    assert native.parse_tree_text("no nodes here\n")[0].size == 0


def test_vlmc_host_mirror_construction():
    assert VLMC.strip_parameters_from_name("PST_AB026117.1_10229_34094") == "AB026117.1"
    assert VLMC.strip_parameters_from_name("pst_vlmc_name_23_23") == "vlmc_name"
    assert VLMC.strip_parameters_from_name("pst_acgtacgt__") == "acgtacgt"
    old = '{"":{"A":0.5,"C":0.5,"G":0,"T":0},"AC":{"A":0.25,"C":0.25,"G":0.25,"T":0.25}}'
    v = VLMC.from_json(old, "x")
    assert v.order == 2 and v.sequence == "" and v.alphabet == ["A", "C", "G", "T"] and str(v) == "x"
    new = json.dumps({"tree": json.loads(old), "occurrence_probability": {"": 1.0, "AC": 0.1}})
    w = VLMC.from_json(new, "x")
    assert w == v and hash(w) == hash(v) and w.occurrence_probability("AC") == 0.1
    with pytest.raises(RuntimeError):
        v.occurrence_probability("")
    assert json.loads(w.to_json())["tree"] == w.tree


class _LookupOnly(object):
    """Stands in for the device handle of ONE model: answers `lookup` the way vgs_lookup does (ctx_id of the longest
    suffix), from a plain dict walk -- so the host arithmetic of VLMC.likelihood / estimated_context_distribution around
    the lookups can be checked without a GPU."""

    def __init__(self, tree):
        self.keys = list(tree.keys())
        self.order = max(len(k) for k in self.keys)

    def lookup(self, models, codes, lengths, want_all=False):
        out = np.zeros(len(codes), dtype=np.int32)
        for i, (c, ln) in enumerate(zip(codes, lengths)):
            hist = flat.decode_context(int(ln), int(c))
            out[i] = -1
            for k in range(min(len(hist), self.order), -1, -1):
                cand = hist[len(hist) - k:] if k else ""
                if cand in self.keys:
                    out[i] = self.keys.index(cand)
                    break
        return out


def test_vlmc_likelihood_and_context_distribution_host_logic():
    """The position -> (history code, length) arithmetic of VLMC._history_ids and what likelihood /
    estimated_context_distribution do with the ids (src/vlmc/vlmc.pyx:103-118, :182-206), against the compiled
    reference's values; the lookups themselves are the device's job (tests/test_gpu_dropin.py)."""
    g = golden_io.fixtures()
    vl = [VLMC(t, n, {}) for t, n in zip(g["trees"], g["names"])]
    for v in vl:
        v._handle = _LookupOnly(v.tree)
        v._keys = list(v.tree.keys())
    for c in golden_io.sampler_from_cases("likelihood"):
        assert vl[c["model"]].likelihood(c["sequence"]) == c["likelihood"]
    with pytest.raises(KeyError):
        vl[0].likelihood("ACGNT")
    for c in golden_io.sampler_from_cases("context_distribution"):
        v = vl[c["model"]]
        v.sequence = c["sequence"]      # cached while the length matches: no sampling, no device
        d = v.estimated_context_distribution(c["length"])
        assert list(d.keys()) == c["contexts"] and list(d.values()) == c["values"]


def test_matrix_layouts_host():
    from clustering_genomic_signatures_b200.clustering import util

    class Fake:
        def distance_matrix(self, vlmcs):
            n = len(vlmcs)
            return (np.arange(n * n, dtype=np.float32).reshape(n, n) / 7).astype(np.float32)

    tri = util.calculate_distances_within_vlmcs([1, 2, 3], Fake())
    assert tri.dtype == np.float32 and tri.shape == (9, 3)
    assert tri[:, 0].tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2] and tri[:, 1].tolist() == [0, 1, 2] * 3
    assert np.array_equal(util.index_distances([1, 2, 3], tri), Fake().distance_matrix([1, 2, 3]))
    with pytest.raises(TypeError):
        util.calculate_distances_within_vlmcs([1, 2], object())


def test_synthetic_sets_are_suffix_closed_and_normalised():
    fm, fam = synthetic.make_family_models(20, 5, 9, 200, 4)
    assert fm.orders().max() <= 9 and abs(fm.prob.sum(axis=1) - 1).max() < 1e-12
    for m in (0, 7, 19):
        t = fm.tree(m)
        assert "" in t and all(c[1:] in t for c in t if c)
    fc, _ = synthetic.make_full_chains(3, 1, 3)
    assert fc.n_ctx(0) == 1 + 4 + 16 + 64
    seqs = synthetic.sample_sequences(fm, 50, 1)
    assert seqs.shape == (20, 50) and seqs.max() <= 3


GLOO_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import tiling_ref
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, world = dist.get_rank(), 2
for n, tile, sym in ((37, 8, True), (21, 4, False)):
    rng = np.random.default_rng(3)
    D = rng.random((n, n)).astype(np.float32)
    if sym:
        D = ((D + D.T) / 2).astype(np.float32)
    mine = torch.from_numpy(tiling_ref.payload_of_rank(D, tile, rank, world, sym).ravel().copy())
    gathered = torch.empty(world * mine.numel(), dtype=torch.float32)
    dist.all_gather_into_tensor(gathered, mine)
    out = tiling_ref.assemble(gathered.numpy(), n, tile, world, sym)
    assert np.array_equal(out, D), "rank %d: assembled matrix differs" % rank
dist.barrier()
print("rank", rank, "ok")
"""


def test_gloo_world2_sharding_round_trip(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "rank %d ok" % r in o, o


GLOO_SHARD_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
root = sys.argv[1]
for p in (root, os.path.join(root, "tests"), os.path.join(root, "oracle")):
    sys.path.insert(0, p)
import c_oracle
from clustering_genomic_signatures_b200 import synthetic
from clustering_genomic_signatures_b200.engine import ShardedNllJob
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, world = dist.get_rank(), 2
n, L = 23, 200
fm, _ = synthetic.make_family_models(n, 3, 4, 30, 3)
seqs = synthetic.sample_sequences(fm, L, 9)
m0, m1 = ShardedNllJob.shard_bounds(n, world)[rank]
per = -(-n // world)
# this rank's band of the transposed score matrix: every sequence against ITS models only (the C oracle on the shard)
band = np.zeros((per, n))
band[: m1 - m0] = c_oracle.COracle(fm.subset(range(m0, m1))).score_matrix(seqs).T
Mt = torch.zeros((per * world, n), dtype=torch.float64)
Mt[rank * per:(rank + 1) * per] = torch.from_numpy(band)
dist.all_gather_into_tensor(Mt, Mt[rank * per:(rank + 1) * per].clone())
M = Mt.numpy()[:n].T.copy()                       # M[s][m]
co = c_oracle.COracle(fm)
assert np.array_equal(M, co.score_matrix(seqs)), "rank %d: gathered score matrix differs" % rank
assert np.array_equal(co.nll_from_scores(M, L), co.nll_matrix(seqs))
dist.barrier()
print("rank", rank, "ok")
"""


def test_gloo_world2_model_sharded_nll(tmp_path):
    """The N > 1 NLL plan on CPU: shard bounds, padded bands of the transposed score matrix, one all-gather that IS the
    assembly, local combine -- against the C oracle on the unsharded set."""
    script = tmp_path / "ws.py"
    script.write_text(GLOO_SHARD_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "rank %d ok" % r in o, o


def test_shard_bounds_cover_the_models():
    from clustering_genomic_signatures_b200.engine import ShardedNllJob
    for n in (1, 5, 23, 1000, 20000):
        for world in (1, 2, 3, 4, 8):
            b = ShardedNllJob.shard_bounds(n, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1)) and all(e >= a for a, e in b)
            per = -(-n // world)
            assert all(a == min(n, r * per) for r, (a, e) in enumerate(b))


@pytest.mark.parametrize("rects,kblocks,pairs,gran", [
    ([(a, 0, 8000) for a in range(4)], 128, 74, 32), ([(a, 0, 1024) for a in range(4)], 128, 74, 32),
    ([(0, 0, 64)], 1, 74, 32), ([(0, 0, 32)], 128, 74, 32), ([(a, 0, 16000) for a in range(8)], 128, 74, 32),
    ([(0, 0, 2400), (1, 1024, 1024)], 16, 74, 32), ([(0, 32, 160032 - 32)], 4, 3, 32),
    # the NLL contraction with 6 / 7 digit rows per model: granules of 16 / 32 models
    ([(a, 0, 6048) for a in range(4)], 128, 74, 96), ([(a, 0, 768) for a in range(4)], 128, 74, 96), ([(0, 768, 96)], 2, 74, 96),
    ([(a, 0, 7168) for a in range(4)], 128, 74, 224), ([(0, 0, 224), (3, 896, 1792)], 16, 5, 224)])
def test_int8_gemm_unit_plan(rects, kblocks, pairs, gran):
    """The unit list of the CTA-pair int8 GEMM (host planner behind vgs_g8_plan): every (256-row block, granule of rows,
    k block) of the requested rectangles is covered exactly once, widths are whole granules of at most two accumulators
    (each the largest multiple of the granule <= 256), k-splits only where they pay, and the dealt load per CTA pair
    stays close to the mean."""
    units, split = native.g8_plan(rects, kblocks, pairs, gran=gran)
    live = units[units[:, 2] > 0]
    seen = set()
    for a, b0, n, k0, kn, part in live:
        assert n % gran == 0 and 0 < n <= 2 * (256 // gran) * gran and kn > 0 and b0 % gran == 0
        assert (part != 0) == (split > 1)
        for r in range(b0, b0 + n, gran):
            for k in range(k0, k0 + kn):
                assert (a, r, k) not in seen
                seen.add((a, r, k))
    assert len(seen) == sum((r[2] // gran) * kblocks for r in rects)
    assert all((a, r, k) in seen for a, b0, rows in rects for r in range(b0, b0 + rows, gran) for k in range(kblocks))
    load = np.zeros(pairs)
    for i, (a, b0, n, k0, kn, part) in enumerate(units):
        load[i % pairs] += n * kn
    if len(live) >= 2 * pairs:
        assert load.max() <= 1.12 * load.mean()
    if split > 1:
        assert min(kn for _, _, _, _, kn, _ in live) >= 8   # a k-split keeps at least 8 k-blocks per unit


def test_bench_reference_arm_runs_on_cpu_and_keeps_the_contract():
    """`bench.py --impl reference` times the compiled reference on the host cores: no GPU needed, one JSON line with
    the contract's keys.  (Skipped where oracle/_ref was not built.)"""
    import json
    import subprocess
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    if not ref_loader.available():
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "frobenius",
                          "--cpu-sample", "24", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "pair-distances/sec" and line["unit"] == "pair-distances/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"] and line["gpu_launches"] == 0
