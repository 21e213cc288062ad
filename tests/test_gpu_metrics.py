"""Row f4 on the B200: retrieval metrics, silhouette and within-cluster standard deviation through the C ABI,
against the golden vectors recorded from the reference (tests/golden/metrics_golden.npz) and the oracle."""
import os

import numpy as np
import pytest
import torch

import golden_io
import vlmc_oracle as O
from clustering_genomic_signatures_b200 import native
from clustering_genomic_signatures_b200.clustering import AverageLinkClustering
from clustering_genomic_signatures_b200.clustering.metrics import ClusteringMetrics
from clustering_genomic_signatures_b200.distance import FrobeniusNorm
from clustering_genomic_signatures_b200.flat import FlatModels
from clustering_genomic_signatures_b200.retrieval import distance_function_metrics, per_model_metrics
from clustering_genomic_signatures_b200.vlmc import VLMC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def z():
    return np.load(os.path.join(golden_io.GOLD, "metrics_golden.npz"))


@pytest.fixture(scope="module")
def h():
    return native.Handle(0)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("dname", ["frobenius", "union"])
def test_retrieval_metrics_match_reference(z, h, dname):
    labels = np.stack([z["genus"], z["family"]])
    in_top, dist_to, avg = h.retrieval_metrics(dev(z[dname + "_D64"]), labels)
    for k, tax in enumerate(("genus", "family")):
        assert np.array_equal(in_top[k], z["%s_procent_%s_in_top" % (dname, tax)])           # exact rationals
        np.testing.assert_allclose(dist_to[k], z["%s_distance_to_%s" % (dname, tax)], rtol=1e-14, atol=0)
    np.testing.assert_allclose(avg, z[dname + "_average_distance"], rtol=1e-14, atol=0)


def test_retrieval_ties_keep_list_order(z, h):
    labels = np.stack([z["genus"], z["family"]])
    in_top, _, _ = h.retrieval_metrics(dev(z["ties_D64"]), labels)
    assert np.array_equal(in_top[0], z["ties_procent_genus_in_top"])
    assert np.array_equal(in_top[1], z["ties_procent_family_in_top"])
    # float32 input, same ranks (one decimal survives the rounding order)
    in_top32, _, _ = h.retrieval_metrics(dev(z["ties_D64"].astype(np.float32)), labels)
    ref32, _, _ = O.retrieval_metrics(z["ties_D64"].astype(np.float32), z["genus"])
    assert np.array_equal(in_top32[0], np.array(ref32))


def test_retrieval_metrics_random_large(h):
    rng = np.random.Generator(np.random.PCG64(5))
    n = 700
    D = np.round(rng.random((n, n)), 3)      # plenty of exact ties
    labels = np.stack([rng.integers(0, 40, n), rng.integers(0, 3, n), np.zeros(n, dtype=np.int64), np.arange(n)]).astype(np.int32)
    in_top, dist_to, avg = h.retrieval_metrics(dev(D), labels)
    for t in range(labels.shape[0]):
        a, b, c = O.retrieval_metrics(D, labels[t])
        assert np.array_equal(in_top[t], np.array(a))
        np.testing.assert_allclose(dist_to[t], np.array(b), rtol=1e-13)
        np.testing.assert_allclose(avg, np.array(c), rtol=1e-13)
    assert np.all(in_top[2] == 1.0)          # one taxon: everything is in the top


@pytest.mark.parametrize("k", [7, 4])
def test_silhouette_matches_reference(z, h, k):
    a, b, s = h.silhouette(dev(z["indexed_f32"]), z["k%d_labels" % k])
    assert np.max(np.abs(s - z["k%d_silhouette" % k])) < 2e-6   # float32 sums in set order in the reference
    ao, bo, so = O.silhouette(z["indexed_f32"], z["k%d_labels" % k])
    assert np.max(np.abs(a - ao)) < 1e-6 and np.max(np.abs(b - bo)) < 1e-6 and np.max(np.abs(s - so)) < 2e-6
    # labels are arbitrary integers: a relabelling changes nothing
    a2, b2, s2 = h.silhouette(dev(z["indexed_f32"]), 1000 - 3 * z["k%d_labels" % k])
    assert np.array_equal(s, s2)


def test_silhouette_single_cluster_and_singletons(z, h):
    _, b, s = h.silhouette(dev(z["indexed_f32"]), z["k1_labels"])
    assert np.all(np.isinf(b)) and np.all(np.isnan(s))
    n = len(z["k1_labels"])
    a, b, s = h.silhouette(dev(z["indexed_f32"]), np.arange(n))
    assert np.all(a == 0.0)
    D = z["indexed_f32"].astype(np.float64).copy()
    np.fill_diagonal(D, np.inf)
    np.testing.assert_array_equal(b, D.min(axis=1))
    assert np.all(s == 1.0)


@pytest.mark.parametrize("k", [7, 4, 1])
def test_cluster_distance_std_matches_reference(z, h, k):
    stds = h.cluster_distance_std(dev(z["frobenius_D64"]), z["k%d_labels" % k])
    want = O.cluster_distance_std(z["frobenius_D64"], z["k%d_labels" % k])
    np.testing.assert_allclose(stds, np.array(want), rtol=1e-12, atol=1e-18)
    assert float(np.float32(sum(float(x) for x in stds) / len(stds))) == float(z["k%d_distance_std" % k])


def test_host_mirror_end_to_end(z):
    """Our VLMC objects + FrobeniusNorm + AverageLinkClustering + ClusteringMetrics / retrieval reproduce the
    reference's test_distance_function_ metrics and clustering metrics on the same models."""
    n = len(z["family"])
    fm = FlatModels(z["ctx_offsets"], z["ctx_len"], z["ctx_code"], z["prob"], ["m%d" % i for i in range(n)])
    vl = [VLMC(fm.tree(m), fm.names[m], {}) for m in range(n)]
    md = {v.name: {"species": v.name, "family": int(z["family"][i]), "genus": int(z["genus"][i])} for i, v in enumerate(vl)}
    d = FrobeniusNorm()
    m = distance_function_metrics(d, vl, md)
    assert m["distance_name"] == "FrobeniusNorm"
    for key in ("average_procent_of_genus_in_top", "average_procent_of_family_in_top", "total_average_distance_to_genus",
                "total_average_distance_to_family", "total_average_distance", "procent_genus_in_top", "distance_to_family",
                "average_distance"):
        # our fp64 matrix differs from the reference's float32 BLAS norm by its rounding (3e-7)
        assert m[key] == pytest.approx(float(z["frobenius_tdf_" + key]), rel=1e-6), key
    per = per_model_metrics(d, vl, md)
    assert np.array_equal(per["procent_genus_in_top"], z["frobenius_procent_genus_in_top"])
    cl = AverageLinkClustering(vl, d, md)
    for k in (7, 4):
        res = cl.cluster(k)
        assert np.array_equal(res.labels, z["k%d_labels" % k])
        met = ClusteringMetrics(res.labels, 0.0, cl.indexed_distances, vl, md, res.get_merge_distances())
        assert met.average_silhouette() == pytest.approx(float(z["k%d_average_silhouette" % k]), abs=2e-6)
        sil = met.silhouette_metric()
        assert max(abs(sil[v.name] - z["k%d_silhouette" % k][i]) for i, v in enumerate(vl)) < 5e-6
        assert met.average_distance_std(d) == pytest.approx(float(z["k%d_distance_std" % k]), rel=1e-5)
        assert met.sensitivity_specificity("genus") == tuple(z["k%d_sens_prec_genus" % k])
        assert met.average_percent_same_taxonomy("family") == pytest.approx(float(z["k%d_same_family" % k]), rel=1e-15)
