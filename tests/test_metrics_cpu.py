"""Row f4 on the CPU: the oracle's restatement of the matrix consumers and the host-side label arithmetic,
both pinned against tests/golden/metrics_golden.npz (recorded from the reference's own distance_metrics.py /
ClusteringMetrics by oracle/make_goldens_metrics.py)."""
import os

import numpy as np
import pytest

import golden_io
import vlmc_oracle as O
from clustering_genomic_signatures_b200.clustering.metrics import ClusteringMetrics


@pytest.fixture(scope="module")
def z():
    return np.load(os.path.join(golden_io.GOLD, "metrics_golden.npz"))


class _Named(object):
    def __init__(self, name):
        self.name = name


def _metadata(z):
    n = len(z["family"])
    vl = [_Named("m%d" % i) for i in range(n)]
    md = {v.name: {"family": int(z["family"][i]), "genus": int(z["genus"][i])} for i, v in enumerate(vl)}
    return vl, md


@pytest.mark.parametrize("dname", ["frobenius", "union"])
def test_oracle_retrieval_metrics_match_reference(z, dname):
    D = z[dname + "_D64"]
    for tax in ("genus", "family"):
        in_top, dist_to, avg = O.retrieval_metrics(D, z[tax])
        assert np.array_equal(np.array(in_top), z["%s_procent_%s_in_top" % (dname, tax)])
        assert np.array_equal(np.array(dist_to), z["%s_distance_to_%s" % (dname, tax)])
        assert np.array_equal(np.array(avg), z[dname + "_average_distance"])


def test_oracle_retrieval_ties_are_stable(z):
    for tax in ("genus", "family"):
        in_top, _, _ = O.retrieval_metrics(z["ties_D64"], z[tax])
        assert np.array_equal(np.array(in_top), z["ties_procent_%s_in_top" % tax])


@pytest.mark.parametrize("k", [7, 4])
def test_oracle_silhouette_matches_reference(z, k):
    a, b, s = O.silhouette(z["indexed_f32"], z["k%d_labels" % k])
    # the reference sums float32 in set order: agreement to float32 rounding of the sums
    assert np.max(np.abs(s - z["k%d_silhouette" % k])) < 2e-6
    assert abs(float(np.mean(s)) - float(z["k%d_average_silhouette" % k])) < 1e-6


def test_oracle_silhouette_single_cluster_is_nan(z):
    _, b, s = O.silhouette(z["indexed_f32"], z["k1_labels"])
    assert np.all(np.isinf(b)) and np.all(np.isnan(s))


@pytest.mark.parametrize("k", [7, 4, 1])
def test_oracle_cluster_std_matches_reference(z, k):
    stds = O.cluster_distance_std(z["frobenius_D64"], z["k%d_labels" % k])
    # cpdef float: the reference returns the mean through a C float
    assert float(np.float32(sum(stds) / len(stds))) == float(z["k%d_distance_std" % k])


@pytest.mark.parametrize("k", [7, 4, 1])
def test_host_label_metrics_match_reference(z, k):
    vl, md = _metadata(z)
    met = ClusteringMetrics(z["k%d_labels" % k], 0.0, z["indexed_f32"], vl, md, z["k%d_merge_distances" % k], device=0)
    for tax in ("genus", "family"):
        assert met.average_percent_same_taxonomy(tax) == pytest.approx(float(z["k%d_same_%s" % (k, tax)]), rel=1e-15)
        if k > 1:
            sens, prec = met.sensitivity_specificity(tax)
            assert (sens, prec) == tuple(z["k%d_sens_prec_%s" % (k, tax)])
    assert [float(x) for x in met.cluster_size_metrics()] == list(z["k%d_size_metrics" % k])
    assert met.get_merge_distances() == list(z["k%d_merge_distances" % k])
    assert met.get_latest_merge_distance() == float(np.float32(z["k%d_merge_distances" % k][-1]))  # cpdef float
    comp = [vl[i] for i in np.nonzero(z["k%d_labels" % k] == 0)[0]]
    tax = [md[v.name]["family"] for v in comp]
    assert met.percent_same_taxonomy(comp, "family") == sum(tax.count(t) for t in tax) / len(comp) ** 2
