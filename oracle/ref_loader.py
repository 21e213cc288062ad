"""Import the compiled reference (``oracle/_ref``) -- TEST INFRASTRUCTURE ONLY.

The product package never imports this module.  It exists so that ``tests/``, ``smoke()`` and
``bench.py``'s CPU-baseline legs can execute the reference's *own* implementation of the hot
path (built by ``oracle/build_ref.py``) without the reference's unrelated dependencies.

Import caveats of the reference that this loader works around:

* ``src/clustering/__init__.py:5-10`` imports modules needing matplotlib / skbio and one that
  crashes at import on numpy 2 (``fuzzy_similarity_clustering.pyx:6``).  We register *empty*
  ``vlmc`` / ``distance`` / ``clustering`` packages whose ``__path__`` points into
  ``oracle/_ref`` and import only the hot-path extension modules.
* ``src/distance/estimate.pyx:99`` calls ``scipy.stats.power_divergence`` whose modern versions
  (scipy >= 1.12) add a sum-agreement check that scipy 1.0.1 (``requirements.txt:8``) did not
  have.  ``patched_power_divergence()`` installs the same Pearson statistic without that check,
  scoped to a ``with`` block.
"""
import contextlib
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")

_loaded = None


def available():
    return os.path.exists(os.path.join(REF, "BUILD_OK"))


def _pkg(name):
    m = sys.modules.get(name)
    if m is None or getattr(m, "__vgs_ref__", False) is False:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, name)]
        m.__package__ = name
        m.__vgs_ref__ = True
        sys.modules[name] = m
    return m


def load():
    """Returns a namespace with the reference classes/functions of the hot path."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("oracle/_ref is not built; run `python oracle/build_ref.py` where /root/reference exists")
    vl = _pkg("vlmc")
    di = _pkg("distance")
    cl = _pkg("clustering")
    ns = types.SimpleNamespace()
    ns.VLMC = importlib.import_module("vlmc.vlmc").VLMC
    vl.VLMC = ns.VLMC
    ns.NegativeLogLikelihood = importlib.import_module("distance.negloglikelihood").NegativeLogLikelihood
    di.NegativeLogLikelihood = ns.NegativeLogLikelihood
    ns.FrobeniusNorm = importlib.import_module("distance.frobenius").FrobeniusNorm
    ns.Projection = importlib.import_module("distance.projection").Projection
    ns.ACGTContent = importlib.import_module("distance.acgt").ACGTContent
    ns.EstimateVLMC = importlib.import_module("distance.estimate").EstimateVLMC
    for k in ("FrobeniusNorm", "Projection", "ACGTContent", "EstimateVLMC"):
        setattr(di, k, getattr(ns, k))
    util = importlib.import_module("clustering.util")
    ns.calculate_distances_within_vlmcs = util.calculate_distances_within_vlmcs
    ns.index_distances = util.index_distances
    ns.ClusteringMetrics = importlib.import_module("clustering.clustering_metrics").ClusteringMetrics
    ns.GraphBasedClustering = importlib.import_module("clustering.graph_based_clustering").GraphBasedClustering
    ns.AverageLinkClustering = importlib.import_module("clustering.average_link_clustering").AverageLinkClustering
    ns.MSTClustering = importlib.import_module("clustering.mst_clustering").MSTClustering
    for k in ("ClusteringMetrics", "GraphBasedClustering", "AverageLinkClustering", "MSTClustering"):
        setattr(cl, k, getattr(ns, k))
    ns.patched_power_divergence = patched_power_divergence
    _loaded = ns
    return ns


@contextlib.contextmanager
def patched_power_divergence():
    """scipy 1.0.1 semantics for ``stats.power_divergence(lambda_="pearson")``: no sum check."""
    import numpy as np
    from scipy import stats

    orig = stats.power_divergence

    def pearson_no_sum_check(f_obs, f_exp=None, ddof=0, axis=0, lambda_=None):
        if lambda_ not in ("pearson", 1):
            return orig(f_obs, f_exp=f_exp, ddof=ddof, axis=axis, lambda_=lambda_)
        return stats.chisquare(np.asarray(f_obs, dtype=np.float64), np.asarray(f_exp, dtype=np.float64),
                               ddof=ddof, axis=axis, sum_check=False)

    stats.power_divergence = pearson_no_sum_check
    try:
        yield
    finally:
        stats.power_divergence = orig


def load_fixture_vlmcs(directory):
    """The 5 ``original_test_trees`` models in sorted-filename order."""
    ns = load()
    out = []
    for f in sorted(os.listdir(directory)):
        if f.endswith(".json"):
            stem = os.path.splitext(f)[0]
            with open(os.path.join(directory, f)) as fh:
                out.append(ns.VLMC.from_json(fh.read(), ns.VLMC.strip_parameters_from_name(stem)))
    return out
