"""Import shim: makes the directory ``clustering-genomic-signatures_b200/`` importable.

The package directory carries the reference's repository name, which contains hyphens and is
therefore not a Python identifier.  This module gives it the importable name
``clustering_genomic_signatures_b200`` by pointing ``__path__`` at that directory and executing
its ``__init__.py`` in this module's namespace, so that

    from clustering_genomic_signatures_b200.distance import NegativeLogLikelihood

works from the repository root (``tests/conftest.py`` and ``bench.py`` put the root on sys.path).
"""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "clustering-genomic-signatures_b200")
__path__ = [_pkg_dir]
__file__ = _os.path.join(_pkg_dir, "__init__.py")
with open(__file__) as _fh:
    exec(compile(_fh.read(), __file__, "exec"))
del _fh
