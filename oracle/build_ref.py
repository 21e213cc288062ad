#!/usr/bin/env python
"""Compile the UNMODIFIED reference hot path into ``oracle/_ref/`` (test infrastructure).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this; only ``tests/``,
``__graft_entry__.build()/smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may use what it produces.

What it does
------------
The reference (``/root/reference/src``) is Cython.  Its own ``src/setup.py:5-19`` lists the
modules and ``cythonize``s them in place; that cannot run here because ``/root/reference`` is
read-only and because the sources rely on Cython-0.2x's default ``language_level=2``
(implicit relative imports, e.g. ``src/clustering/graph_based_clustering.pyx:6-7``).

This recipe therefore cythonizes the sources *where they lie* (no source is copied into the
repository), with the generated C and object files in a scratch directory under ``/tmp`` and only
the resulting extension modules (``*.so``) written to ``oracle/_ref/<package>/``.
``oracle/_ref/`` is git-ignored but travels to the GPU box with the ``gpurun`` snapshot, so the
compiled reference can be timed there as the CPU baseline (``cpu_baseline.kind = "reference"``).

Only the modules on the hot path and its direct consumer are built:
``vlmc.vlmc``, the four ``distance`` classes, ``clustering.util``, and the average-link /
MST consumers used for cluster-assignment identity checks.

Usage:  python oracle/build_ref.py [--ref /root/reference] [--force]
"""
import argparse
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

# (dotted module name, path under <ref>/src) -- subset of src/setup.py:5-15
MODULES = [
    ("vlmc.vlmc", "vlmc/vlmc.pyx"),
    ("distance.negloglikelihood", "distance/negloglikelihood.pyx"),
    ("distance.frobenius", "distance/frobenius.pyx"),
    ("distance.projection", "distance/projection.pyx"),
    ("distance.estimate", "distance/estimate.pyx"),
    ("distance.acgt", "distance/acgt.pyx"),
    ("clustering.util", "clustering/util.pyx"),
    ("clustering.clustering_metrics", "clustering/clustering_metrics.pyx"),
    ("clustering.graph_based_clustering", "clustering/graph_based_clustering.pyx"),
    ("clustering.average_link_clustering", "clustering/average_link_clustering.pyx"),
    ("clustering.mst_clustering", "clustering/mst_clustering.pyx"),
]


def stamp_path():
    return os.path.join(OUT, "BUILD_OK")


def is_built():
    if not os.path.exists(stamp_path()):
        return False
    for mod, _ in MODULES:
        pkg, leaf = mod.split(".")
        d = os.path.join(OUT, pkg)
        if not os.path.isdir(d) or not any(f.startswith(leaf + ".") and f.endswith(".so") for f in os.listdir(d)):
            return False
    return True


def build(ref="/root/reference", force=False, quiet=True):
    """Build oracle/_ref from <ref>/src.  Returns True when oracle/_ref is usable."""
    src = os.path.join(ref, "src")
    if is_built() and not force:
        return True
    if not os.path.isdir(src):
        return is_built()  # e.g. on the GPU box: only the prebuilt files exist

    import numpy
    from Cython.Build import cythonize
    from setuptools import Extension
    from setuptools.dist import Distribution

    scratch = tempfile.mkdtemp(prefix="vgs_refbuild_")
    exts = [
        Extension(mod, [os.path.join(src, rel)], include_dirs=[numpy.get_include()],
                  define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")])
        for mod, rel in MODULES
    ]
    cwd = os.getcwd()
    try:
        # cimports such as "from graph_based_clustering cimport ..." resolve against the
        # package directories themselves (average_link_clustering.pxd:5).
        ext_modules = cythonize(
            exts,
            include_path=[numpy.get_include(), os.path.join(src, "clustering"), os.path.join(src, "distance")],
            compiler_directives={"language_level": 2},
            build_dir=os.path.join(scratch, "c"),
            quiet=quiet,
        )
        dist = Distribution({"name": "vgs_reference_oracle", "ext_modules": ext_modules})
        cmd = dist.get_command_obj("build_ext")
        cmd.build_lib = os.path.join(scratch, "lib")
        cmd.build_temp = os.path.join(scratch, "tmp")
        cmd.inplace = 0
        cmd.ensure_finalized()
        if quiet:
            from distutils import log  # noqa: F401  (setuptools shim)
        cmd.run()
        if os.path.isdir(OUT):
            shutil.rmtree(OUT)
        os.makedirs(OUT)
        for root, _, files in os.walk(cmd.build_lib):
            for f in files:
                if f.endswith(".so"):
                    rel = os.path.relpath(os.path.join(root, f), cmd.build_lib)
                    dst = os.path.join(OUT, rel)
                    os.makedirs(os.path.dirname(dst), exist_ok=True)
                    shutil.copy2(os.path.join(root, f), dst)
        with open(stamp_path(), "w") as fh:
            fh.write("built from %s with Cython language_level=2\n" % src)
    finally:
        os.chdir(cwd)
        shutil.rmtree(scratch, ignore_errors=True)
    return is_built()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    ok = build(a.ref, a.force, quiet=not a.verbose)
    print("oracle/_ref:", "ok" if ok else "UNAVAILABLE")
    sys.exit(0 if ok else 1)
