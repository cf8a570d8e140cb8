#!/usr/bin/env python
"""Build oracle/_ref: the reference's own hot-path sources, compiled where they lie.

TEST INFRASTRUCTURE (oracle).  For each configuration in oracle/configs.py this
compiles the nine hot-path translation units of /root/reference/Src *unmodified*
(mclib.c optical_depth.c mcrat_scattering.c electron.c hot_x_section.c
geometry.c photons.c mc_cyclosynch.c analytic_outflows.c) together with
oracle/ref_harness.c, oracle/gsl_shim/gsl_shim.c and oracle/mc_mathlib.c into
``oracle/_ref/libmcrat_ref_<cfg>.so``.

The reference cannot be built with its own Makefile here (GSL, HDF5 and MPI
are absent), so GSL is replaced by the shim in oracle/gsl_shim and MPI/HDF5 by
two stub headers.  No reference source is copied: a temporary directory of
symlinks to ``/root/reference/Src/*`` is created so that the per-configuration
``mcrat_input.h`` (generated, #define-only) shadows the reference's own, and is
removed afterwards.  Outputs go only to oracle/_ref/ (git-ignored).

Two variants per configuration:

* parity  (``libmcrat_ref_<cfg>.so``): ``gcc -O2 -std=gnu11 -fopenmp -ffp-contract=off`` and no ``-march``,
  i.e. plain x86-64 without FMA contraction, matching the reference's default Makefile arithmetic.  Every
  parity test and golden fixture uses this one.
* timing  (``libmcrat_ref_<cfg>_o3.so``, only for the configurations bench.py times): ``gcc -O3
  -march=x86-64-v3 -fopenmp`` -- AVX2 + FMA with contraction allowed, the fastest arithmetic the
  reference's own sources can be given.  BASELINE.md asks for ``-march=native``; the library is built in the
  build container and travels to the GPU box prebuilt (the reference sources do not exist there), so the
  instruction set has to be one every x86-64 server CPU of the last decade runs: x86-64-v3.  bench.py checks
  ``/proc/cpuinfo`` for avx2 / fma before loading it and says which build was timed.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from configs import CONFIGS, input_header  # noqa: E402

REF_SRC = os.environ.get("MCRAT_REFERENCE_SRC", "/root/reference/Src")
OUT = os.path.join(HERE, "_ref")
HOT_TUS = ["mclib.c", "optical_depth.c", "mcrat_scattering.c", "electron.c", "hot_x_section.c",
           "geometry.c", "photons.c", "mc_cyclosynch.c", "analytic_outflows.c"]
CFLAGS = ["-O2", "-std=gnu11", "-fopenmp", "-fPIC", "-w", "-fno-strict-aliasing", "-ffp-contract=off"]
CFLAGS_TIMING = ["-O3", "-march=x86-64-v3", "-std=gnu11", "-fopenmp", "-fPIC", "-w", "-fno-strict-aliasing"]
TIMING_CONFIGS = ("c2_2d_cyl_stokes", "c5_3d_sph")  # the workloads bench.py's CPU arm runs


def available():
    return os.path.isdir(REF_SRC)


def build_one(name, cfg, force=False, timing=False):
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libmcrat_ref_%s%s.so" % (name, "_o3" if timing else ""))
    deps = [os.path.join(HERE, "ref_harness.c"), os.path.join(HERE, "gsl_shim", "gsl_shim.c"),
            os.path.join(HERE, "mc_mathlib.c"), os.path.join(HERE, "mc_mathlib.h"),
            os.path.join(HERE, "gsl_shim", "gsl", "gsl_shim_all.h"), os.path.join(HERE, "configs.py"),
            os.path.abspath(__file__)]
    if not force and os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in deps):
        return lib
    farm = tempfile.mkdtemp(prefix="mcrat_ref_%s_" % name)
    try:
        for f in os.listdir(REF_SRC):
            if f == "mcrat_input.h":
                continue
            os.symlink(os.path.join(REF_SRC, f), os.path.join(farm, f))
        with open(os.path.join(farm, "mcrat_input.h"), "w") as fh:
            fh.write(input_header(cfg))
        os.symlink(os.path.join(HERE, "ref_harness.c"), os.path.join(farm, "ref_harness.c"))
        srcs = [os.path.join(farm, f) for f in HOT_TUS + ["ref_harness.c"]]
        srcs += [os.path.join(HERE, "gsl_shim", "gsl_shim.c"), os.path.join(HERE, "mc_mathlib.c")]
        cmd = ["gcc"] + (CFLAGS_TIMING if timing else CFLAGS) + ["-shared", "-o", lib, "-I", farm, "-I", os.path.join(HERE, "gsl_shim"),
                                  "-I", HERE] + srcs + ["-lm"]
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(farm, ignore_errors=True)
    return lib


def build_all(names=None, force=False, verbose=True):
    if not available():
        if verbose:
            print("build_ref: %s not present; keeping prebuilt oracle/_ref" % REF_SRC)
        return []
    libs = []
    for name, cfg in CONFIGS.items():
        if names and name not in names:
            continue
        libs.append(build_one(name, cfg, force=force))
        if verbose:
            print("build_ref: %s" % libs[-1])
        if name in TIMING_CONFIGS:
            libs.append(build_one(name, cfg, force=force, timing=True))
            if verbose:
                print("build_ref: %s" % libs[-1])
    return libs


if __name__ == "__main__":
    build_all(names=[a for a in sys.argv[1:] if not a.startswith("--")] or None, force="--force" in sys.argv)
