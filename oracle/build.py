#!/usr/bin/env python3
"""Build the CPU oracle (TEST INFRASTRUCTURE) into oracle/libmchrg_oracle.so.

gcc + OpenMP, linked against the LAPACK/BLAS the reference calls (dsytrf/dsytri/dsytrs/dsymv/
dgemv/dgemm) as shipped in scipy's bundled OpenBLAS (symbols carry a ``scipy_`` prefix).
The reference itself is Fortran with un-vendored dependencies (mctc-lib v0.5.2, mstore v0.3.0)
and this image has no Fortran compiler, so there is no oracle/_ref build -- see DESIGN.md.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCES = ["mchrg_oracle.c", "mchrg_oracle_eeqbc.c"]
OUT = os.path.join(HERE, "libmchrg_oracle.so")


def find_openblas():
    import scipy

    libs = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    cands = sorted(glob.glob(os.path.join(libs, "libscipy_openblas*.so*")))
    if not cands:
        raise RuntimeError("scipy's bundled OpenBLAS not found under " + libs)
    return cands[0]


def build(force=False, verbose=False):
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    deps = srcs + [os.path.join(HERE, "mchrg_oracle.h"), os.path.join(HERE, "param_tables_oracle.inc"),
                   os.path.abspath(__file__)]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    blas = find_openblas()
    cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-fno-fast-math", "-ffp-contract=off",
           "-Wall", "-Wextra", "-Wno-unused-parameter",
           "-o", OUT] + srcs + [blas, "-Wl,-rpath," + os.path.dirname(blas), "-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
