#!/usr/bin/env python3
"""Build libmchrg_b200.so (CUDA C++ for sm_100a only) in-tree with nvcc.

    python -m multicharge_b200.build [--force] [--verbose]

The shared library is the product: a C-ABI (include/mchrg_b200.h) over hand-written kernels.
It is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmchrg_b200.so")
OBJDIR = os.path.join(HERE, "build")
SOURCES = ["context.cu", "comm.cu", "dense.cu", "eeq.cu", "eeq_pbc.cu", "eeqbc.cu", "batch.cu", "api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--fmad=true", "-Xptxas", "-warn-spills"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def build(force=False, verbose=False, defines=(), out=None, objdir=None):
    """defines / out / objdir: variant builds for A/B timing, e.g. build(defines=["LEAF_PROF"], out=".../libx.so",
    objdir=".../build_x"); the default build is the product."""
    global OUT, OBJDIR
    if out or objdir or defines:
        saved = (OUT, OBJDIR)
        OUT, OBJDIR = out or OUT, objdir or OBJDIR
        NVCC_FLAGS.extend("-D" + d for d in defines)
        try:
            return build(force=True, verbose=verbose)
        finally:
            OUT, OBJDIR = saved
            del NVCC_FLAGS[len(NVCC_FLAGS) - len(defines):]
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp", ".inc", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "mchrg_b200.h"))
    hdr_time = max(os.path.getmtime(h) for h in headers + [os.path.abspath(__file__)])
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    objs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJDIR, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            jobs.append([nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0 or (verbose and res.stderr.strip()):
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(OUT):
        run([nvcc, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-ldl"])
    return OUT


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    tag = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--tag=")), None)
    if tag:
        print(build(verbose="-v" in sys.argv, defines=defs, out=os.path.join(HERE, f"libmchrg_b200_{tag}.so"),
                    objdir=os.path.join(HERE, "build", tag)))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv))
