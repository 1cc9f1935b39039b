"""Build libpderead_b200.so in-tree with nvcc for sm_100a.

    python -m pde_read_b200.build [--force] [--verbose]

Every (activation, NX, NT) instantiation of the MLP-jet kernels is its own object file so that
the 39 instantiations compile in parallel; objects are rebuilt only when a source or header is
newer.  No GPU is needed to build (nvcc cross-compiles).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# tuning experiments: PDR_VARIANT=<tag> PDR_EXTRA_CFLAGS="-DX=1 .." builds libpderead_b200_<tag>.so beside the
# product library (load it with PDEREAD_LIB=<path>); the product build uses neither variable.
VARIANT = os.environ.get("PDR_VARIANT", "")
OBJDIR = os.path.join(CSRC, "_build" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(HERE, "libpderead_b200%s.so" % ("_" + VARIANT if VARIANT else ""))
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-I", INCLUDE] + \
    os.environ.get("PDR_EXTRA_CFLAGS", "").split()

# (NX, NT) combinations the kernels are instantiated for; backward only where it is needed:
# plain MLP (0,0) and every combination with a time series (Evaluate_Derivatives always has m >= 1).
FWD_COMBOS = [(0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (1, 1), (2, 1), (3, 1), (4, 1), (1, 2), (2, 2), (3, 2), (4, 2)]
BWD_COMBOS = {(0, 0), (1, 1), (2, 1), (3, 1), (4, 1), (1, 2), (2, 2), (3, 2), (4, 2)}
HEADERS = ["internal.h", "jet_math.cuh", "mlp_jet_kernels.cuh", "mlp_plain_kernels.cuh", "tc_jet_kernels.cuh", "umma.cuh", os.path.join(INCLUDE, "pderead_b200.h")]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), p.stdout))
    return p.stdout


def build(force: bool = False, verbose: bool = False, jobs: int | None = None) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    jobs_list = []
    objs = []
    for act in (0, 1, 2):
        for nx, nt in FWD_COMBOS:
            obj = os.path.join(OBJDIR, "inst_a%d_x%d_t%d.o" % (act, nx, nt))
            objs.append(obj)
            src = os.path.join(CSRC, "inst.cu")
            if force or _newer(obj, [src] + hdrs):
                bwd = 1 if (nx, nt) in BWD_COMBOS else 0
                jobs_list.append([NVCC] + CFLAGS + ["-DPDR_ACT=%d" % act, "-DPDR_NX=%d" % nx, "-DPDR_NT=%d" % nt,
                                                    "-DPDR_BWD=%d" % bwd, "-c", src, "-o", obj])
    for name in ("api", "extraction", "peer"):
        obj = os.path.join(OBJDIR, name + ".o")
        objs.append(obj)
        src = os.path.join(CSRC, name + ".cu")
        if force or _newer(obj, [src] + hdrs):
            jobs_list.append([NVCC] + CFLAGS + ["-c", src, "-o", obj])
    if jobs_list:
        with ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
            list(ex.map(lambda c: _run(c, verbose), jobs_list))
    if force or jobs_list or _newer(LIB, objs):
        _run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs, verbose)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
