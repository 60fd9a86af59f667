"""Builds libsylver_b200.so (host C++ + sm_100a CUDA) in-tree with nvcc.

The shared library is the product: a C-ABI (include/sylver_b200.h) with no
torch types.  Objects land in sylver_b200/_build/, the library next to this
file so that it travels to the GPU box with the source snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# A/B builds of kernel variants: SYLVER_B200_VARIANT=<tag> SYLVER_B200_DEFS="-DX=1 ..." builds
# libsylver_b200_<tag>.so beside the product library (loaded with SYLVER_B200_LIB=<path>)
VARIANT = os.environ.get("SYLVER_B200_VARIANT", "")
OBJ = os.path.join(HERE, "_build" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(HERE, "libsylver_b200" + ("_" + VARIANT if VARIANT else "") + ".so")

SOURCES = ["api.cpp", "analyse.cpp", "scaling.cpp", "clean.cpp", "ordering.cpp", "partition.cpp", "comm.cpp", "engine.cu", "engine_indef.cu", "aux.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# METIS 5 (idx_t = int64) ships with the CUDA toolkit as a static library; options.ordering = 1
# is available when it is found (csrc/ordering.cpp), otherwise that option returns flag -98
METIS_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(NVCC))), "targets", "x86_64-linux", "lib",
                         "libmetis_static.a")
HAVE_METIS = os.path.exists(METIS_LIB)
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
          "-Xptxas", "-v" if os.environ.get("SYLVER_PTXAS_V") else "-O3"] + (["-DSYLVER_HAVE_METIS"] if HAVE_METIS else []) \
    + os.environ.get("SYLVER_B200_DEFS", "").split()


def _deps(src: str):
    d = [os.path.join(CSRC, src)]
    for f in os.listdir(CSRC):
        if f.endswith((".hpp", ".cuh", ".h")):
            d.append(os.path.join(CSRC, f))
    d.append(os.path.join(HERE, "..", "include", "sylver_b200.h"))
    return d


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(src: str) -> str:
    obj = os.path.join(OBJ, src.rsplit(".", 1)[0] + ".o")
    if _stale(obj, _deps(src)):
        cmd = [NVCC, *ARCH, *CFLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if os.environ.get("SYLVER_PTXAS_V"):
            sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(_compile, srcs))
    if _stale(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, *([METIS_LIB] if HAVE_METIS else []),
               "-lcudart_static", "-lpthread", "-ldl", "-lrt", "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
