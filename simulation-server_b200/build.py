"""Builds libbcs.so (hand-written CUDA for sm_100a) in-tree with nvcc; no torch involved.

    python simulation-server_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libbcs.so")

CUDA_SOURCES = ["grid.cu", "cellpass.cu", "collide.cu", "pairs.cu", "vein.cu", "wall.cu", "integrate.cu", "slab.cu", "capi.cu"]
HOST_SOURCES = ["scene_host.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
              "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))]
    headers.append(os.path.join(ROOT, "include", "bcs.h"))
    nvcc = _nvcc()
    jobs = []
    for src in CUDA_SOURCES:
        s, o = os.path.join(CSRC, src), os.path.join(OBJ, src + ".o")
        if force or not _newer(o, [s] + headers):
            jobs.append([nvcc, *ARCH, *NVCC_FLAGS, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", s, "-o", o])
    for src in HOST_SOURCES:
        s, o = os.path.join(CSRC, src), os.path.join(OBJ, src + ".o")
        if force or not _newer(o, [s] + headers):
            # host tables are computed like the reference's host/constexpr code: no FMA contraction
            jobs.append([nvcc, *ARCH, *NVCC_FLAGS, "-Xcompiler", "-ffp-contract=off", "-x", "cu", "-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("compile failed: " + " ".join(cmd))
        return r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, s + ".o") for s in CUDA_SOURCES + HOST_SOURCES]
    if force or jobs or not _newer(LIB, objs):
        run([nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC", "-lnccl"])
    # headless C++ driver over the host mirror of the reference loop (host/bcs_host.hpp)
    host = os.path.join(HERE, "host")
    exe = os.path.join(HERE, "bcs_headless")
    srcs = [os.path.join(host, "headless.cpp"), os.path.join(host, "bcs_host.hpp"), os.path.join(ROOT, "include", "bcsd_io.hpp"), LIB]
    if force or not _newer(exe, srcs):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        run([gxx, "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + host, srcs[0], "-L" + HERE, "-l:libbcs.so",
             "-Wl,-rpath,$ORIGIN", "-o", exe])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
