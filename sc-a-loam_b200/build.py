"""Builds the in-tree native libraries of the scan-to-map engine.

  csrc/libs2m.so            the product: sm_100a kernels + C ABI (include/s2m.h)
  csrc/libs2m_hostmath.so   host build of csrc/s2m_math.cuh for CPU unit tests
nvcc cross-compiles for sm_100a without a GPU; the .so files are git-ignored and
travel to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libs2m.so")
HOSTMATH = os.path.join(CSRC, "libs2m_hostmath.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-O3"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.exists(c) or c == "nvcc"):
            return c
    return "nvcc"


def build_cuda(force=False, verbose=False):
    cu = [os.path.join(CSRC, f) for f in ("s2m_kernels.cu", "s2m_api.cu", "s2m_fx.cu")]
    deps = cu + [os.path.join(CSRC, f) for f in ("s2m_math.cuh", "s2m_internal.h")] + \
        [os.path.join(HERE, "..", "include", "s2m.h")]
    objs = []
    procs = []
    for src in cu:
        obj = src[:-3] + ".o"
        objs.append(obj)
        if force or _newer(obj, deps):
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if force or procs or _newer(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
        subprocess.check_call(cmd)
    return LIB


def build_hostmath(force=False):
    src = os.path.join(CSRC, "s2m_hostmath.cpp")
    deps = [src, os.path.join(CSRC, "s2m_math.cuh"), os.path.join(CSRC, "s2m_internal_host.h")]
    if force or _newer(HOSTMATH, deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                               src, "-o", HOSTMATH])
    return HOSTMATH


if __name__ == "__main__":
    build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_hostmath(force="--force" in sys.argv)
    print("built", LIB, HOSTMATH)
