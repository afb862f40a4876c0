"""In-tree build of the CUDA library (sm_100a only) and of the C host mirror.

`python -m bpp_b200.build` or `__graft_entry__.build()`.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbppgpu.so")
HOST_LIB = os.path.join(HERE, "libbpphost.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--cudart", "static"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(ROOT, "include", "bpp_b200.h"))
    if not force and not _newer(LIB, srcs):
        return LIB
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB, os.path.join(CSRC, "engine.cu")]
    subprocess.check_call(cmd)
    return LIB


def build_host(force=False):
    """The C host-side mirror of the reference's locus seam (bpp_b200/host/*.c), linked against
    the CUDA library."""
    hdir = os.path.join(HERE, "host")
    if not os.path.isdir(hdir):
        return None
    srcs = [os.path.join(hdir, f) for f in sorted(os.listdir(hdir)) if f.endswith(".c")]
    if not srcs:
        return None
    deps = srcs + [os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith(".h")] + [LIB]
    if not force and not _newer(HOST_LIB, deps):
        return HOST_LIB
    cmd = ["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "include"),
           "-o", HOST_LIB] + srcs + ["-L", HERE, "-lbppgpu", "-Wl,-rpath,$ORIGIN", "-lm"]
    subprocess.check_call(cmd)
    return HOST_LIB


if __name__ == "__main__":
    build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_host(force="--force" in sys.argv)
    print("built", LIB)
