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


def _src_hash(sources, extra=()):
    import hashlib
    h = hashlib.sha256()
    for s in sorted(sources):
        h.update(os.path.basename(s).encode())
        with open(s, "rb") as f:
            h.update(f.read())
    for e in extra:
        h.update(str(e).encode())
    return h.hexdigest()


def _stale(target, sources, extra=()):
    """A built library is current only if the hash of the sources it was built from (kept next to it in
    <target>.srchash) equals the hash of the sources present now: modification times do not survive the copy to
    the GPU box, a hash does -- there the driver's build() compiles exactly when the tree differs from what the
    shipped .so was built from."""
    if not os.path.exists(target) or not os.path.exists(target + ".srchash"):
        return True
    with open(target + ".srchash") as f:
        return f.read().strip() != _src_hash(sources, extra)


def _stamp(target, sources, extra=()):
    with open(target + ".srchash", "w") as f:
        f.write(_src_hash(sources, extra) + "\n")


def build_cuda(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(ROOT, "include", "bpp_b200.h"))
    if not force and not _stale(LIB, srcs, NVCC_FLAGS):
        return LIB
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB, os.path.join(CSRC, "engine.cu")]
    subprocess.check_call(cmd)
    _stamp(LIB, srcs, NVCC_FLAGS)
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
    deps = srcs + [os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith(".h")] + \
        [os.path.join(ROOT, "include", "bpp_b200.h")]
    if not force and not _stale(HOST_LIB, deps):
        return HOST_LIB
    cmd = ["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "include"),
           "-o", HOST_LIB] + srcs + ["-L", HERE, "-lbppgpu", "-Wl,-rpath,$ORIGIN", "-lm"]
    subprocess.check_call(cmd)
    _stamp(HOST_LIB, deps)
    return HOST_LIB


if __name__ == "__main__":
    build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_host(force="--force" in sys.argv)
    print("built", LIB)
