#!/usr/bin/env python3
"""
Build libqppfusion.so (the C-ABI CUDA library, sm_100a only) in-tree with nvcc.

    python qpp-fusion-rag_b200/build_native.py [--force] [--verbose]

Output: qpp-fusion-rag_b200/lib/libqppfusion.so  (git-ignored; travels to the GPU box with the snapshot).
nvcc cross-compiles without a GPU.  Objects are rebuilt only when a source or header is newer.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libqppfusion.so")
INCLUDE = os.path.join(ROOT, "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-I", INCLUDE, "-I", CSRC] + os.environ.get("QF_EXTRA_CFLAGS", "").split()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h")]
    return hs


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    hdr_time = max([_mtime(h) for h in headers()] + [_mtime(__file__)])
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _mtime(obj) < max(_mtime(src), hdr_time):
            cmd = [NVCC] + ARCH + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for " + cmd[-3])
    if jobs or force or _mtime(LIB) < max(_mtime(o) for o in objs):
        # cuBLASLt (the encoder's Linear layers outside the own GEMM's envelope) is the one dynamic dependency; inside a PyTorch process the loader
        # binds it to the libcublasLt.so.12 torch has already loaded, otherwise to the toolkit's copy (rpath)
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC", "-lcublasLt",
                                                               "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
