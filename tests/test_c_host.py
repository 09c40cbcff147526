"""The C ABI from a plain C host (examples/c_host_fuse.c): the header is C99-clean, the example links against
libqppfusion.so with gcc alone, fails loudly without a GPU, and on a GPU prints exactly the oracle's fused lists."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "qpp-fusion-rag_b200", "lib")
CUDA_LIB = "/usr/local/cuda/lib64"


def _build(tmp_path):
    if shutil.which("gcc") is None or not os.path.isdir(CUDA_LIB):
        pytest.skip("gcc or the CUDA runtime library is not available")
    if not os.path.isfile(os.path.join(LIBDIR, "libqppfusion.so")):
        import __graft_entry__
        __graft_entry__.build()
    exe = str(tmp_path / "c_host_fuse")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_host_fuse.c"), "-o", exe, "-L", LIBDIR, "-lqppfusion", "-L", CUDA_LIB, "-lcudart",
           f"-Wl,-rpath,{LIBDIR}", f"-Wl,-rpath,{CUDA_LIB}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c99():
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "qppfusion.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_host_links_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_host_output_equals_the_oracle(tmp_path):
    from oracle import fusion_oracle as fo
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    ids = [[[7, 3, 9, 1], [3, 5, 7]], [[2, 4], [4, 2, 8, 6]]]
    micro = [[[900000, 800000, 500000, 100000], [700000, 600000, 200000]], [[1000000, 250000], [750000, 750000, 500000, 125000]]]
    w = [[0.5, 0.25], [1.0, 0.5]]
    want = []
    for q in range(2):
        lists = [[(d, m / 1e6, i + 1) for i, (d, m) in enumerate(zip(ids[q][rr], micro[q][rr]))] for rr in range(2)]
        docs, scores, _ = fo.fuse_weighted_query(lists, w[q], fo.SUM)
        order = fo.rank_order(scores)
        want.append(f"q{q}:" + "".join(f" {docs[i]}:{scores[i] + 0.0:.17g}" for i in order))
    assert r.stdout.splitlines() == want
