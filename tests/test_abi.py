"""CPU-side checks of the C ABI: the library loads and exports every symbol include/qppfusion.h declares;
without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qppfusion.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"QF_API[^;(]*?\b(qf_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ["qf_last_error", "qf_abi_version", "qf_device_check", "qf_fuse", "qf_fuse_max_postings",
              "qf_qpp_stats", "qf_qpp_colnorm", "qf_qpp_colnorm_workspace_bytes", "qf_qpp_select_weight"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    from qppfusion import _native as N
    if not os.path.isfile(N.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(N.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in qppfusion.h but not exported"
    assert set(declared_symbols()) == set(N.PROTOTYPES.keys())
    assert N.lib().qf_abi_version() == 1
    assert N.lib().qf_fuse_max_postings() >= 5000


def test_no_cpu_fallback():
    import torch
    from qppfusion import _native as N, ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    # C ABI level: argument validation works, compute fails with QF_E_NO_DEVICE
    rc = N.lib().qf_device_check(0)
    assert rc == -4 and b"no CPU fallback" in N.lib().qf_last_error()
    one = ctypes.c_void_p(1)   # never dereferenced: device check fails first
    rc = N.lib().qf_fuse(one, one, 0, None, one, None, 1, 1, 1, 0, 0, 60.0, 0, one, one, None, None, one, None, None)
    assert rc == -4
    rc = N.lib().qf_fuse(one, one, 0, None, one, None, 1, 5, 2000, 0, 0, 60.0, 0, one, one, None, None, one, None, None)
    assert rc == -2 and b"envelope" in N.lib().qf_last_error()
    # torch level: CPU tensors are rejected, nothing is computed on the host
    ids = torch.zeros((1, 1, 4), dtype=torch.int32)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.fuse(ids, torch.zeros((1, 1, 4), dtype=torch.float64), torch.ones((1, 1), dtype=torch.int32))
