"""
ctypes binding of libqppfusion.so (include/qppfusion.h) + registration of its entry points as PyTorch
custom ops (torch.ops.qppfusion.*).

STRICT, like the reference ("No fallbacks", src/fusion.py:24, src/qpp.py:25): if the shared library is
missing, or there is no sm_100 CUDA device, every compute call raises -- nothing here falls back to a CPU
or eager-PyTorch implementation.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_PKG_DIR), "lib", "libqppfusion.so")

QF_OK = 0
_ERR_NAMES = {-1: "QF_E_INVALID", -2: "QF_E_UNSUPPORTED", -3: "QF_E_CUDA", -4: "QF_E_NO_DEVICE", -5: "QF_E_WORKSPACE"}

METHOD_SUM, METHOD_MNZ, METHOD_RRF = 0, 1, 2
NORM_NONE, NORM_MINMAX, NORM_ZSCORE = 0, 1, 2
SCORE_F64, SCORE_F32, SCORE_MICRO = 0, 1, 2
FLAG_MNZ_SKIP_ZERO, FLAG_REQUIRE_SORTED, FLAG_UNRANKED_OUTPUT = 1, 2, 4
FLAG_LARGE_UNION = 8              # performance hint: little overlap between the runs (larger hash table)
FLAG_DEBUG_FULLSORT = 0x100      # general kernel, 64-bit sort path (test hook)
FLAG_DEBUG_GENERAL = 0x200       # skip the gather / streaming fusion kernels (test hook)
FLAG_DEBUG_DEEP = 0x800         # ops.fuse: take qf_fuse_deep whatever the size (test hook; never reaches the library)
FLAG_GATHER = 0x400              # opt-in: deposit + gather fusion kernel instead of the streaming one (same results)
QSTAT_DUPLICATES, QSTAT_UNSORTED, QSTAT_BAD_LEN = 1, 2, 4
QPP_BRIDGE_COLS, QPP_RESEARCH_COLS = 13, 5
QPP_COLS = QPP_BRIDGE_COLS + QPP_RESEARCH_COLS
COLNORM_MINMAX, COLNORM_ZSCORE = 1, 2


class NativeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libqppfusion: {_ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


_lib: Optional[ctypes.CDLL] = None

_vp, _i, _d, _u32, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_uint32, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/qppfusion.h declares
PROTOTYPES = {
    "qf_last_error": (ctypes.c_char_p, []),
    "qf_abi_version": (_i, []),
    "qf_device_check": (_i, [_i]),
    "qf_profile_enable": (_i, [_i]),
    "qf_profile_read": (_i, [_i, ctypes.POINTER(ctypes.c_float)]),
    "qf_fuse_max_postings": (_i, []),
    "qf_fuse": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _d, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qf_fuse_deep_max_postings": (_i, []),
    "qf_fuse_deep_workspace_bytes": (_sz, [_i, _i, _i]),
    "qf_fuse_deep": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _d, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "qf_runs_parse": (_i, [_vp, _i, _i, _vp]),
    "qf_runs_free": (None, [_vp]),
    "qf_runs_info": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qf_runs_rows": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "qf_runs_dict": (_i, [_vp, _i, _vp, _vp]),
    "qf_pack_weighted": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "qf_runfile_write": (_i, [ctypes.c_char_p, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, ctypes.c_char_p, _i]),
    "qf_wire_words": (_sz, [ctypes.c_longlong, _i, _i]),
    "qf_wire_plan": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "qf_wire_pack": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, ctypes.c_int32, _vp, _sz, _i]),
    "qf_wire_unpack": (_i, [_vp, ctypes.c_longlong, _i, _i, ctypes.c_int32, _vp, _vp, _vp]),
    "qf_wire_for_blocks": (ctypes.c_longlong, [_i, _i, _i]),
    "qf_wire_for_pack": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "qf_wire_for_unpack": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "qf_compact_workspace_bytes": (_sz, [_i]),
    "qf_compact_fused": (_i, [_vp, _vp, _vp, _i, ctypes.c_longlong, _vp, _vp, _vp, _vp, _sz, _vp]),
    "qf_qpp_stats": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "qf_qpp_colnorm_workspace_bytes": (_sz, [_i, _i]),
    "qf_qpp_colnorm": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "qf_qpp_colstats": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "qf_qpp_colnorm_apply": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "qf_qpp_select_weight": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "qf_dense_workspace_bytes": (_sz, [_i, _i, _i]),
    "qf_dense_max_k": (_i, []),
    "qf_dense_topk": (_i, [_vp, _vp, _i, _i, _i, _i, ctypes.c_longlong, _vp, _vp, _vp, _vp, _sz, _vp]),
    "qf_dense_scores": (_i, [_vp, _vp, _i, _i, _i, _vp, ctypes.c_longlong, _vp]),
    "qf_dense_row_topk": (_i, [_vp, ctypes.c_longlong, _vp, _i, _i, _i, ctypes.c_longlong, _vp, _vp, _vp, _vp]),
    "qf_topk_merge": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "qf_dense_rescore_f32": (_i, [_vp, _vp, _i, ctypes.c_longlong, _i, _vp, _i, ctypes.c_longlong, _vp, _vp]),
    "qf_dense_topk_packed": (_i, [_vp, _vp, _i, _i, _i, _i, ctypes.c_longlong, _vp, _i, ctypes.c_longlong, _vp, _vp, _sz, _vp]),
    "qf_topk_merge_packed": (_i, [_vp, _i, _i, _i, ctypes.c_longlong, _i, _vp, _vp]),
    "qf_topk_merge_check": (_i, [_vp, _i, _i, _i, ctypes.c_longlong, _vp, _i, _vp, _vp]),
    "qf_linear_bf16": (_i, [_vp, _vp, _vp, _vp, ctypes.c_longlong, _i, _i, _i, _vp]),
    "qf_bert_weight_elems": (_sz, [_vp]),
    "qf_bert_workspace_bytes": (_sz, [_vp, _i, _i]),
    "qf_bert_encode": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "qf_bert_encode_packed": (_i, [_vp, _vp, _vp, _vp, _i, ctypes.c_longlong, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "qf_wordpiece_create": (_i, [ctypes.c_char_p, _sz, _i, _vp]),
    "qf_wordpiece_free": (None, [_vp]),
    "qf_wordpiece_encode": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
}

BERT_POOL_CLS, BERT_POOL_MEAN = 0, 1


class BertConfigC(ctypes.Structure):
    """qf_bert_config (include/qppfusion.h)."""
    _fields_ = [("n_layers", _i), ("hidden", _i), ("heads", _i), ("ffn", _i), ("vocab", _i), ("max_pos", _i),
                ("pooling", _i), ("ln_eps", ctypes.c_float)]


def lib() -> ctypes.CDLL:
    """Load libqppfusion.so (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"libqppfusion.so not found at {LIB_PATH}. Build it with "
                f"`python qpp-fusion-rag_b200/build_native.py` (or __graft_entry__.build()). "
                f"There is no CPU / PyTorch fallback for this path.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)      # AttributeError if the library is stale
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != QF_OK:
        raise NativeError(rc, lib().qf_last_error().decode("utf-8", "replace"))


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("qppfusion ops take CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor")
        if not t.is_contiguous():
            raise RuntimeError("qppfusion ops need contiguous tensors")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("qppfusion ops need all tensors on one device")
    if dev is None:
        raise RuntimeError("no tensor arguments")
    return dev


PROF_DENSE_SAMPLE, PROF_DENSE_TAU, PROF_DENSE_MAIN, PROF_DENSE_SELECT = 0, 1, 2, 3
PROF_FUSE_STREAM, PROF_FUSE_GENERAL, PROF_QPP_STATS = 4, 5, 6


def profile_enable(on: bool) -> None:
    check(lib().qf_profile_enable(1 if on else 0))


def profile_read(phase: int) -> float:
    ms = ctypes.c_float(0.0)
    check(lib().qf_profile_read(phase, ctypes.byref(ms)))
    return float(ms.value)
