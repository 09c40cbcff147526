"""
TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/qf_oracle.c (the C restatement of the reference's weighted
fusion and QPP-bridge arithmetic).  Used for full-size parity checks and bench.py's CPU legs; never by the product.
"""
import ctypes
import os

import numpy as np

from .build_oracle import LIB, build

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            build()
        l = ctypes.CDLL(LIB)
        l.qfo_fuse_weighted.restype = ctypes.c_int
        l.qfo_qpp_bridge_batch.restype = ctypes.c_int
        _lib = l
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def fuse_weighted(ids, scores, lens, weights, method, rrf_k=60.0, ranks=None, n_threads=1):
    """ids int32 [Q,R,K], scores int32 micro or float64 [Q,R,K], lens int32 [Q,R], weights float64 [Q,R]
    -> (out_ids int32 [Q,R*K], out_scores f64, out_counts int32, out_pos int32, out_lens int32 [Q]), ranked."""
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    Q, R, K = ids.shape
    micro = np.ascontiguousarray(scores) if scores.dtype == np.int32 else None
    f64 = np.ascontiguousarray(scores, dtype=np.float64) if micro is None else None
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    ranks = None if ranks is None else np.ascontiguousarray(ranks, dtype=np.int32)
    n = R * K
    out_ids = np.full((Q, n), -1, dtype=np.int32)
    out_scores = np.zeros((Q, n), dtype=np.float64)
    out_counts = np.zeros((Q, n), dtype=np.int32)
    out_pos = np.zeros((Q, n), dtype=np.int32)
    out_lens = np.zeros((Q,), dtype=np.int32)
    lib().qfo_fuse_weighted(_p(ids), _p(micro), _p(f64), _p(ranks), _p(lens), _p(weights), Q, R, K, int(method),
                            ctypes.c_double(rrf_k), _p(out_ids), _p(out_scores), _p(out_counts), _p(out_pos), _p(out_lens),
                            int(n_threads))
    return out_ids, out_scores, out_counts, out_pos, out_lens


def qpp_bridge_batch(micro, lens, feats, top_k=100, k_cut=50, n_threads=1):
    """micro int32 [Q,R,K] (sorted desc), lens int32 [Q,R], feats int32 [Q,4] -> float64 [Q,R,13]."""
    micro = np.ascontiguousarray(micro, dtype=np.int32)
    Q, R, K = micro.shape
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    feats = np.ascontiguousarray(feats, dtype=np.int32)
    out = np.zeros((Q, R, 13), dtype=np.float64)
    lib().qfo_qpp_bridge_batch(_p(micro), _p(lens), _p(feats), Q, R, K, int(top_k), int(k_cut), _p(out), int(n_threads))
    return out
