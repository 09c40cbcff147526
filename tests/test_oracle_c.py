"""The C oracle (oracle/qf_oracle.c) agrees with the Python oracles that are pinned on the reference's goldens."""
import numpy as np
import pytest

from helpers import pack_lists, synth_lists
from oracle import c_oracle as co
from oracle import fusion_oracle as fo
from oracle import qpp_oracle as qo


@pytest.mark.parametrize("method", [fo.SUM, fo.MNZ, fo.RRF])
@pytest.mark.parametrize("score_type", ["micro", "f64"])
def test_c_fusion_equals_python_oracle(method, score_type):
    lists = synth_lists(31, n_q=25, n_r=5, k=80, pool=190, ragged=True, neg_ranker=2, ties=True)
    d0 = lists[3][1][2][0]
    lists[3][1][7] = (d0, lists[3][1][7][1], lists[3][1][7][2])          # duplicate doc id inside one list
    ids, sc, ranks, lens = pack_lists(lists, k=80, score_type=score_type)
    w = np.round(np.random.default_rng(32).random((25, 5)), 6)
    oid, osc, ocnt, opos, olen = co.fuse_weighted(ids, sc, lens, w, method, 60.0, ranks=ranks, n_threads=4)
    for q, per_r in enumerate(lists):
        docs, scores, counts = fo.fuse_weighted_query(per_r, w[q], method, 60)
        order = fo.rank_order(scores)
        n = len(docs)
        assert olen[q] == n
        assert oid[q, :n].tolist() == [docs[i] for i in order]
        assert np.array_equal(osc[q, :n], np.array([scores[i] for i in order]) + 0.0)
        assert ocnt[q, :n].tolist() == [counts[i] for i in order]
        assert opos[q, :n].tolist() == order


def test_c_qpp_bridge_equals_python_oracle():
    rng = np.random.default_rng(33)
    Q, R, K = 30, 5, 120
    texts = ["what is machine learning", "  lead", "", "x y z z", "single"] * 6
    feats = np.array([qo.text_features(t) for t in texts], dtype=np.int32)
    micro = np.zeros((Q, R, K), dtype=np.int32)
    lens = np.zeros((Q, R), dtype=np.int32)
    for q in range(Q):
        for r in range(R):
            n = int(rng.integers(0, K + 1))
            v = np.sort(rng.random(n) * [1, 30, 1, 1, 12][r])[::-1]
            if r == 2:
                v = np.sort(-rng.exponential(3.0, n))[::-1]
            micro[q, r, :n] = np.round(v * 1e6).astype(np.int32)
            lens[q, r] = n
    got = co.qpp_bridge_batch(micro, lens, feats, top_k=100, k_cut=50, n_threads=4)
    for q in range(Q):
        for r in range(R):
            lst = [m / 1e6 for m in micro[q, r, :lens[q, r]]]
            exp = np.array(qo.bridge_scores(texts[q], lst[:100]))
            ok = (np.isnan(exp) & np.isnan(got[q, r])) | (np.abs(got[q, r] - exp) <= 1e-12 + 1e-10 * np.abs(exp))
            assert ok.all(), (q, r, got[q, r], exp)


def test_division_free_micro_to_double_is_exact_for_every_int32():
    """The CUDA kernels turn micro-units into the double float("x.xxxxxx") parses to with a multiply and two FMAs;
    exhaustive proof over the whole int32 range (about 6 s on 8 cores)."""
    import ctypes
    import os
    l = co.lib()
    l.qfo_check_micro_division.restype = ctypes.c_longlong
    l.qfo_check_micro_division.argtypes = [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int]
    assert l.qfo_check_micro_division(-2 ** 31, 2 ** 31, os.cpu_count() or 1) == 0
    for v in (0, 1, -1, 123456, 999999, -2147483648, 2147483647):
        assert float(f"{v / 1e6:.6f}") == v / 1e6
