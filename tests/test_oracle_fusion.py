"""Oracle (oracle/fusion_oracle.py) vs golden vectors produced by the REAL reference src/fusion.py."""
import gzip
import hashlib
import json
import os

import pytest

from oracle import fusion_oracle as fo

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fusion_weighted.json.gz")


def load_gold():
    with gzip.open(GOLD, "rt") as f:
        return json.load(f)


def rows_to_runs(case, rankers):
    runs = {}
    for ranker in rankers:
        per_q = {}
        for (qid, _it, docno, rank, score, _tag) in case["rows"][ranker]:
            per_q.setdefault(qid, []).append((docno, float(score), int(rank)))
        runs[ranker] = per_q
    return runs


METHODS = {"wcombsum": (fo.SUM, 60), "wcombmnz": (fo.MNZ, 60), "wrrf": (fo.RRF, 60), "wrrf_k10": (fo.RRF, 10)}


@pytest.mark.parametrize("case_idx", [0, 1, 2])
def test_weighted_matches_reference_bit_exact(case_idx):
    gold = load_gold()
    case = gold["cases"][case_idx]
    runs = rows_to_runs(case, gold["rankers"])
    for key, res in case["results"].items():
        mname, idx = key.split("@")
        method, k = METHODS[mname]
        fused = fo.weighted_fusion(runs, case["qpp"], method, int(idx), k)
        ref = {q: [(d, float(s)) for d, s in lst] for q, lst in res["fused"].items()}
        assert list(fused.keys()) == list(ref.keys()), key
        for q in ref:
            assert fused[q] == ref[q], (case["name"], key, q)   # order, doc ids and float64 bits
        lines = fo.runfile_lines(fused, tag=f"{mname}-x")
        assert len(lines) == res["runfile_n"]
        assert lines[:25] == res["runfile_head"]
        assert hashlib.sha256("\n".join(lines).encode()).hexdigest() == res["runfile_sha256"]


def test_missing_weight_raises_keyerror_even_without_rows():
    runs = {"A": {"1": [("d1", 0.5, 1)]}, "B": {}}
    qpp = {"1": {"A": [0.1] * 13}}
    with pytest.raises(KeyError):
        fo.weighted_fusion(runs, qpp, fo.SUM, 5)
    with pytest.raises(IndexError):
        fo.weighted_fusion({"A": runs["A"]}, {"1": {"A": [0.1] * 3}}, fo.SUM, 5)


def test_ranx_restatement_formulas():
    # hand-computed: two runs, min-max per run, CombSUM / CombMNZ / RRF
    a = [("x", 3.0, 1), ("y", 2.0, 2), ("z", 1.0, 3)]
    b = [("y", 10.0, 1), ("w", 5.0, 2)]
    docs, s, c = fo.fuse_ranx_query([a, b], fo.SUM)
    assert docs == ["x", "y", "z", "w"]
    assert s == [1.0, 1.5, 0.0, 0.0] and c == [1, 2, 1, 1]
    docs, s, c = fo.fuse_ranx_query([a, b], fo.MNZ)
    assert s == [1.0, 3.0, 0.0, 0.0]
    docs, s, c = fo.fuse_ranx_query([a, b], fo.MNZ, mnz_skip_zero=True)
    assert c == [1, 2, 0, 0] and s == [1.0, 3.0, 0.0, 0.0]
    docs, s, c = fo.fuse_ranx_query([a, b], fo.RRF, rrf_k=60)
    assert s == [1 / 61, 1 / 62 + 1 / 61, 1 / 63, 1 / 62]
    # duplicate docno: first position, last score (python dict semantics of _df_runs_to_ranx)
    docs, s, c = fo.fuse_ranx_query([[("x", 1.0, 1), ("y", 0.5, 2), ("x", 0.0, 3)]], fo.SUM)
    assert docs == ["y", "x"] and s == [1.0, 0.0]
    # constant list: denominator guard 1e-9 -> all zeros
    docs, s, c = fo.fuse_ranx_query([[("x", 0.7, 1), ("y", 0.7, 2)]], fo.SUM)
    assert s == [0.0, 0.0]


def test_rank_order_is_stable_descending():
    assert fo.rank_order([0.5, 1.0, 0.5, 1.0, -0.0, 0.0]) == [1, 3, 0, 2, 4, 5]


GOLD_LEARNED = os.path.join(os.path.dirname(__file__), "golden", "fusion_learned.json.gz")


def test_learned_fusion_matches_the_real_reference():
    """oracle learned_fusion vs the outputs of the REAL reference learned_fusion (tools/make_golden_learned.py):
    same qids, same docno order, scores within 2 ulp (pandas sums with Kahan compensation, the oracle sequentially)."""
    import numpy as np
    from helpers import LinearWeightModel
    with gzip.open(GOLD_LEARNED, "rt") as f:
        gold = json.load(f)
    for case in gold["cases"]:
        runs = rows_to_runs(case, gold["rankers"])
        mr = case["model_rankers"]
        fused = fo.learned_fusion(runs, case["qpp"], LinearWeightModel(13 * len(mr), len(mr), case["model_seed"]), mr)
        ref = case["fused"]
        assert sorted(fused.keys()) == sorted(ref.keys())
        for q in ref:
            assert [d for d, _ in fused[q]] == [d for d, _ in ref[q]], (case["name"], q)
            a = np.array([s for _, s in fused[q]])
            b = np.array([float(s) for _, s in ref[q]])
            assert np.all(np.abs(a - b) <= 2 * np.spacing(np.abs(b))), (case["name"], q)


def test_zscore_norm_restatement():
    """QF_NORM_ZSCORE contract: sequential sums, population variance, max(std, 1e-9) guard."""
    import numpy as np
    from oracle import fusion_oracle as fo
    rng = np.random.default_rng(5)
    v = np.sort(rng.normal(3.0, 2.0, 257))[::-1].tolist()
    z = np.array(fo._zscore_norm(v))
    assert abs(z.mean()) < 1e-12 and abs(z.std() - 1.0) < 1e-12
    assert np.allclose(z, (np.array(v) - np.mean(v)) / np.std(v), rtol=1e-12, atol=1e-12)
    assert fo._zscore_norm([0.5] * 8) == [0.0] * 8                # constant list: std 0 -> 1e-9 guard
    assert max(abs(x) for x in fo._zscore_norm([0.7] * 9)) < 1e-6    # (the mean of nine 0.7s is not exactly 0.7)
    assert fo._zscore_norm([2.5]) == [0.0] and fo._zscore_norm([]) == []
    # through the ranx-path restatement: a doc absent from a run contributes nothing
    lists = [[(1, 3.0, 1), (2, 1.0, 2)], [(2, 5.0, 1), (3, 1.0, 2)]]
    docs, scores, counts = fo.fuse_ranx_query(lists, fo.SUM, 60, "zscore")
    assert docs == [1, 2, 3] and scores == [1.0, 0.0, -1.0] and counts == [1, 2, 1]
