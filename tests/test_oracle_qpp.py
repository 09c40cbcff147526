"""Oracle (oracle/qpp_oracle.py) vs the reference's committed .qpp known-answer columns + hand checks."""
import gzip
import json
import math
import os

import numpy as np
import pytest

from oracle import qpp_oracle as qo

GOLD = os.path.join(os.path.dirname(__file__), "golden", "qpp_textcols.json.gz")


@pytest.mark.parametrize("ds", ["scifact", "nq"])
def test_text_columns_match_committed_qpp_files(ds):
    with gzip.open(GOLD, "rt") as f:
        g = json.load(f)[ds]
    raw = {}
    for qid, text in zip(g["qids"], g["texts"]):
        feats = qo.text_features(text)
        raw[qid] = [qo.max_idf_proxy(feats), qo.avg_idf_proxy(feats)]
    norm = qo.normalize_qpp(raw, "minmax")
    lines = qo.qpp_file_lines(norm)
    assert [l.split("\t")[0] for l in lines] == g["qids"]            # file ordering rule
    got7 = [l.split("\t")[1] for l in lines]
    got8 = [l.split("\t")[2] for l in lines]
    assert got7 == g["col7"]
    assert got8 == g["col8"]


def test_java_split_semantics():
    assert qo.java_split_ws("a b") == ["a", "b"]
    assert qo.java_split_ws("  a\tb \n") == ["", "a", "b"]
    assert qo.java_split_ws("") == [""]
    assert qo.java_split_ws(" \t ") == []
    assert qo.java_split_ws("a b") == ["a b"]      # NBSP is not Java \s
    assert qo.text_features("Hello hello  WORLD") == (3, 2, 5, 15)
    assert qo.utf16_len("a\U0001F600") == 3


def test_parse_float32_is_decimal_to_float32():
    for s in ["0.123456", "17.250001", "0.000001", "-3.999999", "1e-3", "16777217"]:
        v = qo.parse_float32_exact(s)
        assert v == float(np.float32(v))
        # nearest float32 to the decimal value
        from fractions import Fraction
        c = np.float32(v)
        for nb in (np.nextafter(c, np.float32(np.inf)), np.nextafter(c, np.float32(-np.inf))):
            assert abs(Fraction(float(c)) - Fraction(s)) <= abs(Fraction(float(nb)) - Fraction(s))


def test_bridge_estimators_hand_checks():
    # the reference's own smoke input (src/qpp.py:397-398); no expected values exist upstream, so
    # check each estimator against an independent numpy formula
    sc = [0.95, 0.82, 0.71, 0.65, 0.58, 0.52, 0.48, 0.41, 0.35, 0.28]
    x = np.array([np.float32(v) for v in sc], dtype=np.float64)
    out = dict(zip(qo.BRIDGE_METHODS, qo.bridge_scores("what is machine learning", sc)))
    assert out["nqc"] == pytest.approx(x.var(), rel=1e-12)
    assert out["smv"] == pytest.approx(np.mean(x * np.abs(np.log(x / x.mean()))), rel=1e-12)
    assert out["wig"] == pytest.approx(np.sum(x - 1 / x.mean()) / (4 * 10), rel=1e-12)
    assert out["SigmaMax"] == pytest.approx(max(x[:i].std() for i in range(2, 11)) / 2.0, rel=1e-12)
    assert out["SigmaX"] == pytest.approx(x[x >= x[0] * 0.5].std(), rel=1e-12)
    z = (x - x.mean()) / x.std()
    assert out["RSD"] == pytest.approx(np.mean(z ** 3), rel=1e-12)
    w = 1 / np.log2(np.arange(10) + 2)
    assert out["UEF"] == pytest.approx(np.sum(x * w) / w.sum(), rel=1e-12)
    assert out["MaxIDF"] == pytest.approx(math.log(5) + 0.5 * math.log(9), rel=1e-15)
    assert out["avgidf"] == pytest.approx(math.log(1 + 21 / 4), rel=1e-15)
    assert out["cumnqc"] == pytest.approx(np.mean([x[:k].var() for k in range(2, 11)]), rel=1e-12)
    assert out["snqc"] == pytest.approx(np.mean((((x - x.mean()) ** 2 / x) ** 0.33) ** 0.33), rel=1e-12)
    assert out["dense-qpp"] == out["dense-qpp-m"] == pytest.approx(math.log(1 + 1 / (x[0] - x[-1])), rel=1e-12)


def test_bridge_edge_cases():
    assert qo.bridge_scores("q", []) == [0.0] * 13
    one = dict(zip(qo.BRIDGE_METHODS, qo.bridge_scores("q", [0.5])))
    assert one["nqc"] == 0.0 and one["SigmaMax"] == 0.0 and one["SigmaX"] == 0.0 and one["RSD"] == 0.0
    assert one["cumnqc"] == 0.0 and one["dense-qpp"] == 0.0 and one["UEF"] == 0.5
    neg = dict(zip(qo.BRIDGE_METHODS, qo.bridge_scores("a b", [-1.0, -2.0, -3.0])))
    assert neg["smv"] == 0.0 and neg["snqc"] == 0.0          # x>0 guards
    assert neg["wig"] == pytest.approx(((-1 - 100) + (-2 - 100) + (-3 - 100)) / (2 * 3))   # baseline 1/max(0.01, mean)
    const = dict(zip(qo.BRIDGE_METHODS, qo.bridge_scores("a", [0.7] * 60)))
    assert const["RSD"] == 0.0 and const["dense-qpp"] == 0.0
    # only the first 50 are used
    a = qo.bridge_scores("a", list(np.linspace(1, 0.1, 80)))
    b = qo.bridge_scores("a", list(np.linspace(1, 0.1, 80))[:50])
    assert a == b


def test_rbo_and_bootstrap_rsd():
    # identical lists: RBO = (1-p) * sum p^(d-1) = 1 - p^k
    ids = list(range(6))
    p = qo.F32_09
    assert qo.rbo(ids, ids, 6) == pytest.approx((1 - p ** 6) * (float(np.float32(1) - np.float32(0.9)) / (1 - p)), rel=1e-12)
    # OverlapStats.java:234-236 example input (no expected output upstream): brute-force check
    a = [1, 3, 4, 5, 8, 9]
    b = [2, 3, 5, 6, 8, 10]
    p8 = float(np.float32(0.8))
    exp = float(np.float32(1) - np.float32(0.8)) * sum(len(set(a[:d]) & set(b[:d])) / d * p8 ** (d - 1) for d in range(1, 7))
    assert qo.rbo(a, b, 6, p8) == pytest.approx(exp, rel=1e-12)
    rng = np.random.default_rng(5)
    scores = sorted(rng.random(120).astype(np.float32).astype(float), reverse=True)
    perms = [rng.permutation(200).tolist() for _ in range(10)]
    v = qo.rsd_bootstrap(scores, perms, 50)
    # brute force
    acc = 0.0
    for perm in perms:
        smp = [i for i in perm if i < 120][:50]
        est = np.var(np.array([scores[i] for i in smp]))
        sim = qo.rbo(list(range(120)), smp, 120)
        acc += sim * est
    assert v == pytest.approx(acc / 10, rel=1e-10)


def test_normalize_and_file_format():
    res = {"10": [1.0, 5.0], "9": [3.0, 5.0], "test1": [2.0, 5.0]}
    mm = qo.normalize_qpp({k: v for k, v in res.items() if k != "test1"}, "minmax")
    assert mm == {"10": [0.0, 0.0], "9": [1.0, 0.0]}
    z = qo.normalize_qpp(res, "zscore")
    assert z["9"][0] == pytest.approx((3 - 2) / math.sqrt(2 / 3))
    assert z["9"][1] == 0.0
    assert qo.qpp_file_lines(mm) == ["9\t1.000000\t0.000000", "10\t0.000000\t0.000000"]


def test_normalize_qpp_bit_equal_to_the_real_reference_function():
    """oracle normalize_qpp vs outputs of the REAL src/qpp.py::_normalize_qpp (tools/make_golden_qpp_norm.py)."""
    import gzip
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qpp_normalize.json.gz")
    with gzip.open(path, "rt") as f:
        gold = json.load(f)
    for case in gold["cases"]:
        table = {q: [float(v) for v in row] for q, row in case["table"].items()}
        for mode in ("minmax", "zscore"):
            got = qo.normalize_qpp(table, mode)
            ref = case[mode]
            assert list(got.keys()) == list(ref.keys())
            for q in ref:
                assert [float(v) for v in got[q]] == [float(v) for v in ref[q]], (case["seed"], mode, q)
