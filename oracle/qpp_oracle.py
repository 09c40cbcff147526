"""
TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's QPP arithmetic.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker.  The product path never imports it.

What is restated (citations relative to /root/reference):
  * QPPBridge.computeQPPScores + the 13 estimators        src/qpp/java/qpp/QPPBridge.java:134-453
  * Gson getAsFloat rounding of the inputs                QPPBridge.java:65-70
  * Java String.toLowerCase().split("\\s+") tokenisation  QPPBridge.java:250,272,369,386
  * compute_qpp_for_res_file: sort/truncate, column order, min-max / z-score, "%.6f" file
                                                          src/qpp.py:245-374
  * research-class predictors the north star names:
      NQCSpecificity.java:44-65, SMVSpecificity.java:22-45, WIGSpecificity.java:18-51,
      SigmaMaxSpecificity.java:18-51, SigmaXSpecificity.java:19-52,
      NQCCalibratedSpecificity.java:31-57, RSDSpecificity.java:22-52 (bootstrap),
      correlation/OverlapStats.java:77-99 (RBO)

Parity status
  * Text-only columns 7 (MaxIDF proxy) and 8 (avgidf proxy) + _normalize_qpp(minmax) + "%.6f" + file
    ordering: PINNED against the committed data/{scifact,nq}/qpp/*.res.mmnorm.qpp
    (tests/golden/qpp_textcols.json.gz, made by tools/make_golden_qpp.py; 300/300 and 3452/3452 rows).
  * The 9 score-dependent bridge columns and every research-class predictor: PARITY UNPINNED.  No JVM
    in the build container, target/classes is not shipped, the .res inputs of the committed .qpp files
    are Git-LFS pointers, and the reference has no tests.  The restatement follows the Java line by line.

Java numerics restated:
  * DoubleStream.average() = Kahan-compensated sum / count (java.util.stream.Collectors
    .sumWithCompensation / computeFinalSum).  JDK <= 16 adds the compensation term with the wrong sign
    in computeFinalSum (JDK-8214761); the difference is <= 1 ulp of the sum and far below the 1e-5
    tolerance of this path.  The fixed (JDK >= 17) form is used here.
  * Math.pow(x, 2) == x*x (HotSpot intrinsic special case), Math.log natural, Math.sqrt exact.
"""
from __future__ import annotations

import math
import struct
from fractions import Fraction
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

BRIDGE_METHODS = ["nqc", "smv", "wig", "SigmaMax", "SigmaX", "RSD", "UEF", "MaxIDF", "avgidf",
                  "cumnqc", "snqc", "dense-qpp", "dense-qpp-m"]           # config/defaults.yaml:164-177
METHOD_INDEX = {"SMV": 0, "Sigma_max": 1, "Sigma(%)": 2, "NQC": 3, "UEF": 4, "RSD": 5, "QPP-PRP": 6,
                "WIG": 7, "SCNQC": 8, "QV-NQC": 9, "DM": 10, "NQA-QPP": 11, "BERTQPP": 12}  # :180-193
RESEARCH_COLUMNS = ["rsd_bootstrap", "smv_research", "wig_research", "sigmax_research", "scnqc_research"]

JAVA_WS = " \t\n\x0b\f\r"      # java.util.regex \s without UNICODE_CHARACTER_CLASS


# --------------------------------------------------------------------------------------------------
# number formats
# --------------------------------------------------------------------------------------------------

def _f32_round(x: float) -> float:
    return struct.unpack("f", struct.pack("f", x))[0]


def parse_float32_exact(text: str) -> float:
    """Float.parseFloat(text): the float32 nearest to the DECIMAL value (ties to even), as a double.

    Gson hands the JSON token to Float.parseFloat (LazilyParsedNumber.floatValue), so the rounding is
    decimal -> float32 directly, not decimal -> double -> float32.
    """
    v = Fraction(text)
    if v == 0:
        return 0.0
    c = np.float32(float(v))                  # within 1 ulp(f32) of the answer
    cands = [c, np.nextafter(c, np.float32(np.inf)), np.nextafter(c, np.float32(-np.inf))]
    best, best_err = None, None
    for x in cands:
        err = abs(Fraction(float(x)) - v)
        even = (int(np.array(x, dtype=np.float32).view(np.uint32)) & 1) == 0
        if best is None or err < best_err or (err == best_err and even):
            best, best_err = x, err
    return float(best)


def json_float32(x: float) -> float:
    """Python json.dumps(float) (repr) -> Gson getAsFloat -> (double).  QPPBridge.java:69,141-142."""
    return parse_float32_exact(repr(float(x)))


def fmt6(v: float) -> float:
    """float(f"{v:.6f}") -- the value every later stage sees after the .qpp / .res text round trip."""
    return float(f"{v:.6f}")


# --------------------------------------------------------------------------------------------------
# Java helpers
# --------------------------------------------------------------------------------------------------

def java_average(xs: Sequence[float], or_else: float = 0.0) -> float:
    """DoubleStream.average().orElse(or_else): Kahan sum (Collectors.sumWithCompensation) / count."""
    if len(xs) == 0:
        return or_else
    s = 0.0
    c = 0.0
    simple = 0.0
    for v in xs:
        tmp = v - c
        velvel = s + tmp
        c = (velvel - s) - tmp
        s = velvel
        simple += v
    tot = s - c
    if math.isnan(tot) and math.isinf(simple):
        tot = simple
    return tot / len(xs)


def java_split_ws(text: str) -> List[str]:
    """text.split("\\s+") with Java semantics: a leading whitespace run yields a leading "" token;
    trailing empty tokens are dropped; no match -> [text] (so "" -> [""]); all-whitespace -> []."""
    toks: List[str] = []
    i, n, start, matched = 0, len(text), 0, False
    while i < n:
        if text[i] in JAVA_WS:
            j = i
            while j < n and text[j] in JAVA_WS:
                j += 1
            toks.append(text[start:i])
            start = j
            i = j
            matched = True
        else:
            i += 1
    if not matched:
        return [text]
    toks.append(text[start:])
    while toks and toks[-1] == "":
        toks.pop()
    return toks


def utf16_len(s: str) -> int:
    return len(s.encode("utf-16-le")) // 2


def text_features(query_text: str) -> Tuple[int, int, int, int]:
    """(n_terms_raw, n_unique, max_len, sum_len) of queryText.toLowerCase().split("\\s+")."""
    terms = java_split_ws(query_text.lower())
    uniq = set(terms)
    max_len = max((utf16_len(t) for t in uniq), default=0)
    sum_len = sum(utf16_len(t) for t in terms)
    return len(terms), len(uniq), max_len, sum_len


# --------------------------------------------------------------------------------------------------
# the 13 bridge estimators (QPPBridge.java:202-453); rsvs = first min(50, n) inputs as float32->double
# --------------------------------------------------------------------------------------------------

def nqc(rsvs: Sequence[float]) -> float:                       # :202-216
    if len(rsvs) == 0:
        return 0.0
    mean = java_average(rsvs)
    acc = 0.0
    for v in rsvs:
        d = v - mean
        acc += d * d
    return (acc / len(rsvs)) * 1.0


def smv(rsvs: Sequence[float]) -> float:                       # :223-239
    if len(rsvs) == 0:
        return 0.0
    mu = java_average(rsvs, 1.0)
    if mu <= 0:
        mu = 1.0
    acc = 0.0
    for s in rsvs:
        if s > 0:
            acc += s * abs(math.log(s / mu))
    return acc / len(rsvs)


def wig(n_terms_raw: int, rsvs: Sequence[float]) -> float:     # :247-262
    if len(rsvs) == 0:
        return 0.0
    nt = max(1, n_terms_raw)
    baseline = 1.0 / max(0.01, java_average(rsvs, 1.0))
    acc = 0.0
    for v in rsvs:
        acc += (v - baseline)
    return acc / (nt * len(rsvs))


def _pop_std(xs: Sequence[float]) -> float:
    mean = java_average(xs)
    return math.sqrt(java_average([(s - mean) * (s - mean) for s in xs]))


def sigma_max(n_terms_raw: int, rsvs: Sequence[float]) -> float:   # :269-293
    if len(rsvs) < 2:
        return 0.0
    nt = max(1, n_terms_raw)
    best = 0.0
    for i in range(2, len(rsvs) + 1):
        best = max(best, _pop_std(rsvs[:i]))
    return best / math.sqrt(max(1, nt))


def sigma_x(rsvs: Sequence[float]) -> float:                   # :299-319
    if len(rsvs) < 2:
        return 0.0
    thr = rsvs[0] * 0.5
    f = [v for v in rsvs if v >= thr]
    if len(f) < 2:
        return 0.0
    return _pop_std(f)


def rsd_skew(rsvs: Sequence[float]) -> float:                  # :324-342
    if len(rsvs) < 3:
        return 0.0
    mean = java_average(rsvs)
    std = math.sqrt(java_average([(s - mean) * (s - mean) for s in rsvs]))
    if std < 1e-10:
        return 0.0
    acc = 0.0
    for v in rsvs:
        acc += math.pow((v - mean) / std, 3)
    return acc / len(rsvs)


def uef(rsvs: Sequence[float]) -> float:                       # :349-363
    if len(rsvs) == 0:
        return 0.0
    util = 0.0
    wsum = 0.0
    for i in range(min(20, len(rsvs))):
        w = 1.0 / (math.log(i + 2) / math.log(2))
        util += rsvs[i] * w
        wsum += w
    return util / wsum


def max_idf_proxy(feats: Tuple[int, int, int, int]) -> float:  # :368-380
    _n, n_unique, max_len, _s = feats
    return math.log(1 + n_unique) + math.log(1 + max_len) * 0.5


def avg_idf_proxy(feats: Tuple[int, int, int, int]) -> float:  # :385-393
    n_terms, n_unique, _m, sum_len = feats
    avg_len = (sum_len / n_terms) if n_terms > 0 else 3.0
    diversity = (n_unique / n_terms) if n_terms > 0 else float("nan")   # (double)0/0 in Java
    return math.log(1 + avg_len) * diversity


def cum_nqc(rsvs: Sequence[float]) -> float:                   # :400-410
    if len(rsvs) < 2:
        return 0.0
    acc = 0.0
    for k in range(2, len(rsvs) + 1):
        acc += nqc(rsvs[:k])
    return acc / (len(rsvs) - 1)


def snqc(rsvs: Sequence[float]) -> float:                      # :417-438
    if len(rsvs) == 0:
        return 0.0
    mean = java_average(rsvs)
    if mean <= 0:
        return 0.0
    alpha = beta = gamma = 0.33
    acc = 0.0
    for v in rsvs:
        if v <= 0:
            continue
        f2 = ((v - mean) * (v - mean)) / v
        prod = math.pow(1.0, alpha) * math.pow(f2, beta)
        acc += math.pow(prod, gamma)
    return (acc / len(rsvs)) * 1.0


def score_spread(rsvs: Sequence[float]) -> float:              # :444-453
    if len(rsvs) < 2:
        return 0.0
    rng = max(rsvs) - min(rsvs)
    return math.log(1 + 1.0 / rng) if rng > 0 else 0.0


def bridge_scores(query_text: str, scores: Sequence[float], k_cut: int = 50,
                  inputs_are_f32: bool = False) -> List[float]:
    """QPPBridge.computeQPPScores (:134-191) -> the 13 values in config.qpp.methods order.

    `scores` are the Python-side floats (already sorted desc and truncated to top_k by
    compute_qpp_for_res_file, src/qpp.py:304-305).  Empty input -> empty map in Java -> 13 x 0.0 after
    `.get(m, 0.0)` (src/qpp.py:321).
    """
    if len(scores) == 0:
        return [0.0] * 13
    k = min(k_cut, len(scores))
    rsvs = [float(s) if inputs_are_f32 else json_float32(s) for s in scores[:k]]
    feats = text_features(query_text)
    spread = score_spread(rsvs)
    return [nqc(rsvs), smv(rsvs), wig(feats[0], rsvs), sigma_max(feats[0], rsvs), sigma_x(rsvs),
            rsd_skew(rsvs), uef(rsvs), max_idf_proxy(feats), avg_idf_proxy(feats), cum_nqc(rsvs),
            snqc(rsvs), spread, spread]


# --------------------------------------------------------------------------------------------------
# research-class predictors (need values an index would supply: n_query_terms, max_idf)
# --------------------------------------------------------------------------------------------------

F32_033 = _f32_round(0.33)      # alpha/beta/gamma are Java floats (experiments/QPPOnPreRetrievedResults.java:64)
F32_09 = _f32_round(0.9)        # RBO p = 0.9f (OverlapStats.java:104)


def rbo(ids_a: Sequence[int], ids_b: Sequence[int], k: int, p32: float = F32_09) -> float:
    """OverlapStats.computeRBO(int[], int[], k, float p) (:93-101) with rbo_overlap (:85-91)."""
    aggr = 0.0
    for d in range(1, k + 1):
        sa = set(ids_a[:d])
        sb = set(ids_b[:d])
        ao = (len(sa & sb) / d) * math.pow(p32, d - 1)
        aggr += ao
    one_minus_p = _f32_round(1.0 - p32)          # (1-p) is evaluated in float
    return one_minus_p * aggr


def rsd_bootstrap(scores_f32: Sequence[float], perms: Sequence[Sequence[int]], k_cut: int = 50) -> float:
    """RSDSpecificity.computeSpecificity (:38-52) over NQCSpecificity(k_cut), with the shuffles supplied
    by the caller: perms[s] is a permutation of range(n_max); entries >= n are skipped and the first
    min(n, k_cut) survivors are the sample (Collections.shuffle + subList, :24-27).  The document ids of
    the original list are 0..n-1 (rank positions)."""
    n = len(scores_f32)
    if n == 0:
        return 0.0
    m = min(n, k_cut)
    acc = 0.0
    for perm in perms:
        sample = [i for i in perm if i < n][:m]
        est = nqc([scores_f32[i] for i in sample])
        sim = rbo(list(range(n)), sample, n)
        acc += sim * est
    return acc / len(perms)


def smv_research(rsvs: Sequence[float], avg_idf: float = 1.0) -> float:      # SMVSpecificity.java:22-45
    if len(rsvs) == 0:
        return 0.0
    mu = java_average(rsvs, 1.0)
    acc = 0.0
    for s in rsvs:
        if s > 0 and mu > 0:
            acc += s * abs(math.log(s / mu))
    return (acc / len(rsvs)) * avg_idf


def wig_research(all_scores_f32: Sequence[float], n_query_terms: int, max_idf: float, wig_k: int = 5) -> float:
    """WIGSpecificity.java:18-51: sum_{i<cutoff}(x_i - 1/maxIDF) / (nTerms*cutoff), cutoff=min(k, n).
    Java double semantics: maxIDF == 0 gives 1/0 = +inf."""
    if len(all_scores_f32) == 0:
        return 0.0
    base = (1.0 / max_idf) if max_idf != 0 else math.inf
    cutoff = min(wig_k, len(all_scores_f32))
    acc = 0.0
    for i in range(cutoff):
        acc += all_scores_f32[i] - base
    den = float(n_query_terms * cutoff)
    if den == 0:
        return math.nan if acc == 0 or math.isnan(acc) else math.copysign(math.inf, acc)
    return acc / den


def sigmax_research(rsvs: Sequence[float], qlen: int, x: float = 0.5) -> float:   # SigmaXSpecificity.java:19-52
    if len(rsvs) == 0:
        return 0.0
    thr = rsvs[0] * x
    f = [v for v in rsvs if v >= thr]
    if len(f) == 0:
        return 0.0
    sig = _pop_std(f)
    root = math.sqrt(qlen)
    if root == 0:
        return math.nan if sig == 0 else math.inf
    return sig / root


def scnqc_research(rsvs: Sequence[float], max_idf: float) -> float:         # NQCCalibratedSpecificity.java:31-57
    if len(rsvs) == 0:
        return 0.0
    mean = java_average(rsvs)
    acc = 0.0
    for v in rsvs:
        if v == 0:
            f2 = math.nan if v == mean else math.inf
        else:
            f2 = (v - mean) * (v - mean) / v
        try:
            prod = math.pow(max_idf, F32_033) * math.pow(f2, F32_033)
            prod = math.pow(prod, F32_033)
        except (ValueError, OverflowError):
            prod = math.nan          # Math.pow of a negative base with a fractional exponent is NaN
        acc += prod
    return (acc / len(rsvs)) * max_idf


def research_scores(all_scores: Sequence[float], n_query_terms: int, max_idf: float,
                    perms: Sequence[Sequence[int]], k_cut: int = 50, wig_k: int = 5,
                    inputs_are_f32: bool = False) -> List[float]:
    """The 5 extra columns (RESEARCH_COLUMNS) for one list.  ScoreDoc.score is a float32
    (qrels/AllRetrievedResults.java:186), so inputs are rounded to float32 first."""
    allf = [float(s) if inputs_are_f32 else _f32_round(float(s)) for s in all_scores]
    rsvs = allf[:k_cut]
    return [rsd_bootstrap(allf, perms, k_cut), smv_research(rsvs), wig_research(allf, n_query_terms, max_idf, wig_k),
            sigmax_research(rsvs, n_query_terms), scnqc_research(rsvs, max_idf)]


# --------------------------------------------------------------------------------------------------
# src/qpp.py: per-file pipeline
# --------------------------------------------------------------------------------------------------

def normalize_qpp(results: Dict[str, List[float]], method: str) -> Dict[str, List[float]]:
    """_normalize_qpp (src/qpp.py:340-374), with numpy's own min/max/mean/std as the reference calls them (pairwise
    summation), so the result is bit-equal to the reference's: pinned by tests/golden/qpp_normalize.json.gz, outputs of
    the REAL function (tools/make_golden_qpp_norm.py)."""
    import numpy as np
    if not results:
        return {}
    n = len(next(iter(results.values())))
    cols = [[v[i] for v in results.values()] for i in range(n)]
    params = []
    for vals in cols:
        arr = np.array(vals)
        if method == "minmax":
            vmin, vmax = arr.min(), arr.max()
            params.append((vmin, vmax - vmin if vmax > vmin else 1.0))
        else:
            params.append((arr.mean(), arr.std() if arr.std() > 0 else 1.0))
    out = {}
    for qid, vals in results.items():
        row = []
        for i, v in enumerate(vals):
            base, scale = params[i]
            if method == "minmax":
                row.append((v - base) / scale if scale > 0 else 0.0)
            else:
                row.append((v - base) / scale)
        out[qid] = row
    return out


def qpp_for_run(run_scores: Dict[str, List[float]], query_texts: Optional[Dict[str, str]] = None,
                top_k: int = 100, normalize: str = "minmax") -> Dict[str, List[float]]:
    """compute_qpp_for_res_file (src/qpp.py:245-337) from already-parsed {qid: [scores in file order]}."""
    query_texts = query_texts or {}
    res: Dict[str, List[float]] = {}
    for qid, sc in run_scores.items():
        top = sorted(sc, reverse=True)[:top_k]
        res[qid] = bridge_scores(query_texts.get(qid, qid), top)
    if normalize != "none" and res:
        res = normalize_qpp(res, normalize)
    return res


def qpp_file_sort_key(x: str):
    return int(x) if x.isdigit() else x          # src/qpp.py:331


def qpp_file_lines(results: Dict[str, List[float]]) -> List[str]:
    """.qpp writer (src/qpp.py:329-335)."""
    lines = []
    for qid in sorted(results.keys(), key=qpp_file_sort_key):
        lines.append(qid + "\t" + "\t".join(f"{s:.6f}" for s in results[qid]))
    return lines
