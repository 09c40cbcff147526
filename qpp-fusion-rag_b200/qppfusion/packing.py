"""
DataFrame runs -> the dense [Q, R, K] device layout of qf_fuse / qf_qpp_stats, and back.

This is the host side of the reference's `_df_runs_to_ranx` (src/fusion.py:455-466) and of the per-query
groupby in `weighted_*` (src/fusion.py:260-264) -- vectorised (pd.factorize + stable argsort) instead of
38 us/row `iterrows`.  Only format conversion happens here; every float operation of the fusion itself
runs in the CUDA kernel.

HBM layout produced (row-major):
    doc_ids int32  [Q, R, K]   dictionary-encoded docno (one dictionary for the whole call), -1 padded
    scores  int32  [Q, R, K]   micro-units (score*1e6) when every score is a 6-decimal value that fits,
            or float64 [Q, R, K] otherwise
    ranks   int32  [Q, R, K]   the file's rank column (only materialised when a method needs it)
    lens    int32  [Q, R]
Q follows sorted(qids) (the reference iterates `sorted(set.union(...))`, src/fusion.py:260); R follows the
order of the `runs` dict.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import pandas as pd


@dataclass
class PackedRuns:
    qids: List[str]                 # length Q, sorted
    rankers: List[str]              # length R, dict order of `runs`
    docnos: np.ndarray              # object array: dictionary code -> docno string
    doc_ids: np.ndarray             # int32 [Q, R, K]
    scores: np.ndarray              # int32 micro-units or float64 [Q, R, K]
    ranks: Optional[np.ndarray]     # int32 [Q, R, K] or None
    lens: np.ndarray                # int32 [Q, R]

    @property
    def shape(self):
        return self.doc_ids.shape


def _micro_or_f64(scores: np.ndarray):
    """int32 micro-units if every score is exactly a 6-decimal value (what .res files carry), else None."""
    if scores.size == 0:
        return scores.astype(np.int32)
    if not np.isfinite(scores).all() or np.abs(scores).max() >= 2147.0:
        return None
    m = np.rint(scores * 1e6)
    if np.array_equal(m / 1e6, scores):       # IEEE division: the double float("d.dddddd") parses to
        return m.astype(np.int32)
    return None


def pack_runs(runs: Dict[str, pd.DataFrame], mode: str = "weighted", need_ranks: bool = False) -> PackedRuns:
    """
    mode = "weighted": rows kept as they are, in file order (every row accumulates, src/fusion.py:275,323).
    mode = "weighted_rrf": rows stably sorted by the rank column inside each (qid, ranker) group
                           (src/fusion.py:371 sort_values("rank")).
    mode = "ranx": python-dict semantics of _df_runs_to_ranx (a repeated docno keeps its first position
                   and its last score, src/fusion.py:460-464), then ranx.Run's sort by score descending
                   (ties keep dict order -- ranx's own tie order is unspecified).
    """
    rankers = list(runs.keys())
    if not rankers:
        raise ValueError("no runs")
    qid_sets = [pd.unique(df["qid"].astype(str)) for df in runs.values()]
    qids = sorted(set().union(*[set(q.tolist()) for q in qid_sets]))
    qindex = pd.Index(qids)
    Q, R = len(qids), len(rankers)

    all_doc = pd.concat([df["docno"].astype(str) for df in runs.values()], ignore_index=True)
    codes, uniques = pd.factorize(all_doc, sort=False)
    if len(uniques) >= 2 ** 31:
        raise ValueError("more than 2^31 distinct documents")
    per = []
    off = 0
    K = 1
    for df in runs.values():
        n = len(df)
        qi = qindex.get_indexer(df["qid"].astype(str).to_numpy())
        di = codes[off:off + n].astype(np.int64)
        off += n
        sc = df["score"].to_numpy(dtype=np.float64, copy=True)
        rk = df["rank"].to_numpy() if "rank" in df.columns else np.zeros(n, dtype=np.int64)
        rk = rk.astype(np.int64, copy=False)
        pos = np.arange(n, dtype=np.int64)
        if mode == "ranx":
            # dict semantics: last score wins, first position kept
            key = qi * (len(uniques) + 1) + di
            order = np.argsort(key, kind="stable")
            ks = key[order]
            first = np.ones(n, dtype=bool)
            first[1:] = ks[1:] != ks[:-1]
            last = np.ones(n, dtype=bool)
            last[:-1] = ks[1:] != ks[:-1]
            keep_pos = pos[order][first]           # first occurrence position
            keep_sc = sc[order][last]              # last occurrence score (groups are contiguous & stable)
            qi, di, rk, sc, pos = qi[order][first], di[order][first], rk[order][first], keep_sc, keep_pos
            o2 = np.lexsort((pos, -sc, qi))        # qid, then score desc, then dict position
            qi, di, rk, sc = qi[o2], di[o2], rk[o2], sc[o2]
        elif mode == "weighted_rrf":
            o2 = np.lexsort((pos, rk, qi))         # stable sort by rank inside each qid
            qi, di, rk, sc = qi[o2], di[o2], rk[o2], sc[o2]
        else:
            o2 = np.argsort(qi, kind="stable")
            qi, di, rk, sc = qi[o2], di[o2], rk[o2], sc[o2]
        cnt = np.bincount(qi, minlength=Q)
        start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
        within = np.arange(len(qi), dtype=np.int64) - start[qi]
        K = max(K, int(cnt.max()) if len(cnt) else 1)
        per.append((qi, within, di, sc, rk, cnt))

    all_scores = np.concatenate([p[3] for p in per]) if per else np.zeros(0)
    micro = _micro_or_f64(all_scores)
    doc_ids = np.full((Q, R, K), -1, dtype=np.int32)
    scores = np.zeros((Q, R, K), dtype=np.int32 if micro is not None else np.float64)
    ranks = np.zeros((Q, R, K), dtype=np.int32) if need_ranks else None
    lens = np.zeros((Q, R), dtype=np.int32)
    off = 0
    for r, (qi, within, di, sc, rk, cnt) in enumerate(per):
        doc_ids[qi, r, within] = di
        scores[qi, r, within] = micro[off:off + len(sc)] if micro is not None else sc
        off += len(sc)
        if ranks is not None:
            if len(rk) and (np.abs(rk).max() >= 2 ** 31):
                raise ValueError("rank column does not fit int32")
            ranks[qi, r, within] = rk
        lens[:, r] = cnt
    return PackedRuns(qids, rankers, np.asarray(uniques, dtype=object), doc_ids, scores, ranks, lens)


def unpack_fused(packed: PackedRuns, ids: np.ndarray, scores: np.ndarray, lens: np.ndarray):
    """Kernel outputs -> the reference's return type {qid: [(docid, score), ...]} (queries with no
    postings are absent, as in the reference where `fused` only gets keys that have documents)."""
    out = {}
    docnos = packed.docnos
    for qi, qid in enumerate(packed.qids):
        n = int(lens[qi])
        if n == 0:
            continue
        out[qid] = list(zip(docnos[ids[qi, :n]].tolist(), scores[qi, :n].tolist()))
    return out
