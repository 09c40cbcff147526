#!/usr/bin/env python3
"""
Rank fusion for multi-retriever RAG -- B200-native drop-in for the reference's src/fusion.py.

Same entry points, argument meaning, return types, printed tags and error behaviour as the reference
(PoisedDok/QPP-Fusion-RAG src/fusion.py), with the arithmetic moved from ranx/numba and Python dict loops
into one CUDA kernel (qf_fuse, csrc/qf_fuse.cu):

    Unweighted (reference: via ranx)        CombSUM   sum_i S_i(d,q)            combsum(runs)
                                            CombMNZ   |{i: d in R_i}| x sum     combmnz(runs)
                                            RRF       sum_i 1/(k + rank_i)      rrf(runs, k)
    QPP-weighted (reference: Python loops)  W-CombSUM / W-CombMNZ / W-RRF       weighted_*(runs, qpp_data, qpp_index)
    Learned                                 W-CombSUM with model-predicted weights   learned_fusion(...)

STRICT, like the reference ("No fallbacks", src/fusion.py:24): no CUDA library or no sm_100 GPU -> the call
raises; there is no CPU path in this package.
"""
from __future__ import annotations

import os
import pickle
from typing import Dict, List, Optional, Tuple

import numpy as np
import pandas as pd
import torch

from . import _native as N
from . import ops
from .config import config
from .packing import PackedRuns, pack_runs, unpack_fused

Fused = Dict[str, List[Tuple[str, float]]]


# =============================================================================
# Data loading
# =============================================================================

def _run_files(res_path: str, use_normalized: bool):
    suffix = ".norm.res" if use_normalized else ".res"
    files = [f for f in os.listdir(res_path) if f.endswith(suffix)]     # use_normalized=False also picks up
    if not files:                                                        # "X.norm.res" (as ranker "X.norm"), as the reference does
        raise FileNotFoundError(f"No {suffix} files found in {res_path}")
    # The reference keeps os.listdir order (src/fusion.py:60), i.e. filesystem dependent.  Ranker order
    # decides float64 accumulation order and the tie-break, so this build fixes it: sorted by name.
    return sorted(files), suffix


def load_runs(res_path: str, use_normalized: bool = True) -> Dict[str, pd.DataFrame]:
    """{retriever_name: DataFrame[qid, iter, docno, rank, score, runid]} (reference src/fusion.py:46-76)."""
    files, suffix = _run_files(res_path, use_normalized)
    runs = {}
    for f in files:
        ranker = f.replace(suffix, "")
        df = pd.read_csv(os.path.join(res_path, f), sep=r"\s+", names=["qid", "iter", "docno", "rank", "score", "runid"],
                         dtype={"qid": str, "docno": str})
        df["qid"] = df["qid"].astype(str)
        runs[ranker] = df
    return runs


def load_runs_as_ranx(res_path: str, use_normalized: bool = True) -> Dict[str, pd.DataFrame]:
    """The reference returns ranx.Run objects (src/fusion.py:79-103); ranx is not part of this build, the
    same data is returned as DataFrames sorted the way ranx.Run sorts (score descending per query)."""
    runs = load_runs(res_path, use_normalized)
    return {k: df.sort_values(["qid", "score"], ascending=[True, False], kind="stable").reset_index(drop=True)
            for k, df in runs.items()}


def load_qpp_scores(qpp_path: str) -> Dict[str, Dict[str, List[float]]]:
    """{qid: {retriever: [13 qpp scores]}} (reference src/fusion.py:106-113)."""
    from .data_utils import load_qpp_scores as _load
    return _load(qpp_path)


def get_qpp_weight(qid: str, ranker: str, qpp_data: Dict[str, Dict[str, List[float]]], qpp_index: int = None,
                   fusion_mode: bool = False) -> float:
    """Weight of one (query, ranker) pair; KeyError / IndexError exactly as the reference (src/fusion.py:116-158)."""
    if qpp_index is None:
        qpp_index = config.qpp.default_index
    if qid not in qpp_data:
        raise KeyError(f"QPP data missing for query '{qid}'. Run QPP computation first.")
    if ranker not in qpp_data[qid]:
        raise KeyError(f"QPP data missing for ranker '{ranker}' on query '{qid}'. Available: {list(qpp_data[qid].keys())}")
    scores = qpp_data[qid][ranker]
    if fusion_mode or qpp_index == -1:
        return sum(scores) / len(scores)
    if qpp_index >= len(scores):
        raise IndexError(f"QPP index {qpp_index} out of range. Available: 0-{len(scores)-1}")
    return scores[qpp_index]


def get_qpp_index(model_name: str) -> int:
    """Resolve QPP model name to index; -1 for 'fusion' (reference src/fusion.py:161-163)."""
    return config.get_qpp_index(model_name)


# =============================================================================
# Device plumbing
# =============================================================================

def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("qppfusion.fusion needs a CUDA (sm_100) device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _weights_matrix(packed: PackedRuns, qpp_data, qpp_index: int) -> np.ndarray:
    """float64 [Q, R]; the weight is looked up for EVERY (qid, ranker), also when that ranker has no rows for
    the qid, so a missing entry raises KeyError exactly as the reference does (src/fusion.py:270-272)."""
    fusion_mode = (qpp_index == -1)
    w = np.empty((len(packed.qids), len(packed.rankers)), dtype=np.float64)
    for qi, qid in enumerate(packed.qids):
        for ri, ranker in enumerate(packed.rankers):
            w[qi, ri] = get_qpp_weight(qid, ranker, qpp_data, qpp_index, fusion_mode)
    return w


# Table-size hint for qf_fuse (results never depend on it): the mean union ratio seen by the previous call with the same
# [R, K] shape decides whether the next one asks for the larger-table configuration (N.FLAG_LARGE_UNION).
_UNION_RATIO: Dict[tuple, float] = {}
LARGE_UNION_RATIO = 0.46


def union_hint_flags(shape_key: tuple) -> int:
    return N.FLAG_LARGE_UNION if _UNION_RATIO.get(shape_key, 0.0) > LARGE_UNION_RATIO else 0


def remember_union(shape_key: tuple, out_lens: torch.Tensor) -> None:
    r, k = shape_key
    if out_lens.numel():
        _UNION_RATIO[shape_key] = float(out_lens.clamp(min=0).float().mean().item()) / float(max(1, r * k))


def fuse_packed(packed: PackedRuns, method: int, weights: Optional[np.ndarray] = None, norm: int = N.NORM_NONE,
                rrf_k: float = 60.0, flags: int = 0, want_counts: bool = False, want_pos: bool = False) -> ops.FuseOut:
    """Array-level entry point: PackedRuns -> kernel outputs on the device (no Python objects built)."""
    dev = _device()
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    shape_key = tuple(packed.doc_ids.shape[1:])
    r_, k_ = shape_key
    max_post = int(N.lib().qf_fuse_deep_max_postings())     # above qf_fuse's envelope ops.fuse takes qf_fuse_deep
    if r_ * k_ > max_post or r_ > 32:
        # K is the deepest (qid, ranker) list of the whole call (every list is padded to it): name it
        qi, ri = np.unravel_index(int(np.argmax(packed.lens)), packed.lens.shape)
        raise ValueError(
            f"qf_fuse: rankers x depth = {r_} x {k_} = {r_ * k_} exceeds the envelope of {max_post} postings per "
            f"query (and 32 rankers). The deepest list is ranker '{packed.rankers[ri]}' on query "
            f"'{packed.qids[qi]}' with {int(packed.lens[qi, ri])} rows; cut the runs to depth <= "
            f"{max_post // max(1, r_)} (e.g. top_k at retrieval time) or fuse fewer runs.")
    out = ops.fuse(t(packed.doc_ids), t(packed.scores), t(packed.lens), t(weights), t(packed.ranks), method, norm,
                   rrf_k, flags | union_hint_flags(shape_key), want_counts, want_pos)
    status = out.status.cpu().numpy()
    remember_union(shape_key, out.lens)
    if (status & N.QSTAT_BAD_LEN).any():
        raise ValueError("qf_fuse: invalid list length")
    if (status & N.QSTAT_UNSORTED).any():
        bad = [packed.qids[i] for i in np.flatnonzero(status & N.QSTAT_UNSORTED)[:5]]
        raise ValueError(f"qf_fuse: run lists violate the ordering contract for queries {bad}")
    return out


def _to_dict(packed: PackedRuns, out: ops.FuseOut) -> Fused:
    return unpack_fused(packed, out.ids.cpu().numpy(), out.scores.cpu().numpy(), out.lens.cpu().numpy())


# =============================================================================
# Unweighted fusion (reference: ranx.fuse with its default norm="min-max")
# =============================================================================

_RUN_NORMS = {"min-max": N.NORM_MINMAX, "minmax": N.NORM_MINMAX, "zscore": N.NORM_ZSCORE, "z-score": N.NORM_ZSCORE,
              None: N.NORM_NONE, "none": N.NORM_NONE}


def _ranx_fuse(runs: Dict[str, pd.DataFrame], method: int, rrf_k: float = 60.0, norm="min-max") -> Fused:
    """ranx.fuse restated (parity unpinned: ranx is not installable offline).  Deliberate
    differences in edge cases the reference pipeline never produces: queries are fused over the UNION of the runs' qids
    (a run without rows for a query contributes nothing; ranx walks the first run's queries and fails on a run that
    lacks one), and a single run is accepted (ranx.fuse wants at least two)."""
    if norm not in _RUN_NORMS:
        raise ValueError(f"Unknown run normalisation '{norm}'. Valid: min-max, zscore, none")
    packed = pack_runs(runs, mode="ranx")
    flags = N.FLAG_REQUIRE_SORTED
    if method == N.METHOD_MNZ and not config.fusion.ranx_mnz_counts_zero_scores:
        flags |= N.FLAG_MNZ_SKIP_ZERO
    norm = N.NORM_NONE if method == N.METHOD_RRF else _RUN_NORMS[norm]
    out = fuse_packed(packed, method, None, norm, rrf_k, flags)
    return _to_dict(packed, out)      # ranked: score descending (as _ranx_to_dict, src/fusion.py:469-477)


def combsum(runs: Dict[str, pd.DataFrame], norm="min-max") -> Fused:
    """CombSUM(d,q) = sum_i S_i(d,q) over min-max normalised runs (reference src/fusion.py:170-186).
    `norm` (extension; the reference always takes ranx's default "min-max"): "min-max", "zscore" or None."""
    print("[fusion] Running CombSUM (qf_fuse)...")
    result = _ranx_fuse(runs, N.METHOD_SUM, norm=norm)
    print(f"[fusion] CombSUM done: {len(result)} queries")
    return result


def combmnz(runs: Dict[str, pd.DataFrame], norm="min-max") -> Fused:
    """CombMNZ(d,q) = |{i: d in R_i}| x sum_i S_i(d,q) (reference src/fusion.py:189-205).  `norm`: see combsum."""
    print("[fusion] Running CombMNZ (qf_fuse)...")
    result = _ranx_fuse(runs, N.METHOD_MNZ, norm=norm)
    print(f"[fusion] CombMNZ done: {len(result)} queries")
    return result


def rrf(runs: Dict[str, pd.DataFrame], k: int = None) -> Fused:
    """RRF(d,q) = sum_i 1/(k + rank_i(d,q)), rank = position in the score-sorted run (reference src/fusion.py:208-230)."""
    k = k if k is not None else config.fusion.rrf_k
    print("[fusion] Running RRF (qf_fuse)...")
    result = _ranx_fuse(runs, N.METHOD_RRF, float(k))
    print(f"[fusion] RRF done: {len(result)} queries")
    return result


# =============================================================================
# QPP-weighted fusion (reference: Python dict loops over iterrows)
# =============================================================================

def _weighted(runs, qpp_data, qpp_index, method: int, rrf_k: float = 60.0) -> Fused:
    packed = pack_runs(runs, mode="weighted_rrf" if method == N.METHOD_RRF else "weighted",
                       need_ranks=(method == N.METHOD_RRF))
    w = _weights_matrix(packed, qpp_data, qpp_index)
    # the reference returns each query's list in dict-insertion (first-seen) order, UNSORTED
    # (src/fusion.py:277-278); ranking happens in write_runfile
    out = fuse_packed(packed, method, w, N.NORM_NONE, rrf_k, N.FLAG_UNRANKED_OUTPUT)
    return _to_dict(packed, out)


def weighted_combsum(runs: Dict[str, pd.DataFrame], qpp_data, qpp_index: int = None) -> Fused:
    """W-CombSUM(d,q) = sum_i w_i(q) x S_i(d,q) (reference src/fusion.py:237-281)."""
    qpp_index = qpp_index if qpp_index is not None else config.qpp.default_index
    print(f"[fusion] Running W-CombSUM (QPP index={qpp_index})...")
    fused = _weighted(runs, qpp_data, qpp_index, N.METHOD_SUM)
    print(f"[fusion] W-CombSUM done: {len(fused)} queries")
    return fused


def weighted_combmnz(runs: Dict[str, pd.DataFrame], qpp_data, qpp_index: int = None) -> Fused:
    """W-CombMNZ(d,q) = |{i}| x sum_i w_i(q) x S_i(d,q) (reference src/fusion.py:284-330)."""
    qpp_index = qpp_index if qpp_index is not None else config.qpp.default_index
    print(f"[fusion] Running W-CombMNZ (QPP index={qpp_index})...")
    fused = _weighted(runs, qpp_data, qpp_index, N.METHOD_MNZ)
    print(f"[fusion] W-CombMNZ done: {len(fused)} queries")
    return fused


def weighted_rrf(runs: Dict[str, pd.DataFrame], qpp_data, qpp_index: int = None, k: int = None) -> Fused:
    """W-RRF(d,q) = sum_i w_i(q) / (k + rank_i(d,q)) with the FILE's rank column (reference src/fusion.py:333-380)."""
    qpp_index = qpp_index if qpp_index is not None else config.qpp.default_index
    k = k if k is not None else config.fusion.rrf_k
    print(f"[fusion] Running W-RRF (QPP index={qpp_index})...")
    fused = _weighted(runs, qpp_data, qpp_index, N.METHOD_RRF, float(k))
    print(f"[fusion] W-RRF done: {len(fused)} queries")
    return fused


# =============================================================================
# Learned fusion (model-predicted weights)
# =============================================================================

def learned_fusion(runs: Dict[str, pd.DataFrame], qpp_data, model_path: str,
                   retrievers: Optional[List[str]] = None) -> Fused:
    """W-CombSUM with per-query weights predicted by a pickled model (reference src/fusion.py:387-448).

    The weight matrix build follows the reference (feature vector = the first n_qpp QPP values of each
    retriever in model order, zeros when missing; weight 1/R for a ranker the model does not know, :435);
    the accumulation runs in qf_fuse.  The reference sums with pandas groupby(["qid","docno"]).sum() and
    returns documents in docno order; this returns the same (docno, score) sets -- scores equal within
    float64 summation-order error (<= 1e-15 relative), list order = docno ascending like the reference.
    """
    print(f"[fusion] Running learned fusion from {model_path}...")
    with open(model_path, "rb") as f:
        model_data = pickle.load(f)
    model = model_data.get("model")
    retrievers = retrievers or model_data.get("retrievers", sorted(runs.keys()))
    n_qpp = model_data.get("n_qpp", config.qpp.n_methods)

    packed = pack_runs(runs, mode="weighted")
    n_ret = len(retrievers)
    X = np.zeros((len(packed.qids), n_qpp * n_ret))
    for i, qid in enumerate(packed.qids):
        for j, retriever in enumerate(retrievers):
            if qid in qpp_data and retriever in qpp_data[qid]:
                X[i, j * n_qpp:(j + 1) * n_qpp] = qpp_data[qid][retriever][:n_qpp]
    pred = np.asarray(model.predict(X), dtype=np.float64)
    w = np.full((len(packed.qids), len(packed.rankers)), 1.0 / n_ret, dtype=np.float64)
    for ri, ranker in enumerate(packed.rankers):
        if ranker in retrievers:
            w[:, ri] = pred[:, retrievers.index(ranker)]
    out = fuse_packed(packed, N.METHOD_SUM, w, N.NORM_NONE, 60.0, N.FLAG_UNRANKED_OUTPUT)
    fused = _to_dict(packed, out)
    fused = {qid: sorted(lst, key=lambda x: x[0]) for qid, lst in fused.items()}
    print(f"[fusion] Learned fusion done: {len(fused)} queries")
    return fused


# =============================================================================
# Output
# =============================================================================

def _qid_sort_key(x: str):
    return int(x.replace("test", "")) if x.startswith("test") else x     # reference src/fusion.py:487


def write_runfile(fused: Fused, output_path: str, tag: str = "fusion"):
    """TREC run file: queries in the reference's order, documents by score descending with ties in list
    order (Python's stable sort, reference src/fusion.py:480-492), "{qid} Q0 {docid} {rank} {score:.6f} {tag}"."""
    with open(output_path, "w") as fout:
        for qid in sorted(fused.keys(), key=_qid_sort_key):
            lst = fused[qid]
            if not lst:
                continue
            scores = np.fromiter((s for _, s in lst), dtype=np.float64, count=len(lst))
            order = np.argsort(-scores, kind="stable")      # == sorted(key=score, reverse=True): stable
            fout.write("".join(f"{qid} Q0 {lst[i][0]} {rank} {lst[i][1]:.6f} {tag}\n"
                               for rank, i in enumerate(order.tolist(), start=1)))
    print(f"Wrote fused run to {output_path}")


def _fused_dict_from_out(packed: PackedRuns, docnos: np.ndarray, out: ops.FuseOut, first_seen_order: bool) -> Fused:
    """Kernel outputs (ranked) -> the reference's return value; weighted_* lists come back in first-seen order."""
    ids, scores, lens = out.ids.cpu().numpy(), out.scores.cpu().numpy(), out.lens.cpu().numpy()
    pos = out.pos.cpu().numpy() if first_seen_order else None
    fused = {}
    for qi, qid in enumerate(packed.qids):
        n = int(lens[qi])
        if n == 0:
            continue
        d, s_ = docnos[ids[qi, :n]], scores[qi, :n]
        if pos is not None:
            inv = np.empty(n, dtype=np.int64)
            inv[pos[qi, :n].astype(np.int64)] = np.arange(n)
            d, s_ = d[inv], s_[inv]
        fused[qid] = list(zip(d.tolist(), s_.tolist()))
    return fused


def run_fusion(method: str, runs_dir: str, qpp_dir: Optional[str] = None, qpp_model: str = None,
               model_path: Optional[str] = None, output_path: Optional[str] = None, rrf_k: int = None,
               return_fused: bool = True) -> Fused:
    """Dispatcher with the reference's method names, aliases, tags and errors (src/fusion.py:495-579).

    File-to-file fast path: the run files are parsed and dictionary-encoded by the native reader
    (native_io.parse_runs), fused in one qf_fuse launch, CSR-compacted on the device and written by the native
    writer; the Python dict the reference returns is still built (return_fused=False skips it, the CLI does)."""
    from . import native_io as nio
    qpp_model = qpp_model or config.qpp.default_method
    rrf_k = rrf_k if rrf_k is not None else config.fusion.rrf_k
    method = method.lower()
    if method == "learned":
        if not model_path:
            raise ValueError("--model_path required for learned fusion")
        if not qpp_dir:
            raise ValueError("--qpp_dir required for learned fusion")
        runs = load_runs(runs_dir, use_normalized=True)
        print(f"Loaded {len(runs)} rankers: {list(runs.keys())}")
        fused = learned_fusion(runs, load_qpp_scores(qpp_dir), model_path)
        if output_path:
            write_runfile(fused, output_path, "learned")
        return fused
    weighted_names = {"wcombsum": N.METHOD_SUM, "wcombmnz": N.METHOD_MNZ, "wrrf": N.METHOD_RRF}
    plain_names = {"combsum": N.METHOD_SUM, "combmnz": N.METHOD_MNZ, "rrf": N.METHOD_RRF}
    base = method.replace("-", "")
    if base not in weighted_names and method not in plain_names:
        files, _ = _run_files(runs_dir, True)          # FileNotFoundError first, as the reference (load_runs runs first)
        raise ValueError(f"Unknown method '{method}'. Valid: {config.fusion.methods}")
    files, suffix = _run_files(runs_dir, True)
    names = [f.replace(suffix, "") for f in files]
    if base in weighted_names and method not in plain_names and not qpp_dir:
        print(f"Loaded {len(names)} rankers: {names}")
        raise ValueError("--qpp_dir required for weighted methods")
    parsed = nio.parse_runs([os.path.join(runs_dir, f) for f in files], names)
    print(f"Loaded {len(names)} rankers: {names}")
    qpp_data = load_qpp_scores(qpp_dir) if method not in plain_names else None
    return _fuse_parsed(parsed, method, qpp_data, qpp_model, output_path, rrf_k, return_fused)


_WEIGHTED = {"wcombsum": N.METHOD_SUM, "wcombmnz": N.METHOD_MNZ, "wrrf": N.METHOD_RRF}
_PLAIN = {"combsum": N.METHOD_SUM, "combmnz": N.METHOD_MNZ, "rrf": N.METHOD_RRF}


def _fuse_parsed(parsed, method: str, qpp_data, qpp_model: str, output_path: Optional[str], rrf_k: int,
                 return_fused: bool, tag: Optional[str] = None) -> Fused:
    """One fusion method over already parsed run files: pack -> qf_fuse -> (CSR compaction -> native writer)."""
    from . import native_io as nio
    base = method.replace("-", "")
    if method in _PLAIN:
        m = _PLAIN[method]
        packed, qorder = nio.pack_parsed(parsed, "ranx")
        flags = N.FLAG_REQUIRE_SORTED | (N.FLAG_MNZ_SKIP_ZERO if (m == N.METHOD_MNZ and not config.fusion.ranx_mnz_counts_zero_scores) else 0)
        print(f"[fusion] Running {method} (qf_fuse)...")
        out = fuse_packed(packed, m, None, N.NORM_NONE if m == N.METHOD_RRF else N.NORM_MINMAX, float(rrf_k), flags)
        tag = tag or (f"rrf-k{rrf_k}" if method == "rrf" else method)
        first_seen = False
    else:
        m = _WEIGHTED[base]
        qpp_index = get_qpp_index(qpp_model)
        packed, qorder = nio.pack_parsed(parsed, "weighted_rrf" if m == N.METHOD_RRF else "weighted", need_ranks=(m == N.METHOD_RRF))
        w = _weights_matrix(packed, qpp_data, qpp_index)
        print(f"[fusion] Running {base} (QPP index={qpp_index}, qf_fuse)...")
        out = fuse_packed(packed, m, w, N.NORM_NONE, float(rrf_k), 0, want_pos=return_fused)
        tag = tag or f"{base}-{qpp_model.lower()}"
        first_seen = True
    if output_path:
        off, cids, csc = ops.compact_fused(out)
        off = off.cpu().numpy()
        total = int(off[-1])
        nio.write_runfile_csr(output_path, parsed, qorder, off, cids[:total].cpu().numpy(), csc[:total].cpu().numpy(), tag)
        print(f"Wrote fused run to {output_path}")
    if not return_fused:
        return {}
    fused = _fused_dict_from_out(packed, parsed.doc_strings(), out, first_seen)
    print(f"[fusion] {method} done: {len(fused)} queries")
    return fused


def run_all_methods(runs_dir, qpp_dir, fused_dir, qpp_model: str = None, rrf_k: int = None) -> Dict[str, "os.PathLike"]:
    """Every fusion method in one go, with the reference's file names and tags (scripts/05_fusion.py:56-152):
    combsum.res, combmnz.res, rrf.res (tag "rrf"), then -- if qpp_dir exists -- wcombsum_{model}.res,
    wcombmnz_{model}.res, wrrf_{model}.res and learned_{name}.res for every model pickle under ../models.
    The run files are parsed and dictionary-encoded ONCE (native reader); each method is one pack + qf_fuse launch +
    device-side CSR compaction + native writer.  -> {method: output path}."""
    from pathlib import Path
    from . import native_io as nio
    runs_dir, qpp_dir, fused_dir = Path(runs_dir), Path(qpp_dir), Path(fused_dir)
    qpp_model = qpp_model or config.qpp.default_method
    rrf_k = rrf_k if rrf_k is not None else config.fusion.rrf_k
    print("[05_fusion] Running all fusion methods...")
    print(f"[05_fusion] QPP model for weighted: {qpp_model}")
    files, suffix = _run_files(str(runs_dir), True)
    names = [f.replace(suffix, "") for f in files]
    parsed = nio.parse_runs([os.path.join(str(runs_dir), f) for f in files], names)
    print(f"[05_fusion] Loaded {len(names)} rankers: {names}")
    fused_dir.mkdir(parents=True, exist_ok=True)
    results: Dict[str, Path] = {}
    for method in ("combsum", "combmnz", "rrf"):
        out = fused_dir / f"{method}.res"
        _fuse_parsed(parsed, method, None, qpp_model, str(out), rrf_k, False, tag=method)
        results[method] = out
    if qpp_dir.exists():
        qpp_data = load_qpp_scores(str(qpp_dir))
        for method in ("wcombsum", "wcombmnz", "wrrf"):
            out = fused_dir / f"{method}_{qpp_model.lower()}.res"
            _fuse_parsed(parsed, method, qpp_data, qpp_model, str(out), rrf_k, False)
            results[method] = out
        models_dir = fused_dir.parent / "models"
        runs = None
        for model_name in ("per_retriever", "multioutput", "mlp"):
            model_file = models_dir / f"fusion_{model_name}.pkl"
            if model_file.exists():
                runs = runs if runs is not None else load_runs(str(runs_dir), use_normalized=True)
                out = fused_dir / f"learned_{model_name}.res"
                write_runfile(learned_fusion(runs, qpp_data, str(model_file)), str(out), f"learned-{model_name}")
                results[f"learned_{model_name}"] = out
            else:
                print(f"[05_fusion] Skipping learned ({model_name}) - model not found")
    else:
        print(f"[05_fusion] Skipping weighted methods (no QPP at {qpp_dir})")
    print(f"\n=== Generated {len(results)} fusion runs ===")
    for method, path in results.items():
        print(f"  {method}: {path}")
    return results


if __name__ == "__main__":
    import argparse

    parser = argparse.ArgumentParser(description="Fusion Methods for Multi-Retriever RAG (B200-native)")
    parser.add_argument("--method", required=True, choices=config.fusion.methods, help="Fusion method")
    parser.add_argument("--runs_dir", required=True, help="Directory with .norm.res files")
    parser.add_argument("--qpp_dir", default=None, help="Directory with .qpp files")
    parser.add_argument("--qpp_model", default=config.qpp.default_method, help="QPP model for weighting")
    parser.add_argument("--model_path", default=None, help="Path to learned model")
    parser.add_argument("--output", required=True, help="Output TREC run file")
    parser.add_argument("--rrf_k", type=int, default=config.fusion.rrf_k, help="RRF k constant")
    a = parser.parse_args()
    run_fusion(method=a.method, runs_dir=a.runs_dir, qpp_dir=a.qpp_dir, qpp_model=a.qpp_model,
               model_path=a.model_path, output_path=a.output, rrf_k=a.rrf_k, return_fused=False)
