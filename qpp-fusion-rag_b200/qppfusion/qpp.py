#!/usr/bin/env python3
"""
Query Performance Prediction -- B200-native drop-in for the reference's src/qpp.py.

The reference spawns a JVM (qpp.QPPBridge) and round-trips JSON over stdin/stdout to compute 13 estimators
per (query, retriever) list (src/qpp.py:117-171 -> src/qpp/java/qpp/QPPBridge.java:134-453).  Here the same
13 columns come from one CUDA kernel (qf_qpp_stats, csrc/qf_qpp.cu), the cross-query normalisation from
qf_qpp_colnorm; the Python surface (QPPResult, QPPBridge.compute / compute_batch,
compute_qpp_for_res_file, the .qpp file format, the CLI) is unchanged.

Column order = config.qpp.methods:
  nqc smv wig SigmaMax SigmaX RSD UEF MaxIDF avgidf cumnqc snqc dense-qpp dense-qpp-m
Extra, explicitly named research-definition outputs (never substituted for the bridge columns):
  rsd_bootstrap smv_research wig_research sigmax_research scnqc_research   (QPPBridge.research_batch)

STRICT, like the reference ("No fallbacks. Java QPP bridge is mandatory", src/qpp.py:25): without the CUDA
library / an sm_100 GPU the constructor raises RuntimeError.
"""
from __future__ import annotations

import json
import os
import sys
import time
from collections import defaultdict
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from . import ops
from .config import config
from .text import text_features_array

RESEARCH_COLUMNS = ["rsd_bootstrap", "smv_research", "wig_research", "sigmax_research", "scnqc_research"]


@dataclass
class QPPResult:
    """QPP computation result (reference src/qpp.py:45-54)."""
    query: str
    retriever_name: str
    qpp_scores: Dict[str, float]
    methods_used: List[str]
    processing_time_ms: float
    predictions: Dict[str, Any]
    error: Optional[str] = None


def _pack_score_lists(score_lists: Sequence[Sequence[float]]):
    """-> (scores float64 [Q, 1, K], lens int32 [Q, 1]) in the given order (the bridge does not sort)."""
    Q = len(score_lists)
    K = max(1, max((len(s) for s in score_lists), default=1))
    sc = np.zeros((Q, 1, K), dtype=np.float64)
    lens = np.zeros((Q, 1), dtype=np.int32)
    for i, s in enumerate(score_lists):
        n = len(s)
        lens[i, 0] = n
        if n:
            sc[i, 0, :n] = np.asarray(s, dtype=np.float64)
    return sc, lens


def default_rsd_perms(n: int, n_samples: int = None, seed: int = None) -> np.ndarray:
    """Caller-supplied resample permutations for the bootstrap RSD, int32 [n_samples, n].  The reference draws
    them from one static java.util.Random(314152) shared by all queries (RSDSpecificity.java:15), which makes
    per-query results depend on processing order; here the stream is explicit and fixed."""
    n_samples = n_samples or config.qpp.rsd_samples
    rng = np.random.default_rng(config.qpp.rsd_seed if seed is None else seed)
    return np.stack([rng.permutation(n) for _ in range(n_samples)]).astype(np.int32)


class QPPBridge:
    """Batch QPP on the GPU behind the reference's bridge interface (src/qpp.py:57-238)."""

    def __init__(self, java_dir: Optional[str] = None, persistent: bool = True):
        # java_dir / persistent are accepted for signature compatibility; there is no JVM.
        self.persistent = persistent
        self.qpp_method_names = list(config.qpp.methods)
        try:
            N.lib()
            if not torch.cuda.is_available():
                raise RuntimeError("no CUDA device")
            N.check(N.lib().qf_device_check(torch.cuda.current_device()))
        except Exception as e:            # same exception type the reference raises when Java is missing
            raise RuntimeError(f"QPP CUDA bridge not available: {e}. Build it with "
                               f"`python qpp-fusion-rag_b200/build_native.py`; there is no CPU fallback.") from e
        self.device = torch.device("cuda", torch.cuda.current_device())
        print(f"[QPP] CUDA QPP (qf_qpp_stats) ready on {torch.cuda.get_device_name(self.device)}", file=sys.stderr)

    # -- array level -----------------------------------------------------------------------------
    def stats(self, scores: np.ndarray, lens: np.ndarray, texts: Sequence[str], top_k: int = None,
              perms: Optional[np.ndarray] = None, query_idf: Optional[np.ndarray] = None) -> np.ndarray:
        """scores [Q, R, K] (float64 / float32 / int32 micro), lens [Q, R], texts[Q] -> float64 [Q, R, 18]."""
        top_k = top_k or config.processing.retrieval.top_k
        dev = self.device
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        feats = text_features_array(texts)
        out, _status = ops.qpp_stats(t(scores), t(lens.astype(np.int32)), t(feats),
                                     t(None if query_idf is None else np.asarray(query_idf, dtype=np.float64)),
                                     t(perms), top_k=top_k, k_cut=config.qpp.k_cut, wig_k=config.qpp.wig_k)
        return out.cpu().numpy()

    # -- reference API -----------------------------------------------------------------------------
    def compute_batch(self, queries_scores: List[tuple]) -> List[QPPResult]:
        """List of (query_text_or_id, scores) -> List[QPPResult], matched by position; the first element is
        tokenised as query TEXT exactly as the reference's Java side does (src/qpp.py:117-171)."""
        if not queries_scores:
            return []
        texts = [str(q) for q, _ in queries_scores]
        sc, lens = _pack_score_lists([s for _, s in queries_scores])
        table = self.stats(sc, lens, texts, top_k=sc.shape[2])
        if np.isnan(table[:, 0, :13]).any():
            # the JVM fails on NaN (Gson refuses to serialise it) and the reference raises RuntimeError
            raise RuntimeError("Java QPP failed: JSON forbids NaN and infinities (all-whitespace query text)")
        results = []
        for i, text in enumerate(texts):
            row = table[i, 0]
            scores = {} if lens[i, 0] == 0 else {m: float(row[j]) for j, m in enumerate(self.qpp_method_names)}
            results.append(QPPResult(query=text, retriever_name="batch", qpp_scores=scores,
                                     methods_used=self.qpp_method_names, processing_time_ms=0, predictions={}))
        return results

    def compute(self, query: str, scores: List[float], retriever_name: str = "unknown",
                methods: Optional[List[str]] = None) -> QPPResult:
        """Single query (reference src/qpp.py:173-238).  Like the Java side, all 13 estimators are always
        computed; `methods` is accepted and ignored there too (QPPBridge.java:89-131)."""
        start = time.time()
        res = self.compute_batch([(query, scores)])[0]
        return QPPResult(query=query, retriever_name=retriever_name, qpp_scores=res.qpp_scores,
                         methods_used=list(res.qpp_scores.keys()),
                         processing_time_ms=(time.time() - start) * 1000, predictions={})

    def research_batch(self, queries_scores: List[tuple], query_idf: Optional[Sequence[float]] = None,
                       perms: Optional[np.ndarray] = None) -> List[Dict[str, float]]:
        """The research-class definitions behind the north-star names (bootstrap RSD with caller-supplied
        permutations, SMV/WIG/Sigma(%)/SCNQC Specificity classes) -- extra outputs, explicitly named."""
        if not queries_scores:
            return []
        texts = [str(q) for q, _ in queries_scores]
        sc, lens = _pack_score_lists([s for _, s in queries_scores])
        if perms is None:
            perms = default_rsd_perms(sc.shape[2])
        table = self.stats(sc, lens, texts, top_k=sc.shape[2], perms=perms, query_idf=query_idf)
        return [{m: float(table[i, 0, 13 + j]) for j, m in enumerate(RESEARCH_COLUMNS)} for i in range(len(texts))]


# ============================================================================
# Batch QPP computation for .res files
# ============================================================================

def compute_qpp_for_res_file(res_path: str, output_path: Optional[str] = None, top_k: int = None,
                             normalize: str = None, queries_path: Optional[str] = None) -> Dict[str, List[float]]:
    """All queries of one TREC .res file -> {qid: [13 QPP scores]} (+ optional .qpp file); same steps as the
    reference (src/qpp.py:245-337): parse, per-qid sort descending + truncate to top_k, bridge, column order
    of config.qpp.methods, cross-query normalisation, "%.6f" tab-separated rows in the reference's qid order."""
    top_k = top_k or config.processing.retrieval.top_k
    normalize = normalize or config.qpp.normalization
    if not os.path.exists(res_path):
        raise FileNotFoundError(f"Run file not found: {res_path}")
    query_texts = {}
    if queries_path and os.path.exists(queries_path):
        print(f"[QPP] Loading query texts from {queries_path}")
        with open(queries_path, "r") as f:
            for line in f:
                q = json.loads(line)
                query_texts[q["_id"]] = q["text"]
        print(f"[QPP] Loaded {len(query_texts)} query texts")
    runs = defaultdict(list)
    with open(res_path, "r") as f:
        for line in f:
            parts = line.strip().split()
            if len(parts) >= 5:
                runs[parts[0]].append(float(parts[4]))
    if not runs:
        raise ValueError(f"No valid entries in run file: {res_path}")
    qids = list(runs.keys())
    lists = [np.sort(np.asarray(runs[q], dtype=np.float64))[::-1][:top_k] for q in qids]
    bridge = QPPBridge()
    sc, lens = _pack_score_lists(lists)
    texts = [query_texts.get(q, q) for q in qids]
    table = bridge.stats(sc, lens, texts, top_k=top_k)[:, :, :13]
    if np.isnan(table).any():
        raise RuntimeError("Java QPP failed: JSON forbids NaN and infinities (all-whitespace query text)")
    if normalize != "none":
        g = torch.from_numpy(np.ascontiguousarray(table)).to(bridge.device)
        ops.qpp_colnorm_(g, normalize, round6=False)
        table = g.cpu().numpy()
    results = {q: [float(v) for v in table[i, 0]] for i, q in enumerate(qids)}
    if output_path:
        with open(output_path, "w") as f:
            for qid in sorted(results.keys(), key=lambda x: int(x) if x.isdigit() else x):
                f.write(qid + "\t" + "\t".join(f"{s:.6f}" for s in results[qid]) + "\n")
        print(f"Wrote QPP scores to {output_path}")
    return results


def compute_qpp_for_runs_dir(runs_dir, qpp_dir=None, queries: Optional[str] = None, top_k: int = None,
                             normalize: str = None, force: bool = False) -> Dict[str, str]:
    """Step 03 for a whole runs directory (reference scripts/03_qpp.py:38-135): every `<ranker>.res` (not `.norm.res`)
    -> `<qpp_dir>/<ranker>.res.mmnorm.qpp`; qpp_dir defaults to `<dataset>/qpp` next to the runs directory; the query
    texts are taken from `queries`, else `<dataset>/BEIR-*/queries.jsonl`, else `<dataset>/queries.jsonl`; existing
    outputs are skipped unless `force`.  The reference spreads the files over up to 8 processes, each spawning a JVM;
    here every file is a handful of kernel launches on the one device, so they run one after the other.
    -> {run file name: output path} of the files computed."""
    from pathlib import Path
    runs_dir = Path(runs_dir)
    dataset_dir = runs_dir.parent
    qpp_dir = Path(qpp_dir) if qpp_dir else dataset_dir / "qpp"
    queries_path = queries
    if not queries_path:
        for sub in sorted(dataset_dir.iterdir()):
            if sub.is_dir() and sub.name.startswith("BEIR") and (sub / "queries.jsonl").exists():
                queries_path = str(sub / "queries.jsonl")
                break
        if not queries_path and (dataset_dir / "queries.jsonl").exists():
            queries_path = str(dataset_dir / "queries.jsonl")
    if queries_path:
        print(f"[03_qpp] Using query texts from: {queries_path}")
    else:
        print("[03_qpp] WARNING: No queries.jsonl found - IDF-based methods will be inaccurate")
    os.makedirs(qpp_dir, exist_ok=True)
    res_files = sorted(f for f in os.listdir(runs_dir) if f.endswith(".res") and not f.endswith(".norm.res"))
    if not res_files:
        print(f"[03_qpp] No .res files found in {runs_dir}")
        return {}
    done: Dict[str, str] = {}
    skipped = 0
    for res_file in res_files:
        out = qpp_dir / res_file.replace(".res", ".res.mmnorm.qpp")
        if out.exists() and not force:
            print(f"[03_qpp] SKIP {res_file} (already exists)")
            skipped += 1
            continue
        print(f"[03_qpp] Processing {res_file}...")
        compute_qpp_for_res_file(str(runs_dir / res_file), str(out), top_k, normalize, queries_path)
        done[res_file] = str(out)
    print(f"\n[03_qpp] Processed: {len(done)}, Skipped: {skipped}")
    return done


def _normalize_qpp(results: Dict[str, List[float]], method: str) -> Dict[str, List[float]]:
    """Normalise QPP scores across queries (reference src/qpp.py:340-374) on the device."""
    if not results:
        return {}
    qids = list(results.keys())
    arr = np.array([results[q] for q in qids], dtype=np.float64)[:, None, :]
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
    ops.qpp_colnorm_(g, "minmax" if method == "minmax" else "zscore", round6=False)
    out = g.cpu().numpy()[:, 0, :]
    return {q: [float(v) for v in out[i]] for i, q in enumerate(qids)}


if __name__ == "__main__":
    import argparse

    parser = argparse.ArgumentParser(description="QPP Bridge - compute 13 QPP methods (CUDA, sm_100a)")
    parser.add_argument("--res_file", help="TREC .res file to process")
    parser.add_argument("--output", help="Output .qpp file")
    parser.add_argument("--top_k", type=int, default=config.processing.retrieval.top_k, help="Top-k for QPP")
    parser.add_argument("--normalize", choices=["none", "minmax", "zscore"], default=config.qpp.normalization)
    parser.add_argument("--test", action="store_true", help="Run test")
    args = parser.parse_args()
    if args.test:
        bridge = QPPBridge()
        result = bridge.compute(query="what is machine learning",
                                scores=[0.95, 0.82, 0.71, 0.65, 0.58, 0.52, 0.48, 0.41, 0.35, 0.28])
        print(f"Query: {result.query}")
        print("QPP Scores:")
        for method, score in result.qpp_scores.items():
            print(f"  {method}: {score:.4f}")
        print(f"Predictions: {result.predictions}")
    elif args.res_file:
        output = args.output or args.res_file.replace(".res", ".mmnorm.qpp")
        compute_qpp_for_res_file(args.res_file, output, args.top_k, args.normalize)
    else:
        parser.print_help()
