"""
Result record and retriever interface of the search call (the return contract of the hot path's first stage).

Field names, the TREC line layout and the per-query min-max rule are the reference's
(src/retrievers/base.py:14-29, :60-73); they are what scripts/02_retrieve.py::write_run serialises into
`.res` / `.norm.res`, i.e. the 6-decimal values every later stage (03_qpp, 05_fusion) reads back.
"""
from __future__ import annotations

import abc
import dataclasses
from typing import Any, Dict, Iterable, List, Optional, Tuple

Hit = Tuple[str, float, int]          # (docno, score, rank)


def trec_line(qid: str, hit: Hit, run_name: str) -> str:
    """One run-file line; the score is printed with 6 decimals -- this rounding defines the downstream values."""
    docno, score, rank = hit
    return "%s Q0 %s %d %.6f %s" % (qid, docno, rank, score, run_name)


def minmax_normalize(hits: Iterable[Hit]) -> List[Hit]:
    """Per-query min-max of the scores: (s - lo) / (hi - lo), with a unit divisor when all scores are equal
    (so a constant or single-document list maps to 0.0).  Docno and rank pass through untouched."""
    hits = list(hits)
    if len(hits) == 0:
        return hits
    lo = hi = hits[0][1]
    for _, s, _ in hits[1:]:
        lo = s if s < lo else lo
        hi = s if s > hi else hi
    span = (hi - lo) if hi > lo else 1.0
    return [(d, (s - lo) / span, r) for d, s, r in hits]


@dataclasses.dataclass
class RetrieverResult:
    """What every retriever returns for one query."""
    qid: str
    results: List[Hit]
    retriever_name: str
    latency_ms: float = 0.0
    metadata: Dict[str, Any] = dataclasses.field(default_factory=dict)

    def to_trec_lines(self, run_name: Optional[str] = None) -> List[str]:
        label = self.retriever_name if not run_name else run_name
        return [trec_line(self.qid, h, label) for h in self.results]


class BaseRetriever(abc.ABC):
    """Interface shared by the retrievers; only BGE's flat inner-product search is implemented in this package."""

    name: str = "base"

    @abc.abstractmethod
    def retrieve(self, query: str, qid: str, top_k: int = 100, **kwargs) -> RetrieverResult:
        raise NotImplementedError

    def retrieve_batch(self, queries: Dict[str, str], top_k: int = 100, **kwargs) -> Dict[str, RetrieverResult]:
        out: Dict[str, RetrieverResult] = {}
        for qid in queries:
            out[qid] = self.retrieve(queries[qid], qid, top_k, **kwargs)
        return out

    # kept as a static method because callers use BaseRetriever.normalize_scores(docs) (scripts/02_retrieve.py:77)
    normalize_scores = staticmethod(minmax_normalize)


def write_run(results: Dict[str, "RetrieverResult"], output_path: str, retriever_name: str, normalize: bool = False) -> None:
    """TREC run file of a retriever's results (reference scripts/02_retrieve.py:65-82): queries in the reference's
    order (numeric for "test<N>" ids, else string order), optional per-query min-max, "%.6f" scores -- the rounding that
    defines the values every later stage sees."""
    import os
    d = os.path.dirname(output_path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(output_path, "w") as f:
        for qid in sorted(results.keys(), key=lambda x: int(x.replace("test", "")) if x.startswith("test") else x):
            docs = results[qid].results
            if normalize:
                docs = BaseRetriever.normalize_scores(docs)
            f.write("".join(f"{qid} Q0 {docno} {rank} {score:.6f} {retriever_name}\n" for docno, score, rank in docs))
    print(f"[02_retrieve] Wrote {output_path}")


def run_retriever(retriever, queries: Dict[str, str], runs_dir, top_k: int, mini_batch_size: Optional[int] = None,
                  name: Optional[str] = None) -> Dict[str, "RetrieverResult"]:
    """One retriever of step 02 (reference scripts/02_retrieve.py run_bge :184-213 / run_tct_colbert :117-150):
    retrieve_batch with the `<name>.checkpoint.jsonl` resume file, then `<name>.res` and `<name>.norm.res` through
    write_run, then the checkpoint is removed.  -> the results dict."""
    import os
    import time
    name = name or retriever.name
    print(f"\n[02_retrieve] === {name} ===")
    start = time.time()
    os.makedirs(str(runs_dir), exist_ok=True)
    checkpoint_path = os.path.join(str(runs_dir), f"{name}.checkpoint.jsonl")
    results = retriever.retrieve_batch(queries, top_k=top_k, checkpoint_path=checkpoint_path,
                                       mini_batch_size=mini_batch_size)
    write_run(results, os.path.join(str(runs_dir), f"{name}.res"), name, normalize=False)
    write_run(results, os.path.join(str(runs_dir), f"{name}.norm.res"), name, normalize=True)
    if os.path.exists(checkpoint_path):
        os.unlink(checkpoint_path)
        print("[02_retrieve] Removed checkpoint file")
    print(f"[02_retrieve] {name} completed in {time.time() - start:.1f}s")
    return results
