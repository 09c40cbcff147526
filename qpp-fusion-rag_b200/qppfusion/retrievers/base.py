"""
Retriever result types (the search call's return contract; reference src/retrievers/base.py).
"""
from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple


@dataclass
class RetrieverResult:
    """Standard retriever output (reference src/retrievers/base.py:14-29)."""
    qid: str
    results: List[Tuple[str, float, int]]  # (docno, score, rank)
    retriever_name: str
    latency_ms: float = 0.0
    metadata: Dict[str, Any] = field(default_factory=dict)

    def to_trec_lines(self, run_name: Optional[str] = None) -> List[str]:
        name = run_name or self.retriever_name
        return [f"{self.qid} Q0 {docno} {rank} {score:.6f} {name}" for docno, score, rank in self.results]


class BaseRetriever(ABC):
    """Abstract base (reference src/retrievers/base.py:32-73)."""

    name: str = "base"

    @abstractmethod
    def retrieve(self, query: str, qid: str, top_k: int = 100, **kwargs) -> RetrieverResult:
        ...

    def retrieve_batch(self, queries: Dict[str, str], top_k: int = 100, **kwargs) -> Dict[str, RetrieverResult]:
        return {qid: self.retrieve(query, qid, top_k, **kwargs) for qid, query in queries.items()}

    @staticmethod
    def normalize_scores(results: List[Tuple[str, float, int]]) -> List[Tuple[str, float, int]]:
        """Per-query min-max producing the .norm.res values: (s - min) / (max - min if max > min else 1.0)
        (reference src/retrievers/base.py:60-73).  A handful of floats per query; stays on the host."""
        if not results:
            return []
        scores = [s for _, s, _ in results]
        min_s, max_s = min(scores), max(scores)
        range_s = max_s - min_s if max_s > min_s else 1.0
        return [(docno, (score - min_s) / range_s, rank) for docno, score, rank in results]
