"""Retrievers on the hot path (reference src/retrievers/__init__.py; the flat inner-product searches of BGE and TCT-ColBERT are in scope)."""
from .base import BaseRetriever, RetrieverResult

__all__ = ["BaseRetriever", "RetrieverResult", "BGERetriever", "TCTColBERTRetriever"]


def __getattr__(name):
    if name == "BGERetriever":          # lazy: importing it touches torch/CUDA
        from .bge import BGERetriever
        return BGERetriever
    if name == "TCTColBERTRetriever":
        from .tct_colbert import TCTColBERTRetriever
        return TCTColBERTRetriever
    raise AttributeError(name)
