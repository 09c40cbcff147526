"""Retrievers on the hot path (reference src/retrievers/__init__.py; only BGE's flat search is in scope)."""
from .base import BaseRetriever, RetrieverResult

__all__ = ["BaseRetriever", "RetrieverResult", "BGERetriever"]


def __getattr__(name):
    if name == "BGERetriever":          # lazy: importing it touches torch/CUDA
        from .bge import BGERetriever
        return BGERetriever
    raise AttributeError(name)
