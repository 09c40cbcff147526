"""
TCT-ColBERT dense retriever -- the reference's second user of a flat inner-product index
(src/retrievers/tct_colbert.py:139-151 builds faiss.IndexFlatIP over the cached document embeddings; :163 and :194 call
index.search).  Here the document embedding matrix lives in HBM as bf16 and the search is the same tcgen05 kernel the
BGE retriever uses (qf_dense_topk, csrc/qf_dense.cu) -- SURVEY.md section 8(f) item 3.

Kept from the reference: the class name and `name = "TCT-ColBERT"`, the constructor arguments
(corpus, cache_dir, batch_size, use_fp16), the embedding cache files it reads
(`tct_colbert_{n_docs}_{fp16|fp32}.npy` + `tct_colbert_{n_docs}_ids.npy`, :76-103), `retrieve`, `_process_mini_batch`,
`retrieve_batch` with the JSONL checkpoint format, and the result tuples (docno, score, rank from 1).
Encoder: `model_dir=` (a local checkpoint directory -> qppfusion.encoder.BertQueryEncoder, mean pooling for a plain
HF checkpoint as SentenceTransformer(model_name) builds it) or `encoder=` (List[str] -> float32 [n, 768], L2-normalised
rows).  With an encoder and a cache_dir a missing cache is computed and written like the reference does (:97-134);
without one a missing cache raises.
"""
from __future__ import annotations

import time
from pathlib import Path
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from ..config import config
from .base import RetrieverResult
from .bge import BGERetriever


class TCTColBERTRetriever(BGERetriever):
    name = "TCT-ColBERT"

    def __init__(self, corpus: Dict[str, Dict[str, str]], cache_dir: Optional[str] = None,
                 batch_size: Optional[int] = None, use_fp16: bool = True, *, doc_embeddings=None,
                 encoder: Optional[Callable[[List[str]], np.ndarray]] = None, **kwargs):
        self.corpus_docs = corpus
        self.doc_ids = list(corpus.keys())
        self.batch_size = batch_size or config.processing.batch_sizes.tct_encoding
        self.use_fp16 = use_fp16
        self.cache_dir = Path(cache_dir) if cache_dir else None
        model_dir = kwargs.pop("model_dir", None)
        if encoder is None and model_dir is not None:
            from ..encoder import BertQueryEncoder
            encoder = BertQueryEncoder.from_pretrained(model_dir)      # plain HF checkpoint -> mean pooling, like
        if doc_embeddings is None:                                      # SentenceTransformer(model_name)
            try:
                doc_embeddings = self._load_cached_embeddings()
            except FileNotFoundError:
                if encoder is None or self.cache_dir is None:
                    raise
                doc_embeddings = self._encode_corpus_chunked(encoder)
        super().__init__(corpus_embeddings=doc_embeddings, docids=self.doc_ids, encoder=encoder, **kwargs)
        self.model_name = config.models.tct_colbert.name
        self.embedding_dim = config.models.tct_colbert.embedding_dim
        if self.corpus.shape[0] != len(self.doc_ids):
            raise ValueError(f"{self.corpus.shape[0]} embeddings for {len(self.doc_ids)} documents")

    def _get_cache_path(self) -> Path:
        dtype = "fp16" if self.use_fp16 else "fp32"
        return self.cache_dir / f"tct_colbert_{len(self.doc_ids)}_{dtype}.npy"

    def _load_cached_embeddings(self) -> np.ndarray:
        """The reference's cache (src/retrievers/tct_colbert.py:82-96); a missing or mismatching cache raises --
        encoding the corpus is not part of this package."""
        if self.cache_dir is None:
            raise FileNotFoundError("TCTColBERTRetriever: pass doc_embeddings= or a cache_dir holding the reference's "
                                    "tct_colbert_*.npy embedding cache")
        cache_path = self._get_cache_path()
        ids_path = self.cache_dir / f"tct_colbert_{len(self.doc_ids)}_ids.npy"
        if not (cache_path.exists() and ids_path.exists()):
            raise FileNotFoundError(f"No embedding cache: {cache_path}")
        emb = np.load(str(cache_path))
        cached_ids = np.load(str(ids_path), allow_pickle=True)
        if list(cached_ids) != self.doc_ids:
            raise ValueError("[TCT-ColBERT] Cache IDs mismatch")
        return emb

    def _encode_corpus_chunked(self, encoder) -> np.ndarray:
        """No cache: encode "title text" of every document in chunks of 8 batches and write the reference's cache
        files (src/retrievers/tct_colbert.py:97-134: same text construction, dtype, file names and id file)."""
        dtype = np.float16 if self.use_fp16 else np.float32
        n_docs = len(self.doc_ids)
        emb = np.zeros((n_docs, config.models.tct_colbert.embedding_dim), dtype=dtype)
        chunk = self.batch_size * 8
        for s0 in range(0, n_docs, chunk):
            ids = self.doc_ids[s0:s0 + chunk]
            texts = [(self.corpus_docs[d].get("title", "") + " " + self.corpus_docs[d].get("text", "")).strip() for d in ids]
            emb[s0:s0 + len(ids)] = np.asarray(encoder(texts), dtype=np.float32).astype(dtype)
        self.cache_dir.mkdir(parents=True, exist_ok=True)
        np.save(str(self._get_cache_path()), emb)
        np.save(str(self.cache_dir / f"tct_colbert_{n_docs}_ids.npy"), np.array(self.doc_ids, dtype=object))
        return emb

    def _process_mini_batch(self, queries: List[Tuple[str, str]], top_k: int) -> List[RetrieverResult]:
        """One mini-batch through the vectorised search (reference :176-208)."""
        qids = [q[0] for q in queries]
        embs = self._encode_queries([q[1] for q in queries])
        all_ids, all_scores = self._search_batch(embs, top_k)
        out = []
        for i, qid in enumerate(qids):
            hits = [(self.docids[idx], float(all_scores[i, j]), j + 1)
                    for j, idx in enumerate(all_ids[i]) if 0 <= idx < len(self.docids)]
            out.append(RetrieverResult(qid=qid, results=hits, retriever_name=self.name, latency_ms=0,
                                       metadata={"model": self.model_name}))
        return out

    def retrieve(self, query: str, qid: str = "", top_k: int = None, **kwargs) -> RetrieverResult:
        start = time.time()
        res = self._process_mini_batch([(qid, query)], top_k or config.processing.retrieval.top_k)[0]
        res.latency_ms = (time.time() - start) * 1000
        return res

    def retrieve_batch(self, queries: Dict[str, str], top_k: int = None, checkpoint_path: Optional[str] = None,
                       mini_batch_size: Optional[int] = None, **kwargs) -> Dict[str, RetrieverResult]:
        out = super().retrieve_batch(queries, top_k=top_k, checkpoint_path=checkpoint_path,
                                     mini_batch_size=mini_batch_size, **kwargs)
        for r in out.values():
            r.metadata = {"model": self.model_name}
        return out
