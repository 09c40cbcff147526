"""
BGE dense retriever -- B200-native drop-in for the search calls of the reference's
src/retrievers/bge.py (BGERetriever._search / _search_batch / retrieve / retrieve_batch).

The reference answers a query batch with faiss.IndexFlatIP.search over a Pyserini prebuilt flat index
(fp32 sgemm + heaps on the CPU, src/retrievers/bge.py:179-194).  Here the corpus embedding matrix lives in
HBM as bf16 [N, 768] and one call scores the whole batch with the tcgen05 kernel behind
qf_dense_topk (csrc/qf_dense.cu); per-GPU results are merged with an NCCL all-gather + qf_topk_merge
when the corpus is row-sharded across ranks.

Query encoder: `encoder=` takes any callable List[str] -> float32 [n, 768] with L2-normalised rows; `model_dir=`
(a local bge-base-en-v1.5 checkpoint directory) builds the on-device BERT encoder (qppfusion.encoder ->
qf_bert_encode), whose bf16 output feeds the search without a host round trip.  Out of scope:
the approximate HNSW segments.
"""
from __future__ import annotations

import json
import struct
import time
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _native as N
from .. import ops
from ..config import config
from .base import BaseRetriever, RetrieverResult


def read_faiss_flat_index(path: str) -> np.ndarray:
    """Read the vectors of a FAISS IndexFlat file ("IxFI"/"IxF2"/"IxFl" fourcc) as float32 [ntotal, d].
    Layout (faiss/impl/index_write.cpp, write_index_header + IndexFlat): fourcc u32, d i32, ntotal i64,
    two i64 dummies, is_trained u8, metric_type i32 [, metric_arg f32 if metric_type > 1], then the
    xb vector as u64 element count + raw float32.  [UNVERIFIED against a real Pyserini index -- no faiss
    and no index in the build container; sizes are validated and a mismatch raises.]"""
    with open(path, "rb") as f:
        head = f.read(4 + 4 + 8 + 8 + 8 + 1 + 4)
        fourcc = head[:4]
        if fourcc not in (b"IxFI", b"IxF2", b"IxFl"):
            raise ValueError(f"{path}: not a FAISS IndexFlat file (fourcc {fourcc!r})")
        d, ntotal = struct.unpack_from("<iq", head, 4)
        metric = struct.unpack_from("<i", head, 33)[0]
        if metric > 1:
            f.read(4)
        (count,) = struct.unpack("<Q", f.read(8))
        if count != ntotal * d:
            raise ValueError(f"{path}: vector block has {count} floats, expected {ntotal}*{d}")
        data = np.fromfile(f, dtype=np.float32, count=count)
        if data.size != count:
            raise ValueError(f"{path}: truncated vector block")
    return data.reshape(ntotal, d)


class BGERetriever(BaseRetriever):
    """Dense retrieval over a flat inner-product index held in HBM."""

    name = "BGE"

    def __init__(self, index_name: Optional[str] = None, dataset: str = "nq", encoder_batch_size: Optional[int] = None,
                 use_mps: bool = True, ef_search: Optional[int] = None, *,
                 corpus_embeddings=None, docids: Optional[Sequence[str]] = None,
                 index_dir: Optional[str] = None, encoder: Optional[Callable[[List[str]], np.ndarray]] = None,
                 model_dir: Optional[str] = None, id_offset: int = 0, process_group=None,
                 rescore_fp32: int = 0, **kwargs):
        """
        index_name / dataset / encoder_batch_size / use_mps / ef_search: reference signature
        (src/retrievers/bge.py:41-85); use_mps and ef_search have no meaning here.
        corpus_embeddings: float32/bf16 [N, 768] (numpy or torch) -- this rank's corpus rows.
        docids: the docno of each GLOBAL row (reference: the Pyserini `docid` file, one per line).
        index_dir: a Pyserini flat index directory holding `index` (FAISS IndexFlat) and `docid`.
        encoder / model_dir: the query encoder (see the module docstring); model_dir is used when encoder is None.
        id_offset / process_group: row offset of this rank's shard and the NCCL group to merge over.
        rescore_fp32: margin m > 0 keeps a float32 copy of the corpus rows in HBM (4 bytes x N x 768) and re-scores
                      the bf16 top (k + m) candidates in float32 (qf_dense_rescore_f32), so the returned scores are
                      float32 inner products of the float32 inputs -- what the reference's faiss.IndexFlatIP computes --
                      and bf16 rounding can no longer change membership or order unless a true top-k document falls
                      outside the bf16 top (k + m).  k + m is capped at qf_dense_max_k() = 1024.  Single-GPU only.
        """
        self.model_name = config.models.bge.name
        self.embedding_dim = config.models.bge.embedding_dim
        self.dataset = dataset
        self.encoder_batch_size = encoder_batch_size or config.processing.batch_sizes.bge_encoding
        self.use_hnsw = False
        self.encoder = encoder
        self.id_offset = int(id_offset)
        self.process_group = process_group
        if not torch.cuda.is_available():
            raise RuntimeError("BGERetriever needs a CUDA (sm_100) device: there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device())
        N.check(N.lib().qf_device_check(self.device.index))
        if index_dir is not None:
            index_dir = Path(index_dir)
            if not index_dir.exists():
                raise FileNotFoundError(f"Index not found: {index_dir}")
            with open(index_dir / "docid", "r") as f:
                docids = [line.strip() for line in f]
            corpus_embeddings = read_faiss_flat_index(str(index_dir / "index"))
        if corpus_embeddings is None:
            raise FileNotFoundError("No index found: pass corpus_embeddings=/docids= or index_dir= "
                                    "(Pyserini flat index directory)")
        self.corpus = self.corpus_to_device(corpus_embeddings)
        self.rescore_margin = int(rescore_fp32)
        self.corpus_f32 = None
        if self.rescore_margin > 0:
            if process_group is not None:
                raise ValueError("rescore_fp32 is not offered for the sharded search")
            if isinstance(corpus_embeddings, torch.Tensor):
                if corpus_embeddings.dtype != torch.float32:
                    raise TypeError("rescore_fp32 needs the float32 corpus embeddings")
                self.corpus_f32 = corpus_embeddings.to(self.device).contiguous()
            else:
                self.corpus_f32 = torch.from_numpy(np.ascontiguousarray(corpus_embeddings, dtype=np.float32)).to(self.device)
        if self.encoder is None and model_dir is not None:
            from ..encoder import BertQueryEncoder
            self.encoder = BertQueryEncoder.from_pretrained(model_dir, device=self.device)
        self.docids = list(docids) if docids is not None else [str(i) for i in range(self.id_offset + self.corpus.shape[0])]
        print(f"[BGE] Total vectors on {self.device}: {self.corpus.shape[0]:,} (bf16, {self.corpus.numel() * 2 / 1e9:.2f} GB)")

    def corpus_to_device(self, emb) -> torch.Tensor:
        """fp32 index -> bf16 [N, D] in HBM (once, at load time), in row blocks to bound host/device staging."""
        if isinstance(emb, torch.Tensor):
            if emb.is_cuda and emb.dtype == torch.bfloat16:
                return emb.contiguous()
            out = torch.empty(emb.shape, dtype=torch.bfloat16, device=self.device)
            for s in range(0, emb.shape[0], 1 << 18):
                out[s:s + (1 << 18)] = emb[s:s + (1 << 18)].to(self.device).to(torch.bfloat16)
            return out
        emb = np.asarray(emb)
        out = torch.empty(emb.shape, dtype=torch.bfloat16, device=self.device)
        for s in range(0, emb.shape[0], 1 << 18):
            out[s:s + (1 << 18)] = torch.from_numpy(np.ascontiguousarray(emb[s:s + (1 << 18)], dtype=np.float32)).to(self.device).to(torch.bfloat16)
        return out

    # ---------------------------------------------------------------------------------------- search
    def _encode_queries(self, queries: List[str]) -> np.ndarray:
        if self.encoder is None:
            raise RuntimeError("BGERetriever has no query encoder: pass encoder= (the SentenceTransformer "
                               "BERT forward is outside this package's hot path)")
        return np.asarray(self.encoder(queries), dtype=np.float32)

    def _encode_queries_device(self, queries: List[str]):
        """bf16 rows on the device when the encoder can produce them (BertQueryEncoder), else the host array."""
        enc = getattr(self.encoder, "encode_device", None)
        return enc(queries) if enc is not None else self._encode_queries(queries)

    def _search_device(self, query_embs: torch.Tensor, top_k: int):
        if self.process_group is not None:
            from .. import distributed as qd
            scores, ids, status = qd.dense_search_sharded(query_embs, self.corpus, top_k, self.id_offset, self.process_group)
            if int(status.max().item()) != 0:      # some shard's sampled threshold failed (flags are identical on all
                scores, ids, status = qd.dense_search_sharded(query_embs, self.corpus, top_k, self.id_offset,   # ranks)
                                                              self.process_group, exact=True)
        else:
            scores, ids, status = ops.dense_topk(query_embs, self.corpus, top_k, self.id_offset)
        if int(status.max().item()) != 0:      # cannot happen with the exact fallback enabled; never silently accept
            bad = torch.nonzero(status).flatten()[:5].tolist()
            raise RuntimeError(f"qf_dense_topk: candidate threshold failed for queries {bad}; results would not be exact")
        return scores, ids

    def _search_batch(self, query_embs: np.ndarray, top_k: int) -> Tuple[np.ndarray, np.ndarray]:
        """(ids int64 [n, k], scores float32 [n, k]) sorted by score descending, -1 ids when fewer than k hits
        -- note the (ids, scores) order, as the reference (src/retrievers/bge.py:191-194)."""
        if isinstance(query_embs, torch.Tensor):
            q = query_embs.to(self.device).to(torch.bfloat16).contiguous()
        else:
            q = torch.from_numpy(np.ascontiguousarray(query_embs, dtype=np.float32)).to(self.device).to(torch.bfloat16)
        if self.corpus_f32 is not None:
            q32 = (query_embs.to(self.device).float() if isinstance(query_embs, torch.Tensor)
                   else torch.from_numpy(np.ascontiguousarray(query_embs, dtype=np.float32)).to(self.device))
            k2 = min(int(N.lib().qf_dense_max_k()), int(top_k) + self.rescore_margin)
            _s, cand = self._search_device(q, k2)
            scores, ids = ops.dense_rescore_topk(q32.contiguous(), self.corpus_f32, cand, int(top_k), self.id_offset)
        else:
            scores, ids = self._search_device(q, int(top_k))
        return ids.cpu().numpy().astype(np.int64), scores.cpu().numpy()

    _search_faiss_batch = _search_batch

    def _search(self, query_emb: np.ndarray, top_k: int) -> Tuple[np.ndarray, np.ndarray]:
        """Single query [1, d] -> (ids[k], scores[k]) (reference src/retrievers/bge.py:172-189)."""
        ids, scores = self._search_batch(np.asarray(query_emb).reshape(1, -1), top_k)
        return ids[0], scores[0]

    _search_faiss = _search

    def retrieve(self, query: str, qid: str = "", top_k: int = None, **kwargs) -> RetrieverResult:
        top_k = top_k or config.processing.retrieval.top_k
        start = time.time()
        indices, scores = self._search(self._encode_queries([query]), top_k)
        results = [(self.docids[idx], float(score), rank)
                   for rank, (idx, score) in enumerate(zip(indices, scores), start=1) if 0 <= idx < len(self.docids)]
        return RetrieverResult(qid=qid, results=results, retriever_name=self.name,
                               latency_ms=(time.time() - start) * 1000,
                               metadata={"dataset": self.dataset, "index_type": "flat-ip-bf16-tcgen05"})

    def retrieve_batch(self, queries: Dict[str, str], top_k: int = None, checkpoint_path: Optional[str] = None,
                       mini_batch_size: Optional[int] = None, **kwargs) -> Dict[str, RetrieverResult]:
        """Batch retrieval with the reference's JSONL checkpoint format ({"qid", "results":[{"docno","score",
        "rank"}]}, src/retrievers/bge.py:262-359).  mini_batch_size defaults to the whole pending set: the
        kernel streams the corpus once per 128-query tile, so bigger batches are strictly better."""
        top_k = top_k or config.processing.retrieval.top_k
        completed: Dict[str, RetrieverResult] = {}
        if checkpoint_path and Path(checkpoint_path).exists():
            with open(checkpoint_path, "r") as f:
                for line in f:
                    data = json.loads(line)
                    completed[data["qid"]] = RetrieverResult(
                        qid=data["qid"], results=[(d["docno"], d["score"], d["rank"]) for d in data["results"]],
                        retriever_name=self.name, latency_ms=0, metadata={"dataset": self.dataset})
            print(f"[BGE] Resumed: {len(completed)}/{len(queries)}")
        pending = [(qid, text) for qid, text in queries.items() if qid not in completed]
        if not pending:
            return completed
        mini_batch_size = mini_batch_size or len(pending)
        for i in range(0, len(pending), mini_batch_size):
            batch = pending[i:i + mini_batch_size]
            qids = [q[0] for q in batch]
            t0 = time.time()
            query_embs = self._encode_queries_device([q[1] for q in batch])
            encode_time = time.time() - t0
            t0 = time.time()
            all_ids, all_scores = self._search_batch(query_embs, top_k)
            search_time = time.time() - t0
            print(f"[BGE] Batch {i // mini_batch_size + 1}: encode={encode_time:.2f}s, search={search_time:.3f}s")
            batch_results = []
            for j, qid in enumerate(qids):
                results = [(self.docids[idx], float(score), rank)
                           for rank, (idx, score) in enumerate(zip(all_ids[j], all_scores[j]), start=1)
                           if 0 <= idx < len(self.docids)]
                batch_results.append(RetrieverResult(qid=qid, results=results, retriever_name=self.name, latency_ms=0,
                                                     metadata={"dataset": self.dataset}))
            if checkpoint_path:
                with open(checkpoint_path, "a") as f:
                    for r in batch_results:
                        f.write(json.dumps({"qid": r.qid, "results": [{"docno": d[0], "score": d[1], "rank": d[2]}
                                                                      for d in r.results]}) + "\n")
            for r in batch_results:
                completed[r.qid] = r
        return completed
