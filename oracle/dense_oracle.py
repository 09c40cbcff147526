"""
TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's flat inner-product search.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker.  The product path never imports it.

Restates BGERetriever._search_faiss_batch (reference src/retrievers/bge.py:191-194): faiss.IndexFlatIP
.search(query_embs, top_k) = exact inner product of every query with every corpus row + top-k, returned
as (ids, scores) -- note the reference swaps FAISS's (scores, ids) order.  faiss-cpu (requirements.txt:10,
">=1.7.4", not vendored, not installed here) computes sgemm blocks in fp32 and keeps per-query heaps; its
tie order is implementation-defined.  This restatement ranks by (score descending, id ascending).

PARITY UNPINNED: the reference has no test or golden output for this call (data/*/runs/BGE.res are
Git-LFS pointers) and faiss is not importable in the build container.
"""
import numpy as np


def flat_ip_topk(queries: np.ndarray, corpus: np.ndarray, k: int, block: int = 65536):
    """fp32 exact search.  -> (ids int64 [Q, k] (-1 padded), scores fp32 [Q, k]), FAISS conventions."""
    q = np.ascontiguousarray(queries, dtype=np.float32)
    Q, N = q.shape[0], corpus.shape[0]
    kk = min(k, N)
    best_s = np.full((Q, 0), 0, dtype=np.float32)
    best_i = np.full((Q, 0), 0, dtype=np.int64)
    for b0 in range(0, N, block):
        c = np.ascontiguousarray(corpus[b0:b0 + block], dtype=np.float32)
        s = q @ c.T
        ids = np.broadcast_to(np.arange(b0, b0 + c.shape[0], dtype=np.int64), s.shape)
        s = np.concatenate([best_s, s], axis=1)
        ids = np.concatenate([best_i, ids], axis=1)
        order = np.lexsort((ids, -s), axis=1)[:, :kk]          # score desc, then id asc
        best_s = np.take_along_axis(s, order, axis=1)
        best_i = np.take_along_axis(ids, order, axis=1)
    out_s = np.full((Q, k), -np.inf, dtype=np.float32)
    out_i = np.full((Q, k), -1, dtype=np.int64)
    out_s[:, :kk] = best_s
    out_i[:, :kk] = best_i
    return out_i, out_s


def merge_topk(scores: np.ndarray, ids: np.ndarray, k: int):
    """[G, Q, k] per-shard results -> merged [Q, k] by (score desc, id asc); ids < 0 are padding."""
    G, Q, kk = scores.shape
    s = np.transpose(scores, (1, 0, 2)).reshape(Q, G * kk).astype(np.float32).copy()
    i = np.transpose(ids, (1, 0, 2)).reshape(Q, G * kk).astype(np.int64).copy()
    s[i < 0] = -np.inf
    key_i = np.where(i < 0, np.iinfo(np.int64).max, i)
    order = np.lexsort((key_i, -s), axis=1)[:, :k]
    return np.take_along_axis(s, order, axis=1), np.take_along_axis(i, order, axis=1)
