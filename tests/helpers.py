"""Shared test helpers: synthetic run lists (seeded) and packing to the [Q, R, K] device layout."""
import numpy as np


def synth_lists(seed, n_q, n_r, k, pool, n_docs=2_681_468, ragged=False, neg_ranker=None, ties=False):
    """-> lists[q][r] = [(doc_id:int, score:float(6 decimals), rank:int)], rank order, distinct docs."""
    rng = np.random.default_rng(seed)
    out = []
    for q in range(n_q):
        cand = rng.choice(n_docs, size=pool, replace=False)
        per_r = []
        for r in range(n_r):
            n = k
            if ragged and rng.random() < 0.3:
                n = int(rng.integers(0, k + 1))
            n = min(n, pool)
            docs = rng.choice(cand, size=n, replace=False)
            scale = [1.0, 30.0, 1.0, 1.0, 12.0][r % 5]
            sc = np.sort(rng.beta(2, 5, size=n) * scale)[::-1]
            if neg_ranker is not None and r == neg_ranker:
                sc = np.sort(-rng.exponential(3.0, size=n))[::-1]
            sc = np.round(sc, 6)
            if ties and n > 8:
                sc[3:7] = sc[3]
            per_r.append([(int(d), float(s), i + 1) for i, (d, s) in enumerate(zip(docs, sc))])
        out.append(per_r)
    return out


def pack_lists(lists, k=None, score_type="f64"):
    """lists[q][r] rows -> (doc_ids int32 [Q,R,K], scores [Q,R,K], ranks int32 [Q,R,K], lens int32 [Q,R])."""
    Q = len(lists)
    R = len(lists[0])
    K = k or max(1, max(len(rows) for per_r in lists for rows in per_r))
    ids = np.full((Q, R, K), -1, dtype=np.int32)
    ranks = np.zeros((Q, R, K), dtype=np.int32)
    lens = np.zeros((Q, R), dtype=np.int32)
    if score_type == "micro":
        sc = np.zeros((Q, R, K), dtype=np.int32)
    elif score_type == "f32":
        sc = np.zeros((Q, R, K), dtype=np.float32)
    else:
        sc = np.zeros((Q, R, K), dtype=np.float64)
    for q, per_r in enumerate(lists):
        for r, rows in enumerate(per_r):
            lens[q, r] = len(rows)
            for i, (d, s, rk) in enumerate(rows):
                ids[q, r, i] = d
                ranks[q, r, i] = rk
                sc[q, r, i] = int(round(s * 1e6)) if score_type == "micro" else s
    return ids, sc, ranks, lens


class LinearWeightModel:
    """Stand-in for the reference's pickled fusion-weight model (src/fusion.py:408-429 only needs `.predict(X)` ->
    [n_queries, n_retrievers]): softmax of a seeded linear map of the QPP feature vector.  Lives here so that the
    pickle written by tools/make_golden_learned.py resolves (`helpers.LinearWeightModel`) in the tests."""

    def __init__(self, n_features: int, n_retrievers: int, seed: int = 0):
        self.W = np.random.default_rng(seed).standard_normal((n_features, n_retrievers))

    def predict(self, X):
        z = np.asarray(X, dtype=np.float64) @ self.W
        z = z - z.max(axis=1, keepdims=True)
        e = np.exp(z)
        return e / e.sum(axis=1, keepdims=True)
