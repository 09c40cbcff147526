"""GPU fuzz: random shapes / options through the qf_fuse kernel paths (gather, streaming, each with the larger table;
general; qf_fuse_deep) must agree bit for bit; a sample of the queries is also checked against the Python oracle."""
import numpy as np
import pytest
import torch

from oracle import fusion_oracle as fo

pytestmark = pytest.mark.gpu

PATHS = {"stream": 0, "stream-large": 8, "gather": 0x400, "gather-large": 0x408, "general": 0x200, "deep": 0x800}


def _random_case(rng):
    R = int(rng.integers(1, 17))
    K = int(rng.choice([7, 16, 50, 100, 128, 257, 1000]))
    while R * K > 5120:
        R = max(1, R - 1)
    Q = int(rng.integers(3, 40))
    pool = max(K, int(K * rng.uniform(1.0, 3.0)))
    score_type = rng.choice(["f64", "f32", "micro"])
    ids = np.full((Q, R, K), -1, dtype=np.int32)
    lens = np.zeros((Q, R), dtype=np.int32)
    sc = np.zeros((Q, R, K), dtype=np.float64)
    for q in range(Q):
        cand = rng.choice(3_000_000, size=pool, replace=False)
        for r in range(R):
            n = K if rng.random() < 0.7 else int(rng.integers(0, K + 1))
            lens[q, r] = n
            ids[q, r, :n] = rng.choice(cand, size=n, replace=False)
            kind = rng.integers(0, 4)
            if kind == 0:
                v = rng.beta(2, 5, size=n) * rng.choice([1.0, 30.0])
            elif kind == 1:
                v = -rng.exponential(3.0, size=n)
            elif kind == 2:
                v = np.round(rng.random(n), 2)                      # many exact ties
            else:
                v = rng.normal(size=n)
            sc[q, r, :n] = np.round(np.sort(v)[::-1], 6)
    if score_type == "micro":
        dev_sc = np.round(sc * 1e6).astype(np.int32)
        sc = dev_sc / 1e6
    elif score_type == "f32":
        dev_sc = sc.astype(np.float32)
        sc = dev_sc.astype(np.float64)
    else:
        dev_sc = sc
    weighted = rng.random() < 0.6
    w = np.round(rng.random((Q, R)) * rng.choice([1.0, 1.0, 5.0]), 6) if weighted else None
    if weighted and rng.random() < 0.3:
        w[rng.integers(0, Q), rng.integers(0, R)] = 0.0
    method = int(rng.integers(0, 3))
    norm = int(rng.integers(0, 2)) if (not weighted and method != 2) else 0
    flags = (2 if (not weighted) else 0) | (1 if (method == 1 and not weighted and R <= 8 and rng.random() < 0.5) else 0)
    if rng.random() < 0.15:
        flags |= 4                                                    # unranked output
    return dict(ids=ids, lens=lens, sc=sc, dev_sc=dev_sc, w=w, method=method, norm=norm, flags=flags, R=R, K=K, Q=Q)


@pytest.mark.parametrize("seed", range(6))
def test_kernel_paths_agree_and_match_the_oracle(seed):
    from qppfusion import ops
    rng = np.random.default_rng(1000 + seed)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for _ in range(8):
        c = _random_case(rng)
        outs = {}
        for name, extra in PATHS.items():
            o = ops.fuse(t(c["ids"]), t(c["dev_sc"]), t(c["lens"]), None if c["w"] is None else t(c["w"]), None,
                         c["method"], c["norm"], 60.0, c["flags"] | extra)
            torch.cuda.synchronize()
            outs[name] = [x.cpu().numpy() for x in o]
        ref = outs["general"]
        n = ref[4]
        assert (n >= 0).all()
        m = np.arange(c["R"] * c["K"])[None, :] < n[:, None]
        for name in ("gather", "gather-large", "stream", "stream-large", "deep"):
            got = outs[name]
            desc = (name, {k: c[k] for k in ("R", "K", "Q", "method", "norm", "flags")})
            assert np.array_equal(got[4], n) and np.array_equal(got[5], ref[5]), desc
            assert np.array_equal(got[0][m], ref[0][m]), desc
            assert np.array_equal(got[1][m].view(np.int64), ref[1][m].view(np.int64)), desc
            assert np.array_equal(got[2][m], ref[2][m]) and np.array_equal(got[3][m], ref[3][m]), desc
        # oracle on a few queries
        for q in rng.choice(c["Q"], size=min(3, c["Q"]), replace=False):
            per_r = [[(int(c["ids"][q, r, i]), float(c["sc"][q, r, i]), i + 1) for i in range(c["lens"][q, r])] for r in range(c["R"])]
            if c["w"] is None:
                docs, scores, counts = fo.fuse_ranx_query(per_r, c["method"], 60.0, "min-max" if c["norm"] else None,
                                                          bool(c["flags"] & 1))
            else:
                docs, scores, counts = fo.fuse_weighted_query(per_r, c["w"][q], c["method"], 60.0)
            order = list(range(len(docs))) if (c["flags"] & 4) else fo.rank_order(scores)
            k = len(docs)
            assert n[q] == k
            assert [int(x) for x in ref[0][q, :k]] == [docs[i] for i in order]
            assert np.array_equal(ref[1][q, :k], np.array([scores[i] for i in order], dtype=np.float64) + 0.0)
            assert [int(x) for x in ref[2][q, :k]] == [counts[i] for i in order]
