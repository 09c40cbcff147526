#!/usr/bin/env python3
"""
Generate tests/golden/fusion_weighted.json.gz by running the REAL reference
(/root/reference/src/fusion.py: weighted_combsum / weighted_combmnz / weighted_rrf, get_qpp_weight,
write_runfile) on small seeded synthetic runs.  Run in the build container only:

    python tools/make_golden_fusion.py

The fixture stores inputs (rows per ranker, qpp_data) and the reference's outputs (first-seen-order
lists with float64 scores serialised by repr -> exact round trip, plus the sha256 / head of
write_runfile's output).
"""
import gzip
import hashlib
import io
import json
import os
import sys
import contextlib
import tempfile

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_import import import_reference_fusion  # noqa: E402

RANKERS = ["BGE", "BM25", "BM25_MonoT5", "BM25_TCT", "Splade"]


def make_case(seed, n_q, k, pool, qid_fmt, adversarial):
    rng = np.random.default_rng(seed)
    rows = {r: [] for r in RANKERS}
    qids = [qid_fmt.format(i) for i in range(n_q)]
    for qi, qid in enumerate(qids):
        cand = rng.choice(100000, size=pool, replace=False)
        for ri, ranker in enumerate(RANKERS):
            n = k
            if adversarial:
                if qi == 1 and ri == 2:
                    continue                      # query absent from one ranker
                if qi == 2:
                    n = int(rng.integers(0, k))   # ragged (possibly empty) lists
                if qi == 3 and ri == 0:
                    n = 1                         # single-doc list
            docs = rng.choice(cand, size=min(n, pool), replace=False)
            if ranker == "BM25":
                sc = np.sort(rng.beta(2, 5, size=len(docs)) * 30.0)[::-1]
            elif ranker == "BM25_MonoT5":
                sc = np.sort(-rng.exponential(3.0, size=len(docs)))[::-1]   # negative log-probs
            else:
                sc = np.sort(rng.beta(2, 5, size=len(docs)))[::-1]
            sc = np.round(sc, 6)
            if adversarial and qi == 4 and len(sc) > 6:
                sc[2:6] = sc[2]                   # tied scores
                if ri == 1:
                    sc[:] = 0.5                   # all-equal list
                if ri == 3:
                    sc[-3:] = 0.0                 # exact zeros
            ranks = np.arange(1, len(docs) + 1)
            if adversarial and qi == 5 and len(docs) > 8:
                if ri == 0:
                    docs[7] = docs[1]             # duplicate docno inside one list
                if ri == 4:
                    perm = rng.permutation(len(docs))   # file order != rank order
                    docs, sc, ranks = docs[perm], sc[perm], ranks[perm]
            for d, s, rk in zip(docs, sc, ranks):
                rows[ranker].append((qid, "Q0", f"doc{int(d)}", int(rk), float(s), ranker))
    qpp = {}
    for qid in qids:
        qpp[qid] = {}
        for ranker in RANKERS:
            qpp[qid][ranker] = [float(x) for x in np.round(rng.random(13), 6)]
    if adversarial:
        qpp[qids[0]]["BM25"][5] = 0.0             # a zero weight
    return qids, rows, qpp


def to_frames(rows):
    frames = {}
    for ranker, rr in rows.items():
        df = pd.DataFrame(rr, columns=["qid", "iter", "docno", "rank", "score", "runid"])
        df["qid"] = df["qid"].astype(str)
        df["docno"] = df["docno"].astype(str)
        df["rank"] = df["rank"].astype("int64")
        df["score"] = df["score"].astype("float64")
        frames[ranker] = df
    return frames


def main():
    fusion = import_reference_fusion()
    out = {"rankers": RANKERS, "cases": []}
    specs = [
        dict(name="small_adversarial", seed=11, n_q=7, k=20, pool=45, qid_fmt="test{}", adversarial=True),
        dict(name="scifact_like", seed=12, n_q=5, k=100, pool=220, qid_fmt="{}", adversarial=False),
        dict(name="deep_k300", seed=13, n_q=2, k=300, pool=660, qid_fmt="test{}", adversarial=False),
    ]
    for spec in specs:
        name = spec.pop("name")
        qids, rows, qpp = make_case(**spec)
        frames = to_frames(rows)
        case = {"name": name, "rows": rows, "qpp": qpp, "results": {}}
        for qpp_index in (5, -1, 2):
            for mname, fn, kw in (
                ("wcombsum", fusion.weighted_combsum, {}),
                ("wcombmnz", fusion.weighted_combmnz, {}),
                ("wrrf", fusion.weighted_rrf, {"k": 60}),
                ("wrrf_k10", fusion.weighted_rrf, {"k": 10}),
            ):
                if mname == "wrrf_k10" and qpp_index != 5:
                    continue
                with contextlib.redirect_stdout(io.StringIO()):
                    fused = fn(frames, qpp, qpp_index, **kw)
                    with tempfile.NamedTemporaryFile("r", suffix=".res") as tf:
                        fusion.write_runfile(fused, tf.name, tag=f"{mname}-x")
                        lines = open(tf.name).read().splitlines()
                case["results"][f"{mname}@{qpp_index}"] = {
                    "fused": {q: [[d, repr(float(s))] for d, s in lst] for q, lst in fused.items()},
                    "runfile_sha256": hashlib.sha256("\n".join(lines).encode()).hexdigest(),
                    "runfile_head": lines[:25],
                    "runfile_n": len(lines),
                }
        out["cases"].append(case)
    path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fusion_weighted.json.gz")
    with gzip.open(path, "wt", compresslevel=9) as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", os.path.abspath(path), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
