#!/usr/bin/env python3
"""
Generate tests/golden/qpp_textcols.json.gz from the reference's COMMITTED artefacts
(data/{scifact,nq}/qpp/*.res.mmnorm.qpp + data/*/BEIR-*/queries.jsonl).  Build container only.

These files are the only known-answer vectors the reference holds for the QPP path: their columns 7
(MaxIDF proxy) and 8 (avgidf proxy) depend only on the query text, so they can be checked although the
.res inputs are Git-LFS pointers.  The fixture keeps, per dataset: the qids in FILE order, the query
texts, the two columns as the printed 6-decimal strings, and which columns are identical (11 == 12).
"""
import gzip
import json
import os

REF = "/root/reference/data"
RANKERS = ["BGE", "BM25", "BM25_MonoT5", "BM25_TCT", "Splade"]


def main():
    out = {}
    for ds in ("scifact", "nq"):
        texts = {}
        with open(f"{REF}/{ds}/BEIR-{ds}/queries.jsonl") as f:
            for line in f:
                q = json.loads(line)
                texts[q["_id"]] = q["text"]
        per_ranker = {}
        for rk in RANKERS:
            rows = []
            with open(f"{REF}/{ds}/qpp/{rk}.res.mmnorm.qpp") as f:
                for line in f:
                    parts = line.rstrip("\n").split("\t")
                    rows.append(parts)
            per_ranker[rk] = rows
        base = per_ranker["BGE"]
        for rk in RANKERS:
            rows = per_ranker[rk]
            assert [r[0] for r in rows] == [r[0] for r in base], (ds, rk, "qid order differs")
            assert all(r[8] == b[8] and r[9] == b[9] for r, b in zip(rows, base)), (ds, rk)
            assert all(r[12] == r[13] for r in rows), (ds, rk, "col 11 != col 12")
            assert all(len(r) == 14 for r in rows)
        qids = [r[0] for r in base]
        out[ds] = {
            "qids": qids,
            "texts": [texts[q] for q in qids],
            "col7": [r[8] for r in base],
            "col8": [r[9] for r in base],
        }
        print(ds, len(qids), "rows")
    path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "qpp_textcols.json.gz")
    with gzip.open(path, "wt", compresslevel=9) as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", os.path.abspath(path), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
