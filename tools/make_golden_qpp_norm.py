#!/usr/bin/env python3
"""
Generate tests/golden/qpp_normalize.json.gz with the REAL reference _normalize_qpp (/root/reference/src/qpp.py:340-374;
the module imports here once src/__init__ is bypassed, tools/ref_import.py) on seeded [Q, 13] tables: both modes, a
constant column (range 0 / std 0 -> scale 1.0), a single-query table.  Values stored by repr (exact round trip).
Build container only:  python tools/make_golden_qpp_norm.py
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import import_reference_fusion  # noqa: E402


def main():
    import_reference_fusion()
    from src import qpp as rq  # type: ignore
    out = {"cases": []}
    for seed, nq in ((31, 60), (32, 1), (33, 7)):
        rng = np.random.default_rng(seed)
        table = {f"q{i}": [float(x) for x in np.round(rng.standard_normal(13) * rng.random(13) * 10, 6)] for i in range(nq)}
        for row in table.values():
            row[11] = row[12] = 1.5
        case = {"seed": seed, "table": {q: [repr(v) for v in row] for q, row in table.items()}}
        for mode in ("minmax", "zscore"):
            res = rq._normalize_qpp(table, mode)
            case[mode] = {q: [repr(float(v)) for v in row] for q, row in res.items()}
        out["cases"].append(case)
    path = os.path.join(HERE, "..", "tests", "golden", "qpp_normalize.json.gz")
    with gzip.open(path, "wt", compresslevel=9) as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", os.path.abspath(path), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
