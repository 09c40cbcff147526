#!/usr/bin/env python3
"""
Generate tests/golden/fusion_learned.json.gz by running the REAL reference learned_fusion
(/root/reference/src/fusion.py:387-448) on small seeded synthetic runs with a stand-in weight model
(tests/helpers.py::LinearWeightModel -- the reference only calls `.predict(X)` on the unpickled object).
Build container only:  python tools/make_golden_learned.py

Cases: model knows all 5 rankers; model knows 4 of the 5 (the fifth gets the reference's 1/n_retrievers default,
:435); a (qid, ranker) pair missing from qpp_data (zero features, :420-423).  Scores are stored by repr (exact).
"""
import contextlib
import gzip
import io
import json
import os
import pickle
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
from helpers import LinearWeightModel  # noqa: E402
from make_golden_fusion import RANKERS, make_case, to_frames  # noqa: E402
from ref_import import import_reference_fusion  # noqa: E402


def main():
    fusion = import_reference_fusion()
    out = {"rankers": RANKERS, "cases": []}
    specs = [
        dict(name="all_rankers", seed=21, n_q=6, k=30, pool=70, qid_fmt="test{}", adversarial=True, model_rankers=RANKERS),
        dict(name="four_of_five", seed=22, n_q=4, k=100, pool=230, qid_fmt="{}", adversarial=False,
             model_rankers=["Splade", "BGE", "BM25", "BM25_TCT"]),
    ]
    for spec in specs:
        name, model_rankers = spec.pop("name"), spec.pop("model_rankers")
        qids, rows, qpp = make_case(**spec)
        if name == "all_rankers":
            del qpp[qids[2]]["BM25"]                       # features of a missing (qid, ranker) stay zero
        frames = to_frames(rows)
        model_seed = spec["seed"] + 100
        blob = {"model": LinearWeightModel(13 * len(model_rankers), len(model_rankers), model_seed),
                "model_type": "PerRetrieverLGBM", "retrievers": model_rankers, "n_qpp": 13}
        with tempfile.NamedTemporaryFile("wb", suffix=".pkl", delete=False) as tf:
            pickle.dump(blob, tf)
        with contextlib.redirect_stdout(io.StringIO()):
            fused = fusion.learned_fusion(frames, qpp, tf.name)
        os.unlink(tf.name)
        out["cases"].append({"name": name, "rows": rows, "qpp": qpp, "model_rankers": model_rankers,
                             "model_seed": model_seed,
                             "fused": {q: [[d, repr(float(s))] for d, s in lst] for q, lst in fused.items()}})
    path = os.path.join(HERE, "..", "tests", "golden", "fusion_learned.json.gz")
    with gzip.open(path, "wt", compresslevel=9) as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", os.path.abspath(path), os.path.getsize(path), "bytes",
          [(c["name"], len(c["fused"])) for c in out["cases"]])


if __name__ == "__main__":
    main()
