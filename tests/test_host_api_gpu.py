"""GPU tests through the reference-facing Python API (fusion.py / qpp.py / retrievers/bge.py drop-ins)."""
import gzip
import hashlib
import json
import os

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import dense_oracle as do
from oracle import fusion_oracle as fo
from oracle import qpp_oracle as qo
from test_host_api import GOLD, gold_frames

pytestmark = pytest.mark.gpu


def test_weighted_functions_equal_the_real_reference_outputs(tmp_path):
    """weighted_combsum / weighted_combmnz / weighted_rrf on DataFrames: same dict (keys, list order, float64
    bits) and same write_runfile bytes as the REAL reference produced for tests/golden/."""
    from qppfusion import fusion
    with gzip.open(GOLD, "rt") as f:
        gold = json.load(f)
    fns = {"wcombsum": lambda r, q, i: fusion.weighted_combsum(r, q, i),
           "wcombmnz": lambda r, q, i: fusion.weighted_combmnz(r, q, i),
           "wrrf": lambda r, q, i: fusion.weighted_rrf(r, q, i, k=60),
           "wrrf_k10": lambda r, q, i: fusion.weighted_rrf(r, q, i, k=10)}
    for case in gold["cases"]:
        frames = gold_frames(case, gold["rankers"])
        for key, res in case["results"].items():
            mname, idx = key.split("@")
            fused = fns[mname](frames, case["qpp"], int(idx))
            ref = {q: [(d, float(s) + 0.0) for d, s in lst] for q, lst in res["fused"].items()}
            assert list(fused.keys()) == list(ref.keys())
            for q in ref:
                assert fused[q] == ref[q], (case["name"], key, q)
            out = tmp_path / "o.res"
            fusion.write_runfile(fused, str(out), tag=f"{mname}-x")
            assert hashlib.sha256(out.read_text().rstrip("\n").encode()).hexdigest() == res["runfile_sha256"]


def test_unweighted_functions_equal_the_ranx_restatement():
    from qppfusion import fusion
    with gzip.open(GOLD, "rt") as f:
        gold = json.load(f)
    case = gold["cases"][0]
    frames = gold_frames(case, gold["rankers"])
    runs = {rk: {} for rk in gold["rankers"]}
    for rk in gold["rankers"]:
        for (qid, _i, docno, rank, score, _t) in case["rows"][rk]:
            runs[rk].setdefault(qid, []).append((docno, float(score), int(rank)))
    for name, fn, method in (("combsum", fusion.combsum, fo.SUM), ("combmnz", fusion.combmnz, fo.MNZ),
                             ("rrf", lambda r: fusion.rrf(r, k=60), fo.RRF)):
        got = fn(frames)
        exp = fo.ranx_fusion(runs, method, 60)
        assert set(got.keys()) == set(exp.keys())
        for q in exp:
            assert got[q] == [(d, s + 0.0) for d, s in exp[q]], (name, q)


def test_missing_qpp_weight_raises_keyerror():
    from qppfusion import fusion
    df = pd.DataFrame({"qid": ["1"], "iter": "Q0", "docno": ["d"], "rank": [1], "score": [0.5], "runid": "A"})
    with pytest.raises(KeyError):
        fusion.weighted_combsum({"A": df, "B": df.iloc[:0]}, {"1": {"A": [0.1] * 13}}, 5)


def test_run_fusion_end_to_end_files(tmp_path):
    from qppfusion import fusion
    from qppfusion.qpp import compute_qpp_for_res_file
    rng = np.random.default_rng(5)
    runs_dir, qpp_dir = tmp_path / "runs", tmp_path / "qpp"
    runs_dir.mkdir(); qpp_dir.mkdir()
    rankers = ["BGE", "BM25", "Splade"]
    raw = {}
    for rk in rankers:
        lines, nlines = [], []
        for q in range(12):
            docs = rng.choice(400, size=60, replace=False)
            sc = np.sort(rng.random(60) * (30 if rk == "BM25" else 1))[::-1]
            nsc = (sc - sc.min()) / (sc.max() - sc.min())
            for i, (d, s, ns) in enumerate(zip(docs, sc, nsc)):
                lines.append(f"test{q} Q0 doc{d} {i+1} {s:.6f} {rk}")
                nlines.append(f"test{q} Q0 doc{d} {i+1} {ns:.6f} {rk}")
        (runs_dir / f"{rk}.res").write_text("\n".join(lines) + "\n")
        (runs_dir / f"{rk}.norm.res").write_text("\n".join(nlines) + "\n")
        raw[rk] = lines
    (tmp_path / "queries.jsonl").write_text("".join(json.dumps({"_id": f"test{q}", "text": f"what is topic number {q} about"}) + "\n" for q in range(12)))
    # 03_qpp
    for rk in rankers:
        res = compute_qpp_for_res_file(str(runs_dir / f"{rk}.res"), str(qpp_dir / f"{rk}.res.mmnorm.qpp"),
                                       queries_path=str(tmp_path / "queries.jsonl"))
        per_q = {}
        for ln in raw[rk]:
            p = ln.split()
            per_q.setdefault(p[0], []).append(float(p[4]))
        exp = qo.qpp_for_run(per_q, {f"test{q}": f"what is topic number {q} about" for q in range(12)}, 100, "minmax")
        for q in exp:
            assert np.allclose(res[q], exp[q], rtol=1e-9, atol=1e-12)
        got_lines = (qpp_dir / f"{rk}.res.mmnorm.qpp").read_text().splitlines()
        assert got_lines == qo.qpp_file_lines(exp)
    # scripts/03_qpp.py: the whole runs directory in one call -- same files, skip / force behaviour, query auto-detection
    from qppfusion.qpp import compute_qpp_for_runs_dir
    (tmp_path / "BEIR-x").mkdir()
    (tmp_path / "BEIR-x" / "queries.jsonl").write_text((tmp_path / "queries.jsonl").read_text())
    qpp_dir2 = tmp_path / "qpp"                                         # the default location next to runs/
    before = {rk: (qpp_dir / f"{rk}.res.mmnorm.qpp").read_text() for rk in rankers}
    assert compute_qpp_for_runs_dir(runs_dir) == {}                     # everything exists already: skipped
    for rk in rankers:
        (qpp_dir2 / f"{rk}.res.mmnorm.qpp").unlink()
    done = compute_qpp_for_runs_dir(runs_dir)
    assert sorted(done) == [f"{rk}.res" for rk in rankers]
    assert {rk: (qpp_dir / f"{rk}.res.mmnorm.qpp").read_text() for rk in rankers} == before
    assert sorted(compute_qpp_for_runs_dir(runs_dir, force=True)) == [f"{rk}.res" for rk in rankers]
    # 05_fusion
    out = tmp_path / "fused.res"
    fused = fusion.run_fusion("wcombsum", str(runs_dir), qpp_dir=str(qpp_dir), qpp_model="RSD", output_path=str(out))
    runs = {rk: {} for rk in rankers}
    for rk in rankers:
        for ln in (runs_dir / f"{rk}.norm.res").read_text().splitlines():
            p = ln.split()
            runs[rk].setdefault(p[0], []).append((p[2], float(p[4]), int(p[3])))
    qpp_data = fusion.load_qpp_scores(str(qpp_dir))
    exp = fo.weighted_fusion(runs, qpp_data, fo.SUM, 5)
    assert fused == {q: [(d, s + 0.0) for d, s in l] for q, l in exp.items()}
    assert out.read_text().splitlines() == fo.runfile_lines(exp, "wcombsum-rsd")
    for m, tag in (("combsum", "combsum"), ("rrf", "rrf-k60"), ("w-rrf", "wrrf-rsd")):
        fusion.run_fusion(m, str(runs_dir), qpp_dir=str(qpp_dir), output_path=str(out))
        assert out.read_text().splitlines()[0].endswith(tag)
    # scripts/05_fusion.py::run_all_methods: one parse, six methods, the reference's file names and tags
    fused_dir = tmp_path / "fused"
    res = fusion.run_all_methods(runs_dir, qpp_dir, fused_dir, qpp_model="RSD", rrf_k=60)
    assert list(res.keys()) == ["combsum", "combmnz", "rrf", "wcombsum", "wcombmnz", "wrrf"]
    assert sorted(p.name for p in fused_dir.iterdir()) == ["combmnz.res", "combsum.res", "rrf.res", "wcombmnz_rsd.res",
                                                          "wcombsum_rsd.res", "wrrf_rsd.res"]
    strip = lambda text: [ln.rsplit(" ", 1)[0] for ln in text.splitlines()]
    for m, name, tag in (("combsum", "combsum.res", "combsum"), ("combmnz", "combmnz.res", "combmnz"), ("rrf", "rrf.res", "rrf"),
                         ("wcombsum", "wcombsum_rsd.res", "wcombsum-rsd"), ("wcombmnz", "wcombmnz_rsd.res", "wcombmnz-rsd"),
                         ("wrrf", "wrrf_rsd.res", "wrrf-rsd")):
        fusion.run_fusion(m, str(runs_dir), qpp_dir=str(qpp_dir), qpp_model="RSD", output_path=str(out), rrf_k=60)
        got = (fused_dir / name).read_text()
        assert strip(got) == strip(out.read_text()), m
        assert all(ln.endswith(" " + tag) for ln in got.splitlines()), m
    res2 = fusion.run_all_methods(runs_dir, tmp_path / "no_qpp", tmp_path / "fused2")
    assert list(res2.keys()) == ["combsum", "combmnz", "rrf"]


def test_qpp_bridge_api():
    from qppfusion.qpp import QPPBridge, RESEARCH_COLUMNS, default_rsd_perms
    b = QPPBridge()
    sc = [0.95, 0.82, 0.71, 0.65, 0.58, 0.52, 0.48, 0.41, 0.35, 0.28]          # the reference's --test vector
    r = b.compute("what is machine learning", sc, retriever_name="BGE")
    exp = dict(zip(qo.BRIDGE_METHODS, qo.bridge_scores("what is machine learning", sc)))
    assert list(r.qpp_scores.keys()) == qo.BRIDGE_METHODS and r.retriever_name == "BGE"
    for k in exp:
        assert r.qpp_scores[k] == pytest.approx(exp[k], rel=1e-9, abs=1e-12)
    rs = b.compute_batch([("q one", sc), ("q two words", []), ("third", sc[:1])])
    assert [x.query for x in rs] == ["q one", "q two words", "third"] and rs[1].qpp_scores == {}
    assert rs[0].retriever_name == "batch" and rs[2].qpp_scores["nqc"] == 0.0
    perms = default_rsd_perms(10)
    ext = b.research_batch([("what is machine learning", sc)], query_idf=[4.0], perms=perms)[0]
    feats = qo.text_features("what is machine learning")
    e = qo.research_scores(sc, feats[1], 4.0, [p.tolist() for p in perms])
    for name, v in zip(RESEARCH_COLUMNS, e):
        assert ext[name] == pytest.approx(v, rel=1e-9, abs=1e-12)


def test_bge_retriever_search_calls(tmp_path):
    from qppfusion.retrievers.bge import BGERetriever
    rng = np.random.default_rng(11)
    N, D = 6000, 768
    c = rng.standard_normal((N, D)).astype(np.float32)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    vocab = {}

    def encoder(texts):
        out = []
        for t in texts:
            r = np.random.default_rng(abs(hash(t)) % (2 ** 32)).standard_normal(D).astype(np.float32)
            out.append(r / np.linalg.norm(r))
        return np.stack(out)

    ret = BGERetriever(corpus_embeddings=c, docids=[f"doc{i}" for i in range(N)], encoder=encoder)
    q = encoder(["alpha", "beta", "gamma"])
    ids, scores = ret._search_batch(q, 100)
    assert ids.dtype == np.int64 and scores.dtype == np.float32 and ids.shape == (3, 100)
    cb = torch.from_numpy(c).to(torch.bfloat16).float().numpy()
    qb = torch.from_numpy(q).to(torch.bfloat16).float().numpy()
    oi, os_ = do.flat_ip_topk(qb, cb, 100)
    assert np.abs(scores - os_).max() <= 2e-5
    assert np.mean(ids == oi) > 0.98
    i1, s1 = ret._search(q[:1], 10)
    assert i1.shape == (10,) and np.array_equal(i1, ids[0, :10])
    ck = tmp_path / "BGE.checkpoint.jsonl"
    res = ret.retrieve_batch({"q1": "alpha", "q2": "beta"}, top_k=5, checkpoint_path=str(ck))
    assert [d for d, _, _ in res["q1"].results] == [f"doc{i}" for i in ids[0, :5]]
    assert [r for _, _, r in res["q2"].results] == [1, 2, 3, 4, 5]
    lines = [json.loads(l) for l in ck.read_text().splitlines()]
    assert lines[0]["qid"] == "q1" and set(lines[0]["results"][0].keys()) == {"docno", "score", "rank"}
    res2 = ret.retrieve_batch({"q1": "alpha", "q2": "beta", "q3": "gamma"}, top_k=5, checkpoint_path=str(ck))
    assert len(res2) == 3 and res2["q1"].results[0][0] == res["q1"].results[0][0]
    one = ret.retrieve("alpha", "q1", top_k=5)
    assert [d for d, _, _ in one.results] == [d for d, _, _ in res["q1"].results]


def test_tct_colbert_retriever_uses_the_same_flat_ip_search(tmp_path):
    """TCT-ColBERT's faiss.IndexFlatIP search (reference src/retrievers/tct_colbert.py:139-151,163,194) through the
    dense kernel: embedding cache files in the reference's format, result tuples, checkpoint."""
    from qppfusion.retrievers import TCTColBERTRetriever
    rng = np.random.default_rng(12)
    N, D = 5000, 768
    c = rng.standard_normal((N, D)).astype(np.float32)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    corpus = {f"d{i}": {"text": "x", "title": ""} for i in range(N)}
    np.save(str(tmp_path / f"tct_colbert_{N}_fp16.npy"), c.astype(np.float16))
    np.save(str(tmp_path / f"tct_colbert_{N}_ids.npy"), np.array(list(corpus.keys()), dtype=object))

    def encoder(texts):
        out = []
        for t in texts:
            r = np.random.default_rng(abs(hash(t)) % (2 ** 32)).standard_normal(D).astype(np.float32)
            out.append(r / np.linalg.norm(r))
        return np.stack(out)

    ret = TCTColBERTRetriever(corpus, cache_dir=str(tmp_path), encoder=encoder)
    assert ret.name == "TCT-ColBERT"
    q = encoder(["alpha", "beta"])
    res = ret._process_mini_batch([("q1", "alpha"), ("q2", "beta")], 50)
    c16 = torch.from_numpy(c.astype(np.float16).astype(np.float32)).to(torch.bfloat16).float().numpy()
    qb = torch.from_numpy(q).to(torch.bfloat16).float().numpy()
    oi, os_ = do.flat_ip_topk(qb, c16, 50)
    got_ids = [[int(d[1:]) for d, _, _ in r.results] for r in res]
    assert np.mean(np.array(got_ids) == oi) > 0.98
    assert np.abs(np.array([[s for _, s, _ in r.results] for r in res]) - os_).max() <= 2e-5
    assert [rk for _, _, rk in res[0].results] == list(range(1, 51))
    ck = tmp_path / "TCT.checkpoint.jsonl"
    out = ret.retrieve_batch({"q1": "alpha", "q2": "beta"}, top_k=5, checkpoint_path=str(ck))
    assert [d for d, _, _ in out["q1"].results] == [d for d, _, _ in res[0].results[:5]]
    with pytest.raises(FileNotFoundError):        # no cache and no encoder to compute one with
        TCTColBERTRetriever(corpus, cache_dir=str(tmp_path / "missing"))


def test_learned_fusion_equals_the_real_reference_outputs(tmp_path):
    """qppfusion.fusion.learned_fusion (weights from a pickled model -> qf_fuse W-CombSUM) against the outputs of the
    REAL reference learned_fusion (tests/golden/fusion_learned.json.gz): same qids, same docno-ascending lists, scores
    within 2 ulp (the reference sums with pandas' Kahan-compensated groupby, the kernel sequentially in list order)."""
    import os
    import pickle
    import numpy as np
    from helpers import LinearWeightModel
    from qppfusion import fusion
    path = os.path.join(os.path.dirname(GOLD), "fusion_learned.json.gz")
    with gzip.open(path, "rt") as f:
        gold = json.load(f)
    for case in gold["cases"]:
        frames = gold_frames(case, gold["rankers"])
        mr = case["model_rankers"]
        pkl = tmp_path / f"{case['name']}.pkl"
        with open(pkl, "wb") as f:
            pickle.dump({"model": LinearWeightModel(13 * len(mr), len(mr), case["model_seed"]),
                         "model_type": "PerRetrieverLGBM", "retrievers": mr, "n_qpp": 13}, f)
        fused = fusion.learned_fusion(frames, case["qpp"], str(pkl))
        ref = case["fused"]
        assert sorted(fused.keys()) == sorted(ref.keys())
        for q in ref:
            assert [d for d, _ in fused[q]] == [d for d, _ in ref[q]], (case["name"], q)
            a = np.array([s for _, s in fused[q]])
            b = np.array([float(s) for _, s in ref[q]])
            assert np.all(np.abs(a - b) <= 2 * np.spacing(np.abs(b))), (case["name"], q)
