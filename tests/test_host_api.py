"""CPU-side tests of the host mirror of the reference API: packing, loaders, writers, error behaviour."""
import gzip
import hashlib
import json
import os

import numpy as np
import pandas as pd
import pytest

from oracle import fusion_oracle as fo
from qppfusion import fusion
from qppfusion.config import ConfigError, config
from qppfusion.data_utils import load_qpp_scores, load_run_as_dict, load_run_file
from qppfusion.packing import pack_runs
from qppfusion.retrievers.base import BaseRetriever, RetrieverResult

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fusion_weighted.json.gz")


def gold_frames(case, rankers):
    frames = {}
    for rk in rankers:
        df = pd.DataFrame(case["rows"][rk], columns=["qid", "iter", "docno", "rank", "score", "runid"])
        frames[rk] = df.astype({"qid": str, "docno": str, "rank": "int64", "score": "float64"})
    return frames


def packed_to_lists(p):
    Q, R, K = p.shape
    lists = []
    for q in range(Q):
        per_r = []
        for r in range(R):
            n = p.lens[q, r]
            sc = p.scores[q, r, :n]
            sc = sc / 1e6 if sc.dtype == np.int32 else sc
            rk = p.ranks[q, r, :n] if p.ranks is not None else np.arange(1, n + 1)
            per_r.append([(p.docnos[p.doc_ids[q, r, i]], float(sc[i]), int(rk[i])) for i in range(n)])
        lists.append(per_r)
    return lists


def test_pack_runs_reproduces_reference_visiting_order():
    with gzip.open(GOLD, "rt") as f:
        gold = json.load(f)
    case = gold["cases"][0]                      # adversarial: ragged, missing ranker, duplicate docno, shuffled rows
    frames = gold_frames(case, gold["rankers"])
    for mode, method in (("weighted", fo.SUM), ("weighted", fo.MNZ), ("weighted_rrf", fo.RRF)):
        p = pack_runs(frames, mode=mode, need_ranks=True)
        assert p.scores.dtype == np.int32          # 6-decimal scores -> micro-units
        assert p.qids == sorted(p.qids)
        lists = packed_to_lists(p)
        key = {fo.SUM: "wcombsum@5", fo.MNZ: "wcombmnz@5", fo.RRF: "wrrf@5"}[method]
        for qi, qid in enumerate(p.qids):
            w = [case["qpp"][qid][rk][5] for rk in p.rankers]
            docs, scores, _ = fo.fuse_weighted_query(lists[qi], w, method, 60)
            ref = [(d, float(s)) for d, s in case["results"][key]["fused"].get(qid, [])]
            assert list(zip(docs, scores)) == ref, (mode, qid)


def test_pack_runs_ranx_mode_dict_semantics_and_f64_fallback():
    df = pd.DataFrame({"qid": ["1"] * 4 + ["2"], "iter": "Q0", "docno": ["a", "b", "a", "c", "z"],
                       "rank": [1, 2, 3, 4, 1], "score": [0.2, 0.9, 0.5, 0.9, 1.0 / 3.0], "runid": "x"})
    p = pack_runs({"A": df}, mode="ranx")
    assert p.scores.dtype == np.float64            # 1/3 is not a 6-decimal value
    l = packed_to_lists(p)
    # "a" keeps first dict position but last score (0.5); sorted by score desc, ties keep dict order (b before c)
    assert [(d, s) for d, s, _ in l[0][0]] == [("b", 0.9), ("c", 0.9), ("a", 0.5)]
    assert l[1][0][0][0] == "z"


def test_write_runfile_matches_reference_bytes(tmp_path):
    with gzip.open(GOLD, "rt") as f:
        gold = json.load(f)
    for case in gold["cases"]:
        for key, res in case["results"].items():
            fused = {q: [(d, float(s)) for d, s in lst] for q, lst in res["fused"].items()}
            out = tmp_path / "o.res"
            fusion.write_runfile(fused, str(out), tag=f"{key.split('@')[0]}-x")
            lines = out.read_text().splitlines()
            assert len(lines) == res["runfile_n"] and lines[:25] == res["runfile_head"]
            assert hashlib.sha256("\n".join(lines).encode()).hexdigest() == res["runfile_sha256"]


def test_loaders_and_errors(tmp_path):
    runs = tmp_path / "runs"
    runs.mkdir()
    with pytest.raises(FileNotFoundError):
        fusion.load_runs(str(runs))
    (runs / "B.norm.res").write_text("q1 Q0 d1 1 1.000000 B\nq1 Q0 d2 2 0.000000 B\n")
    (runs / "A.norm.res").write_text("q1 Q0 d2 1 0.750000 A\n007 Q0 d9 1 0.5 A\n")
    (runs / "A.res").write_text("q1 Q0 d2 1 12.5 A\nshort line\n")
    r = fusion.load_runs(str(runs))
    assert list(r.keys()) == ["A", "B"]                      # fixed, documented order
    assert r["A"]["qid"].tolist() == ["q1", "007"] and r["A"]["score"].dtype == np.float64
    # use_normalized=False: every "*.res" file, "X.norm.res" included under the ranker name "X.norm" -- what the
    # reference's endswith(".res") + replace(".res", "") does (src/fusion.py:57-64)
    assert list(fusion.load_runs(str(runs), use_normalized=False).keys()) == ["A.norm", "A", "B.norm"]
    assert load_run_file(runs / "A.res") == {"q1": [("d2", 12.5, 1)]}
    assert load_run_as_dict(runs / "B.norm.res") == {"q1": {"d1": 1.0, "d2": 0.0}}
    qdir = tmp_path / "qpp"
    qdir.mkdir()
    with pytest.raises(FileNotFoundError):
        load_qpp_scores(qdir)
    (qdir / "A.res.mmnorm.qpp").write_text("q1\t" + "\t".join(["0.500000"] * 13) + "\n\n")
    q = fusion.load_qpp_scores(str(qdir))
    assert q == {"q1": {"A": [0.5] * 13}}
    assert fusion.get_qpp_weight("q1", "A", q, 5) == 0.5 and fusion.get_qpp_weight("q1", "A", q, -1) == 0.5
    with pytest.raises(KeyError):
        fusion.get_qpp_weight("q2", "A", q)
    with pytest.raises(KeyError):
        fusion.get_qpp_weight("q1", "B", q)
    with pytest.raises(IndexError):
        fusion.get_qpp_weight("q1", "A", q, 13)
    with pytest.raises(ValueError, match="qpp_dir required"):
        fusion.run_fusion("wcombsum", str(runs))
    with pytest.raises(ValueError, match="Unknown method"):
        fusion.run_fusion("borda", str(runs))
    with pytest.raises(ConfigError):
        fusion.get_qpp_index("nope")


def test_method_id_tables_are_the_reference_quirk():
    assert [fusion.get_qpp_index(n) for n in ("SMV", "Sigma_max", "Sigma(%)", "NQC", "RSD", "WIG", "SCNQC")] == [0, 1, 2, 3, 5, 7, 8]
    assert fusion.get_qpp_index("rsd") == 5 and fusion.get_qpp_index("fusion") == -1
    assert config.qpp.methods[3] == "SigmaMax" and config.qpp.methods[0] == "nqc"     # id 3 ("NQC") selects SigmaMax
    assert config.qpp.default_index == 5 and config.fusion.rrf_k == 60 and config.processing.retrieval.top_k == 100


def test_retriever_result_and_normalize_scores():
    r = RetrieverResult("q1", [("d1", 3.0, 1), ("d2", 1.0, 2)], "BGE")
    assert r.to_trec_lines() == ["q1 Q0 d1 1 3.000000 BGE", "q1 Q0 d2 2 1.000000 BGE"]
    assert BaseRetriever.normalize_scores(r.results) == [("d1", 1.0, 1), ("d2", 0.0, 2)]
    assert BaseRetriever.normalize_scores([("d", 2.0, 1)]) == [("d", 0.0, 1)]
    assert BaseRetriever.normalize_scores([]) == []


def test_write_run_is_the_reference_format(tmp_path):
    """retrievers.base.write_run == scripts/02_retrieve.py:65-82: qid order, optional min-max, "%.6f"."""
    from qppfusion.retrievers.base import write_run
    results = {"test10": RetrieverResult("test10", [("d1", 3.0, 1), ("d2", 1.0, 2)], "BGE"),
               "test9": RetrieverResult("test9", [("d7", 0.1234567, 1)], "BGE"),
               "test2": RetrieverResult("test2", [], "BGE")}
    out = tmp_path / "sub" / "BGE.res"
    write_run(results, str(out), "BGE")
    assert out.read_text().splitlines() == ["test9 Q0 d7 1 0.123457 BGE", "test10 Q0 d1 1 3.000000 BGE",
                                           "test10 Q0 d2 2 1.000000 BGE"]
    write_run(results, str(tmp_path / "BGE.norm.res"), "BGE", normalize=True)
    assert (tmp_path / "BGE.norm.res").read_text().splitlines() == [
        "test9 Q0 d7 1 0.000000 BGE", "test10 Q0 d1 1 1.000000 BGE", "test10 Q0 d2 2 0.000000 BGE"]
    plain = {"10": RetrieverResult("10", [("a", 1.0, 1)], "X"), "9": RetrieverResult("9", [("b", 2.0, 1)], "X")}
    write_run(plain, str(tmp_path / "x.res"), "X")
    assert [ln.split()[0] for ln in (tmp_path / "x.res").read_text().splitlines()] == ["10", "9"]     # string order


def test_run_retriever_writes_both_run_files_and_clears_the_checkpoint(tmp_path):
    """retrievers.base.run_retriever == the per-retriever step of scripts/02_retrieve.py (run_bge :184-213)."""
    from qppfusion.retrievers.base import run_retriever

    class Fake(BaseRetriever):
        name = "BGE"
        calls = []

        def retrieve(self, query, qid="", top_k=100, **kw):
            raise NotImplementedError

        def retrieve_batch(self, queries, top_k=100, checkpoint_path=None, mini_batch_size=None, **kw):
            Fake.calls.append((top_k, checkpoint_path, mini_batch_size))
            open(checkpoint_path, "w").write("{}\n")
            return {q: RetrieverResult(q, [(f"d{i}", float(top_k - i), i + 1) for i in range(top_k)], self.name)
                    for q in queries}

    runs = tmp_path / "runs"
    res = run_retriever(Fake(), {"test1": "a", "test0": "b"}, runs, top_k=3, mini_batch_size=7)
    assert sorted(p.name for p in runs.iterdir()) == ["BGE.norm.res", "BGE.res"]          # checkpoint removed
    assert Fake.calls == [(3, str(runs / "BGE.checkpoint.jsonl"), 7)] and sorted(res) == ["test0", "test1"]
    assert (runs / "BGE.res").read_text().splitlines()[:3] == ["test0 Q0 d0 1 3.000000 BGE", "test0 Q0 d1 2 2.000000 BGE",
                                                               "test0 Q0 d2 3 1.000000 BGE"]
    assert (runs / "BGE.norm.res").read_text().splitlines()[:3] == ["test0 Q0 d0 1 1.000000 BGE", "test0 Q0 d1 2 0.500000 BGE",
                                                                    "test0 Q0 d2 3 0.000000 BGE"]


def _write_runs(tmp_path, seed=3, n_q=9, k=30, adversarial=True):
    rng = np.random.default_rng(seed)
    d = tmp_path / "runs"
    d.mkdir()
    for rk in ("BGE", "BM25", "Splade"):
        lines = []
        for q in range(n_q):
            docs = rng.choice(200, size=k, replace=False)
            sc = np.sort(rng.random(k) * (25 if rk == "BM25" else 1))[::-1]
            if adversarial and q == 2 and rk == "BM25":
                docs[5] = docs[1]                                  # duplicate docno
            if adversarial and q == 3:
                sc[4:8] = sc[4]                                    # ties
            order = rng.permutation(k) if (adversarial and q == 4 and rk == "Splade") else np.arange(k)
            for i in order:
                lines.append(f"test{q}\tQ0  doc{docs[i]} {i + 1} {sc[i]:.6f} {rk}")
        (d / f"{rk}.norm.res").write_text("\n".join(lines) + "\n")
    return d


@pytest.mark.parametrize("mode", ["weighted", "weighted_rrf", "ranx"])
def test_native_parser_and_packing_equal_the_pandas_path(tmp_path, mode):
    from qppfusion import native_io as nio
    d = _write_runs(tmp_path)
    runs = fusion.load_runs(str(d))
    ref = pack_runs(runs, mode=mode, need_ranks=True)
    files = sorted(f for f in os.listdir(d))
    parsed = nio.parse_runs([str(d / f) for f in files], [f.replace(".norm.res", "") for f in files], n_threads=3)
    got, qorder = nio.pack_parsed(parsed, mode, need_ranks=True)
    assert parsed.all_micro and got.scores.dtype == np.int32
    assert got.qids == ref.qids and got.rankers == ref.rankers
    assert np.array_equal(got.lens, ref.lens) and np.array_equal(got.scores, ref.scores) and np.array_equal(got.ranks, ref.ranks)
    docs = parsed.doc_strings()
    valid = got.doc_ids >= 0
    assert np.array_equal(valid, ref.doc_ids >= 0)
    assert (docs[got.doc_ids[valid]] == ref.docnos[ref.doc_ids[valid]]).all()


def test_native_parser_f64_fallback_and_errors(tmp_path):
    from qppfusion import _native as N_
    from qppfusion import native_io as nio
    p = tmp_path / "a.res"
    p.write_text("q1 Q0 d1 1 0.1234567 A\nq1 Q0 d2 2 1e-3 A\n\nq2 Q0 d1 1 -7 A\n")
    parsed = nio.parse_runs([str(p)])
    assert not parsed.all_micro
    assert parsed.cols[0]["score"].tolist() == [0.1234567, 1e-3, -7.0]
    assert parsed.qid_strings() == ["q1", "q2"] and parsed.doc_strings().tolist() == ["d1", "d2"]
    bad = tmp_path / "b.res"
    bad.write_text("q1 Q0 d1\n")
    with pytest.raises(N_.NativeError, match="fields"):
        nio.parse_runs([str(bad)])
    with pytest.raises(N_.NativeError, match="cannot open"):
        nio.parse_runs([str(tmp_path / "missing.res")])


def test_native_writer_matches_python_write_runfile(tmp_path):
    from qppfusion import native_io as nio
    d = _write_runs(tmp_path, adversarial=False)
    files = sorted(os.listdir(d))
    parsed = nio.parse_runs([str(d / f) for f in files])
    packed, qorder = nio.pack_parsed(parsed, "weighted")
    docs = parsed.doc_strings()
    rng = np.random.default_rng(1)
    fused, offs, ids, scs = {}, [0], [], []
    for qi, qid in enumerate(packed.qids):
        n = int(rng.integers(0, 40))
        di = rng.choice(len(docs), size=n, replace=False).astype(np.int32)
        sc = np.sort(rng.normal(size=n))[::-1]
        if n:
            fused[qid] = list(zip(docs[di].tolist(), sc.tolist()))
        ids.append(di); scs.append(sc); offs.append(offs[-1] + n)
    out_py, out_c = tmp_path / "py.res", tmp_path / "c.res"
    fusion.write_runfile(fused, str(out_py), "tag-x")
    nio.write_runfile_csr(str(out_c), parsed, qorder, np.array(offs), np.concatenate(ids), np.concatenate(scs), "tag-x", n_threads=3)
    assert out_c.read_bytes() == out_py.read_bytes()


def test_native_parser_multithreaded_shards_equal_single_thread(tmp_path):
    """A file large enough (> 1 MiB) for the chunk-parallel tokeniser and the shard-parallel docno dictionary: every
    thread count yields the same rows; docno codes may be numbered differently but must name the same strings."""
    from qppfusion import native_io as nio
    rng = np.random.default_rng(5)
    lines = []
    for q in range(120):
        docs = rng.choice(30_000, size=400, replace=False)
        sc = np.sort(rng.normal(size=400) * 3)[::-1]
        for i, (dno, s_) in enumerate(zip(docs, sc)):
            lines.append(f"q{q} Q0 D-{dno:07d} {i + 1} {s_:.6f} runA")
    p = tmp_path / "big.res"
    p.write_text("\n".join(lines) + "\n")
    assert p.stat().st_size > (1 << 20)
    ref = nio.parse_runs([str(p)], n_threads=1)
    ref_docs = ref.doc_strings()
    df = pd.read_csv(p, sep=r"\s+", names=["qid", "iter", "docno", "rank", "score", "runid"], dtype={"qid": str, "docno": str})
    assert (ref_docs[ref.cols[0]["doc"]] == df["docno"].to_numpy()).all()
    assert np.array_equal(ref.cols[0]["score"], df["score"].to_numpy()) and np.array_equal(ref.cols[0]["rank"], df["rank"].to_numpy())
    assert np.array_equal(ref.cols[0]["micro"], np.round(df["score"].to_numpy() * 1e6).astype(np.int32))
    for nt in (2, 5, 8):
        got = nio.parse_runs([str(p)], n_threads=nt)
        assert got.all_micro and got.qid_strings() == ref.qid_strings()
        for col in ("qid", "rank", "micro", "score"):
            assert np.array_equal(got.cols[0][col], ref.cols[0][col]), (nt, col)
        gd = got.doc_strings()
        assert sorted(gd.tolist()) == sorted(ref_docs.tolist())
        assert (gd[got.cols[0]["doc"]] == ref_docs[ref.cols[0]["doc"]]).all()


def test_native_parser_blank_lines_between_chunks_and_missing_final_newline(tmp_path):
    """The columns are allocated from a line count and written in place; blank / whitespace-only lines leave gaps between
    the threads' chunks that must be closed, and the last line may lack its newline."""
    from qppfusion import native_io as nio
    rng = np.random.default_rng(9)
    lines, exp = [], []
    for q in range(90):
        docs = rng.choice(50_000, size=500, replace=False)
        for i, dno in enumerate(docs):
            lines.append(f"q{q}\tQ0  doc{dno} {i + 1} {1.0 / (i + 1):.6f} runB")
            exp.append((f"q{q}", f"doc{dno}", i + 1, round(1.0 / (i + 1), 6)))
            if rng.random() < 0.02:
                lines.append("" if rng.random() < 0.5 else "   \t ")
    p = tmp_path / "gaps.res"
    p.write_text("\n\n" + "\n".join(lines))                     # leading blank lines, no trailing newline
    assert p.stat().st_size > (1 << 20)
    for nt in (1, 4, 7):
        got = nio.parse_runs([str(p)], n_threads=nt)
        c = got.cols[0]
        assert len(c["qid"]) == len(exp) and got.all_micro
        qs, ds = got.qid_strings(), got.doc_strings()
        assert [qs[i] for i in c["qid"]] == [e[0] for e in exp]
        assert ds[c["doc"]].tolist() == [e[1] for e in exp]
        assert c["rank"].tolist() == [e[2] for e in exp]
        assert np.array_equal(c["score"], np.array([e[3] for e in exp]))


def test_native_parser_fuzz_separators_blank_lines_and_thread_counts(tmp_path):
    """Random run files: mixed separators, leading / trailing blanks, CR line ends, blank lines, empty files, missing
    final newline, small and multi-chunk sizes -- every thread count returns exactly the rows that were written."""
    import random
    from qppfusion import native_io as nio
    rnd = random.Random(7)
    for trial in range(12):
        big = trial % 4 == 0
        n = rnd.randrange(40000, 50000) if big else rnd.randrange(0, 200)
        lines, exp, q = [], [], 0
        for _ in range(n):
            q += rnd.random() < 0.02
            if rnd.random() < 0.03:
                lines.append(rnd.choice(["", " ", "\t", "  \t  ", "\r"]))
                continue
            sep = rnd.choice([" ", "  ", "\t", " \t "])
            doc = "d" + str(rnd.randrange(0, 5000 if big else 50)) + rnd.choice(["", "-x", "__"])
            sc = rnd.choice([f"{rnd.uniform(-50, 50):.6f}", f"{rnd.uniform(0, 1):.3f}", str(rnd.randrange(-5, 5))])
            rank = rnd.randrange(1, 2000)
            lines.append(rnd.choice(["", " "]) + sep.join([f"q{q}", "Q0", doc, str(rank), sc, "tag"]) + rnd.choice(["", " ", "\r"]))
            exp.append((f"q{q}", doc, rank, float(sc)))
        p = tmp_path / f"f{trial}.res"
        with open(p, "w", newline="") as f:
            f.write("\n".join(lines) + rnd.choice(["\n", "", "\n\n"]))
        for nt in (1, 3, 8):
            got = nio.parse_runs([str(p)], n_threads=nt)
            c = got.cols[0]
            assert len(c["qid"]) == len(exp) and got.all_micro, (trial, nt)
            qs, ds = got.qid_strings(), got.doc_strings()
            assert [qs[i] for i in c["qid"]] == [e[0] for e in exp]
            assert (ds[c["doc"]].tolist() if exp else []) == [e[1] for e in exp]
            assert c["rank"].tolist() == [e[2] for e in exp]
            assert np.array_equal(c["score"], np.array([e[3] for e in exp], dtype=np.float64))


def test_native_writer_huge_and_non_finite_scores(tmp_path):
    """"%.6f" of 1e300 is 308 characters: the writer's per-line bound must hold (reference: Python's f-string)."""
    from qppfusion import native_io as nio
    p = tmp_path / "one.res"
    p.write_text("q1 Q0 d0 1 0.5 A\n")
    parsed = nio.parse_runs([str(p)])
    _, qorder = nio.pack_parsed(parsed)
    vals = np.array([1e300, -1.7976931348623157e308, 1e22, float("inf"), float("-inf"), float("nan"), 3.25, 1e9])
    out = tmp_path / "out.res"
    nio.write_runfile_csr(str(out), parsed, qorder, np.array([0, len(vals)]), np.zeros(len(vals), dtype=np.int32), vals, "t")
    got = [ln.split(" ")[4] for ln in out.read_text().splitlines()]
    assert got == [f"{v:.6f}" for v in vals]


def test_native_writer_fixed6_is_python_format(tmp_path):
    """The printf-free "%.6f" of qf_runfile_write against Python's formatting: random magnitudes, exact ties of the
    seventh decimal (k + 0.5) * 1e-6 representable in binary, negative zero, values at and beyond the 1e9 fallback."""
    from qppfusion import native_io as nio
    rng = np.random.default_rng(6)
    vals = np.concatenate([
        rng.normal(size=20000) * 10.0 ** rng.integers(-8, 9, size=20000),
        (rng.integers(-2_000_000, 2_000_000, size=20000) + 0.5) / 1e6,            # decimal ties (inexact in binary)
        rng.integers(-4096, 4096, size=4000) / 64.0 + 2.0 ** -21 * rng.integers(-3, 4, size=4000),   # binary ties region
        np.array([0.0, -0.0, 1e-7, -1e-7, 4.9999995e-7, 5e-7, 5.0000005e-7, 0.9999995, 0.99999949999,
                  123456789.1234565, 999999999.9999995, 1e9, -1e9, 1.5e12, 2147.4836475, -2147.4836475]),
    ])
    p = tmp_path / "one.res"
    p.write_text("q1 Q0 d0 1 0.5 A\n")
    parsed = nio.parse_runs([str(p)])
    _, qorder = nio.pack_parsed(parsed)
    n = len(vals)
    out = tmp_path / "out.res"
    nio.write_runfile_csr(str(out), parsed, qorder, np.array([0, n]), np.zeros(n, dtype=np.int32), vals, "t", n_threads=4)
    got = [ln.split(" ")[4] for ln in out.read_text().splitlines()]
    exp = ["%.6f" % v for v in vals]
    bad = [(g, e, float(v).hex()) for g, e, v in zip(got, exp, vals) if g != e]
    assert not bad, bad[:10]


def test_run_all_methods_orchestration_without_gpu(tmp_path, monkeypatch):
    """File names, tags, the qpp_dir switch and the learned-model discovery of run_all_methods (scripts/05_fusion.py:56-152),
    with the device work stubbed out (the GPU test runs the real thing)."""
    from qppfusion import fusion
    runs_dir = _write_runs(tmp_path, n_q=3, k=5, adversarial=False)
    qpp_dir, fused_dir = tmp_path / "qpp", tmp_path / "data" / "fused"
    qpp_dir.mkdir()
    for rk in ("BGE", "BM25", "Splade"):
        (qpp_dir / f"{rk}.res.mmnorm.qpp").write_text("".join(f"test{q}\t" + "\t".join(["0.500000"] * 13) + "\n" for q in range(3)))
    (tmp_path / "data" / "models").mkdir(parents=True)
    (tmp_path / "data" / "models" / "fusion_mlp.pkl").write_bytes(b"stub")
    calls = []

    def fake_fuse_parsed(parsed, method, qpp_data, qpp_model, output_path, rrf_k, return_fused, tag=None):
        calls.append((method, qpp_data is not None, qpp_model, rrf_k, tag))
        open(output_path, "w").write("x\n")
        return {}

    monkeypatch.setattr(fusion, "_fuse_parsed", fake_fuse_parsed)
    monkeypatch.setattr(fusion, "learned_fusion", lambda runs, qpp, path: {"test0": [("d1", 1.0)]})
    res = fusion.run_all_methods(runs_dir, qpp_dir, fused_dir, qpp_model="NQC", rrf_k=10)
    assert list(res) == ["combsum", "combmnz", "rrf", "wcombsum", "wcombmnz", "wrrf", "learned_mlp"]
    assert sorted(p.name for p in fused_dir.iterdir()) == ["combmnz.res", "combsum.res", "learned_mlp.res", "rrf.res",
                                                          "wcombmnz_nqc.res", "wcombsum_nqc.res", "wrrf_nqc.res"]
    assert calls[:3] == [("combsum", False, "NQC", 10, "combsum"), ("combmnz", False, "NQC", 10, "combmnz"),
                         ("rrf", False, "NQC", 10, "rrf")]
    assert [c[0] for c in calls[3:]] == ["wcombsum", "wcombmnz", "wrrf"] and all(c[1] and c[4] is None for c in calls[3:])
    assert (fused_dir / "learned_mlp.res").read_text() == "test0 Q0 d1 1 1.000000 learned-mlp\n"
    calls.clear()
    res = fusion.run_all_methods(runs_dir, tmp_path / "absent", tmp_path / "f2")
    assert list(res) == ["combsum", "combmnz", "rrf"] and len(calls) == 3
    with pytest.raises(FileNotFoundError):
        fusion.run_all_methods(tmp_path / "no_runs", qpp_dir, tmp_path / "f3")
