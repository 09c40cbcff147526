"""GPU: steps 02 -> 03 -> 05 of the reference pipeline chained through the drop-in modules only
(query strings -> device encoder -> flat-IP search -> .res/.norm.res -> .qpp -> six fused runs), every artefact on
disk in the reference's formats and file names; the QPP and fused files are checked against the oracles."""
import json

import numpy as np
import pytest
import torch

from make_golden_bert import CFG, bert_small_state, synthetic_vocab
from oracle import fusion_oracle as fo
from oracle import qpp_oracle as qo

pytestmark = pytest.mark.gpu


def _encoder(seed, pooling):
    from qppfusion.encoder import BertConfig, BertQueryEncoder, WordPieceTokenizer
    vocab = {t: i for i, t in enumerate(synthetic_vocab())}
    cfg_hf = dict(CFG, hidden_size=768, num_attention_heads=12, intermediate_size=1024, vocab_size=len(vocab))
    sd = bert_small_state(seed=seed, cfg=cfg_hf)
    return BertQueryEncoder({k: torch.from_numpy(v) for k, v in sd.items()}, BertConfig.from_hf(cfg_hf, pooling),
                            WordPieceTokenizer(vocab, max_length=64))


def test_retrieve_qpp_fusion_steps_end_to_end(tmp_path):
    from qppfusion import fusion
    from qppfusion.qpp import compute_qpp_for_runs_dir
    from qppfusion.retrievers.base import run_retriever
    from qppfusion.retrievers.bge import BGERetriever
    from qppfusion.retrievers.tct_colbert import TCTColBERTRetriever
    texts = ["what is the capital of france?", "who wrote 'hamlet'", "unbelievably long-winded antidisestablishmentarianism",
             "a", "tower height", "cafe au lait costs $4.50!", "the capital", "hamlet resume"]
    queries = {f"test{i}": t for i, t in enumerate(texts)}
    (tmp_path / "queries.jsonl").write_text("".join(json.dumps({"_id": q, "text": t}) + "\n" for q, t in queries.items()))
    runs_dir = tmp_path / "runs"
    rng = np.random.default_rng(1)
    words = ["what", "is", "the", "capital", "of", "france", "cafe", "au", "lait", "naive", "resume", "costs", "who", "wrote",
             "hamlet", "long", "tower", "height"]
    docs = {f"d{i}": {"title": " ".join(rng.choice(words, size=3)), "text": " ".join(rng.choice(words, size=int(rng.integers(3, 30))))}
            for i in range(400)}
    # 02_retrieve: BGE over the encoded corpus (CLS pooling) and TCT-ColBERT (mean pooling, corpus cache written)
    enc_bge, enc_tct = _encoder(31, "cls"), _encoder(32, "mean")
    corpus_emb = enc_bge([(d["title"] + " " + d["text"]).strip() for d in docs.values()])
    bge = BGERetriever(corpus_embeddings=corpus_emb, docids=list(docs.keys()), encoder=enc_bge)
    tct = TCTColBERTRetriever(docs, cache_dir=str(tmp_path / "cache"), batch_size=16, encoder=enc_tct)
    for r in (bge, tct):
        res = run_retriever(r, queries, runs_dir, top_k=60)
        assert sorted(res) == sorted(queries) and all(len(v.results) == 60 for v in res.values())
    assert sorted(p.name for p in runs_dir.iterdir()) == ["BGE.norm.res", "BGE.res", "TCT-ColBERT.norm.res", "TCT-ColBERT.res"]
    # 03_qpp: both run files -> .qpp, checked against the oracle on the same file contents
    done = compute_qpp_for_runs_dir(runs_dir)
    assert sorted(done) == ["BGE.res", "TCT-ColBERT.res"]
    for name in ("BGE", "TCT-ColBERT"):
        per_q = {}
        for ln in (runs_dir / f"{name}.res").read_text().splitlines():
            p = ln.split()
            per_q.setdefault(p[0], []).append(float(p[4]))
        exp = qo.qpp_for_run(per_q, queries, 100, "minmax")
        got = (tmp_path / "qpp" / f"{name}.res.mmnorm.qpp").read_text().splitlines()
        want = qo.qpp_file_lines(exp)
        assert [ln.split("\t")[0] for ln in got] == [ln.split("\t")[0] for ln in want]
        g = np.array([[float(x) for x in ln.split("\t")[1:]] for ln in got])
        w = np.array([[float(x) for x in ln.split("\t")[1:]] for ln in want])
        assert np.abs(g - w).max() <= 2e-6                      # "%.6f" of values equal within 1e-5 relative
    # 05_fusion: all six methods; the QPP-weighted CombSUM file equals the oracle's lines on the same inputs
    out = fusion.run_all_methods(runs_dir, tmp_path / "qpp", tmp_path / "fused")
    assert list(out) == ["combsum", "combmnz", "rrf", "wcombsum", "wcombmnz", "wrrf"]
    runs = {}
    for name in ("BGE", "TCT-ColBERT"):
        runs[name] = {}
        for ln in (runs_dir / f"{name}.norm.res").read_text().splitlines():
            p = ln.split()
            runs[name].setdefault(p[0], []).append((p[2], float(p[4]), int(p[3])))
    qpp_data = fusion.load_qpp_scores(str(tmp_path / "qpp"))
    for fname, method, tag in (("wcombsum_rsd.res", fo.SUM, "wcombsum-rsd"), ("wcombmnz_rsd.res", fo.MNZ, "wcombmnz-rsd"),
                               ("wrrf_rsd.res", fo.RRF, "wrrf-rsd")):
        exp = fo.weighted_fusion(runs, qpp_data, method, 5, 60)
        assert (tmp_path / "fused" / fname).read_text().splitlines() == fo.runfile_lines(exp, tag), fname
    for fname in ("combsum.res", "combmnz.res", "rrf.res"):
        lines = (tmp_path / "fused" / fname).read_text().splitlines()
        assert len({ln.split()[0] for ln in lines}) == len(queries)
        ranks = [int(ln.split()[3]) for ln in lines if ln.startswith("test0 ")]
        assert ranks == list(range(1, len(ranks) + 1))
