"""Query encoder on the GPU (qf_bert_encode through torch.ops.qppfusion.bert_encode and BertQueryEncoder) against
the numpy oracle and the committed transformers.BertModel outputs.

Tolerances (bf16 weights and activations, fp32 accumulation, against an fp32/float64 reference; rows are unit
vectors): cosine >= 0.9995 and max |diff| <= 4e-3 for the 2-layer fixture, cosine >= 0.999 for the 12-layer
BGE-base shape against bf16-rounded weights in float64 arithmetic (>= 0.998 against the fp32 weights: the
weight rounding itself, a storage decision) -- far below what moves a top-1000 ranking (the corpus itself is held in bf16)."""
import os

import numpy as np
import pytest
import torch

from make_golden_bert import CFG, TEXTS, bert_small_state, synthetic_vocab
from oracle import bert_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bert_small.npz")


def _encode(sd, cfg_hf, tokens, lens, pooling="cls"):
    from qppfusion import ops
    from qppfusion.encoder import BertConfig, pack_bert_weights
    cfg = BertConfig.from_hf(cfg_hf, pooling)
    dev = torch.device("cuda", 0)
    blob = pack_bert_weights({k: torch.from_numpy(v) for k, v in sd.items()}, cfg).to(dev)
    return ops.bert_encode(blob, torch.from_numpy(np.ascontiguousarray(tokens, dtype=np.int32)).to(dev),
                           torch.from_numpy(np.ascontiguousarray(lens, dtype=np.int32)).to(dev), n_layers=cfg.n_layers,
                           hidden=cfg.hidden, heads=cfg.heads, ffn=cfg.ffn, vocab=cfg.vocab, max_pos=cfg.max_pos,
                           pooling=cfg.pooling_id, ln_eps=cfg.ln_eps)


@pytest.mark.parametrize("pooling", ["cls", "mean"])
def test_small_model_matches_golden_bertmodel(pooling):
    g = np.load(GOLD)
    out_b, out_f, status = _encode(bert_small_state(), CFG, g["tokens"], g["lens"], pooling)
    got = out_f.double().cpu().numpy()
    assert int(status.max().item()) == 0
    assert ((got * g[pooling]).sum(1) >= 0.9995).all()
    assert np.abs(got - g[pooling]).max() <= 4e-3
    np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)
    # the bf16 rows handed to the dense search are the rounded fp32 rows
    assert torch.equal(out_b, out_f.to(torch.bfloat16))


@pytest.mark.parametrize("hidden,heads,L", [(256, 4, 8), (512, 8, 40), (768, 12, 64), (1024, 16, 24), (256, 4, 160),
                                            (256, 4, 128), (256, 4, 129), (256, 4, 384), (256, 4, 512), (256, 4, 7), (256, 4, 33), (256, 4, 1)])
def test_shapes_against_oracle(hidden, heads, L):
    cfg = dict(CFG, hidden_size=hidden, num_attention_heads=heads, intermediate_size=2 * hidden, num_hidden_layers=2,
               max_position_embeddings=max(64, L))
    B = 9 if L <= 160 else 4
    sd = bert_small_state(seed=hidden + L, cfg=cfg)
    rng = np.random.default_rng(L)
    tokens = rng.integers(0, cfg["vocab_size"], size=(B, L))
    lens = rng.integers(1, L + 1, size=B)
    lens[0], lens[1] = L, 1
    sd_bf = {k: bert_oracle._bf16_round(v) for k, v in sd.items()}
    ref = bert_oracle.bert_encode(sd_bf, tokens, lens, n_layers=2, heads=heads)
    _, out_f, status = _encode(sd, cfg, tokens, lens)
    got = out_f.double().cpu().numpy()
    assert int(status.max().item()) == 0
    assert ((got * ref).sum(1) >= 0.9995).all()
    assert np.abs(got - ref).max() <= 4e-3


def test_padding_and_batch_independence():
    g = np.load(GOLD)
    sd = bert_small_state()
    _, a, _ = _encode(sd, CFG, g["tokens"], g["lens"])
    junk = np.pad(g["tokens"], ((0, 0), (0, 16)), constant_values=77)
    _, b, _ = _encode(sd, CFG, junk, g["lens"])
    assert (a - b).abs().max().item() <= 2e-3             # masked keys contribute nothing (other L: other GEMM algorithm)
    _, c, _ = _encode(sd, CFG, g["tokens"][2:3], g["lens"][2:3])
    assert (a[2:3] - c).abs().max().item() <= 2e-3        # cuBLASLt may pick another algorithm for another M


def test_bad_token_is_reported_not_read():
    g = np.load(GOLD)
    toks = g["tokens"].copy()
    toks[3, 2] = CFG["vocab_size"] + 5
    toks[4, 0] = -1
    _, _, status = _encode(bert_small_state(), CFG, toks, g["lens"])
    assert status.cpu().tolist() == [0, 0, 0, 1, 1, 0]


def test_bge_base_shape_against_oracle():
    base = dict(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                max_position_embeddings=512)
    sd = bert_small_state(seed=42, cfg=base)
    rng = np.random.default_rng(1)
    B, L = 70, 32
    tokens = rng.integers(0, base["vocab_size"], size=(B, L))
    lens = np.clip(rng.poisson(11, size=B) + 3, 3, L)
    _, out_f, status = _encode(sd, base, tokens, lens)
    n = 6
    ref = bert_oracle.bert_encode(sd, tokens[:n], lens[:n], n_layers=12, heads=12)     # fp32 weights, float64 math
    ref_bf = bert_oracle.bert_encode({k: bert_oracle._bf16_round(v) for k, v in sd.items()}, tokens[:n], lens[:n],
                                     n_layers=12, heads=12)
    got = out_f[:n].double().cpu().numpy()
    assert int(status.max().item()) == 0
    assert ((got * ref_bf).sum(1) >= 0.999).all(), (got * ref_bf).sum(1)
    assert ((got * ref).sum(1) >= 0.998).all(), (got * ref).sum(1)


def test_packed_layout_matches_padded_layout_and_reports_bad_lengths():
    """qf_bert_encode_packed (tokens back to back + cu_seqlens) against the padded call on the same sequences: the
    per-token arithmetic is identical, only the GEMM row count differs; length errors are reported per sequence."""
    from qppfusion import ops
    from qppfusion.encoder import BertConfig, pack_bert_weights
    cfg_hf = dict(CFG, max_position_embeddings=256)
    sd = bert_small_state(seed=4, cfg=cfg_hf)
    cfg = BertConfig.from_hf(cfg_hf)
    dev = torch.device("cuda", 0)
    blob = pack_bert_weights({k: torch.from_numpy(v) for k, v in sd.items()}, cfg).to(dev)
    kw = dict(n_layers=cfg.n_layers, hidden=cfg.hidden, heads=cfg.heads, ffn=cfg.ffn, vocab=cfg.vocab, max_pos=cfg.max_pos)
    rng = np.random.default_rng(12)
    for L, pooling in ((24, 0), (97, 1), (200, 0)):               # mma kernel (32 / 128 wide) and the long-sequence kernel
        B = 13
        lens = rng.integers(1, L + 1, size=B).astype(np.int32)
        lens[0], lens[1] = L, 1
        tokens = rng.integers(0, cfg.vocab, size=(B, L)).astype(np.int32)
        _, pad_f, st0 = ops.bert_encode(blob, torch.from_numpy(tokens).to(dev), torch.from_numpy(lens).to(dev), pooling=pooling, **kw)
        flat = np.concatenate([tokens[i, :lens[i]] for i in range(B)])
        cu = np.concatenate(([0], np.cumsum(lens))).astype(np.int32)
        out_b, out_f, st = ops.bert_encode_packed(blob, torch.from_numpy(flat).to(dev), torch.from_numpy(cu).to(dev), L,
                                                  pooling=pooling, **kw)
        assert st.cpu().tolist() == [0] * B and st0.cpu().tolist() == [0] * B
        assert (out_f - pad_f).abs().max().item() <= 2e-3, (L, pooling)
        assert torch.equal(out_b, out_f.to(torch.bfloat16))
        ref = bert_oracle.bert_encode({k: bert_oracle._bf16_round(v) for k, v in sd.items()}, tokens, lens, n_layers=2,
                                      heads=4, pooling="mean" if pooling else "cls")
        assert ((out_f.double().cpu().numpy() * ref).sum(1) >= 0.9995).all()
    # a sequence longer than max_len and an empty one: flagged (bit 1), the others unaffected
    lens = np.array([5, 40, 0, 7], dtype=np.int32)
    flat = rng.integers(0, cfg.vocab, size=int(lens.sum())).astype(np.int32)
    flat[2] = cfg.vocab + 9                                        # bad token id in sequence 0 (bit 0)
    cu = np.concatenate(([0], np.cumsum(lens))).astype(np.int32)
    _, f, st = ops.bert_encode_packed(blob, torch.from_numpy(flat).to(dev), torch.from_numpy(cu).to(dev), 32, **kw)
    assert st.cpu().tolist() == [1, 2, 2, 0]
    assert float(f[2].abs().max()) == 0.0 and torch.isfinite(f).all()
    solo = ops.bert_encode_packed(blob, torch.from_numpy(flat[45:]).to(dev),
                                  torch.tensor([0, 7], dtype=torch.int32, device=dev), 32, **kw)[1]
    assert (solo[0] - f[3]).abs().max().item() <= 2e-3


def test_bucketed_token_lists_match_one_padded_call():
    from qppfusion.encoder import BertConfig, BertQueryEncoder, WordPieceTokenizer
    vocab = {t: i for i, t in enumerate(synthetic_vocab())}
    cfg_hf = dict(CFG, vocab_size=len(vocab))
    sd = bert_small_state(seed=3, cfg=cfg_hf)
    enc = BertQueryEncoder({k: torch.from_numpy(v) for k, v in sd.items()}, BertConfig.from_hf(cfg_hf),
                           WordPieceTokenizer(vocab, max_length=64), batch_size=16)
    rng = np.random.default_rng(5)
    rows = [rng.integers(5, len(vocab), size=int(n)).tolist() for n in rng.integers(1, 50, size=57)]
    _, f = enc.encode_token_lists(rows)
    L = 56
    toks = np.zeros((len(rows), L), dtype=np.int32)
    for i, r in enumerate(rows):
        toks[i, :len(r)] = r
    lens = np.array([len(r) for r in rows], dtype=np.int32)
    _, g = enc.encode_tokens(torch.from_numpy(toks).cuda(), torch.from_numpy(lens).cuda())
    assert (f - g).abs().max().item() <= 2e-3
    ref = bert_oracle.bert_encode({k: bert_oracle._bf16_round(v) for k, v in sd.items()}, toks, lens, n_layers=2, heads=4)
    assert ((f.double().cpu().numpy() * ref).sum(1) >= 0.9995).all()
    with pytest.raises(ValueError):
        enc.encode_token_lists([[1, 2, len(vocab) + 3]])
    with pytest.raises(ValueError):
        enc.encode_token_lists([[]])
    assert enc.encode_token_lists([])[1].shape == (0, 256)


def test_query_encoder_end_to_end_into_dense_search():
    """Strings -> host WordPiece -> device BERT -> bf16 rows -> tcgen05 flat search, through BGERetriever."""
    from qppfusion.encoder import BertConfig, BertQueryEncoder, WordPieceTokenizer
    from qppfusion.retrievers.bge import BGERetriever
    vocab = {t: i for i, t in enumerate(synthetic_vocab())}
    cfg_hf = dict(CFG, hidden_size=768, num_attention_heads=12, intermediate_size=1024, vocab_size=len(vocab))
    sd = bert_small_state(seed=9, cfg=cfg_hf)
    enc = BertQueryEncoder({k: torch.from_numpy(v) for k, v in sd.items()}, BertConfig.from_hf(cfg_hf),
                           WordPieceTokenizer(vocab, max_length=64), batch_size=3)
    emb = enc(TEXTS)
    assert emb.shape == (len(TEXTS), 768) and emb.dtype == np.float32
    tok = enc.tokenizer
    toks, lens = tok.batch(TEXTS)
    ref = bert_oracle.bert_encode({k: bert_oracle._bf16_round(v) for k, v in sd.items()}, toks, lens, n_layers=2, heads=12)
    assert ((emb.astype(np.float64) * ref).sum(1) >= 0.9995).all()
    rng = np.random.default_rng(0)
    corpus = rng.standard_normal((5000, 768)).astype(np.float32)
    corpus /= np.linalg.norm(corpus, axis=1, keepdims=True)
    corpus[1234] = emb[0]
    r = BGERetriever(corpus_embeddings=corpus, encoder=enc)
    res = r.retrieve_batch({f"q{i}": t for i, t in enumerate(TEXTS)}, top_k=10)
    assert res["q0"].results[0][0] == "1234" and abs(res["q0"].results[0][1] - 1.0) < 1e-2
    assert all(len(v.results) == 10 for v in res.values())


def test_from_pretrained_and_tct_colbert_corpus_encoding(tmp_path):
    """A checkpoint directory (config.json, vocab.txt, model.safetensors) -> BertQueryEncoder.from_pretrained (mean
    pooling for a plain HF checkpoint, as SentenceTransformer(name) builds it) -> TCTColBERTRetriever encodes the
    corpus, writes the reference's cache files, a second instance reads them back."""
    import json
    from safetensors.torch import save_file
    from qppfusion.encoder import BertQueryEncoder
    from qppfusion.retrievers.tct_colbert import TCTColBERTRetriever
    vocab = synthetic_vocab()
    cfg_hf = dict(CFG, hidden_size=768, num_attention_heads=12, intermediate_size=1024, vocab_size=len(vocab),
                  layer_norm_eps=1e-12, hidden_act="gelu", max_position_embeddings=64)
    sd = bert_small_state(seed=17, cfg=cfg_hf)
    d = tmp_path / "ckpt"
    d.mkdir()
    (d / "config.json").write_text(json.dumps(cfg_hf))
    (d / "vocab.txt").write_text("\n".join(vocab) + "\n", encoding="utf-8")
    save_file({"bert." + k: torch.from_numpy(v) for k, v in sd.items()}, str(d / "model.safetensors"))
    enc = BertQueryEncoder.from_pretrained(str(d))
    assert enc.cfg.pooling == "mean" and enc.tokenizer.max_length == 64
    emb = enc(TEXTS)
    toks, lens = enc.tokenizer.batch(TEXTS)
    ref = bert_oracle.bert_encode({k: bert_oracle._bf16_round(v) for k, v in sd.items()}, toks, lens, n_layers=2,
                                  heads=12, pooling="mean")
    assert ((emb.astype(np.float64) * ref).sum(1) >= 0.9995).all()
    (d / "1_Pooling").mkdir()
    (d / "1_Pooling" / "config.json").write_text(json.dumps({"pooling_mode_cls_token": True, "pooling_mode_mean_tokens": False}))
    assert BertQueryEncoder.from_pretrained(str(d)).cfg.pooling == "cls"
    corpus = {f"d{i}": {"title": TEXTS[i % len(TEXTS)], "text": "what is the capital of france " * (1 + i % 9)} for i in range(40)}
    cache = tmp_path / "cache"
    r = TCTColBERTRetriever(corpus, cache_dir=str(cache), batch_size=4, use_fp16=True, encoder=enc)
    files = sorted(p.name for p in cache.iterdir())
    assert files == ["tct_colbert_40_fp16.npy", "tct_colbert_40_ids.npy"]
    saved = np.load(str(cache / "tct_colbert_40_fp16.npy"))
    assert saved.dtype == np.float16 and saved.shape == (40, 768)
    want = enc([(corpus["d7"]["title"] + " " + corpus["d7"]["text"]).strip()])[0]
    assert np.abs(saved[7].astype(np.float32) - want).max() <= 2e-3
    hit = r.retrieve((corpus["d7"]["title"] + " " + corpus["d7"]["text"]).strip(), qid="q", top_k=3)
    assert abs(hit.results[0][1] - 1.0) < 2e-2
    r2 = TCTColBERTRetriever(corpus, cache_dir=str(cache), batch_size=4, use_fp16=True, encoder=enc)     # from the cache
    assert torch.equal(r2.corpus, r.corpus)
    with pytest.raises(FileNotFoundError):
        TCTColBERTRetriever(corpus, cache_dir=str(tmp_path / "none"), use_fp16=True)                     # no encoder


def test_more_than_65535_sequences_in_one_call():
    """Sequences ride on grid.x: a call with more sequences than gridDim.y allows (65 535) works."""
    cfg = dict(CFG, num_hidden_layers=1)
    sd = bert_small_state(seed=8, cfg=cfg)
    rng = np.random.default_rng(3)
    B, L = 70000, 2
    tokens = rng.integers(0, cfg["vocab_size"], size=(B, L))
    lens = rng.integers(1, L + 1, size=B)
    _, out_f, status = _encode(sd, cfg, tokens, lens)
    assert int(status.max().item()) == 0
    pick = np.array([0, 1, 65535, 65536, B - 1])
    ref = bert_oracle.bert_encode({k: bert_oracle._bf16_round(v) for k, v in sd.items()}, tokens[pick], lens[pick],
                                  n_layers=1, heads=4)
    got = out_f[torch.from_numpy(pick).cuda()].double().cpu().numpy()
    assert ((got * ref).sum(1) >= 0.9995).all()
