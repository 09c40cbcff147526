"""Query encoder, CPU side: the numpy oracle (oracle/bert_oracle.py) against the committed outputs of the real
transformers.BertModel (tests/golden/bert_small.npz, tools/make_golden_bert.py) and against a live BertModel when
transformers is importable; the host tokeniser (qppfusion.encoder.WordPieceTokenizer) against the token ids the real
transformers.BertTokenizer produced; the weight packer's layout and error behaviour."""
import os

import numpy as np
import pytest

from make_golden_bert import CFG, TEXTS, bert_small_state, synthetic_vocab
from oracle import bert_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bert_small.npz")


def test_oracle_matches_golden_bertmodel_outputs():
    g = np.load(GOLD)
    sd = bert_small_state()
    for pooling in ("cls", "mean"):
        got = bert_oracle.bert_encode(sd, g["tokens"], g["lens"], n_layers=CFG["num_hidden_layers"],
                                      heads=CFG["num_attention_heads"], pooling=pooling)
        # the golden rows are float32 torch outputs: 2e-6 absolute on unit vectors is fp32 round-off over 2 layers
        np.testing.assert_allclose(got, g[pooling], atol=2e-6, rtol=0)
        np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-12)


def test_oracle_matches_live_bertmodel():
    torch = pytest.importorskip("torch")
    tr = pytest.importorskip("transformers")
    cfg = dict(CFG, num_hidden_layers=3, hidden_size=512, num_attention_heads=8, intermediate_size=1024)
    sd = bert_small_state(seed=5, cfg=cfg)
    model = tr.BertModel(tr.BertConfig(**cfg), add_pooling_layer=False).eval()
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    rng = np.random.default_rng(11)
    tokens = rng.integers(0, cfg["vocab_size"], size=(5, 40))
    lens = np.array([40, 3, 17, 33, 1])
    mask = torch.from_numpy((np.arange(40)[None, :] < lens[:, None]).astype(np.int64))
    with torch.no_grad():
        hid = model(input_ids=torch.from_numpy(tokens), attention_mask=mask).last_hidden_state
    ref = torch.nn.functional.normalize(hid[:, 0], p=2, dim=1).numpy()
    got = bert_oracle.bert_encode(sd, tokens, lens, n_layers=3, heads=8)
    np.testing.assert_allclose(got, ref, atol=3e-6, rtol=0)
    # padding never leaks: the same rows padded to another length give the same embedding
    got2 = bert_oracle.bert_encode(sd, np.pad(tokens, ((0, 0), (0, 8))), lens, n_layers=3, heads=8)
    np.testing.assert_allclose(got2, got, atol=1e-12)


def test_bf16_emulation_rounding():
    x = np.array([1.0, 1.00390625, 1.01171875, -3.1415927, 0.0], dtype=np.float32)
    torch = pytest.importorskip("torch")
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    np.testing.assert_array_equal(bert_oracle._bf16_round(x), want)


def test_tokenizer_matches_real_berttokenizer_ids():
    from qppfusion.encoder import WordPieceTokenizer
    g = np.load(GOLD)
    tok = WordPieceTokenizer({t: i for i, t in enumerate(synthetic_vocab())}, max_length=32)
    for i, text in enumerate(TEXTS):
        want = g["tok_flat"][g["tok_offs"][i]:g["tok_offs"][i + 1]].tolist()
        assert tok.encode(text) == want, text
    toks, lens = tok.batch(TEXTS)
    assert toks.shape[1] % 8 == 0 and toks.dtype == np.int32 and lens.tolist() == np.diff(g["tok_offs"]).tolist()
    assert (toks[5, 2:] == tok.pad_id).all() and toks[5, 0] == tok.cls_id and toks[5, 1] == tok.sep_id
    long = WordPieceTokenizer(tok.vocab, max_length=16).encode("what is the capital of france " * 10)
    assert len(long) == 16 and long[0] == tok.cls_id and long[-1] == tok.sep_id


def test_tokenizer_matches_live_berttokenizer(tmp_path):
    tr = pytest.importorskip("transformers")
    from qppfusion.encoder import WordPieceTokenizer
    vp = tmp_path / "vocab.txt"
    vp.write_text("\n".join(synthetic_vocab()) + "\n", encoding="utf-8")
    real = tr.BertTokenizer(str(vp), do_lower_case=True)
    mine = WordPieceTokenizer(vp, max_length=48)
    rng = np.random.default_rng(3)
    alphabet = list("abcdefghij  \t!?'-.,éÉñ$5京") + ["what", "capital", "unbelievably", "resume", "Hamlet", " "]
    for _ in range(200):
        text = "".join(rng.choice(alphabet, size=int(rng.integers(0, 40))))
        assert mine.encode(text) == real(text, truncation=True, max_length=48)["input_ids"], repr(text)


def test_pack_weights_layout_and_errors():
    torch = pytest.importorskip("torch")
    from qppfusion.encoder import BertConfig, pack_bert_weights
    sd = {k: torch.from_numpy(v) for k, v in bert_small_state().items()}
    cfg = BertConfig.from_hf(dict(CFG, layer_norm_eps=1e-12, hidden_act="gelu"))
    blob = pack_bert_weights(sd, cfg)
    H, F, V, P = cfg.hidden, cfg.ffn, cfg.vocab, cfg.max_pos
    per_layer = 4 * H * H + 2 * F * H + 9 * H + F
    assert blob.dtype == torch.bfloat16 and blob.numel() == (V + P + 2) * H + 2 * H + cfg.n_layers * per_layer
    o = (V + P + 2) * H + 2 * H
    assert torch.equal(blob[o + H * H:o + 2 * H * H].float().reshape(H, H),
                       sd["encoder.layer.0.attention.self.key.weight"].to(torch.bfloat16).float())
    assert torch.equal(blob[o + 3 * H * H + H:o + 3 * H * H + 2 * H].float(),
                       sd["encoder.layer.0.attention.self.key.bias"].to(torch.bfloat16).float())
    # a 'bert.'-prefixed dict packs the same; a missing or mis-shaped tensor raises
    assert torch.equal(pack_bert_weights({"bert." + k: v for k, v in sd.items()}, cfg), blob)
    bad = dict(sd)
    del bad["encoder.layer.1.output.dense.bias"]
    with pytest.raises(KeyError):
        pack_bert_weights(bad, cfg)
    bad = dict(sd)
    bad["embeddings.LayerNorm.weight"] = torch.zeros(H + 1)
    with pytest.raises(ValueError):
        pack_bert_weights(bad, cfg)
    with pytest.raises(ValueError):
        BertConfig.from_hf(dict(CFG, hidden_act="relu"))


def test_encoder_fails_loudly_without_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from qppfusion import ops
    from qppfusion.encoder import BertConfig, BertQueryEncoder, WordPieceTokenizer
    sd = {k: torch.from_numpy(v) for k, v in bert_small_state().items()}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BertQueryEncoder(sd, BertConfig.from_hf(CFG), WordPieceTokenizer({t: i for i, t in enumerate(synthetic_vocab())}))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.bert_encode(torch.zeros(8, dtype=torch.bfloat16), torch.zeros((1, 8), dtype=torch.int32),
                        torch.ones(1, dtype=torch.int32), n_layers=1, hidden=256, heads=4, ffn=512, vocab=10, max_pos=8)


def test_bucket_token_rows_groups_by_padded_length():
    from qppfusion.encoder import bucket_token_rows
    rng = np.random.default_rng(2)
    rows = [rng.integers(0, 1000, size=int(n)).tolist() for n in rng.integers(1, 41, size=300)]
    seen = []
    for idx, toks, lens in bucket_token_rows(rows, 8, 40, 64, pad_id=7):
        assert toks.dtype == np.int32 and lens.dtype == np.int32 and len(idx) <= 64
        for j, i in enumerate(idx):
            assert toks[j, :lens[j]].tolist() == rows[i] and (toks[j, lens[j]:] == 7).all()
            assert toks.shape[1] == -(-len(rows[i]) // 8) * 8
        seen += idx.tolist()
    assert sorted(seen) == list(range(300))
    assert list(bucket_token_rows([], 8, 40, 64)) == []
    with pytest.raises(ValueError):
        list(bucket_token_rows([[1] * 41], 8, 40, 64))
    with pytest.raises(ValueError):
        list(bucket_token_rows([[1], []], 8, 40, 64))


def test_native_tokenizer_matches_python_and_real_berttokenizer(tmp_path):
    """qf_wordpiece_encode (C++, 7-bit fast path) against the Python tokeniser and the live transformers.BertTokenizer;
    non-ASCII texts take the Python path inside encode_batch."""
    from qppfusion.encoder import WordPieceTokenizer
    vocab = synthetic_vocab() + ["x" * 101, "##" + "y" * 30, "hello", "##llo", "he", "world", "##ld", "wor"]
    tok = WordPieceTokenizer({t: i for i, t in enumerate(vocab)}, max_length=24)
    rng = np.random.default_rng(8)
    alphabet = list("abcdefghijxy  \t\n\r!?'-.,;:()[]{}$5Z\x0b\x01\x7f") + ["what", "Capital", "unbelievably", "hello", "world",
                                                                      "helloworld", " ", "é", "京"]
    texts = ["".join(rng.choice(alphabet, size=int(rng.integers(0, 60)))) for _ in range(1500)]
    texts += ["", " ", "x" * 101, "x" * 100, "a" * 300, "hello " * 40, "\x00a\x00", "A.B,C"]
    rows = tok.encode_batch(texts, n_threads=3)
    assert len(rows) == len(texts)
    for t, r in zip(texts, rows):
        assert r.dtype == np.int32 and r.tolist() == tok.encode(t), repr(t)
        assert len(r) <= 24 and r[0] == tok.cls_id and r[-1] == tok.sep_id
    assert [r.tolist() for r in tok.encode_batch(texts[:40], n_threads=1)] == [r.tolist() for r in rows[:40]]
    assert tok.encode_batch([]) == []
    tr = pytest.importorskip("transformers")
    vp = tmp_path / "vocab.txt"
    vp.write_text("\n".join(vocab) + "\n", encoding="utf-8")
    real = tr.BertTokenizer(str(vp), do_lower_case=True)
    for t, r in list(zip(texts, rows))[::7]:
        assert r.tolist() == real(t, truncation=True, max_length=24)["input_ids"], repr(t)
    # a vocabulary without the special tokens is rejected by the C side as by the Python side
    with pytest.raises(ValueError):
        WordPieceTokenizer({"a": 0, "b": 1})


def test_encode_batch_packed_equals_rows():
    from qppfusion.encoder import WordPieceTokenizer
    tok = WordPieceTokenizer({t: i for i, t in enumerate(synthetic_vocab())}, max_length=32)
    texts = TEXTS + ["what is the capital of france", "a b c", "", "who wrote hamlet?"] * 3
    flat, lens = tok.encode_batch_packed(texts)
    rows = [tok.encode(t) for t in texts]
    assert flat.dtype == np.int32 and lens.tolist() == [len(r) for r in rows]
    assert flat.tolist() == [i for r in rows for i in r]
    ascii_only = [t for t in texts if t.isascii()]
    flat2, lens2 = tok.encode_batch_packed(ascii_only)                     # the all-native branch (no Python rows)
    assert flat2.tolist() == [i for t in ascii_only for i in tok.encode(t)] and lens2.sum() == flat2.size
    assert tok.encode_batch_packed([])[0].size == 0


def test_native_tokenizer_unicode_scope(tmp_path):
    """Accented Latin, Greek, Cyrillic, typographic punctuation, NBSP / soft hyphen / line separator go through the C++
    table-driven path and equal the Python tokeniser and the live transformers.BertTokenizer; combining marks, capital
    sigma, CJK and malformed input are flagged and take the Python path -- same result either way."""
    import ctypes
    from qppfusion import _native as N
    from qppfusion.encoder import WordPieceTokenizer
    vocab = synthetic_vocab() + ["e", "##e", "cafe2", "naive2", "##ς", "α", "β", "##β", "αβ", "д", "##а", "да", "—", "“", "”", "¿", "ss",
                                 "ß", "##ß", "æ", "##æ", "ø", "i", "##i"]
    vocab = list(dict.fromkeys(vocab))
    tok = WordPieceTokenizer({t: i for i, t in enumerate(vocab)}, max_length=40)
    rng = np.random.default_rng(21)
    alphabet = list("abcdeinoszCAFÉéèêëñÑçÇüÜöïîİıßæøÆØ¿¡—–“”‘’«»…· ­  αβγΑΒΣσςдаДЖж 	!?.,'-") + \
        ["cafe", "naive", "résumé", " ", "é", "京", "́", "ﬁ", "Å", "Å"]
    texts = ["".join(rng.choice(alphabet, size=int(rng.integers(0, 30)))) for _ in range(1500)]
    texts += ["Café au lait — “naïve” résumé…", "¿Qué?", "ΑΒ αβ", "Да", "İstanbul", "straße", "a­b", "a b", "x y"]
    toks, lens, flags = tok._native_encode(texts)
    assert flags.sum() > 0 and (flags == 0).sum() > len(texts) // 3          # both branches exercised
    for i, t in enumerate(texts):
        want = tok.encode(t)
        if flags[i] == 0:
            assert toks[i, :lens[i]].tolist() == want, repr(t)
        else:
            assert any(ord(c) >= 0x3000 or c in "́ΣÅ" for c in t), repr(t)      # flagged for a reason
    special = ["a\u00adb", "a\u00a0b", "a\u2028b", "a\u2029b", "a\u200bb", "x\u0085y", "caf\u00e9", "\u0130i", "\u00bfa?"]
    t3, l3, f3 = tok._native_encode(special)
    assert f3.tolist() == [0] * len(special)
    assert [t3[i, :l3[i]].tolist() for i in range(len(special))] == [tok.encode(t) for t in special]
    assert tok.encode("a\u00adb") == tok.encode("ab") and tok.encode("a\u2028b") == tok.encode("a b")
    rows = tok.encode_batch(texts)
    assert [r.tolist() for r in rows] == [tok.encode(t) for t in texts]
    flat, ln = tok.encode_batch_packed(texts)
    assert flat.tolist() == [x for r in rows for x in r.tolist()] and ln.tolist() == [len(r) for r in rows]
    # malformed UTF-8 straight at the C ABI: flagged, never read past the end
    bad = [b"ab\xff", b"\xc3", b"\xe2\x82", b"\xc0\xaf", b"\xed\xa0\x80", b"ok"]
    blob = b"".join(bad)
    offs = np.zeros(len(bad) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in bad], out=offs[1:])
    t2 = np.empty((len(bad), 8), dtype=np.int32); l2 = np.empty(len(bad), dtype=np.int32); f2 = np.empty(len(bad), dtype=np.uint8)
    N.check(N.lib().qf_wordpiece_encode(tok._native_handle(), blob, offs.ctypes.data, len(bad), 8, 8, 1, t2.ctypes.data,
                                        l2.ctypes.data, f2.ctypes.data))
    assert f2.tolist() == [1, 1, 1, 1, 1, 0]
    tr = pytest.importorskip("transformers")
    vp = tmp_path / "vocab.txt"
    vp.write_text("\n".join(vocab) + "\n", encoding="utf-8")
    real = tr.BertTokenizer(str(vp), do_lower_case=True)
    for i in range(0, len(texts), 3):
        if flags[i] == 0:
            assert toks[i, :lens[i]].tolist() == real(texts[i], truncation=True, max_length=40)["input_ids"], repr(texts[i])
