#!/usr/bin/env python3
"""
Generate tests/golden/bert_small.npz: outputs of the REAL third-party encoder stack the reference runs
(transformers.BertModel -- what SentenceTransformer("BAAI/bge-base-en-v1.5") instantiates, reference
src/retrievers/bge.py:153-170) on a small seeded random-init model, plus the token ids a real
transformers.BertTokenizer produces for a handful of strings over a synthetic vocabulary.

The weights themselves are NOT stored: bert_small_state(seed) below regenerates them bit-identically from
numpy's PCG64 stream (independent of the torch version), so the fixture stays a few kB.
"""
import os

import numpy as np

CFG = dict(vocab_size=512, hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
           max_position_embeddings=64)
TEXTS = ["what is the capital of France?", "  Café au lait -- naïve résumé\tcosts $4.50!", "who wrote 'hamlet'",
         "unbelievably long-winded antidisestablishmentarianism", "東京 tower height", "", "a"]


def bert_small_state(seed=20250104, cfg=CFG):
    """HF BertModel state dict (float32 numpy), drawn in a fixed key order from default_rng(seed)."""
    rng = np.random.default_rng(seed)
    H, F, V, P = cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"], cfg["max_position_embeddings"]
    sd = {}

    def w(name, *shape, scale=0.05):
        sd[name] = (rng.standard_normal(shape) * scale).astype(np.float32)

    def ln(prefix):
        sd[prefix + ".weight"] = (1.0 + 0.1 * rng.standard_normal(H)).astype(np.float32)
        sd[prefix + ".bias"] = (0.1 * rng.standard_normal(H)).astype(np.float32)

    w("embeddings.word_embeddings.weight", V, H, scale=0.5)
    w("embeddings.position_embeddings.weight", P, H, scale=0.2)
    w("embeddings.token_type_embeddings.weight", 2, H, scale=0.2)
    ln("embeddings.LayerNorm")
    for l in range(cfg["num_hidden_layers"]):
        p = f"encoder.layer.{l}."
        for n in ("query", "key", "value"):
            w(p + f"attention.self.{n}.weight", H, H, scale=0.08)
            w(p + f"attention.self.{n}.bias", H)
        w(p + "attention.output.dense.weight", H, H)
        w(p + "attention.output.dense.bias", H)
        ln(p + "attention.output.LayerNorm")
        w(p + "intermediate.dense.weight", F, H)
        w(p + "intermediate.dense.bias", F)
        w(p + "output.dense.weight", H, F)
        w(p + "output.dense.bias", H)
        ln(p + "output.LayerNorm")
    return sd


def synthetic_vocab():
    """A small WordPiece vocabulary: specials, ASCII letters/digits/punctuation, '##' continuations, a few words."""
    toks = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    chars = list("abcdefghijklmnopqrstuvwxyz0123456789") + list("!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~") + ["東", "京"]
    toks += chars + ["##" + c for c in chars[:36]]
    toks += ["what", "is", "the", "capital", "of", "france", "cafe", "au", "lait", "naive", "resume", "costs", "who",
             "wrote", "hamlet", "un", "##believ", "##ably", "long", "wind", "##ed", "anti", "##dis", "##establish",
             "##ment", "##arian", "##ism", "tower", "height", "##s", "##ing"]
    return toks


def main():
    import torch
    from transformers import BertConfig, BertModel, BertTokenizer
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sd = bert_small_state()
    model = BertModel(BertConfig(**CFG), add_pooling_layer=False).eval()
    missing = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not missing.unexpected_keys and all("position_ids" in k for k in missing.missing_keys), missing
    rng = np.random.default_rng(7)
    B, L = 6, 24
    tokens = rng.integers(0, CFG["vocab_size"], size=(B, L)).astype(np.int32)
    lens = np.array([24, 1, 7, 16, 23, 2], dtype=np.int32)
    mask = (np.arange(L)[None, :] < lens[:, None]).astype(np.int64)
    with torch.no_grad():
        hid = model(input_ids=torch.from_numpy(tokens).long(), attention_mask=torch.from_numpy(mask)).last_hidden_state
    cls = torch.nn.functional.normalize(hid[:, 0], p=2, dim=1).numpy()
    m = torch.from_numpy(mask).float()[:, :, None]
    mean = torch.nn.functional.normalize((hid * m).sum(1) / m.sum(1), p=2, dim=1).numpy()
    # tokeniser ids from the real BertTokenizer
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        vp = os.path.join(d, "vocab.txt")
        open(vp, "w", encoding="utf-8").write("\n".join(synthetic_vocab()) + "\n")
        tok = BertTokenizer(vp, do_lower_case=True)
        tok_ids = [tok(t, truncation=True, max_length=32)["input_ids"] for t in TEXTS]
    flat = np.array([i for r in tok_ids for i in r], dtype=np.int32)
    offs = np.cumsum([0] + [len(r) for r in tok_ids]).astype(np.int32)
    out = os.path.join(root, "tests", "golden", "bert_small.npz")
    np.savez_compressed(out, tokens=tokens, lens=lens, cls=cls.astype(np.float32), mean=mean.astype(np.float32),
                        tok_flat=flat, tok_offs=offs)
    print(out, os.path.getsize(out), "bytes; tokenizer class:", type(tok).__name__)


if __name__ == "__main__":
    main()
