"""
TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's query encoder.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module, and only as the checker.  The product path (qppfusion/encoder.py -> qf_bert_encode) never imports it.

Restates BGERetriever._encode_queries (reference src/retrievers/bge.py:160-170):
    SentenceTransformer("BAAI/bge-base-en-v1.5").encode(queries, normalize_embeddings=True)
The arithmetic lives in third-party code that is absent from /root/reference: sentence-transformers
(requirements.txt, unpinned) running HF transformers' BertModel.  Its published algorithm, restated here in
numpy: BertEmbeddings (word + token_type(0) + position, LayerNorm) -> 12 x BertLayer (post-LN: self-attention
with an additive -inf key mask, scale 1/sqrt(64); output.dense + residual + LayerNorm; intermediate.dense +
exact erf GELU; output.dense + residual + LayerNorm) -> Pooling(CLS token; bge's 1_Pooling/config.json has
pooling_mode_cls_token = true) -> Normalize (x / max(||x||_2, 1e-12)).

PINNED against transformers.BertModel (the same class sentence-transformers instantiates; transformers 5.5 is in
this image) on seeded random-init weights: tests/test_oracle_bert.py, and against the committed outputs of that
model in tests/golden/bert_small.npz (tools/make_golden_bert.py).  Not pinned against the BAAI checkpoint itself
(no network, no weights in the container).
"""
import math

import numpy as np


def _bf16_round(x: np.ndarray) -> np.ndarray:
    """Round float32 to the nearest bfloat16 (ties to even), returned as float32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


try:
    from scipy.special import erf as _erf
except ImportError:                                              # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float64])


def _layer_norm(x, g, b, eps):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)           # biased, as torch.nn.LayerNorm
    return (x - mu) / np.sqrt(var + eps) * g + b


def bert_encode(sd: dict, tokens: np.ndarray, lens: np.ndarray, *, n_layers: int, heads: int, pooling: str = "cls",
                ln_eps: float = 1e-12, normalize: bool = True, emulate_bf16: bool = False) -> np.ndarray:
    """sd: HF BertModel state dict (names without the 'bert.' prefix) as float arrays; tokens int [B, L];
    lens int [B] (number of real tokens per row).  -> float64 [B, H] sentence embeddings.
    emulate_bf16: round the activations to bf16 at the points where the device path stores them (after every
    Linear, LayerNorm, GELU and attention output) -- the weights are whatever precision `sd` holds."""
    f = (lambda a: _bf16_round(a).astype(np.float64)) if emulate_bf16 else (lambda a: a)
    W = {k: np.asarray(v, dtype=np.float64) for k, v in sd.items()}
    tokens = np.asarray(tokens)
    B, L = tokens.shape
    lens = np.clip(np.asarray(lens, dtype=np.int64), 1, L)
    H = W["embeddings.word_embeddings.weight"].shape[1]
    hd = H // heads
    x = (W["embeddings.word_embeddings.weight"][tokens] + W["embeddings.token_type_embeddings.weight"][0]
         + W["embeddings.position_embeddings.weight"][:L][None])
    x = f(_layer_norm(x, W["embeddings.LayerNorm.weight"], W["embeddings.LayerNorm.bias"], ln_eps))
    key_ok = np.arange(L)[None, :] < lens[:, None]                # [B, L]
    for l in range(n_layers):
        p = f"encoder.layer.{l}."
        lin = lambda t, name: t @ W[p + name + ".weight"].T + W[p + name + ".bias"]
        q = f(lin(x, "attention.self.query")).reshape(B, L, heads, hd).transpose(0, 2, 1, 3)
        k = f(lin(x, "attention.self.key")).reshape(B, L, heads, hd).transpose(0, 2, 1, 3)
        v = f(lin(x, "attention.self.value")).reshape(B, L, heads, hd).transpose(0, 2, 1, 3)
        s = q @ k.transpose(0, 1, 3, 2) / math.sqrt(hd)
        s = np.where(key_ok[:, None, None, :], s, -np.inf)
        s = s - s.max(axis=-1, keepdims=True)
        e = np.exp(s)
        a = e / e.sum(axis=-1, keepdims=True)
        ctx = f((a @ v).transpose(0, 2, 1, 3).reshape(B, L, H))
        t = f(lin(ctx, "attention.output.dense"))
        x = f(_layer_norm(t + x, W[p + "attention.output.LayerNorm.weight"], W[p + "attention.output.LayerNorm.bias"], ln_eps))
        h = f(lin(x, "intermediate.dense"))
        h = f(0.5 * h * (1.0 + _erf(h / math.sqrt(2.0))))
        t = f(lin(h, "output.dense"))
        x = f(_layer_norm(t + x, W[p + "output.LayerNorm.weight"], W[p + "output.LayerNorm.bias"], ln_eps))
    if pooling == "cls":
        emb = x[:, 0, :]
    elif pooling == "mean":
        m = key_ok[:, :, None].astype(np.float64)
        emb = (x * m).sum(axis=1) / m.sum(axis=1)
    else:
        raise ValueError(f"unknown pooling {pooling!r}")
    if normalize:
        emb = emb / np.maximum(np.sqrt((emb ** 2).sum(axis=1, keepdims=True)), 1e-12)
    return emb

