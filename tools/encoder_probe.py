#!/usr/bin/env python3
"""GPU probe for the query encoder (qf_bert_encode): accuracy against the numpy oracle and launch timing on the
BGE-base shape (12 layers, hidden 768, 3 452 NQ-sized query batch).  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qpp-fusion-rag_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)

from make_golden_bert import bert_small_state  # noqa: E402
from oracle import bert_oracle  # noqa: E402
from qppfusion import ops  # noqa: E402
from qppfusion.encoder import BertConfig, pack_bert_weights  # noqa: E402

BASE = dict(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
            max_position_embeddings=512)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=3452)
    ap.add_argument("--len", type=int, default=32)
    ap.add_argument("--check", type=int, default=12)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    sd = bert_small_state(seed=42, cfg=BASE)
    cfg = BertConfig.from_hf(BASE)
    blob = pack_bert_weights({k: torch.from_numpy(v) for k, v in sd.items()}, cfg).to(dev)
    rng = np.random.default_rng(1)
    B, L = a.queries, a.len
    tokens = rng.integers(0, cfg.vocab, size=(B, L)).astype(np.int32)
    lens = np.clip(rng.poisson(11, size=B) + 3, 3, L).astype(np.int32)
    lens[0] = L
    t_d, l_d = torch.from_numpy(tokens).to(dev), torch.from_numpy(lens).to(dev)
    kw = dict(n_layers=cfg.n_layers, hidden=cfg.hidden, heads=cfg.heads, ffn=cfg.ffn, vocab=cfg.vocab, max_pos=cfg.max_pos)
    out_b, out_f, status = ops.bert_encode(blob, t_d, l_d, **kw)
    torch.cuda.synchronize()
    res = {"B": B, "L": L, "status_max": int(status.max().item())}
    n = a.check
    sd_bf = {k: bert_oracle._bf16_round(v) for k, v in sd.items()}
    t0 = time.time()
    ref32 = bert_oracle.bert_encode(sd, tokens[:n], lens[:n], n_layers=12, heads=12)
    refbf = bert_oracle.bert_encode(sd_bf, tokens[:n], lens[:n], n_layers=12, heads=12)
    refem = bert_oracle.bert_encode(sd_bf, tokens[:n], lens[:n], n_layers=12, heads=12, emulate_bf16=True)
    res["oracle_s"] = round(time.time() - t0, 2)
    got = out_f[:n].double().cpu().numpy()
    for name, ref in (("fp32_weights", ref32), ("bf16_weights", refbf), ("bf16_emulated", refem)):
        res[name] = {"min_cos": float((got * ref).sum(1).min()), "max_abs": float(np.abs(got - ref).max())}
    res["bf16_out_vs_f32_out_max_abs"] = float((out_b.float() - out_f).abs().max().item())
    res["norm_err"] = float((out_f.double().norm(dim=1) - 1).abs().max().item())
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.bert_encode(blob, t_d, l_d, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    per_tok = 2 * 12 * (4 * 768 * 768 + 2 * 768 * 3072) + 12 * 4 * L * 768
    res["ms"] = ms
    res["queries_per_s"] = B / ms * 1e3
    res["tflops"] = per_tok * B * L / ms / 1e9
    # the host layer's packed (variable-length) path on the same rows (H2D of the token ids included, tokenisation not)
    from qppfusion.encoder import BertQueryEncoder
    enc = BertQueryEncoder.__new__(BertQueryEncoder)
    enc.cfg, enc.tokenizer, enc.device, enc.batch_size, enc.weights = cfg, None, dev, 8192, blob
    rows = [tokens[i, :lens[i]].tolist() for i in range(B)]
    b2, f2 = enc.encode_token_lists(rows)
    torch.cuda.synchronize()
    res["packed_vs_padded_max_abs"] = float((f2 - out_f).abs().max().item())
    ts = []
    for _ in range(a.iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        enc.encode_token_lists(rows)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    res["packed_ms"] = float(np.median(ts))
    res["packed_queries_per_s"] = B / res["packed_ms"] * 1e3
    res["real_tokens"] = int(lens.sum())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
