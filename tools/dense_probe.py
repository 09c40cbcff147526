#!/usr/bin/env python3
"""Time qf_dense_topk on the BGE/NQ shape (Q=3452, N=2,681,468, D=768, k=1000) -- dev probe."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qpp-fusion-rag_b200"))
import torch
from qppfusion import ops

Q = int(os.environ.get("Q", 3452)); N = int(os.environ.get("N", 2_681_468)); D = 768; k = 1000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(3)
corpus = torch.empty((N, D), dtype=torch.bfloat16, device=dev)
for s in range(0, N, 262144):
    x = torch.randn((min(262144, N - s), D), generator=g, device=dev)
    corpus[s:s + x.shape[0]] = (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16)
x = torch.randn((Q, D), generator=g, device=dev)
queries = (x / x.norm(dim=1, keepdim=True)).to(torch.bfloat16)
for _ in range(2):
    sc, ids, st = ops.dense_topk(queries, corpus, k)
torch.cuda.synchronize()
print("status nonzero:", int((st != 0).sum()), "min/max cand? n/a")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 5
e0.record()
for _ in range(iters):
    sc, ids, st = ops.dense_topk(queries, corpus, k)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
fl = 2.0 * Q * N * D
print(f"dense_topk: {ms:.3f} ms  {Q/ms*1e3:.0f} QPS  {fl/ms/1e9:.1f} TFLOP/s end-to-end (all kernels)")
# spot check 4 queries against torch fp32 on the bf16 inputs
qs = queries[:4].float(); ref = qs @ corpus.float().T if N <= 3_000_000 else None
if ref is not None:
    rs, ri = torch.topk(ref, k, dim=1)
    print("max |score diff| vs torch fp32 (4 queries):", float((rs - sc[:4]).abs().max()),
          "id overlap:", sum(len(set(ri[i].tolist()) & set(ids[i].tolist())) for i in range(4)) / (4 * k))
