#!/usr/bin/env python3
"""GPU probe: per-phase times of qf_dense_topk (in-library CUDA events) on the BGE/NQ shape or Q/N from the env."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qpp-fusion-rag_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from qppfusion import _native as N_, ops  # noqa: E402

Q = int(os.environ.get("Q", 3452))
N = int(os.environ.get("N", 2_681_468))
dev = torch.device("cuda:0")
q32, corpus = bench.synth_dense(Q, N, 768, 20250103, dev)
qb = q32.to(torch.bfloat16)
for _ in range(3):
    sc, ids, st = ops.dense_topk(qb, corpus, 1000, 0)
torch.cuda.synchronize()
N_.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8):
    sc, ids, st = ops.dense_topk(qb, corpus, 1000, 0)
e1.record()
torch.cuda.synchronize()
ph = {n: N_.profile_read(i) for i, n in enumerate(["sample", "tau", "main", "select"])}
print(f"Q={Q} N={N}: {e0.elapsed_time(e1) / 8:.3f} ms/call  " + "  ".join(f"{k}={v:.3f}" for k, v in ph.items()),
      f" status!=0: {int((st != 0).sum())}", flush=True)
