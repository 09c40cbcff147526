#!/usr/bin/env python3
"""GPU probe: BASELINE configs[0] (SciFact-shaped: 300 queries x 5 retrievers x top-100) -- the whole QPP -> min-max ->
W-CombSUM step eager and as a CUDA graph, plus the three ranx-style methods."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qpp-fusion-rag_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from qppfusion import ops  # noqa: E402
from qppfusion.pipeline import FusionStepGraph  # noqa: E402

dev = torch.device("cuda:0")
bench.N_DOCS, bench.POOL = 5183, 230
ids, micro, lens, feats = bench.synth_device(300, 5, 100, 11, dev)


def timed(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def eager():
    qpp, _ = ops.qpp_stats(micro, lens, feats, None, None, top_k=100, k_cut=50)
    ops.qpp_colnorm_(qpp, "minmax", round6=True)
    w = ops.qpp_select_weight(qpp, 5)
    return ops.fuse(ids, micro, lens, w, None, 0, 0, 60.0, 0, want_counts=False, want_pos=False)


g = FusionStepGraph(ids, micro, lens, feats, qpp_index=5)
print(f"C1 300 x 5 x top-100 step: eager {timed(eager):.1f} us, CUDA graph {timed(g.replay):.1f} us "
      f"({300 / timed(g.replay):.2f} M fused q/s); union mean {float(g.out.lens.float().mean()):.0f}")
for name, m, norm, fl in (("CombSUM", 0, 1, 2), ("CombMNZ", 1, 1, 2), ("RRF", 2, 0, 2)):
    us = timed(lambda: ops.fuse(ids, micro, lens, None, None, m, norm, 60.0, fl, want_counts=False, want_pos=False))
    print(f"  qf_fuse {name} (ranx semantics): {us:.1f} us per 300 queries")
