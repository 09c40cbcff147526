#!/usr/bin/env python3
"""GPU probe: qf_fuse time for the six fusion entry points on the NQ shape (3452 x 5 x 1000), and qf_qpp_stats with
and without the research columns (bootstrap RSD from caller-supplied permutations)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qpp-fusion-rag_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from qppfusion import _native as N_, ops  # noqa: E402

dev = torch.device("cuda:0")
Q = bench.Q_NQ
sets = [bench.synth_device(Q, 5, 1000, 20250103 + 1000 * i, dev) for i in range(3)]
g = torch.Generator(device=dev)
g.manual_seed(5)
w = (torch.rand((Q, 5), generator=g, device=dev, dtype=torch.float64) * 1e6).round() / 1e6
ranks = torch.arange(1, 1001, device=dev, dtype=torch.int32).repeat(Q, 5, 1).contiguous()


def timed(fn, n=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


cases = [("W-CombSUM (weighted_combsum)", 0, w, None, 0, 0), ("W-CombMNZ (weighted_combmnz)", 1, w, None, 0, 0),
         ("W-RRF (weighted_rrf, file ranks)", 2, w, ranks, 0, 0), ("CombSUM (ranx sum, min-max)", 0, None, None, 1, 2),
         ("CombMNZ (ranx mnz, min-max)", 1, None, None, 1, 2), ("RRF (ranx rrf)", 2, None, None, 0, 2)]
for name, method, ww, rk, norm, flags in cases:
    def run(i, method=method, ww=ww, rk=rk, norm=norm, flags=flags):
        ids, micro, lens, _ = sets[i % 3]
        return ops.fuse(ids, micro, lens, ww, rk, method, norm, 60.0, flags, want_counts=False, want_pos=False)
    us = timed(run)
    o = run(0)
    torch.cuda.synchronize()
    print(f"{name:36s} {us:7.1f} us per {Q} queries  ({Q / us:.2f} M q/s)  status max {int(o.status.max())}", flush=True)

perms = torch.stack([torch.randperm(1000, generator=g, device=dev) for _ in range(10)]).to(torch.int32).contiguous()
idf = torch.rand((Q,), generator=g, device=dev, dtype=torch.float64) * 8 + 1
for label, pm, qi in (("bridge columns only", None, None), ("bridge + research columns (bootstrap RSD, 10 samples)", perms, idf)):
    def runq(i, pm=pm, qi=qi):
        ids, micro, lens, feats = sets[i % 3]
        return ops.qpp_stats(micro, lens, feats, qi, pm, top_k=100, k_cut=50)
    print(f"qpp_stats, {label}: {timed(runq, 10):.1f} us per {Q * 5} lists", flush=True)
