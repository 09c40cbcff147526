#!/usr/bin/env python3
"""GPU probe: qf_fuse time vs union size per query (pool size of the synthetic generator), for the streaming kernel
variants (QF_STREAM_VARIANT) and the general kernel (argument `general`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qpp-fusion-rag_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from qppfusion import ops  # noqa: E402

dev = torch.device("cuda:0")
flags = 0x200 if "general" in sys.argv else (8 if "large" in sys.argv else 0)
Q = 2048
for pool in (1400, 2200, 2600, 2800, 3200, 3600, 4200):
    bench.POOL = pool
    ids, micro, lens, _ = bench.synth_device(Q, 5, 1000, 7, dev)
    w = torch.rand((Q, 5), device=dev, dtype=torch.float64)
    run = lambda: ops.fuse(ids, micro, lens, w, None, 0, 0, 60.0, flags, want_counts=False, want_pos=False)
    for _ in range(3):
        o = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        o = run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"variant={os.environ.get('QF_STREAM_VARIANT', '0')} flags={flags:#x} pool {pool}: union mean {float(o.lens.float().mean()):.0f} "
          f"max {int(o.lens.max())}  {us:.0f} us per {Q} queries = {us * 1e3 / Q:.0f} ns/query", flush=True)
