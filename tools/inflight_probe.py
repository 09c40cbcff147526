#!/usr/bin/env python3
"""GPU probe: throughput of the graph-captured step with 1, 2 or 3 steps in flight (one stream per in-flight step)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qpp-fusion-rag_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from qppfusion.pipeline import FusionStepGraph  # noqa: E402

dev = torch.device("cuda:0")
Q = bench.Q_NQ
NG = 6
sets = [bench.synth_device(Q, 5, 1000, 20250103 + 1000 * i, dev) for i in range(NG)]
graphs = [FusionStepGraph(*s, qpp_index=5, compact=False) for s in sets]
K = 60
for inflight in (1, 2, 3):
    streams = [torch.cuda.Stream(dev) for _ in range(inflight)]
    cur = torch.cuda.current_stream(dev)
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for s_ in streams:
            s_.wait_stream(cur)
        for i in range(K):
            with torch.cuda.stream(streams[i % inflight]):
                graphs[i % NG].replay()
        for s_ in streams:
            cur.wait_stream(s_)
        e1.record(cur)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"in flight {inflight}: {ms * 1e3:.1f} us/step  {Q / ms * 1e3 / 1e6:.2f} M fused q/s", flush=True)
