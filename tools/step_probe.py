#!/usr/bin/env python3
"""GPU probe: per-kernel CUDA-event times of the benchmark step (qpp_stats -> colnorm -> select_weight -> fuse) on
the NQ shape.  Under ncu, `-k regex:qpp_kernel` etc. pick the launch to capture."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qpp-fusion-rag_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from qppfusion import ops  # noqa: E402

dev = torch.device("cuda:0")
Q = int(os.environ.get("PROBE_Q", bench.Q_NQ))
sets = [bench.synth_device(Q, 5, 1000, 20250103 + 1000 * i, dev) for i in range(3)]
n = int(os.environ.get("PROBE_N", "20"))
names = ["qpp_stats", "colnorm", "select_weight", "fuse"]
acc = [0.0] * 4
for it in range(n + 3):
    ids, micro, lens, feats = sets[it % 3]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    qpp, _ = ops.qpp_stats(micro, lens, feats, None, None, top_k=100, k_cut=50)
    ev[1].record()
    ops.qpp_colnorm_(qpp, "minmax", round6=True)
    ev[2].record()
    w = ops.qpp_select_weight(qpp, 5)
    ev[3].record()
    out = ops.fuse(ids, micro, lens, w, None, 0, 0, 60.0, 0, want_counts=False, want_pos=False)
    ev[4].record()
    torch.cuda.synchronize()
    if it >= 3:
        for i in range(4):
            acc[i] += ev[i].elapsed_time(ev[i + 1])
print("  ".join(f"{nm}={1e3 * a / n:.1f}us" for nm, a in zip(names, acc)), f"total={1e3 * sum(acc) / n:.1f}us", flush=True)
