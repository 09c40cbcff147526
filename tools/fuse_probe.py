#!/usr/bin/env python3
"""GPU probe: time qf_fuse on the NQ shape for the general kernel and the streaming-kernel variants, count the
queries the streaming kernel defers, and check the variants agree bit for bit.  Run one variant per process:
    QF_STREAM_VARIANT=0 python tools/fuse_probe.py [general]"""
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
g = torch.Generator(device=dev)
g.manual_seed(5)
w = (torch.rand((Q, 5), generator=g, device=dev, dtype=torch.float64) * 1e6).round() / 1e6
flags = (0x200 if "general" in sys.argv else 0) | int(os.environ.get("PROBE_FLAGS", "0"), 0)
method = int(os.environ.get("PROBE_METHOD", "0"))


def run(i):
    ids, micro, lens, _ = sets[i % 3]
    return ops.fuse(ids, micro, lens, w, None, method, 0, 60.0, flags, want_counts=False, want_pos=False)


for i in range(5):
    out = run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 30
e0.record()
for i in range(n):
    out = run(i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
o0 = run(0)
torch.cuda.synchronize()
union = int(o0.lens.clamp(min=0).sum())
alg = 8 * Q * 5000 + Q * 5 * 12 + 12 * union + 8 * Q
print(f"variant={os.environ.get('QF_STREAM_VARIANT', '0')} flags={flags:#x} method={method} Q={Q}: {ms * 1e3:.1f} us/launch, "
      f"{alg / ms / 1e6:.0f} GB/s algorithmic, deferred={int((o0.lens < 0).sum())}, "
      f"checksum={int(o0.ids.to(torch.int64)[:, :16].sum())} {float(o0.scores[:, :16].sum()):.9f}", flush=True)
