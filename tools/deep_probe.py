#!/usr/bin/env python3
"""GPU probe: qf_fuse_deep (workspace-resident kernel for run lists above the shared-memory envelope) -- time per query
on a few deep shapes, next to the shared-memory kernels on the deepest shape both take, plus the bit-packed upload
expansion (qf_wire_unpack) on the NQ shape.  python tools/deep_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qpp-fusion-rag_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from qppfusion import _native as N, ops  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (Q, R, K, flags, label) in [(592, 5, 1000, 0, "5 x 1000 shared-memory streaming kernel"),
                                (592, 5, 1000, N.FLAG_DEBUG_DEEP, "5 x 1000 qf_fuse_deep"),
                                (592, 6, 1000, 0, "6 x 1000 micro: shared-memory general kernel"),
                                (592, 6, 1000, N.FLAG_DEBUG_DEEP, "6 x 1000 qf_fuse_deep"),
                                (592, 8, 2000, 0, "8 x 2000 qf_fuse_deep"),
                                (296, 32, 2000, 0, "32 x 2000 qf_fuse_deep")]:
    ids, micro, lens, _ = bench.synth_device(Q, R, K, 7, dev)
    w = torch.rand((Q, R), dtype=torch.float64, device=dev).mul(1e6).round().div(1e6)
    f = lambda: ops.fuse(ids, micro, lens, w, None, 0, 0, 60.0, flags, want_counts=False, want_pos=False)
    ms = timed(f)
    out = f()
    union = float(out.lens.float().mean())
    print(f"{label:48s} Q={Q}: {ms:8.3f} ms/launch = {1e3 * ms / Q:8.2f} us/query ({1e6 * ms / Q / (R * K):6.3f} ns/posting), "
          f"union {union:.0f}, {Q / ms * 1e3:10.0f} queries/s", flush=True)
    del ids, micro, lens, out

Q, R, K = bench.Q_NQ, 5, 1000
ids, micro, lens, _ = bench.synth_device(Q, R, K, 3, dev)
wire = ops.wire_pack(ids.cpu(), micro.cpu(), lens.cpu(), pin=True).to(dev)
i2, m2 = torch.empty_like(ids), torch.empty_like(micro)
ms = timed(lambda: ops.wire_unpack_(wire, i2, m2), n=20)
assert torch.equal(i2, ids) and torch.equal(m2, micro)
nbytes = wire.words.numel() * 8 + 2 * ids.numel() * 4
print(f"qf_wire_unpack NQ shape ({wire.id_bits}+{wire.sc_bits} bits): {ms * 1e3:.1f} us, {nbytes / ms / 1e6:.0f} GB/s "
      f"({wire.words.numel() * 8 / 1e6:.1f} MB in, {2 * ids.numel() * 4 / 1e6:.1f} MB out)", flush=True)
