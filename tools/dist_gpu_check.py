#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/dist_gpu_check.py -- NCCL multi-GPU check of the sharded dense search and
the sharded QPP normalisation against single-GPU results computed on rank 0."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qpp-fusion-rag_b200"))
import numpy as np
import torch
import torch.distributed as dist
from qppfusion import distributed as qd, ops

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator(device="cpu"); g.manual_seed(5)
Q, N, D, k = 300, 200_000, 768, 1000
c = torch.randn((N, D), generator=g); c = (c / c.norm(dim=1, keepdim=True)).to(torch.bfloat16)
q = torch.randn((Q, D), generator=g); q = (q / q.norm(dim=1, keepdim=True)).to(torch.bfloat16).to(dev)
lo, hi = qd.shard_range(N, world, rank)
shard = c[lo:hi].to(dev).contiguous()
ms, mi, st = qd.dense_search_sharded(q, shard, k, lo)              # packed one-exchange path (truncated lists at G >= 3)
assert st.shape == (world,) and int(st.max()) == 0, st
ms2, mi2, _ = qd.dense_search_sharded(q, shard, k, lo, merge="all_gather")
assert torch.equal(ms, ms2) and torch.equal(mi, mi2), "packed exchange and all_gather merges differ"
ms3, mi3, _ = qd.dense_search_sharded(q, shard, k, lo, exact=True)
assert torch.equal(ms, ms3) and torch.equal(mi, mi3), "packed exchange and the exact all_to_all path differ"
res = qd.dense_search_packed(q, shard, k, lo, gather=False)          # rank-local slice only
assert res.ok() and torch.equal(res.scores, ms[res.lo:res.hi]) and torch.equal(res.ids, mi[res.lo:res.hi])
# a corpus whose best documents all sit in ONE shard: the truncated exchange must notice and say so
if world >= 3:
    c2 = c.clone()
    nq2 = min(Q, (N // world) // k)                       # queries whose k best documents all fit into shard 0
    q2 = q[:nq2].contiguous()
    c2[: nq2 * k] = (q2.cpu().float()[torch.arange(nq2 * k) % nq2] + 0.02 * torch.randn((nq2 * k, D), generator=g)).to(torch.bfloat16)
    shard2 = c2[lo:hi].to(dev).contiguous()
    res2 = qd.dense_search_packed(q2, shard2, k, lo, gather=True)
    assert int(res2.flags.max()) == 0 and int(res2.incomplete.max()) != 0, "truncation loss was not detected"
    ms4, mi4, st4 = qd.dense_search_sharded(q2, shard2, k, lo, exact=True)
    if rank == 0:
        fs2, fi2, _ = ops.dense_topk(q2, c2.to(dev), k, 0)
        assert torch.equal(fs2, ms4) and torch.equal(fi2, mi4), "exact path differs from single-GPU search"
if rank == 0:
    fs, fi, st0 = ops.dense_topk(q, c.to(dev), k, 0)
    assert torch.equal(fs, ms) and torch.equal(fi, mi), "sharded dense search differs from single-GPU search"
vals = torch.randn((1000, 5, 13), generator=g, dtype=torch.float64)
qlo, qhi = qd.shard_range(1000, world, rank)
loc = vals[qlo:qhi].to(dev).contiguous()
qd.qpp_colnorm_sharded_(loc, "minmax", True)
if rank == 0:
    full = vals.to(dev).contiguous(); ops.qpp_colnorm_(full, "minmax", round6=True)
    assert torch.equal(full[qlo:qhi], loc), "sharded colnorm differs"
dist.barrier()
if rank == 0:
    print(f"dist_gpu_check OK on {world} GPUs")
dist.destroy_process_group()
