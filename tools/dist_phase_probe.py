#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/dist_phase_probe.py -- where the sharded dense search spends its time (NQ shape):
CUDA-event times of the four stages of distributed.dense_search_packed, max over ranks."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qpp-fusion-rag_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

import bench
from qppfusion import distributed as qd, ops

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
Q, N, D, k = bench.Q_NQ, bench.N_DENSE, bench.D_DENSE, bench.K_DENSE
n_local = (N + world - 1) // world
lo = rank * n_local
n_local = max(0, min(n_local, N - lo))
q32, corpus = bench.synth_dense(Q, n_local, D, 20250103 + rank, dev)
dist.broadcast(q32, src=0)
qb = q32.to(torch.bfloat16)
k_in = qd.send_depth(k, world)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
acc = [0.0] * 5
steps = 12
for it in range(steps + 3):
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    send, _st = ops.dense_topk_packed(qb, corpus, k_in, lo, chunks=world)
    ev[1].record()
    rows = send.shape[1] - 1
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    ev[2].record()
    flags = recv[:, rows, 0, 0].contiguous()
    merged = torch.empty((rows + 1, 2, k), dtype=torch.int32, device=dev)
    merged[rows, 0, :1] = 0
    ops.topk_merge_packed(recv, rows, k_out=k, out=merged, flag=merged[rows, 0, :1])
    ev[3].record()
    full = torch.empty((world, rows + 1, 2, k), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(full, merged)
    ev[4].record()
    torch.cuda.synchronize()
    if it >= 3:
        for i in range(4):
            acc[i] += ev[i].elapsed_time(ev[i + 1])
        acc[4] += ev[0].elapsed_time(ev[4])
t = torch.tensor(acc, dtype=torch.float64, device=dev) / steps
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    names = ["local search (packed, k_in=%d)" % k_in, "all_to_all", "merge + check", "all_gather", "total"]
    print("  ".join(f"{n}: {v:.3f} ms" for n, v in zip(names, t.tolist())), flush=True)
dist.destroy_process_group()
