#!/usr/bin/env python3
"""[torchrun --nproc-per-node N] tools/dense_leg_probe.py [hotpot] -- bench.py's dense leg alone (NQ shape, or the
HotpotQA shape), one JSON line from rank 0.  BENCH_DENSE_INFLIGHT=1|2 selects the number of searches in flight."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qpp-fusion-rag_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

import bench

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
args = argparse.Namespace(steps=10, warmup=3, no_cpu_baseline=True)
hot = len(sys.argv) > 1 and sys.argv[1] == "hotpot"
d = bench.dense_leg(args, dev, world, rank, Q=7405, n_docs=5_233_329, label="HotpotQA shape", max_steps=4) if hot \
    else bench.dense_leg(args, dev, world, rank)
if rank == 0:
    keep = {k: d[k] for k in ("value", "ms_per_step", "kernel_ms", "threshold_failures", "parity")}
    keep["e2e"] = d["e2e"]["value"]
    keep["frac"] = d["roofline"]["frac"]
    keep["inflight"] = d["config"]["steps_in_flight"]
    print(json.dumps(keep), flush=True)
if world > 1:
    dist.destroy_process_group()
