#!/usr/bin/env python
"""Concurrent D2H bandwidth of N ranks on one box (what bounds the e2e leg at N > 1: every rank copies ~0.5 GB of index lists
per evaluator pass): page-locked destination allocated with / without the rank bound to its GPU's NUMA node.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 profiles/d2h_bandwidth.py [bind]"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fc_planner_b200.distributed import bind_to_gpu_numa_node  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
bound = None
if len(sys.argv) > 1 and sys.argv[1] == "bind":
    bound = bind_to_gpu_numa_node(local, world)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 512 << 20
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
host.copy_(dev); torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    host.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = torch.tensor([10 * n / dt / 1e9], dtype=torch.float64, device="cuda")
tot = gbs.clone()
mn = gbs.clone()
if world > 1:
    dist.all_reduce(tot, op=dist.ReduceOp.SUM); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"n_gpus": world, "bound_to_numa_node": bound is not None, "host_threads_if_bound": bound, "d2h_gbs_sum": float(tot), "d2h_gbs_min_rank": float(mn),
                      "d2h_gbs_rank0": float(gbs)}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
