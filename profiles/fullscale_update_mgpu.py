#!/usr/bin/env python
"""BASELINE.json configs[4] in full through the manager on N GPUs of one box: P = 1 000 000 points, V = 200 000 initial
candidate viewpoints, 0.1 m grid; setInitViewpoints + updateViewpoints with the viewpoints sharded over the ranks
(fc-planner_b200/distributed.py).  No CPU number beside it (the oracle would need days); every rank must end with the same
selected set (digest compared), and size-independent properties are checked.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 profiles/fullscale_update_mgpu.py [views]"""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fc_planner_b200 import distributed as fd, launch_params, scenes  # noqa: E402
from fc_planner_b200.manager import ViewpointManager  # noqa: E402

views = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cores = len(os.sched_getaffinity(0))
pts, nrm, prims = scenes.make_plant(n_points=1_000_000, scale=1.0, seed=0)
prm = launch_params("synthetic", seed=0, resolution=0.1, device=local, host_threads=max(1, cores // world))
vps = scenes.plant_viewpoints(pts, nrm, prims, views, prm.viewpoints_distance, seed=100)
vm = ViewpointManager(prm)
if world > 1:
    fd.attach(vm)
vm.setMapPointCloud(pts); vm.setModel(pts); vm.setNormals(pts.astype(np.float64), nrm)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter(); vm.setInitViewpoints(vps); t_init = time.perf_counter() - t0
t0 = time.perf_counter(); rc = vm.updateViewpoints(); t_upd = time.perf_counter() - t0
assert rc == 0
gi, gp, gx = vm.final_arrays()
dig = int.from_bytes(hashlib.sha256(gi.tobytes() + gp.tobytes() + gx.tobytes()).digest()[:7], "little")
t = torch.tensor([t_init, t_upd], dtype=torch.float64, device="cuda")
lo = torch.tensor([dig], dtype=torch.int64, device="cuda"); hi = lo.clone()
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
c = vm.counters()
cov, con, cid = vm.cover_state()
vox = vm.vps_voxcount()
if rank == 0:
    print(json.dumps({"n_gpus": world, "P": len(pts), "V_initial": len(vps), "host_threads_per_rank": max(1, cores // world),
                      "setInitViewpoints_s": float(t[0]), "updateViewpoints_s": float(t[1]), "final_viewpoints": int(len(gi)),
                      "uncovered_points": int(len(vm.getUncoveredModelCloud())), "candidates_total": int(len(vm.vps_pose())),
                      "rays_this_rank": int(c[1]), "cap_trips": int(c[4]), "ranks_agree": bool(int(lo) == int(hi)),
                      "owned_points_equal_sum_of_counts": bool(int(cov.sum()) == int(vox.sum())),
                      "final_counts_positive": bool((gx > 0).all())}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
