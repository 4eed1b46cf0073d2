"""World-size-2 `gloo` worker for tests/test_distributed_gloo.py (CPU only).  Emulates the multi-GPU
flow of fc-planner_b200/distributed.py with the oracle standing in for the evaluator kernel: each
rank evaluates its contiguous shard of the viewpoint batch, the bitset rows are all-gathered through
the very callback the C library would call (here on a host buffer), and every rank must end up with
exactly the rows of the unsharded pass."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, port = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from fc_planner_b200 import distributed as fd, launch_params, scenes
    from oracle import oracle

    ex = fd.Exchange()
    assert ex.rank == rank and ex.world == world and not ex.on_device

    # 1. the callback on a host buffer: slice r must come from rank r
    B = 1000
    buf = np.zeros(world * B, np.uint8)
    buf[rank * B:(rank + 1) * B] = rank + 1
    assert ex(None, buf.ctypes.data, B, None) == 0
    assert all((buf[r * B:(r + 1) * B] == r + 1).all() for r in range(world))

    # 2. sharded evaluator pass == unsharded pass (an odd batch size exercises the padding)
    pts, nrm, prims = scenes.make_plant(n_points=2500, scale=0.3, seed=2)
    prm = launch_params("mbs", z_ground=0, seed=1)
    vps = scenes.plant_viewpoints(pts, nrm, prims, 37, prm.viewpoints_distance, seed=2)
    poses = scenes.poses_from_viewpoints(vps)
    orc = oracle.Oracle(prm)
    orc.setMapPointCloud(pts); orc.setModel(pts)
    n = len(poses)
    W = (len(pts) + 31) // 32
    per = (n + world - 1) // world
    lo, hi = fd.shard_bounds(n, rank, world)
    assert (lo, hi) == (min(n, per * rank), min(n, per * (rank + 1)))
    rows = np.zeros((per * world, W), np.uint32)
    if hi > lo:
        _, mine, _ = orc.evaluate(1, poses[lo:hi])
        rows[lo:hi] = mine
    assert ex(None, rows.ctypes.data, per * W * 4, None) == 0
    _, full, _ = orc.evaluate(1, poses)
    assert np.array_equal(rows[:n], full) and not rows[n:].any()
    assert full.any()

    # 3. timing / throughput reductions used by bench.py
    assert fd.max_over_ranks([float(rank), 5.0 - rank]) == [float(world - 1), 5.0]
    assert fd.sum_over_ranks([1.0, float(rank)]) == [float(world), float(sum(range(world)))]
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok", flush=True)


if __name__ == "__main__":
    main()
