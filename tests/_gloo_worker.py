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

    from fc_planner_b200._lib import COMM_ALLGATHER, COMM_ALLREDUCE_MAX
    # 1. the callback on a host buffer: slice r must come from rank r
    B = 1000
    buf = np.zeros(world * B, np.uint8)
    buf[rank * B:(rank + 1) * B] = rank + 1
    assert ex(None, COMM_ALLGATHER, buf.ctypes.data, B, None) == 0
    assert all((buf[r * B:(r + 1) * B] == r + 1).all() for r in range(world))
    assert ex(None, 7, buf.ctypes.data, B, None) == -1          # unknown op: refused, not raised
    # attach() refuses a backend that cannot take the device pointers the library hands out
    class _M:
        def set_shard(self, *a): raise AssertionError("must not be reached")
    try:
        fd.attach(_M(), native=False)
        raise SystemExit("attach(native=False) accepted gloo")
    except RuntimeError as e:
        assert "nccl" in str(e)

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
    assert ex(None, COMM_ALLGATHER, rows.ctypes.data, per * W * 4, None) == 0
    ocnt, full, _ = orc.evaluate(1, poses)
    assert np.array_equal(rows[:n], full) and not rows[n:].any()
    assert full.any()

    # 2b. the sharded ownership exchange: every rank reduces ITS rows to one key per point
    # (contribute_voxel_num << 32 | ~position), one all-reduce MAX, and the winners must be exactly those of the
    # reference's ordered rule (viewpoint_manager.cpp:500-528) run over all rows on one process.
    P = len(pts)
    bits = np.unpackbits(full.view(np.uint8), axis=1, bitorder="little")[:, :P].astype(bool)      # [n, P]
    keys = np.zeros(P, np.int64)
    for r in range(lo, hi):
        if ocnt[r] > 0:
            k = fd.ownership_key(ocnt[r], r)
            keys[bits[r]] = np.maximum(keys[bits[r]], k)
    assert ex(None, COMM_ALLREDUCE_MAX, keys.ctypes.data, P, None) == 0
    owner = np.full(P, -1, np.int64); contrib = np.zeros(P, np.int64)
    for r in range(n):                                     # the ordered rule, all rows, processing order = row order
        if ocnt[r] <= 0:
            continue
        take = bits[r] & ((owner < 0) | (ocnt[r] > contrib))
        owner[take] = r; contrib[take] = ocnt[r]
    for j in range(P):
        if keys[j] == 0:
            assert owner[j] == -1
        else:
            c, pos = fd.decode_ownership_key(keys[j])
            assert (c, pos) == (contrib[j], owner[j]), (j, c, pos, contrib[j], owner[j])
    assert (keys > 0).any() and keys.max() < 2 ** 63

    # 3. timing / throughput reductions used by bench.py
    assert fd.max_over_ranks([float(rank), 5.0 - rank]) == [float(world - 1), 5.0]
    assert fd.sum_over_ranks([1.0, float(rank)]) == [float(world), float(sum(range(world)))]
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok", flush=True)


if __name__ == "__main__":
    main()
