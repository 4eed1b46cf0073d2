"""Per-rank body of tests/test_multi_gpu.py: launched by torch.distributed.run, one rank per GPU.
Viewpoints sharded over the ranks, grid + model replicated, bitset rows all-gathered over NCCL
(fc-planner_b200/distributed.py); every rank must reproduce the oracle's result bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from fc_planner_b200 import distributed as fd, launch_params, scenes
    from fc_planner_b200.manager import ViewpointManager
    from oracle import oracle

    pts, nrm, prims = scenes.make_plant(n_points=6000, scale=0.4, seed=0)
    prm = launch_params("mbs", z_ground=0, seed=7, device=local)
    vps = scenes.plant_viewpoints(pts, nrm, prims, 401, prm.viewpoints_distance, seed=0)   # odd: exercises shard padding
    vm, orc = ViewpointManager(prm), oracle.Oracle(prm)
    ex = fd.attach(vm)
    for m in (vm, orc):
        m.setMapPointCloud(pts); m.setModel(pts); m.setNormals(pts.astype(np.float64), nrm)
    poses = scenes.poses_from_viewpoints(vps)
    for stage in (1, 3, 4):
        cnt, rows = vm.evaluate(stage, poses)
        ocnt, orows, _ = orc.evaluate(stage, poses)
        assert np.array_equal(rows[:, :orows.shape[1]], orows) and np.array_equal(cnt, ocnt), f"stage {stage} masks differ on rank {rank}"
    calls_eval = ex.calls
    assert calls_eval >= 3
    # this rank's kernel only saw its own shard: 1/world of the viewpoint evaluations (+ padding)
    evals = int(vm.counters()[6])
    assert evals <= 3 * ((len(poses) + world - 1) // world), evals
    for m in (vm, orc):
        m.setInitViewpoints(vps)
        assert m.updateViewpoints() == 0
    gi, gp, gx = vm.final_arrays()
    oi, op, ox = orc.getUpdatedViewpoints()
    assert len(oi) > 5
    assert np.array_equal(gi, oi) and np.array_equal(gp, op) and np.array_equal(gx, ox), f"pruned set differs on rank {rank}"
    assert np.array_equal(vm.vps_pose(), orc.vps_pose())
    assert np.array_equal(vm.getUncoveredModelCloud(), orc.getUncoveredModelCloud())
    assert ex.calls > calls_eval
    dist.barrier()
    print(f"rank {rank}/{world} ok: {len(gi)} final viewpoints, {ex.calls} all-gathers, {ex.bytes / 1e6:.1f} MB gathered", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
