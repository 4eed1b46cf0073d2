"""Per-rank body of tests/test_multi_gpu.py: launched by torch.distributed.run, one rank per GPU.
Viewpoints sharded over the ranks, grid + model replicated; per stage an all-gather of the per-viewpoint results and one
all-reduce MAX of the per-point ownership keys (fc-planner_b200/distributed.py); every rank must reproduce the oracle's
result bit for bit.  argv[1] = "native" (the library's own NCCL communicator, default), "torch" (collectives through
torch.distributed via the fcvm_comm_fn callback) or "staged" (every rank on device 0, gloo backend, each collective staged
through host memory: the sharded algorithm verified on a box with ONE GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    one_gpu = len(sys.argv) > 1 and sys.argv[1] == "staged"       # every rank on device 0, gloo, collectives staged through the host
    if one_gpu:
        local = 0
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from fc_planner_b200 import distributed as fd, launch_params, scenes
    from fc_planner_b200.manager import ViewpointManager
    from oracle import oracle

    pts, nrm, prims = scenes.make_plant(n_points=6000, scale=0.4, seed=0)
    prm = launch_params("mbs", z_ground=0, seed=7, device=local)
    vps = scenes.plant_viewpoints(pts, nrm, prims, 401, prm.viewpoints_distance, seed=0)   # odd: exercises shard padding
    vm, orc = ViewpointManager(prm), oracle.Oracle(prm)
    native = not (len(sys.argv) > 1 and sys.argv[1] in ("torch", "staged"))
    fd.attach(vm, native=native, staged=one_gpu)
    for m in (vm, orc):
        m.setMapPointCloud(pts); m.setModel(pts); m.setNormals(pts.astype(np.float64), nrm)
    poses = scenes.poses_from_viewpoints(vps)
    for stage in (1, 3, 4):
        cnt, rows = vm.evaluate(stage, poses)
        ocnt, orows, _ = orc.evaluate(stage, poses)
        assert np.array_equal(rows[:, :orows.shape[1]], orows) and np.array_equal(cnt, ocnt), f"stage {stage} masks differ on rank {rank}"
    calls_eval = int(vm.comm_stats()[0])
    assert calls_eval >= 3                                   # the evaluator-level API gathers the rows of every pass
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
    # golden fixtures of the shipped scenes through the sharded manager (BASELINE.json configs[2], configs[3])
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import compare_with_golden, load_golden, run_sequence
    for scene in ("christ", "pipe4x"):
        gprm, z = load_golden(scene)
        gprm.device = local
        gm = ViewpointManager(gprm)
        fd.attach(gm, native=native, staged=one_gpu)
        run_sequence(gm, z, trace=False)
        ids, pose, vox = gm.final_arrays()
        c = gm.counters()
        assert np.array_equal(ids, z["final_ids"]) and np.array_equal(pose, z["final_pose"]) and np.array_equal(vox, z["final_vox"]), f"{scene} differs on rank {rank}"
        assert np.array_equal(gm.vps_pose(), z["vps_pose"]) and np.array_equal(gm.getUncoveredModelCloud(), z["uncovered"])
        cov, con, cid = gm.cover_state()
        assert np.array_equal(cov, z["covered"]) and np.array_equal(con, z["cover_contrib"]) and np.array_equal(cid, z["contrib_id"])
        t = torch.tensor([int(c[1]), int(c[2])], dtype=torch.int64, device="cpu" if one_gpu else "cuda")
        dist.all_reduce(t)                                  # rays and lookups are per rank: their sum is the fixture's count
        assert int(t[0]) == int(z["counters"][1]) and int(t[1]) == int(z["counters"][2]), f"{scene}: ray / lookup totals differ"
        gm.close()
    st = vm.comm_stats()
    assert st[0] > calls_eval and st[2] > 0                  # all-gathers of per-viewpoint results + all-reduces of the ownership keys
    dist.barrier()
    print(f"rank {rank}/{world} ok ({'native NCCL' if native else ('gloo, staged through the host, one GPU' if one_gpu else 'torch.distributed')}): {len(gi)} final viewpoints, {int(st[0])} all-gathers "
          f"({st[1] / 1e6:.1f} MB), {int(st[2])} all-reduces ({st[3] / 1e6:.1f} MB)", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
