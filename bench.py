#!/usr/bin/env python
"""bench.py — visibility rays/s of the candidate-viewpoint x surface-point evaluator pass, plus the legs `pruning`
(viewpoint-pruning wall time, BASELINE.json's second metric), `coverage_evaluation` and `planner_grid`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[4], the synthetic scaled scene): P = 1 000 000 surface points on a
0.1 m grid (100 MB bit-packed), MBS camera parameters; every GPU evaluates its own 25 000 candidate
viewpoints per step (200 000 / 8: at N = 8 one step is exactly configs[4]; weak scaling).
A step = one E1 evaluator pass (viewpoint_manager.cpp:358-425 for every viewpoint of the batch):
FOV + range test of every (viewpoint, point) pair, a bidirectional voxel ray cast for every pair
that passes, coverage bitsets and per-viewpoint counts out.  A ray = one BiRC invocation.

  value  rays/s with poses, points and grid resident in HBM (bitset rows stay on the device)
  e2e    rays/s through the C ABI call fcvm_evaluate_csr() with HOST (page-locked) buffers: host pose -> FOV normals,
         H2D, kernels, extraction of the visible sets as index lists, D2H of the per-viewpoint counts, offsets and lists
         (what the manager consumes from a pass); e2e_rows: the same pass returning the raw coverage bitsets
  roofline  k_eval alone: algorithmic bytes / CUDA-event time / measured HBM peak
  cpu_baseline  the CPU oracle (port of the reference) on a bounded sample of the same batch
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "visibility_rays_per_s"
UNIT = "rays/s"


def build_scene(args, rank):
    from fc_planner_b200 import launch_params, scenes
    pts, nrm, prims = scenes.make_plant(n_points=args.points, scale=args.scale, seed=0)
    prm = launch_params("synthetic", seed=0, resolution=args.resolution)
    vps = scenes.plant_viewpoints(pts, nrm, prims, args.views, prm.viewpoints_distance, seed=100 + rank)
    poses = scenes.poses_from_viewpoints(vps)
    return pts, nrm, prm, poses


def workload_name(args):
    return (f"synthetic plant seed 0 (BASELINE configs[4] geometry): P={args.points} points, {args.resolution} m grid, "
            f"{args.views} candidate viewpoints per GPU per step, E{args.stage} evaluator pass")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            self.path = tempfile.mktemp(suffix=".csv")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}",
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx = float(f[1])
                except ValueError:
                    continue
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}
        return out


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


_THREADS_PER_RANK = None      # set by main() when the ranks are bound to their GPUs' NUMA nodes


def threads_per_rank(world):
    return _THREADS_PER_RANK or max(1, host_cores() // max(1, world))


def cpu_evaluator(prm, pts, kind):
    """kind 'reference': the reference's own RayCaster / PerceptionUtils (raycast.cpp, perception_utils.cpp compiled from
    /root/reference into oracle/_ref, driven by viewpoint_manager.cpp's loops); 'port': the oracle's restatement."""
    from oracle import oracle
    orc = oracle.RefEvaluator(prm) if kind == "reference" else oracle.Oracle(prm)
    orc.setMapPointCloud(pts); orc.setModel(pts)
    return orc


def cpu_kind():
    from oracle import oracle
    return "reference" if oracle.ref_available() else "port"


def cpu_oracle_rate(pts, nrm, prm, poses, seconds, threads, stage=1, kind="port", orc=None):
    """Time the CPU implementation's evaluator pass on a bounded prefix of the batch.  Returns (rays/s, sample, s, counters, counts)."""
    orc = orc or cpu_evaluator(prm, pts, kind)
    n0 = min(len(poses), max(threads, 4))
    t = time.perf_counter(); _, _, c = orc.evaluate(stage, poses[:n0], threads=threads); dt = time.perf_counter() - t
    n = int(min(len(poses), max(n0, n0 * seconds / max(dt, 1e-3))))
    t = time.perf_counter(); cnt, rows, c = orc.evaluate(stage, poses[:n], threads=threads); dt = time.perf_counter() - t
    del rows
    return c[1] / dt, n, dt, c, cnt


def pruning_leg(args, device, with_cpu, world=1, rank=0):
    """BASELINE.json's second metric: viewpoint-pruning wall time = setInitViewpoints + updateViewpoints
    (the whole coverage-driven update: E1/E2, ownership, gravitation prune with E3/E4, uncovered-area
    rounds) through the public call sequence, next to the single-threaded CPU oracle (the reference is
    single-threaded on this path) on the same inputs, with the selected viewpoint sets compared."""
    from fc_planner_b200 import launch_params, scenes
    from fc_planner_b200.manager import ViewpointManager
    pts, nrm, prims = scenes.make_plant(n_points=args.prune_points, scale=0.5, seed=0)
    # the ranks of one box share its cores: each rank reduces only its own shard of the viewpoints on its share of the threads
    prm = launch_params("synthetic", seed=0, resolution=0.2, device=device, host_threads=threads_per_rank(world))
    vps = scenes.plant_viewpoints(pts, nrm, prims, args.prune_views, prm.viewpoints_distance, seed=5)

    def run(m):
        m.reset(); m.setMapPointCloud(pts); m.setModel(pts); m.setNormals(pts.astype(np.float64), nrm)
        t = time.perf_counter()
        m.setInitViewpoints(vps)
        rc = m.updateViewpoints()
        dt = time.perf_counter() - t
        assert rc == 0
        return dt

    # The first run on a fresh handle is the one compared with the oracle: the reference's prune maps keep
    # their bucket arrays across reset() (viewpoint_manager.h:169-177), so the iteration order - and with it
    # the selected set - of a handle depends on its history, and the oracle below is a fresh handle too.
    vm = ViewpointManager(prm)
    if world > 1:
        # viewpoints sharded over the ranks, grid + model replicated; per stage an all-gather of the per-viewpoint results
        # and one all-reduce MAX of the ownership keys (native NCCL inside the library); every rank obtains the same result
        from fc_planner_b200 import distributed as fd
        fd.attach(vm)
    first_s = run(vm)                         # includes every device allocation of the handle
    gi, gp, gx = vm.final_arrays()
    c = vm.counters()
    n_unc = int(len(vm.getUncoveredModelCloud()))
    gpu_s = min(run(vm), run(vm))
    st = vm.comm_stats()
    if world > 1:
        gpu_s, first_s = fd.max_over_ranks([gpu_s, first_s], device="cuda")     # wall time: max over ranks
        ranks_agree = ranks_agree_on(gi, gp, gx)
    out = {"workload": f"synthetic plant scale 0.5: P={len(pts)} points, {len(vps)} initial candidate viewpoints, 0.2 m grid, "
                       f"max_iter_num={prm.max_iter_num}; setInitViewpoints + updateViewpoints",
           "gpu_ms": 1e3 * gpu_s, "gpu_ms_first_call": 1e3 * first_s, "final_viewpoints": int(len(gi)), "uncovered_points": n_unc,
           "rays": int(c[1]), "kernel_launches": int(c[5]), "n_gpus": world,
           "scaling": "strong (one fixed problem sharded over the ranks; latency-bound at this size)"}
    if world > 1:
        out.update({"ranks_agree": ranks_agree, "allgathers_per_run": int(st[0]), "allgather_mb_per_run": st[1] / 1e6,
                    "allreduces_per_run": int(st[2]), "allreduce_mb_per_run": st[3] / 1e6})
    if with_cpu and rank == 0:
        from oracle import oracle
        orc = oracle.Oracle(prm)
        cpu_s = run(orc)
        oi, op, ox = orc.getUpdatedViewpoints()
        out.update({"cpu_ms": 1e3 * cpu_s, "cpu_cores": 1, "cpu_kind": "port", "speedup": cpu_s / gpu_s,
                    "selected_set_identical": bool(np.array_equal(gi, oi) and np.array_equal(gp, op) and np.array_equal(gx, ox))})
    return out


def ranks_agree_on(*arrays):
    """Every rank must hold the same result: a digest of the arrays, MIN and MAX over the ranks compared."""
    import hashlib
    import torch
    import torch.distributed as dist
    dig = int.from_bytes(hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in arrays)).digest()[:7], "little")
    t = torch.tensor([dig], dtype=torch.int64, device="cuda")
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(int(lo) == int(hi))


def pruning_full_leg(args, device, world=1, rank=0):
    """BASELINE.json configs[4] in full through the public call sequence: P = 10^6 surface points x 200 000 initial candidate
    viewpoints on a 0.1 m grid, setInitViewpoints + updateViewpoints (hcplanner.cpp:709-716), STRONG scaling: the same
    problem at every N, viewpoints sharded over the ranks.  Reported with the library's own phase split (fcvm_get_timers,
    rank 0's view) and the bytes the collectives moved.  The CPU oracle would need days here; parity at this size is
    covered by size-independent properties (tests/test_gpu_fullsize.py) and by `ranks_agree`."""
    from fc_planner_b200 import launch_params, scenes
    from fc_planner_b200.manager import ViewpointManager
    pts, nrm, prims = scenes.make_plant(n_points=args.points, scale=args.scale, seed=0)
    prm = launch_params("synthetic", seed=0, resolution=args.resolution, device=device, host_threads=threads_per_rank(world))
    vps = scenes.plant_viewpoints(pts, nrm, prims, args.full_views, prm.viewpoints_distance, seed=100)
    vm = ViewpointManager(prm)
    if world > 1:
        import torch.distributed as dist
        from fc_planner_b200 import distributed as fd
        fd.attach(vm)
    nd = pts.astype(np.float64)

    def run():
        vm.reset(); vm.setMapPointCloud(pts); vm.setModel(pts); vm.setNormals(nd, nrm)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter(); vm.setInitViewpoints(vps); t1 = time.perf_counter()
        rc = vm.updateViewpoints(); t2 = time.perf_counter()
        assert rc == 0
        return t1 - t0, t2 - t1

    first = run()                                   # device allocations, page-locking, NCCL channel set-up
    init_s, upd_s = min((run() for _ in range(2)), key=sum)     # the host-serial parts pick up the noise of a shared box: best of two
    tm, st, c = vm.timers(), vm.comm_stats(), vm.counters()
    gi, gp, gx = vm.final_arrays()
    cov, con, cid = vm.cover_state()
    vox = vm.vps_voxcount()
    agree = True
    if world > 1:
        init_s, upd_s, f0, f1 = fd.max_over_ranks([init_s, upd_s, first[0], first[1]], device="cuda")
        first = (f0, f1)
        agree = ranks_agree_on(gi, gp, gx, cid)
    out = {"workload": f"BASELINE configs[4]: synthetic plant seed 0, P={len(pts)} points x {len(vps)} initial candidate viewpoints, {args.resolution} m grid, "
                       f"max_iter_num={prm.max_iter_num}; setInitViewpoints + updateViewpoints",
           "n_gpus": world, "scaling": "strong", "seconds": init_s + upd_s, "setInitViewpoints_s": init_s, "updateViewpoints_s": upd_s,
           "first_call_seconds": first[0] + first[1], "final_viewpoints": int(len(gi)), "uncovered_points": int((cid < 0).sum()),
           "candidates_total": int(len(vm.vps_pose())), "rays_this_rank": int(c[1]), "cap_trips": int(c[4]), "kernel_launches_this_rank": int(c[5]),
           "host_threads_per_rank": prm.host_threads, "ranks_agree": agree,
           "owned_points_equal_sum_of_counts": bool(int(cov.sum()) == int(vox.sum())), "final_counts_positive": bool((gx > 0).all()),
           "phase_ms_rank0": {k: round(v, 2) for k, v in tm.items()},
           "collectives": {"allgathers": int(st[0]), "allgather_mb": st[1] / 1e6, "allreduces": int(st[2]), "allreduce_mb": st[3] / 1e6}}
    vm.close()
    return out


def golden_sharded_leg(device, world=1, rank=0):
    """BASELINE.json configs[2] (Christ the Redeemer) and configs[3] (the Pacifico substitute: pipe with 4x the candidate
    density) through the manager sharded over this run's ranks, against the committed fixtures of the CPU oracle
    (tests/golden/*.npz, themselves checked against the reference's own viewpoint_manager.cpp compiled here)."""
    from fc_planner_b200 import launch_params
    from fc_planner_b200.manager import ViewpointManager
    out = {}
    for scene in ("christ", "pipe4x"):
        try:
            z = np.load(os.path.join(ROOT, "tests", "golden", f"{scene}.npz"))
        except Exception as e:
            out[scene] = {"unavailable": f"fixture not found: {e}"}
            continue
        prm = launch_params({"pipe4x": "pipe"}.get(scene, scene), seed=int(z["seed"]), device=device)
        vm = ViewpointManager(prm)
        if world > 1:
            from fc_planner_b200 import distributed as fd
            fd.attach(vm)

        def run():
            vm.reset(); vm.setMapPointCloud(z["map_cloud"]); vm.setModel(z["model"]); vm.setNormals(z["model"].astype(np.float64), z["normals"])
            t = time.perf_counter(); rc = vm.updateViewpoints(); dt = time.perf_counter() - t
            assert rc == 0
            return dt
        first = run()
        ids, pose, vox = vm.final_arrays()
        cov, con, cid = vm.cover_state()
        ok = bool(np.array_equal(ids, z["final_ids"]) and np.array_equal(pose, z["final_pose"]) and np.array_equal(vox, z["final_vox"])
                  and np.array_equal(vm.vps_pose(), z["vps_pose"]) and np.array_equal(cid, z["contrib_id"]) and np.array_equal(con, z["cover_contrib"])
                  and np.array_equal(vm.getUncoveredModelCloud(), z["uncovered"]))
        warm = min(run(), run())
        if world > 1:
            import torch
            import torch.distributed as dist
            t = torch.tensor([1 if ok else 0], dtype=torch.int64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = bool(int(t))
            warm, first = fd.max_over_ranks([warm, first], device="cuda")
        out[scene] = {"matches_golden": ok, "n_gpus": world, "P": int(len(z["model"])), "candidates": int(len(z["vps_pose"])), "final_viewpoints": int(len(ids)),
                      "updateViewpoints_ms": 1e3 * warm, "first_call_ms": 1e3 * first}
        vm.close()
    return out


def coverage_leg(args, device, with_cpu):
    """the planner's pin-hole coverage evaluation (hcplanner.cpp:1032-1211) over the reference's
    own shipped MBS trajectory (9 311 poses x the 75 620-point full cloud; tests/golden/cov_mbs.npz), through the public
    call with host buffers, next to the single-threaded CPU oracle (the reference runs this loop on one thread)."""
    from fc_planner_b200 import launch_camera, launch_params
    from fc_planner_b200.coverage import CoverageEvaluator
    g = os.path.join(ROOT, "tests", "golden")
    try:
        cloud = np.load(os.path.join(g, "mbs.npz"))["map_cloud"]
        z = np.load(os.path.join(g, "cov_mbs.npz"))
    except Exception as e:                                   # fixtures missing: say so instead of inventing a workload
        return {"unavailable": f"fixture not found: {e}"}
    prm, cam = launch_params("mbs", device=device), launch_camera("mbs")
    poses = z["poses"]
    ce = CoverageEvaluator(prm, cam)
    ce.setCloud(cloud)
    ce.CoverageEvaluation(poses)                             # allocations
    best, kern, scan = 1e9, 0.0, 0.0
    for _ in range(5):
        t = time.perf_counter(); rate, offs, ids, ever = ce.CoverageEvaluation(poses); dt = time.perf_counter() - t
        if dt < best:
            best, kern, scan = dt, ce.last_kernel_ms(), ce.last_scan_ms()
    c0 = ce.counters()
    ce.CoverageEvaluation(poses)
    c = ce.counters() - c0
    m, n = len(poses), len(cloud)
    # algorithmic bytes of one k_cov_scan launch: one 16-B PointXYZ per insideFOV call + the three image-table entries
    # (state 8 B, distance 8 B, point id 4 B) per projection + the two bitset rows per pose
    b_alg = 16.0 * c[0] + 20.0 * c[1] + 2.0 * m * (n / 8.0)
    peak = 6554.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out = {"workload": f"MBS shipped trajectory: {m} poses x {n} points, 640x480 pin-hole image; CoverageEvaluation (rate + visible_group per pose)",
           "gpu_ms": 1e3 * best, "kernel_ms": kern, "scan_kernel_ms": scan, "coverage_rate": rate, "group_ids": int(len(ids)),
           "pair_tests": int(c[0]), "projections": int(c[1]), "chunks_scanned": int(c[4]), "chunks_total": int(m * ((n + ce.chunk_points() - 1) // ce.chunk_points())), "chunk_points": ce.chunk_points(),
           "pair_tests_per_s": c[0] / (scan * 1e-3), "kernel_launches": int(c[3]),
           "matches_fixture": bool(rate == float(z["rate"]) and np.array_equal(offs, z["group_offsets"]) and np.array_equal(ids, z["group_ids"])),
           "reference_output_agreement": {"same": int(z["agreement"][0]), "only_reference": int(z["agreement"][1]),
                                          "only_ours": int(z["agreement"][2]),
                                          "note": "vs the reference's shipped solution/Traj/CloudInfoMBS.txt; its poses are printed with 6 decimals"},
           "roofline": {"bound": "hbm", "kernel": "k_cov_scan", "achieved": b_alg / (scan * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": b_alg / (scan * 1e-3) / 1e9 / peak, "algorithmic_bytes": b_alg,
                        "note": "algorithmic = what the reference loop touches; whole chunks of consecutive points are culled exactly, so fewer bytes move"}}
    if with_cpu:
        from oracle import oracle
        orc = oracle.CoverageOracle(prm, cam)
        orc.setCloud(cloud)
        t = time.perf_counter(); o_rate, o_offs, o_ids, o_ever = orc.CoverageEvaluation(poses); cpu_s = time.perf_counter() - t
        out.update({"cpu_ms": 1e3 * cpu_s, "cpu_cores": 1, "cpu_kind": "port", "speedup": cpu_s / best,
                    "identical_to_oracle": bool(rate == o_rate and np.array_equal(offs, o_offs) and np.array_equal(ids, o_ids)
                                                and np.array_equal(ever, o_ever))})
    return out


def planner_grid_leg(args, device, with_cpu):
    """HCSolver::search_Path's straight-line test for all ordered pairs of n positions
    (hcsolver.cpp:556-583) and lightStateDet (hcplanner.cpp:805-863) on the MBS grid (hc = initHCMap of the shipped full
    cloud, inflate = its 1-voxel dilation), next to the CPU oracle on all host threads / one thread."""
    from fc_planner_b200 import launch_params
    from fc_planner_b200.planner_grid import PlannerGrid
    try:
        z = np.load(os.path.join(ROOT, "tests", "golden", "mbs.npz"))
    except Exception as e:
        return {"unavailable": f"fixture not found: {e}"}
    cloud = z["map_cloud"]
    prm = launch_params("mbs")
    res, infl = prm.resolution, prm.visible_range
    mn, mx = cloud.min(0), cloud.max(0)
    origin = mn.astype(np.float64) - 2 * infl
    dims = np.ceil((4 * infl + mx.astype(np.float64) - mn) / res).astype(np.int32)
    idx = np.floor((cloud.astype(np.float64) - origin) / res).astype(np.int64)
    hc = np.zeros(tuple(dims), np.int8)
    hc[idx[:, 0], idx[:, 1], idx[:, 2]] = 1
    inflate = hc.copy()
    for ax in range(3):
        inflate = np.maximum(inflate, np.maximum(np.roll(inflate, 1, ax), np.roll(inflate, -1, ax)))
    rng = np.random.default_rng(0)
    n = args.los_points
    P = np.concatenate([z["final_pose"][:, :3], rng.uniform(mn - 6.0, mx + 6.0, (max(0, n - len(z["final_pose"])), 3))])[:n]
    pg = PlannerGrid(origin, res, dims, hc, inflate, None, device=device)
    pg.lineOfSight(P)
    best, kern = 1e9, 0.0
    for _ in range(3):
        t = time.perf_counter(); safe, dist = pg.lineOfSight(P); dt = time.perf_counter() - t
        if dt < best:
            best, kern = dt, pg.last_kernel_ms()
    c0 = pg.counters(); pg.lineOfSight(P, want_dist=False); c = pg.counters() - c0
    rosa = cloud[rng.integers(0, len(cloud), 400)]
    cand = np.repeat(z["vps_pose"][:, :3], 8, axis=0) + rng.normal(0, 0.5, (8 * len(z["vps_pose"]), 3))
    pg.lightStateDet(cand, rosa)
    t = time.perf_counter(); det, cross, near = pg.lightStateDet(cand, rosa); ls_s = time.perf_counter() - t
    out = {"workload": f"MBS grid {dims[0]}x{dims[1]}x{dims[2]} at {res} m; search_Path straight-line test for all {n}x{n} ordered pairs; "
                       f"lightStateDet for {len(cand)} candidates x {len(rosa)} ROSA vertices",
           "los_gpu_ms": 1e3 * best, "los_kernel_ms": kern, "los_pairs": int(n * n), "los_safe_pairs": int(safe.sum()), "los_voxel_lookups": int(c[0]),
           "los_pairs_per_s": n * n / (kern * 1e-3), "light_state_gpu_ms": 1e3 * ls_s, "light_state_kernel_ms": pg.last_kernel_ms(),
           "light_state_outside": int(det.sum()), "cap_trips": int(pg.counters()[1])}
    if with_cpu:
        from oracle import oracle
        orc = oracle.PlanOracle(origin, res, dims, hc, inflate, None)
        threads = host_cores()
        m = min(n, 1024)
        t = time.perf_counter(); o_safe, o_dist = orc.lineOfSight(P[:m], threads=threads); cpu_s = time.perf_counter() - t
        t = time.perf_counter(); orc.lineOfSight(P[:min(m, 256)], threads=1); cpu1_s = time.perf_counter() - t
        t = time.perf_counter(); o_det, o_cross, o_near = orc.lightStateDet(cand, rosa); cpu_ls = time.perf_counter() - t
        out.update({"los_cpu_pairs_per_s": m * m / cpu_s, "los_cpu_cores": threads, "los_cpu_single_thread_pairs_per_s": min(m, 256) ** 2 / cpu1_s,
                    "cpu_kind": "port", "los_identical_to_oracle": bool(np.array_equal(safe[:m, :m], o_safe) and np.array_equal(dist[:m, :m], o_dist)),
                    "light_state_cpu_ms": 1e3 * cpu_ls, "light_state_identical_to_oracle": bool(np.array_equal(det, o_det) and np.array_equal(cross, o_cross)
                                                                                               and np.array_equal(near, o_near))})
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on all host cores, same metric and config.
    The reference's own raycast.cpp + perception_utils.cpp, compiled from /root/reference into oracle/_ref and driven by the
    loops of viewpoint_manager.cpp:358-425 ('kind': 'reference'), when that library travelled with the repo; the oracle's
    restatement ('port') otherwise.  The reference is single-threaded on this path; here its viewpoints are handed out to
    every host thread."""
    if rank != 0:
        return
    pts, nrm, prm, poses = build_scene(args, 0)
    threads = host_cores()
    kind = cpu_kind()
    orc = cpu_evaluator(prm, pts, kind)
    # size one step so that warmup + steps end within a few minutes
    n0 = max(threads, 4)
    t = time.perf_counter(); orc.evaluate(args.stage, poses[:n0], threads=threads); dt0 = time.perf_counter() - t
    per_step_s = min(20.0, 150.0 / max(1, args.steps + args.warmup))
    n = int(min(len(poses), max(n0, n0 * per_step_s / max(dt0, 1e-3))))
    for _ in range(args.warmup):
        orc.evaluate(args.stage, poses[:n], threads=threads)
    rays, t0 = 0, time.perf_counter()
    for _ in range(args.steps):
        _, _, c = orc.evaluate(args.stage, poses[:n], threads=threads)
        rays += int(c[1])
    dt = time.perf_counter() - t0
    value = rays / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_name(args), "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"first {n} of {len(poses)} viewpoints x all {len(pts)} points per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--views", type=int, default=25_000, help="candidate viewpoints per GPU per step")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--resolution", type=float, default=0.1)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--prune-points", type=int, default=40_000)
    ap.add_argument("--prune-views", type=int, default=4_000)
    ap.add_argument("--no-prune", action="store_true")
    ap.add_argument("--full-views", type=int, default=200_000, help="initial candidate viewpoints of the pruning_full leg (configs[4]: 200 000 in total)")
    ap.add_argument("--no-full", action="store_true")
    ap.add_argument("--no-golden", action="store_true")
    ap.add_argument("--no-stages", action="store_true", help="skip the E2 / E3 / E4 kernel timings")
    ap.add_argument("--no-cov", action="store_true")
    ap.add_argument("--no-plan", action="store_true")
    ap.add_argument("--los-points", type=int, default=2048)
    ap.add_argument("--stage", type=int, default=1, choices=[1, 3, 4],
                    help="evaluator pass to time: 1 = getPose pass 1 (headline), 3 = gravitation pass, 4 = updateViewpointInfo (offset walk + BiRC)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # one rank per GPU on one box: keep each rank, and with it its page-locked staging buffers, on its GPU's NUMA node
        from fc_planner_b200.distributed import bind_to_gpu_numa_node
        global _THREADS_PER_RANK
        _THREADS_PER_RANK = bind_to_gpu_numa_node(local_rank, world)
        print(f"[bench] rank {rank}: {'bound to its GPU node, ' + str(len(os.sched_getaffinity(0))) + ' cores, ' + str(_THREADS_PER_RANK) + ' host threads' if _THREADS_PER_RANK else 'NUMA topology not readable, not bound'}",
              file=sys.stderr, flush=True)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from fc_planner_b200.manager import ViewpointManager
    pts, nrm, prm, poses = build_scene(args, rank)
    prm.device = local_rank
    if world > 1:
        prm.host_threads = threads_per_rank(world)
    vm = ViewpointManager(prm)
    vm.setMapPointCloud(pts)
    vm.setModel(pts)
    words = vm.words()
    # a non-default torch stream: the kernels are launched on it and torch's events are recorded on it
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: everything resident ------------------------------------------------
    vm.stage_poses(poses)
    for _ in range(args.warmup):
        vm.evaluate_resident(args.stage, stream=stream, sync=True)
    c0 = vm.counters()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = int(vm.counters()[5])
    ev0.record()
    for _ in range(args.steps):
        vm.evaluate_resident(args.stage, stream=stream, sync=False)
    ev1.record()
    launches_timed = int(vm.counters()[5]) - launches0      # the library's own count: k_ray_ends + k_eval + k_popc_rows per step
    barrier()
    clocks = sampler.stop()
    dev_ms = ev0.elapsed_time(ev1)
    # one more, synchronising, launch for the kernel-alone event time and the counters of one step
    rays_step = vm.evaluate_resident(args.stage, stream=stream, sync=True)
    kernel_ms = vm.last_kernel_ms()
    c1 = vm.counters()
    lookups_step = int(c1[2] - c0[2])        # only the synchronising launch accumulates counters
    pairs_step = int(c1[0] - c0[0])
    # the other three evaluators on the same resident batch, kernel alone (E2 re-tests the E1 sets): what updateViewpoints spends
    # its kernel time in besides E1
    other_stages = {}
    if args.stage == 1 and not args.no_stages:
        for st in (2, 3, 4):
            if st == 2:
                vm.evaluate_resident(1, stream=stream, sync=True)
            vm.evaluate_resident(st, stream=stream, sync=True)            # warm
            ca = vm.counters()
            r = vm.evaluate_resident(st, stream=stream, sync=True)
            cb = vm.counters() - ca
            other_stages[f"E{st}"] = {"kernel_ms": vm.last_kernel_ms(), "rays": int(r), "voxel_lookups": int(cb[2]), "pair_tests": int(cb[0])}
        vm.evaluate_resident(1, stream=stream, sync=True)

    # ---------------- e2e: host buffers through the C ABI ---------------------------------------
    # poses come from, and results go to, page-locked host memory (fcvm_host_alloc); every step does the host pose ->
    # FOV-normal conversion, H2D, the kernels and the D2H of the result.  The headline call returns what the manager
    # consumes from an evaluator pass - per-viewpoint counts + the visible sets as ascending index lists (the reference's
    # `visible_structural_voxels`, viewpoint_manager.cpp:358-425); `e2e_rows` is the same pass returning the raw bitset
    # rows (P / 8 bytes per viewpoint).
    from fc_planner_b200.manager import pinned_empty
    h_poses = pinned_empty(poses.shape, np.float64)
    h_poses[:] = poses
    vis_cap = int(vm.evaluate(args.stage, h_poses, want_rows=False)[0].sum())
    out_csr = (pinned_empty((len(poses),), np.int32), pinned_empty((len(poses) + 1,), np.int64), pinned_empty((vis_cap + 1024,), np.uint32))
    barrier()
    for _ in range(min(args.warmup, 3)):
        vm.evaluate_csr(args.stage, h_poses, out=out_csr)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cnt, offs, idx = vm.evaluate_csr(args.stage, h_poses, out=out_csr)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_visible = int(cnt.sum())
    assert e2e_visible == len(idx) == int(offs[-1])
    del out_csr
    out = (pinned_empty((len(poses),), np.int32), pinned_empty((len(poses), words), np.uint32))
    vm.evaluate(args.stage, h_poses, out=out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cnt_r, rows = vm.evaluate(args.stage, h_poses, out=out)
    torch.cuda.synchronize()
    e2e_rows_s = time.perf_counter() - t0
    assert int(cnt_r.sum()) == e2e_visible

    t = torch.tensor([dev_ms, e2e_s * 1e3, float(rays_step)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms_max, e2e_ms_max, rays_total = float(mx[0]), float(mx[1]), float(sm[2])
    else:
        dev_ms_max, e2e_ms_max, rays_total = dev_ms, e2e_s * 1e3, float(rays_step)

    # second metric (collective when world > 1: every rank takes part)
    prune = None if args.no_prune else pruning_leg(args, local_rank, with_cpu=not args.no_cpu, world=world, rank=rank)
    prune_full = None if args.no_full else pruning_full_leg(args, local_rank, world=world, rank=rank)
    golden = None if args.no_golden else golden_sharded_leg(local_rank, world=world, rank=rank)

    if rank == 0:
        P, V = len(pts), len(poses)
        value = rays_total * args.steps / (dev_ms_max * 1e-3)
        e2e_value = rays_total * args.steps / (e2e_ms_max * 1e-3)
        # algorithmic bytes of one k_eval launch (SURVEY.md 8d / DESIGN.md):
        #   points streamed once per 32-viewpoint tile + one 32 B sector per voxel lookup + pose records + bitset rows
        b_alg = 16.0 * P * np.ceil(V / 32.0) + 32.0 * lookups_step + 136.0 * V + (P / 8.0) * V
        peak, peak_src = 6554.6, "fallback"
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak, peak_src = float(mp["hbm_gbs"]), "measured"
        except Exception:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = b_alg / (kernel_ms * 1e-3) / 1e9
        # Per-launch figures of one `ncu --set full` capture of this kernel (profiles/traffic.json, written by
        # profiles/summarize_ncu.py), stamped with the hash of the kernel sources they were taken on: stale when that differs.
        tr = {}
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        except Exception:
            pass
        build_id = None
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("fcvm_build", os.path.join(ROOT, "fc-planner_b200", "build.py"))
            bm = importlib.util.module_from_spec(spec); spec.loader.exec_module(bm)
            build_id = bm.kernel_build_id()
        except Exception:
            pass
        traffic = tr.get("dram_bytes_per_launch")
        stale = bool(tr) and (tr.get("kernel_build") != build_id or args.stage != 1 or V != 25000 or P != 1000000)
        l2_peak, fp64_peak = None, None
        try:
            pk = json.load(open(os.path.join(ROOT, "profiles", "r2", "measured_peaks_l2_fp64.json")))
            l2_peak = max(v for k, v in pk["read_gbs"].items() if int(k.split("_")[0]) <= 96)
            fp64_peak = pk["fp64"]["dadd_tops"]
        except Exception:
            pass
        sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_stale": stale,
                "kernel": f"k_eval<E{args.stage}>", "kernel_ms": kernel_ms, "algorithmic_bytes": b_alg, "peak_source": peak_src,
                "kernel_build": build_id, "traffic_kernel_build": tr.get("kernel_build"),
                "note": "achieved = ALGORITHMIC bytes (SURVEY.md 8d: 32 B per occCheck the reference makes + streamed points + records + bitset rows) per second; "
                        "the kernel is built not to move them (bit-packed grid in L2, brick words in registers, certified half-rays counted instead of "
                        "marched), so frac > 1 is possible and says nothing about DRAM utilisation - see dram / l2 / issue below"}
        if tr:
            l2b, ins = tr.get("l2_bytes_per_launch"), tr.get("warp_instructions")
            roof["dram"] = {"bytes_per_launch": traffic, "achieved_gbs": (traffic / (kernel_ms * 1e-3) / 1e9) if traffic else None, "peak_gbs": peak,
                            "frac": (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if traffic else None}
            roof["l2"] = {"bytes_per_launch": l2b, "achieved_gbs": (l2b / (kernel_ms * 1e-3) / 1e9) if l2b else None, "peak_gbs": l2_peak,
                          "frac": (l2b / (kernel_ms * 1e-3) / 1e9 / l2_peak) if (l2b and l2_peak) else None, "hit_rate_pct": tr.get("l2_hit_rate_pct"),
                          "peak_source": "profiles/scripts/peaks.cu: streaming 128-bit __ldcg reads of an L2-resident buffer, measured on this pool"}
            roof["l1_sectors"] = tr.get("l1_sectors_global_ld_per_launch")
            # the bound that actually binds: instruction issue.  warp instructions of the capture / live kernel time against 4 issue slots x SMs x clock
            roof["issue"] = {"bound": "issue", "warp_instructions_per_launch": ins, "achieved_ginst_s": (ins / (kernel_ms * 1e-3) / 1e9) if ins else None,
                             "peak_ginst_s": 148 * 4 * sm_clock / 1e9, "frac": (ins / (kernel_ms * 1e-3) / (148 * 4 * sm_clock)) if ins else None,
                             "issue_active_pct_ncu": tr.get("issue_active_pct"), "warps_per_scheduler": tr.get("warps_per_scheduler"),
                             "active_lanes_per_instruction": tr.get("active_lanes_per_instruction"), "top_stalls_pct": tr.get("top_stalls_pct"),
                             "fp64_pipe_pct": tr.get("fp64_pipe_pct"), "fp64_dadd_peak_tops": fp64_peak}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "P": P, "V_per_gpu": V, "grid_voxels_mb_packed": None,
                       "l2": f"working set larger than L2: {V * words * 4 / 1e9:.2f} GB of bitset rows rewritten every step",
                       "parallelism": f"viewpoint shards x{world}, grid and points replicated, no data-path collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(V * 128), "d2h_bytes_per_step": int(V * 4 + (V + 4) * 8 + e2e_visible * 4),
                    "ms_per_step": e2e_ms_max / args.steps, "visible_pairs_per_step_per_gpu": e2e_visible,
                    "call": f"fcvm_evaluate_csr(E{args.stage}, pinned host poses) -> pinned host counts + offsets + visible index lists; the lists of a "
                            "finished sub-batch are compacted and copied out while the next is evaluated"},
            "e2e_rows": {"value": rays_step / (e2e_rows_s / args.steps), "unit": UNIT, "ms_per_step": 1e3 * e2e_rows_s / args.steps,
                         "d2h_bytes_per_step": int(V * 4 + V * words * 4), "note": "rank 0's own rate",
                         "call": "fcvm_evaluate(...) -> counts + raw bitset rows"},
            "gpu_launches": launches_timed,
            "rays_per_step_per_gpu": rays_step, "pair_tests_per_step_per_gpu": pairs_step, "voxel_lookups_per_step_per_gpu": lookups_step,
            "viewpoint_evals_per_s": V * world * args.steps / (dev_ms_max * 1e-3),
            "other_evaluators_rank0": other_stages,
            "roofline": roof,
        }
        if not args.no_cpu and world == 1:      # the CPU baseline is timed on rank 0 at N = 1 only
            threads = host_cores()
            kind = cpu_kind()
            orc = cpu_evaluator(prm, pts, kind)
            rate, n, dt, c, ocnt = cpu_oracle_rate(pts, nrm, prm, poses, args.cpu_seconds, threads, args.stage, kind, orc)
            # the kernel's per-viewpoint visible counts and its ray / look-up counters must equal the CPU implementation's
            g0 = vm.counters()
            vm2_cnt, _ = vm.evaluate(args.stage, poses[:n], want_rows=False)
            g1 = vm.counters() - g0
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": f"first {n} of {V} viewpoints x all {P} points, {dt:.1f} s, {threads} threads",
                                    "counts_match_gpu": bool(np.array_equal(vm2_cnt, ocnt)),
                                    "rays_and_lookups_match_gpu": bool(int(g1[1]) == int(c[1]) and int(g1[2]) == int(c[2]))}
            rate1, n1, dt1, _, _ = cpu_oracle_rate(pts, nrm, prm, poses[:max(8, n // max(threads, 1))], args.cpu_seconds / 3, 1, args.stage, kind, orc)
            line["cpu_baseline"]["single_thread_value"] = rate1
            del orc
            if kind == "reference":      # the oracle's restatement beside the reference's own code, same sample size
                ratep, npp, dtp, _, _ = cpu_oracle_rate(pts, nrm, prm, poses[:n], args.cpu_seconds / 2, threads, args.stage, "port")
                line["cpu_baseline"]["port_value"] = ratep
        line["pruning"] = prune
        line["pruning_full"] = prune_full
        line["golden_sharded"] = golden
        if not args.no_cov:
            line["coverage_evaluation"] = coverage_leg(args, local_rank, with_cpu=not args.no_cpu and world == 1)
        if not args.no_plan:
            line["planner_grid"] = planner_grid_leg(args, local_rank, with_cpu=not args.no_cpu and world == 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
