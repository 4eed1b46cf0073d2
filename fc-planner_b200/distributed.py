"""Viewpoint sharding across the GPUs of one box: one process per GPU, the occupancy
grid and the model cloud replicated on every GPU, each rank evaluating a contiguous block of every
viewpoint batch, and ONE exchange per evaluator stage — an all-gather of the coverage bitset rows —
after which every rank holds all rows and runs the (cheap, order-sensitive) host logic and the
ownership sweep redundantly, so the result is bit-identical to the single-GPU run.

`torch.distributed` is plumbing only: the C library calls back into `Exchange.__call__` with the
device buffer that holds `world` equal slices (this rank's slice already filled) and the CUDA
stream its kernels run on; the all-gather is enqueued behind them on that stream (NCCL over
NVLink / NVSwitch).  With the `gloo` backend (CPU tests) the same callback gathers a host buffer.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist


class _DevBuf:
    """A raw device pointer as a `__cuda_array_interface__` object (uint8, 1-D)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3, "strides": None}


def shard_bounds(n, rank, world):
    """Contiguous shard [lo, hi) of n items: equal blocks of ceil(n / world) — the layout an
    all-gather of equal slices needs (include/fcvm.h: fcvm_shard_range)."""
    per = (n + world - 1) // world
    return min(n, per * rank), min(n, per * (rank + 1))


class Exchange:
    """All-gather callback for `fcvm_set_shard` (fcvm_allgather_fn in include/fcvm.h)."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.on_device = dist.get_backend(group) == "nccl"
        self.calls = 0
        self.bytes = 0

    def gather(self, buf_ptr, bytes_per_rank, stream):
        total = bytes_per_rank * self.world
        if self.on_device:
            whole = torch.as_tensor(_DevBuf(buf_ptr, total), device=torch.device("cuda", torch.cuda.current_device()))
            mine = whole[self.rank * bytes_per_rank:(self.rank + 1) * bytes_per_rank]
            ext = torch.cuda.ExternalStream(stream) if stream else torch.cuda.current_stream()
            with torch.cuda.stream(ext):
                dist.all_gather_into_tensor(whole, mine, group=self.group)   # in place: slice r of the output is rank r's input
        else:
            host = np.ctypeslib.as_array(ctypes.cast(buf_ptr, ctypes.POINTER(ctypes.c_uint8)), shape=(total,))
            whole = torch.from_numpy(host)
            mine = whole[self.rank * bytes_per_rank:(self.rank + 1) * bytes_per_rank].clone()
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine, group=self.group)
            for r, p in enumerate(parts):
                whole[r * bytes_per_rank:(r + 1) * bytes_per_rank] = p
        self.calls += 1
        self.bytes += total
        return 0

    def __call__(self, user, buf, bytes_per_rank, stream):
        try:
            return self.gather(buf, bytes_per_rank, stream)
        except Exception as e:   # never raise through the C ABI
            import traceback
            traceback.print_exc()
            return -1


def attach(manager, group=None):
    """Shard `manager` (a fc_planner_b200.manager.ViewpointManager) over the ranks of `group`."""
    ex = Exchange(group)
    manager.set_shard(ex.rank, ex.world, ex if ex.world > 1 else None)
    manager._exchange = ex
    return ex


def max_over_ranks(values, device=None, group=None):
    """Element-wise max of a small list of floats over all ranks (timings: max over ranks)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(x) for x in t]


def sum_over_ranks(values, device=None, group=None):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(x) for x in t]
