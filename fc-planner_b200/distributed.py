"""Viewpoint sharding across the GPUs of one box: one process per GPU, the occupancy grid and the model
cloud replicated on every GPU, each rank evaluating a contiguous block of every viewpoint batch.  Bitset rows stay on the
GPU that produced them; per stage the ranks exchange the per-viewpoint results of their blocks (all-gather, a few bytes
per viewpoint) and run the ownership rule as one all-reduce MAX over per-point keys (8 bytes per model point), after which
every rank applies the same winners to its replica of the cover state - bit-identical to the single-GPU run.

`torch.distributed` is plumbing only.  Default (`attach`): the C library drives NCCL itself over NVLink / NVSwitch
(`fcvm_set_shard_nccl`); torch merely broadcasts the 128-byte ncclUniqueId.  `attach(..., native=False)` routes the two
collectives through torch.distributed instead (`Exchange`, the `fcvm_comm_fn` callback of include/fcvm.h) - the form any
other communication backend would plug into.  With the `gloo` backend (CPU tests) the callback works on host buffers.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from ._lib import COMM_ALLGATHER, COMM_ALLREDUCE_MAX


class _DevBuf:
    """A raw device pointer as a `__cuda_array_interface__` object (1-D)."""

    def __init__(self, ptr, n, typestr="|u1"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


def shard_bounds(n, rank, world):
    """Contiguous shard [lo, hi) of n items: equal blocks of ceil(n / world) - the layout an
    all-gather of equal slices needs (include/fcvm.h: fcvm_shard_range)."""
    per = (n + world - 1) // world
    return min(n, per * rank), min(n, per * (rank + 1))


def ownership_key(contrib, pos):
    """The per-point key of the arg-max ownership exchange: larger contribute_voxel_num wins, the earlier position in
    processing order among equals (k_own_best in fcvm_kernels.cu); 0 = no seer.  Fits a signed 64-bit MAX."""
    return (int(contrib) << 32) | (~int(pos) & 0xFFFFFFFF)


def decode_ownership_key(key):
    return int(key) >> 32, (~int(key)) & 0xFFFFFFFF


class Exchange:
    """`fcvm_comm_fn` over torch.distributed: op 0 = in-place all-gather of `world` equal byte slices, op 1 = in-place
    all-reduce MAX over int64 values."""

    def __init__(self, group=None, host_buffers=None, staged=False):
        self.group = group
        # staged=True: the library's DEVICE pointers with a CPU backend (gloo): every collective is staged through host memory.
        # Slow by design - it exists so that the sharded path can be verified with several ranks on ONE GPU
        # (tests/test_multi_gpu.py::test_sharded_update_on_one_gpu), where NCCL refuses two ranks on the same device.
        self.staged = staged
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.on_device = dist.get_backend(group) == "nccl"
        # host_buffers=True: the pointers handed to the callback are host memory (the gloo CPU tests call it directly)
        self.host_buffers = (not self.on_device) if host_buffers is None else host_buffers
        self.calls = 0
        self.bytes = 0
        self.reduces = 0

    def _stream(self, stream):
        return torch.cuda.ExternalStream(stream) if stream else torch.cuda.current_stream()

    def gather(self, buf_ptr, bytes_per_rank, stream):
        total = bytes_per_rank * self.world
        if self.on_device:
            whole = torch.as_tensor(_DevBuf(buf_ptr, total), device=torch.device("cuda", torch.cuda.current_device()))
            mine = whole[self.rank * bytes_per_rank:(self.rank + 1) * bytes_per_rank]
            with torch.cuda.stream(self._stream(stream)):
                dist.all_gather_into_tensor(whole, mine, group=self.group)   # in place: slice r of the output is rank r's input
        elif self.staged:
            st = self._stream(stream)
            whole = torch.as_tensor(_DevBuf(buf_ptr, total), device=torch.device("cuda", torch.cuda.current_device()))
            with torch.cuda.stream(st):
                host = whole.cpu()                       # ordered after everything the library has queued on its stream
            st.synchronize()
            mine = host[self.rank * bytes_per_rank:(self.rank + 1) * bytes_per_rank].clone()
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine, group=self.group)
            with torch.cuda.stream(st):
                whole.copy_(torch.cat(parts))
            st.synchronize()
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

    def reduce_max(self, buf_ptr, count, stream):
        if self.on_device:
            t = torch.as_tensor(_DevBuf(buf_ptr, count, "<i8"), device=torch.device("cuda", torch.cuda.current_device()))
            with torch.cuda.stream(self._stream(stream)):
                dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        elif self.staged:
            st = self._stream(stream)
            t = torch.as_tensor(_DevBuf(buf_ptr, count, "<i8"), device=torch.device("cuda", torch.cuda.current_device()))
            with torch.cuda.stream(st):
                host = t.cpu()
            st.synchronize()
            dist.all_reduce(host, op=dist.ReduceOp.MAX, group=self.group)
            with torch.cuda.stream(st):
                t.copy_(host)
            st.synchronize()
        else:
            host = np.ctypeslib.as_array(ctypes.cast(buf_ptr, ctypes.POINTER(ctypes.c_int64)), shape=(count,))
            t = torch.from_numpy(host)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        self.reduces += 1
        return 0

    def __call__(self, user, op, buf, n, stream):
        try:
            if op == COMM_ALLGATHER:
                return self.gather(buf, n, stream)
            if op == COMM_ALLREDUCE_MAX:
                return self.reduce_max(buf, n, stream)
            return -1
        except Exception:   # never raise through the C ABI
            import traceback
            traceback.print_exc()
            return -1


def broadcast_unique_id(group=None):
    """ncclUniqueId made on rank 0 of `group` by the library, broadcast with whatever backend the group has."""
    from .manager import nccl_unique_id
    rank = dist.get_rank(group)
    payload = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(payload, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return payload[0]


def attach(manager, group=None, native=True, staged=False):
    """Shard `manager` (a fc_planner_b200.manager.ViewpointManager) over the ranks of `group`.
    native=True: the library's own NCCL communicator (one per handle).  native=False: collectives through torch.distributed;
    the library always passes DEVICE pointers, so that needs the nccl backend - or staged=True, which copies every collective
    through host memory and works with any backend (several ranks on one GPU: verification only)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        manager.set_shard(0, 1, None)
        return None
    if native:
        manager.set_shard_nccl(broadcast_unique_id(group), rank, world)
        return None
    ex = Exchange(group, host_buffers=False, staged=staged)
    if not ex.on_device and not staged:
        raise RuntimeError("fcvm hands device pointers to the collectives callback: attach(native=False) needs the nccl backend "
                           f"(got {dist.get_backend(group)}); use attach(native=True) or a CUDA-aware backend")
    manager.set_shard(ex.rank, ex.world, ex)
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



def bind_to_gpu_numa_node(device_index, local_world=1):
    """Pin this process to the cores of the NUMA node its GPU hangs off (sysfs: the PCI device's `numa_node`, the node's
    `cpulist`), before any page-locked buffer is allocated: page-locked memory lands on the node of the allocating thread, and
    a rank whose staging buffers sit on the other socket pays the inter-socket link for every D2H byte - with 8 ranks each
    copying 0.5 GB of index lists per evaluator pass that link, not PCIe, sets the pace.  Returns the number of host threads
    this rank should use (the node's cores shared between the `local_world` ranks whose GPUs sit on the same node), or None
    when the topology cannot be read (nothing is changed then)."""
    import os
    try:
        import torch
        def node_of(i):
            p = torch.cuda.get_device_properties(i)
            bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
            return int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        node = node_of(device_index)
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        sharing = sum(1 for i in range(min(local_world, torch.cuda.device_count())) if node_of(i) == node)
        return max(1, len(cpus) // max(1, sharing))
    except Exception:
        return None
