"""Host-side multi-GPU plumbing: one process per GPU, reads sharded across ranks, the library replicated, and one
all-reduce(sum) of the per-(sample, taxon) read counters at the end (the GPU twin of Slacken's
groupBy(sampleId, taxon).count across executors, slacken/Classifier.scala:214-217).
Uses torch.distributed only: NCCL over NVLink on the GPU box, gloo in the CPU tests."""
from __future__ import annotations

from typing import Tuple


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of [0, n_items): the first (n_items % world) ranks get one item more."""
    if not (0 <= rank < world):
        raise ValueError("rank outside the world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_counts(counts, group=None):
    """Sum the counter tensor over all ranks in place (int64; CPU tensor with gloo, CUDA tensor with NCCL)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def max_over_ranks(value: float, device=None, group=None) -> float:
    """The slowest rank decides (device-side timings are reported as the max over ranks)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


class DeviceCounterView:
    """Zero-copy torch view of the device counter matrix of a ReportCounts (for the NCCL all-reduce)."""

    def __init__(self, device_ptr: int, n_samples: int, n_taxa: int):
        self.__cuda_array_interface__ = {"shape": (n_samples, n_taxa), "typestr": "<i8", "data": (device_ptr, False),
                                         "version": 2}

    def tensor(self, device):
        import torch
        return torch.as_tensor(self, device=device)


def bind_to_gpu_numa(device: int) -> list:
    """Pins the calling process to the CPUs that are local to `device` (NVML's CPU affinity of the GPU, intersected with
    the CPUs the process may use), so that the pinned host buffers it allocates afterwards sit on the GPU's NUMA node and
    its H2D/D2H copies do not cross the socket interconnect. Matters with one process per GPU on a two-socket box: the
    copies of 4-8 ranks otherwise share the inter-socket links. Returns the CPU list ([] = left unchanged)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (max(ncpu, 1024) + 63) // 64)
        local = {i for i in range(64 * len(mask)) if (mask[i // 64] >> (i % 64)) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(local & allowed)
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return []
