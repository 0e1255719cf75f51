"""Multi-GPU plumbing: one process per GPU, impulse responses sharded by index.

Every channel is independent, so the data path has no collective.  The only exchange is one
gather of the per-row metric tables to rank 0: NCCL over NVLink on the
GPU box, gloo in the CPU tests.  EDC curves and CSD slices stay on the GPU that made them.
"""

from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from torchrun's environment; no-op for a single process."""
    size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if size > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=size)
    return rank, size, local


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(device) -> Optional[List[int]]:
    """Pin this process to the CPUs that sit next to `device` (its PCIe root's NUMA node), so that the pinned staging
    buffers it allocates afterwards are local to the GPU's copy engine.  With one process per GPU on a two-socket box
    the host-to-device streams otherwise cross the socket interconnect.  Returns the CPU list, or None when the
    topology cannot be read (then nothing is changed)."""
    try:
        props = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except (AttributeError, OSError, ValueError, RuntimeError, AssertionError):   # no CUDA, no sysfs entry, old torch
        return None


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [begin, end) of item indices owned by `rank`: ceil(n / world) per rank."""
    per = (int(n_items) + world_size - 1) // world_size
    begin = min(int(n_items), rank * per)
    return begin, min(int(n_items), begin + per)


def gather_tables(tables: Dict[str, torch.Tensor], n_items: int, dst: int = 0) -> Optional[Dict[str, torch.Tensor]]:
    """Gather per-row tables (first dimension = this rank's rows, in shard order) to rank `dst`.

    Returns {name: tensor with n_items rows} on `dst`, None elsewhere.  Shards may have different
    sizes (the last one is short); rows are padded to the common shard size for the collective.
    """
    rank, size = world()
    if size == 1:
        return {k: v[:n_items] for k, v in tables.items()}
    per = (int(n_items) + size - 1) // size
    out: Dict[str, torch.Tensor] = {}
    for name in sorted(tables):
        t = tables[name]
        pad = torch.zeros((per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        if rank == dst:
            parts: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(size)]
            dist.gather(pad, parts, dst=dst)
            out[name] = torch.cat(parts, dim=0)[:n_items]
        else:
            dist.gather(pad, None, dst=dst)
    return out if rank == dst else None
