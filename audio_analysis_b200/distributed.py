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


class TableGatherer:
    """The path's one collective, made cheap: the per-row tables of a rank are packed into ONE pre-allocated byte
    buffer (a fixed-size record per row) and travel in ONE `dist.gather` to rank `dst`, whose receive buffer is
    allocated once.  Build it during warm-up (the first NCCL call on a communicator sets up the peer channels and
    costs hundreds of milliseconds), then call `gather` every step.

    `like`: {name: tensor} giving dtype and trailing shape of every table; `rows_per_rank`: the common (padded) shard
    size.  Tables are packed in sorted-name order.
    """

    def __init__(self, like: Dict[str, torch.Tensor], rows_per_rank: int, dst: int = 0):
        self.rank, self.size = world()
        self.dst = int(dst)
        self.per = int(rows_per_rank)
        self.names = sorted(like)
        self.meta = {}
        off = 0
        for name in self.names:
            t = like[name]
            tail = tuple(int(v) for v in t.shape[1:])
            width = t.element_size()
            for v in tail:
                width *= v
            off = (off + 7) // 8 * 8          # keep every field 8-byte aligned inside the record
            self.meta[name] = (off, width, t.dtype, tail)
            off += width
        self.record = (off + 15) // 16 * 16
        dev = like[self.names[0]].device
        self.send = torch.zeros((self.per, self.record), dtype=torch.uint8, device=dev)
        self.recv = None
        if self.rank == self.dst and self.size > 1:
            self.recv = torch.empty((self.size * self.per, self.record), dtype=torch.uint8, device=dev)
            self.parts = list(self.recv.view(self.size, self.per, self.record).unbind(0))

    def pack(self, tables: Dict[str, torch.Tensor]) -> torch.Tensor:
        for name in self.names:
            off, width, dtype, tail = self.meta[name]
            t = tables[name]
            rows = int(t.shape[0])
            if rows > self.per or t.dtype != dtype or tuple(int(v) for v in t.shape[1:]) != tail:
                raise ValueError(f"table {name!r} does not match the gatherer's layout")
            self.send[:rows, off:off + width].copy_(t.contiguous().view(rows, -1).view(torch.uint8))
        return self.send

    def unpack(self, buf: torch.Tensor, n_items: int) -> Dict[str, torch.Tensor]:
        out = {}
        for name in self.names:
            off, width, dtype, tail = self.meta[name]
            out[name] = buf[:n_items, off:off + width].contiguous().view(dtype).view((n_items,) + tail)
        return out

    def gather(self, tables: Dict[str, torch.Tensor], n_items: Optional[int] = None,
               unpack: bool = True) -> Optional[Dict[str, torch.Tensor]]:
        """One collective.  Returns the gathered tables on `dst` (None elsewhere, or when `unpack` is False: the
        packed records then sit in `self.recv`)."""
        send = self.pack(tables)
        if self.size == 1:
            n = int(tables[self.names[0]].shape[0]) if n_items is None else int(n_items)
            return self.unpack(send, n) if unpack else None
        if self.rank == self.dst:
            dist.gather(send, self.parts, dst=self.dst)
            if not unpack:
                return None
            n = self.size * self.per if n_items is None else int(n_items)
            return self.unpack(self.recv, n)
        dist.gather(send, None, dst=self.dst)
        return None


def gather_tables(tables: Dict[str, torch.Tensor], n_items: int, dst: int = 0) -> Optional[Dict[str, torch.Tensor]]:
    """Gather per-row tables (first dimension = this rank's rows, in shard order) to rank `dst`.

    Returns {name: tensor with n_items rows} on `dst`, None elsewhere.  Shards may have different sizes (the last
    one is short); rows are padded to the common shard size and packed into one record per row, so the exchange is
    a single collective (TableGatherer; build one and reuse it when gathering every step).
    """
    rank, size = world()
    if size == 1:
        return {k: v[:n_items] for k, v in tables.items()}
    per = (int(n_items) + size - 1) // size
    return TableGatherer(tables, per, dst).gather(tables, n_items)
