"""Data-parallel plumbing: rank/world discovery, parameter broadcast, bucketed gradient allreduce
over the flat gradient arena, window sharding.

Replaces ``DistributedDataParallel(model)`` + ``DistributedSampler`` as used at
/root/reference/src/cli/train.py:99-102,143-150,175,281: one process per GPU, NCCL over NVLink, the
per-step collective is an allreduce of the gradients, overlapped with the rest of backward.  Because
gradients already live in one contiguous fp32 arena in reverse-free layer order, a bucket is just a
slice ``grad[a:b]``: no flatten/unflatten copies, one ``all_reduce`` per bucket enqueued on a side
stream as soon as the last layer of the bucket has finished its weight-gradient GEMMs.  The 1/W mean
is folded into the fused optimizer kernel (``grad_scale``), so the collective is a plain SUM.

Everything here is device-agnostic (works on CPU tensors with the gloo backend — that is how the
N>1 path is unit-tested without GPUs).
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT)."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend)
    return rank, ws, local


def shard_indices(n: int, rank: int, world_size: int) -> range:
    """DistributedSampler(shuffle=False, drop_last=True): r::W of the first floor(N/W)*W (train.py:143)."""
    total = (n // world_size) * world_size
    return range(rank, total, world_size)


def contiguous_shard(n: int, rank: int, world_size: int) -> range:
    """Contiguous ranges for sampling / analysis (windows independent, no collective)."""
    per = (n + world_size - 1) // world_size
    return range(min(n, rank * per), min(n, (rank + 1) * per))


def make_buckets(boundaries: Sequence[int], total: int, bucket_elems: int) -> List[Tuple[int, int, int]]:
    """Group consecutive layer slices into buckets of >= bucket_elems elements.

    boundaries: ascending arena offsets where a new "layer group" starts (first must be 0).  Backward
    finishes groups from the LAST to the first, so buckets are built from the end.  Returns
    [(start, end, ready_group)] in firing order; a bucket is ready when backward has finished group
    ``ready_group`` (its lowest-numbered member)."""
    edges = list(boundaries) + [total]
    buckets: List[Tuple[int, int, int]] = []
    end = total
    g = len(boundaries) - 1
    while g >= 0:
        start_g = g
        while start_g > 0 and end - edges[start_g] < bucket_elems:
            start_g -= 1
        buckets.append((edges[start_g], end, start_g))
        end = edges[start_g]
        g = start_g - 1
    return buckets


class GradBucketer:
    """Fires one allreduce(SUM) per bucket of the flat gradient tensor, optionally on a side stream.

    Two schedules (``mode``, default from IBM_ALLREDUCE):
      * ``"overlap"`` — a bucket's allreduce goes to a side stream as soon as backward has finished its last layer group
        (what DistributedDataParallel's hooks do, train.py:175,281);
      * ``"tail"`` — nothing is enqueued during backward; ``finish()`` issues ONE allreduce of the whole arena on the compute
        stream.  The collective is exposed (the arena crosses NVSwitch once, a few hundred microseconds for 100 MB) but no
        NCCL CTA ever shares the machine with the persistent GEMMs, whose statically assigned CTA pairs lose whole rounds
        when a co-resident kernel holds one of their SMs; and a step holds exactly one collective on one stream, which is
        the shape a CUDA graph of the small-batch steps captures."""

    def __init__(self, flat_grad: torch.Tensor, buckets: Sequence[Tuple[int, int, int]], group=None, mode: Optional[str] = None):
        self.mode = mode or os.environ.get("IBM_ALLREDUCE", "overlap")
        if self.mode not in ("overlap", "tail"):
            raise ValueError(f"IBM_ALLREDUCE / mode must be 'overlap' or 'tail', got {self.mode!r}")
        self.flat = flat_grad
        self.buckets = list(buckets)
        self.group = group
        self.rank, self.world = world()
        self.cuda = flat_grad.is_cuda
        self.comm_stream = torch.cuda.Stream(device=flat_grad.device) if self.cuda else None
        self._fired = 0
        self.collectives = 0
        # tail mode only: called INSTEAD of the collective (the trainer's graph capture ends one CUDA graph and begins the
        # next here, and issues the allreduce itself between the two replays)
        self.finish_hook: Optional[Callable[[], None]] = None

    def begin_step(self) -> None:
        self._fired = 0

    def group_done(self, group_index: int) -> None:
        """Call when backward has produced all gradients of layer group ``group_index`` (and all later ones)."""
        if self.mode == "tail":
            return
        while self._fired < len(self.buckets) and self.buckets[self._fired][2] >= group_index:
            self._launch(self.buckets[self._fired])
            self._fired += 1

    def _launch(self, b: Tuple[int, int, int]) -> None:
        if self.world == 1:
            return
        view = self.flat[b[0]:b[1]]
        self.collectives += 1
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)

    def finish(self) -> None:
        """Flush remaining buckets and make the compute stream wait for the collectives."""
        if self.mode == "tail":
            if self.finish_hook is not None:
                self.finish_hook()
            else:
                self.allreduce_all()
            return
        self.group_done(-1)
        if self.cuda and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def allreduce_all(self) -> None:
        """One allreduce(SUM) of the whole arena on the current stream."""
        if self.world > 1:
            self.collectives += 1
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)

    def broadcast_(self, flat_param: torch.Tensor, src: int = 0) -> None:
        if self.world > 1:
            dist.broadcast(flat_param, src=src, group=self.group)


def mean_over_ranks(t: torch.Tensor, group=None) -> torch.Tensor:
    """Plain mean of a small per-rank tensor over all ranks (one allreduce of ~40 floats).  The reference keeps metrics per
    rank (one wandb run per process, train.py:103-118); this is the optional rank-0 aggregate of SURVEY §8e/§8f-4 — with the
    DistributedSampler's equal shard sizes the mean of per-rank means of batch means is the global mean of batch means."""
    rank, ws = world()
    if ws == 1:
        return t.clone()
    out = t.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out / ws


def bind_to_gpu_numa(local_rank: int) -> Optional[Sequence[int]]:
    """Pin this process to the CPU cores local to its GPU's PCIe root (``/sys/bus/pci/devices/<bdf>/local_cpulist``) BEFORE it
    allocates pinned staging memory: pinned pages are placed on the NUMA node of the allocating thread, and a rank whose host
    buffers sit on the far socket shares one inter-socket link with its neighbours (the T = 200 analysis stream measured
    55 GB/s of H2D alone at 1-2 GPUs but 29-35 GB/s per GPU at 4-8).  Returns the core list, or None when the topology cannot
    be read (containers without /sys access): nothing is changed then."""
    try:
        props = torch.cuda.get_device_properties(local_rank)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = []
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus += list(range(int(a), int(b) + 1))
            elif part:
                cpus.append(int(part))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
