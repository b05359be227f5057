"""Multi-GPU plumbing for the hot path: one process per GPU, mixtures sharded by batch index.

Every stage of the path is per-mixture, so ranks never exchange tensors.
The only collective is an all-reduce of a handful of scalars: [loss_sum, element_count] per train
step and [sdr, sir, sar, si_sdr sums, clip_count] per evaluation -- replacing nn.DataParallel's
scatter / gather of the whole `outputs` dict (main.py:661, 108-111).  Backend: NCCL over
NVLink / NVSwitch on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun's contract).
    Returns (rank, world_size, local_rank); a plain single-process run returns (0, 1, 0) without a group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def bind_to_gpu_numa(local_rank: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so pinned host buffers allocated afterwards
    are node-local (first touch) and H2D copies do not cross the socket interconnect.  Returns the node or None."""
    try:
        prop = torch.cuda.get_device_properties(local_rank)
        bdf = f"{getattr(prop, 'pci_domain_id', 0):04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of `total` mixtures owned by `rank` (what DataParallel's scatter does along dim 0);
    the first total % world ranks get one extra mixture."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(total), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_windows(num_windows: int, rank: int, world: int) -> Sequence[int]:
    """Round-robin window ownership for long-clip separation."""
    return list(range(rank, int(num_windows), world))


def _reduce(vec: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec


def global_mean_loss(local_loss_mean: torch.Tensor, local_count: int, group=None) -> torch.Tensor:
    """Batch-size-weighted mean of the per-rank mean losses = the loss DataParallel + err.mean() would
    report on the un-sharded batch when shards are equal (main.py:562)."""
    vec = torch.stack([local_loss_mean.detach().reshape(()).double() * float(local_count),
                       torch.tensor(float(local_count), dtype=torch.float64, device=local_loss_mean.device)])
    vec = _reduce(vec, group)
    return (vec[0] / vec[1]).float()


def global_metrics(local_sums: Dict[str, float], local_clips: int, device=None, group=None) -> Dict[str, float]:
    """All-reduce of the metric sums and the clip count; returns the global means (AverageMeter semantics)."""
    keys = sorted(local_sums)
    vec = torch.tensor([float(local_sums[k]) for k in keys] + [float(local_clips)], dtype=torch.float64, device=device)
    vec = _reduce(vec, group)
    n = max(float(vec[-1]), 1.0)
    return {k: float(vec[i]) / n for i, k in enumerate(keys)}


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Device-timed durations are reported as the max over ranks (bench contract)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])


def gather_windows(local: torch.Tensor, num_windows: int, rank: int, world: int, group=None) -> torch.Tensor:
    """Long-clip separation (config C4): every rank reconstructed the windows rank, rank + world, ... ;
    returns all `num_windows` windows in order on every rank.  The one place tensors cross GPUs (waveforms out)."""
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        return local
    per = -(-num_windows // world)
    pad = local.new_zeros((per,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = local.new_zeros((num_windows,) + tuple(local.shape[1:]))
    for r in range(world):
        idx = list(range(r, num_windows, world))
        if idx:
            out[idx] = bufs[r][:len(idx)]
    return out
