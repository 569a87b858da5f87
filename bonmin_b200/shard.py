"""Multi-GPU plumbing: children are independent, so ranks own contiguous blocks of the child index
space and the only exchange is (1) an all-gather of per-child {objective, status} so every rank can replay the
branching decision and (2) an all-reduce(min) of the incumbent bound.  Backend-agnostic (nccl on GPUs, gloo in tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous block [lo, hi) of rank; both directions of a strong-branching candidate (children 2i, 2i+1) stay on
    one rank because block boundaries are rounded to even indices."""
    per = (total + world - 1) // world
    per += per % 2
    lo = min(rank * per, total)
    return lo, min(lo + per, total)


def gather_children(obj: torch.Tensor, status: torch.Tensor, total: int, group=None):
    """all-gather of the per-child results of every rank's shard -> (obj[total], status[total]) on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return obj[:total], status[:total]
    per = shard_range(total, 0, world)[1]
    pad_o = torch.full((per,), float("inf"), dtype=obj.dtype, device=obj.device)
    pad_s = torch.zeros((per,), dtype=status.dtype, device=status.device)
    pad_o[: obj.numel()] = obj
    pad_s[: status.numel()] = status
    go = [torch.empty_like(pad_o) for _ in range(world)]
    gs = [torch.empty_like(pad_s) for _ in range(world)]
    dist.all_gather(go, pad_o, group=group)
    dist.all_gather(gs, pad_s, group=group)
    return torch.cat(go)[:total], torch.cat(gs)[:total]


def allreduce_incumbent(value: float, device, group=None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return float(t.item())
