"""Host-side data-parallel protocol of the A-FAN step (one process per GPU; no device code here).

PGD is per-sample and mix_feature per-pixel, so the batch shards across ranks with NO collective inside
the ascent.  Exactly two exchanges exist:
  1. dual-BN statistics: every rank contributes its LOCAL per-(group, channel) sums
     {sum x, sum x^2} (forward) / {sum dy, sum dy*xhat} (backward) as ONE float64 [G, C, 2] message that
     carries the clean and the adversarial statistics together; the all-reduced sums with the GLOBAL count
     n*hw*world give the statistics of the global batch (what a single process at batch n*world computes).
  2. the gradient arena: one SUM all-reduce per iteration; the 1/world mean is folded into the SGD kernel.
"""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_global: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, equal shards (the global batch must divide evenly: the loss is a mean of shard means)."""
    if n_global % world:
        raise ValueError(f"global batch {n_global} is not divisible by world size {world}")
    per = n_global // world
    return rank * per, (rank + 1) * per


def global_count(n_local: int, hw: int, world: int) -> float:
    return float(n_local) * float(hw) * float(world)


def allreduce_sums_(sums: torch.Tensor, process_group=None) -> torch.Tensor:
    """SUM all-reduce of the [G, C, 2] float64 statistics message (NCCL on GPUs, gloo in CPU tests)."""
    if sums.dtype != torch.float64:
        raise TypeError("BN statistics travel as float64 sums")
    if process_group is not None and dist.get_world_size(process_group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=process_group)
    return sums


def stats_from_sums(sums: torch.Tensor, count: float, eps: float):
    """(mean, biased var, unbiased var, invstd) per (group, channel) from all-reduced sums -- the maths of the
    device finaliser (csrc/afan_bn.cu: fwd_finalize_channel), restated on the host for tests and docs."""
    mean = sums[..., 0] / count
    var = (sums[..., 1] / count - mean * mean).clamp_min(0.0)
    unbiased = var * (count / (count - 1.0)) if count > 1 else var
    return mean, var, unbiased, torch.rsqrt(var + eps)


def allreduce_grad_arena_(flat_grad: torch.Tensor, process_group=None) -> float:
    """SUM all-reduce of the flat gradient arena; returns the scale (1/world) the SGD kernel applies."""
    world = dist.get_world_size(process_group) if process_group is not None else 1
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=process_group)
    return 1.0 / world
