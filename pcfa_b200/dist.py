"""Multi-GPU plumbing: one process per GPU (torchrun), NCCL for the only exchange step the path has.

  * disjoint / joint perturbations: image pairs are independent optimisation problems
    (attack_PCFA.py:668-670) → `shard_indices`, no data-path collective;
  * universal perturbation: every closure evaluation all-reduces ONE fused buffer
    [grad_delta1 | grad_delta2 | loss] (sum, then / world) so that all ranks hold identical gradients and
    losses and therefore take identical L-BFGS decisions → `pack_reduce_unpack`.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_indices(n: int, rank: int, world: int):
    """Pair indices owned by `rank` (round robin)."""
    return list(range(rank, n, world))


def pack_reduce_unpack(flat: torch.Tensor, loss: torch.Tensor, g1: torch.Tensor, g2: torch.Tensor | None = None):
    """All-reduce (mean) of the gradients and the loss through one preallocated flat buffer.
    Gradients are overwritten in place; the reduced loss is returned as a view of `flat`."""
    n1 = g1.numel()
    in_place = g1.data_ptr() == flat.data_ptr()          # gradients already are slices of `flat` (DeviceLBFGS layout)
    if not in_place:
        flat[:n1].copy_(g1.reshape(-1))
        if g2 is not None:
            flat[n1:n1 + g2.numel()].copy_(g2.reshape(-1))
    flat[-1:].copy_(loss.reshape(1))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat)
        flat.div_(dist.get_world_size())
    if not in_place:
        g1.reshape(-1).copy_(flat[:n1])
        if g2 is not None:
            g2.reshape(-1).copy_(flat[n1:n1 + g2.numel()])
    return flat[-1]
