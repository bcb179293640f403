"""Bitwise-reproducible cross-rank sum of per-chain FP64 scalars.

north_star's only collective is "a single NCCL allreduce of per-chain
log-likelihood scalars over NVLink per step" (SURVEY.md §8e).  A ring or tree
all-reduce adds the ranks' contributions in an order that depends on the
algorithm NCCL picks, so two runs (or two world sizes) need not agree in the
last bit — and the accept/reject decision downstream compares that sum with a
uniform draw.  `ordered_allreduce_sum` fixes the order instead: all-gather the
`world x chains` partials (8 B x chains per rank — latency-bound either way),
then every rank adds rows 0..world-1 left to right.  Every rank gets the same
bits, and the result is independent of NCCL's algorithm choice.

This is reference-independent plumbing: it does not depend on what the partials
mean.  It works on any backend (`nccl` on the GPU box, `gloo` in CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def ordered_allreduce_sum(partial: torch.Tensor, group=None) -> torch.Tensor:
    """Sum `partial` (shape [chains], float64) over ranks in rank order 0,1,...,W-1."""
    if partial.dtype != torch.float64 or partial.dim() != 1:
        raise ValueError("partial must be a 1-D float64 tensor of per-chain scalars")
    world = dist.get_world_size(group)
    n = partial.numel()
    flat = torch.empty(world * n, dtype=partial.dtype, device=partial.device)  # gloo wants it flat
    dist.all_gather_into_tensor(flat, partial.contiguous(), group=group)
    gathered = flat.view(world, n)
    total = gathered[0].clone()
    for r in range(1, world):  # explicit left-to-right order; torch.sum's order is unspecified
        total += gathered[r]
    return total
