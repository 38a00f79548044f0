"""Multi-GPU partition of one enumeration level.

Within a level every candidate is independent (the reference itself maps them over a process pool,
/root/reference/src/ppopt/mp_solvers/mpqp_parrallel_combinatorial.py:116); the only exchange is the per-candidate
status byte that next-level generation needs.  Rank g owns the contiguous slice g of the (lexicographically ordered)
candidate array; one all-gather of status bytes per level makes every rank hold all statuses, after which each rank
regenerates the identical next level (K6 is deterministic and replicated).  No numerical data is ever reduced.
"""
from typing import Tuple

import torch


def slice_bounds(n: int, rank: int, world: int) -> Tuple[int, int, int]:
    """(lo, hi, per): rank's slice [lo, hi) of n candidates cut into `world` contiguous slices of `per` (last ragged)."""
    per = (n + world - 1) // world if world > 0 else n
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi, per


def gather_status(status: torch.Tensor, n: int, dist, rank: int, world: int) -> torch.Tensor:
    """All ranks end with the full status vector: rank r contributes status[lo_r:hi_r] (padded to `per`)."""
    lo, hi, per = slice_bounds(n, rank, world)
    padded = torch.zeros((per,), dtype=torch.uint8, device=status.device)
    padded[:hi - lo] = status[lo:hi]
    parts = [torch.empty((per,), dtype=torch.uint8, device=status.device) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat(parts)[:n].contiguous()


def owned(indices: torch.Tensor, n: int, rank: int, world: int) -> torch.Tensor:
    """subset of (ascending) candidate indices that fall in rank's slice"""
    lo, hi, _ = slice_bounds(n, rank, world)
    return indices[(indices >= lo) & (indices < hi)].contiguous()
