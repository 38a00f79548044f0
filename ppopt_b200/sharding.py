"""Multi-GPU partition of one enumeration level.

Within a level every candidate is independent (the reference itself maps them over a process pool,
/root/reference/src/ppopt/mp_solvers/mpqp_parrallel_combinatorial.py:116); the only exchange is the per-candidate
status byte that next-level generation needs.  The candidate array (lexicographic order) is cut into world x CHUNKS
contiguous chunks dealt to the ranks in snake order (0..G-1, G-1..0, ...) - the expensive candidates cluster in index
ranges (sets containing the same leading rows) and the cost drifts along the index, so plain contiguous slices left the
slowest rank 15 % behind (measured at 4 GPUs) and a plain round-robin still favoured the low ranks (measured at 8).  Every rank
evaluates its chunks in place in a full-length status vector that is zero elsewhere; one all-reduce(SUM) of the byte
vector then gives every rank all statuses, after which each rank regenerates the identical next level (K6 is
deterministic and replicated).  No numerical data is ever reduced.
"""
from typing import List, Tuple

import torch

CHUNKS_PER_RANK = 16
MIN_CHUNK = 65536


def chunks(n: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """[lo, hi) ranges of the candidates rank evaluates; the ranges of all ranks tile [0, n) exactly"""
    if world <= 1:
        return [(0, n)] if n > 0 else []
    per_rank = max(1, min(CHUNKS_PER_RANK, n // (world * MIN_CHUNK)))
    total = world * per_rank
    size = (n + total - 1) // total if n > 0 else 0
    out = []
    for rnd in range(per_rank):
        j = rnd * world + (rank if rnd % 2 == 0 else world - 1 - rank)
        lo, hi = min(n, j * size), min(n, (j + 1) * size)
        if hi > lo:
            out.append((lo, hi))
    return out


def gather_status(status: torch.Tensor, dist) -> torch.Tensor:
    """status holds this rank's chunks and zeros elsewhere; afterwards every rank holds every byte"""
    dist.all_reduce(status, op=dist.ReduceOp.SUM)
    return status


def owned(indices: torch.Tensor, n: int, rank: int, world: int) -> torch.Tensor:
    """subset of (ascending) candidate indices that fall in rank's chunks"""
    keep = torch.zeros_like(indices, dtype=torch.bool)
    for lo, hi in chunks(n, rank, world):
        keep |= (indices >= lo) & (indices < hi)
    return indices[keep].contiguous()
