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
import os
from typing import List, Tuple

import torch

CHUNKS_PER_RANK = int(os.environ.get('PPGPU_CHUNKS_PER_RANK', '16'))
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


def gather_regions(eng, mine: torch.Tensor, bufs, k_act: int, dist):
    """Region payloads of all ranks, in candidate order, on EVERY rank: counts first, then the raw K5 buffers
    (laws / rows / flags / info) and the owning candidate indices through NCCL all_gather on padded tensors - no Python
    objects cross the wire (the object all-gather of round 1 pickled ~48 MB per level)."""
    world = dist.get_world_size()
    dev = eng.tdev
    n_mine = int(mine.shape[0])
    counts = torch.zeros((world,), dtype=torch.int64, device=dev)
    counts[dist.get_rank()] = n_mine
    dist.all_reduce(counts)
    counts_h = counts.cpu().tolist()
    cap = max(counts_h)
    if cap == 0:
        return mine, None
    N = eng.n + eng.n_eq + k_act
    shapes = [(N, eng.t + 1), (eng.R0, eng.t + 1), (eng.R0,), (4,)]
    dtypes = [torch.float64, torch.float64, torch.int32, torch.float64]

    def pad(x, shape, dtype):
        out = torch.zeros((cap,) + shape, dtype=dtype, device=dev)
        if x is not None and x.shape[0]:
            out[:x.shape[0]] = x
        return out
    parts = [pad(None if bufs is None else b, sh, dt) for b, sh, dt in zip(bufs or [None] * 4, shapes, dtypes)]
    parts.append(pad(mine, (), torch.int64))
    gathered = []
    for x in parts:
        buf = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=dev)   # concatenated layout
        dist.all_gather_into_tensor(buf, x)
        buf = buf.view((world,) + tuple(x.shape))
        gathered.append(torch.cat([buf[r, :counts_h[r]] for r in range(world)]))
    idx = gathered[4]
    order = torch.argsort(idx)
    return idx[order].contiguous(), [g[order].contiguous() for g in gathered[:4]]


def gather_region_bits(status: torch.Tensor, mine: torch.Tensor, dist, already_global: bool = False, bufs=None):
    """K5 sets the region bit (8) only on the emitting rank; make every rank's status vector agree (the digest and the
    collected status arrays are then identical on all ranks)."""
    if already_global:
        if bufs is not None and mine.shape[0]:
            is_region = bufs[3][:, 0] == 1.0
            status[mine[is_region]] |= 8
        return status
    bits = torch.zeros_like(status)
    if mine.shape[0]:
        bits[mine] = status[mine] & 8
    dist.all_reduce(bits, op=dist.ReduceOp.MAX)
    return status | bits
