"""Program presolve with the reference's semantics (SURVEY.md section 8f row 1 - the step right before the hot path).

Restates, in this package's own code, what the reference's program constructor does to the constraint data
(/root/reference/src/ppopt/mplp_program.py:110-134,276-306 and utils/constraint_utilities.py:38-97,186-200,
286-466): equalities first, theta-only rows moved to A_t, L2 row scaling of [A|-F], implicit equalities, dependent
equalities, and finally redundancy removal - one feasibility LP per row with that row forced active.  The numpy steps run
on the host exactly once per program; the per-row LPs (find_redundant_constraints, m + q of them, the same LP family as
check_feasibility) are evaluated as ONE batch by the GPU engine's K2a/K2 kernels instead of m + q solver calls.
"""
from typing import List, Tuple

import numpy


def _rows_not_in(M, drop):
    drop = set(drop)
    return M[[i for i in range(M.shape[0]) if i not in drop]]


def equalities_first(A, b, F, eq: List[int]):
    if len(eq) == 0:
        return A, b, F, []
    A = numpy.vstack([A[eq], _rows_not_in(A, eq)])
    b = numpy.vstack([b[eq], _rows_not_in(b, eq)])
    F = numpy.vstack([F[eq], _rows_not_in(F, eq)])
    return A, b, F, list(range(len(eq)))


def _split_by_norm(M, eps):
    nrm = numpy.array([numpy.linalg.norm(r) for r in M]) if M.shape[0] else numpy.zeros(0)
    keep = [i for i in range(M.shape[0]) if nrm[i] >= eps]
    move = [i for i in range(M.shape[0]) if not nrm[i] >= eps]
    return keep, move


def move_parametric_rows(A, b, F, A_t, b_t, eps=1e-6):
    """rows with [A|-F] ~ 0 and rows with A ~ 0 go to the parametric set as  -F_i theta <= b_i
    (process_program_constraints, constraint_utilities.py:365-401)"""
    for block in (lambda: numpy.hstack([A, -F]), lambda: A):
        keep, move = _split_by_norm(block(), eps)
        if move:
            A_t = numpy.vstack([A_t, -F[move]])
            b_t = numpy.vstack([b_t, b[move]])
        A, b, F = A[keep], b[keep], F[keep]
    return A, b, F, A_t, b_t


def scale_rows(A, b, F):
    """||[A_i | -F_i]||_2 = 1 (mplp_program.py:276-283)"""
    norm = numpy.linalg.norm(numpy.block([A, -F]), axis=1, keepdims=True)
    return A / norm, b / norm, F / norm


def implicit_equality_pairs(M, v) -> List[Tuple[int, int]]:
    """pairs (i, j >= i) of rows of [M | v] that are numerically opposite: two of the reference's three tests must hold
    (constraint_utilities.py:38-97)"""
    blk = numpy.hstack([M, v.reshape(-1, 1)])
    blk = blk / numpy.linalg.norm(blk, axis=1, keepdims=True)
    blk = blk / numpy.linalg.norm(blk, axis=1, keepdims=True)
    pairs = []
    for i in range(blk.shape[0]):
        for j in range(i, blk.shape[0]):
            hits = int(abs(blk[i].T @ blk[j] + 1) <= 1e-8)
            hits += int(numpy.linalg.norm(blk[i] - blk[j], 2) <= 1e-12)
            hits += int(numpy.allclose(blk[i], -blk[j]))
            if hits >= 2:
                pairs.append((i, j))
    return pairs


def promote_implicit_equalities(A, b, F, eq: List[int]):
    """(constraint_utilities.py:404-466)"""
    pairs = implicit_equality_pairs(numpy.hstack([A, -F]), b)
    keep = sorted(set(p[0] for p in pairs))
    drop = [i for i in sorted(set(p[1] for p in pairs)) if i not in keep]
    active = [*eq, *keep]
    rest = [i for i in range(A.shape[0]) if i not in active and i not in drop]
    order = active + rest
    return A[order], b[order], F[order], list(range(len(active)))


def independent_rows(M) -> List[int]:
    """indices where the running rank increases; like the reference, the last row is never examined
    (constraint_utilities.py:320-331)"""
    n = M.shape[0]
    ranks = numpy.zeros(n)
    for i in range(n - 1):
        ranks[i] = numpy.linalg.matrix_rank(M[:i + 1])
    grew = numpy.diff(ranks, prepend=0) > 0
    return [i for i in range(n) if grew[i]]


def reduce_equalities(A, b, F, eq: List[int]):
    """(constraint_utilities.py:334-362)"""
    if len(eq) == 0:
        return A, b, F, []
    if int(numpy.linalg.matrix_rank(A[eq])) == len(eq):
        return A, b, F, eq
    sel = independent_rows(A[eq])
    A2 = numpy.vstack([A[sel], _rows_not_in(A, eq)])
    b2 = numpy.vstack([b[sel], _rows_not_in(b, eq)])
    F2 = numpy.vstack([F[sel], _rows_not_in(F, eq)])
    return A2, b2, F2, sel


def base_processing(A, b, F, A_t, b_t, eq: List[int]):
    """MPLP_Program.base_constraint_processing (mplp_program.py:110-134)"""
    f = lambda a: numpy.asarray(a).astype('float64')
    A, b, F, A_t, b_t = f(A), f(b).reshape(-1, 1), f(F), f(A_t), f(b_t).reshape(-1, 1)
    A, b, F, eq = equalities_first(A, b, F, list(eq))
    A, b, F, A_t, b_t = move_parametric_rows(A, b, F, A_t, b_t)
    A, b, F = scale_rows(A, b, F)
    A, b, F, eq = promote_implicit_equalities(A, b, F, eq)
    A, b, F, eq = reduce_equalities(A, b, F, eq)
    return A, b, F, A_t, b_t, eq


def nonredundant_rows_gpu(A, b, F, A_t, b_t, n_eq: int) -> List[int]:
    """find_redundant_constraints over P = {(x,theta): A x - F theta <= b, A_t theta <= b_t} (mplp_program.py:285-306,
    constraint_utilities.py:186-200): row i survives iff P with rows [eq..., i] as equalities is non-empty.
    The m + q LPs run as one level-1 batch of the engine: parametric rows are appended to the main body as
    0 x - (-A_t) theta <= b_t so that every row owns a bit of the candidate mask."""
    import torch
    from . import engine
    from ._lib import ST_FEAS, ST_RANK
    m, n = A.shape
    q, t = A_t.shape[0], F.shape[1]
    A2 = numpy.vstack([A, numpy.zeros((q, n))])
    F2 = numpy.vstack([F, -A_t])
    b2 = numpy.vstack([b, b_t])
    arrays = dict(A=numpy.ascontiguousarray(A2), b=numpy.ascontiguousarray(b2.ravel()), F=numpy.ascontiguousarray(F2),
                  A_t=numpy.zeros((0, t)), b_t=numpy.zeros(0), c=numpy.zeros(n), H=numpy.zeros((n, t)), Q=None, n_eq=n_eq,
                  is_qp=False)
    eng = engine.Engine(arrays)
    try:
        rows = list(range(n_eq, m + q))
        if not rows:
            return list(range(n_eq))
        masks = eng.masks_from_lists([[i] for i in rows])
        status = torch.full((len(rows),), ST_RANK, dtype=torch.uint8, device=eng.tdev)  # no rank screen in the reference here
        st = eng.level_eval(masks, 1, status, stages=2).cpu().numpy()
    finally:
        eng.close()
    return list(range(n_eq)) + [r for r, s in zip(rows, st) if s & ST_FEAS]


def remove_redundant(A, b, F, A_t, b_t, n_eq: int):
    """MPLP_Program.process_constraints (mplp_program.py:285-306)"""
    kept = nonredundant_rows_gpu(A, b, F, A_t, b_t, n_eq)
    m = A.shape[0]
    up = [i for i in kept if i < m]
    lo = [i - m for i in kept if i >= m]
    return A[up], b[up], F[up], A_t[lo], b_t[lo]
