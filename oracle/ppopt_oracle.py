"""CPU ORACLE: a numpy/scipy restatement of the reference's combinatorial path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module; nothing under ppopt_b200/
does, and the product has no CPU fallback.

What it restates (PPOPT v1.6.12, all paths relative to /root/reference/src/ppopt/):
    solve / check_child_feasibility      mp_solvers/mpqp_combinatorial.py:10-92
    CombinationTester, children          mp_solvers/solver_utils.py:15-55,154-166
    check_feasibility                    mplp_program.py:411-444
    check_optimality                     mpqp_program.py:203-322, mplp_program.py:446-569
    optimal_control_law                  mpqp_program.py:146-198, mplp_program.py:372-395
    gen_cr_from_active_set(_1d)          utils/mpqp_utils.py:89-320
    is_full_dimensional, chebyshev_ball  utils/mpqp_utils.py:323-344, utils/chebyshev_ball.py:10-63
    is_full_rank, scaling, row filters   utils/constraint_utilities.py:14-34,125-134,222-236,469-476
    solve_lp seam                        solver.py:211-246, solver_interface/cvxopt_interface.py:153-208
The LP arithmetic itself lives in a third-party dependency that is NOT under /root/reference: cvxopt -> GLPK (or
gurobipy); neither is vendored or version-pinned (requirements.txt:1-8, setup.py:23-34) and neither is installed here.
The oracle therefore calls scipy 1.18 HiGHS (dual simplex with presolve, scipy defaults) - the same backend the golden
vectors in tests/golden/ were produced with, by running the UNMODIFIED reference under oracle/ref_shim.

PINNING: tests/test_oracle.py checks this module against every golden fixture (candidate lists, status bytes, region
sets and matrices) - i.e. against outputs of the reference itself run in the build container.  Region MATRICES are
pinned by no test of the reference's own suite; the golden vectors are the only pin for those.
"""
import multiprocessing
import os
from typing import Dict, List, Optional, Sequence

import numpy
from scipy.optimize import linprog


class Program:
    """Post-presolve program data (the attributes of a reference program object after its constructor ran)."""

    def __init__(self, A, b, c, H, A_t, b_t, F, Q=None, n_eq=0):
        f = lambda a: numpy.asarray(a, dtype=float)
        self.A, self.F = f(A), f(F)
        self.b = f(b).reshape(-1, 1)
        self.c = f(c).reshape(-1, 1)
        self.H = f(H)
        self.A_t = f(A_t).reshape(-1, self.F.shape[1])
        self.b_t = f(b_t).reshape(-1, 1)
        self.Q = None if Q is None else f(Q)
        self.n_eq = int(n_eq)
        self.n, self.t, self.m, self.q = self.A.shape[1], self.F.shape[1], self.A.shape[0], self.A_t.shape[0]
        self.is_qp = self.Q is not None

    @classmethod
    def from_npz(cls, path):
        g = numpy.load(path)
        Q = g['Q'] if str(g['kind']) == 'qp' else None
        return cls(g['A'], g['b'], g['c'], g['H'], g['A_t'], g['b_t'], g['F'], Q, int(g['n_eq']))

    @property
    def equality_indices(self):
        return list(range(self.n_eq))


# ------------------------------------------------------------------ the LP seam (solver.py:211-246)
def solve_lp(c, A, b, equality_rows: Sequence[int] = ()):
    """min c'x s.t. A x <= b with the listed rows as equalities; returns x or None (cvxopt_interface.py:153-208)."""
    if A is None or A.shape[0] == 0 or A.shape[1] == 0:
        return None
    nv = A.shape[1]
    cost = numpy.zeros(nv) if c is None else numpy.asarray(c, dtype=float).ravel()
    eq = list(equality_rows)
    if len(eq) == A.shape[0]:  # "fully constrained" shortcut, cvxopt_interface.py:80-104,196-197
        return numpy.linalg.solve(A, b).ravel()
    eqset = set(eq)
    ineq = [i for i in range(A.shape[0]) if i not in eqset]
    kw = {}
    if eq:
        kw.update(A_eq=A[eq], b_eq=b[eq].ravel())
    res = linprog(cost, A_ub=A[ineq], b_ub=b[ineq].ravel(), bounds=(None, None), method='highs', **kw)
    return res.x if res.status == 0 else None


# ------------------------------------------------------------------ constraint utilities
def is_full_rank(A, rows) -> bool:
    if len(rows) == 0:
        return True
    return int(numpy.linalg.matrix_rank(A[list(rows)])) == len(rows)


def nonzero_rows(M) -> List[int]:
    return [i for i in range(M.shape[0]) if not numpy.allclose(M[i], 0, atol=1e-8)]


def scale_rows(M, v):
    s = 1.0 / numpy.linalg.norm(M, axis=1, keepdims=True)
    return M * s, v * s


def drop_duplicate_rows(M, v):
    if M.size == 0 or v.size == 0:
        return M, v
    stacked = numpy.hstack((M, v.reshape(v.size, 1)))
    first = numpy.sort(numpy.unique(stacked, axis=0, return_index=True)[1])
    return M[first], v[first]


def chebyshev_radius(E, f) -> Optional[float]:
    """max r : E x + ||E_i|| r <= f, r >= 0 (chebyshev_ball.py:10-63); None when the LP is not solved to optimality."""
    nv = E.shape[1]
    norms = numpy.linalg.norm(E, axis=1, keepdims=True)
    cost = numpy.zeros(nv + 1)
    cost[-1] = -1.0
    A = numpy.vstack([numpy.hstack([E, norms]), cost.reshape(1, -1)])
    b = numpy.vstack([f.reshape(-1, 1), numpy.zeros((1, 1))])
    x = solve_lp(cost, A, b)
    return None if x is None else float(x[-1])


# ------------------------------------------------------------------ per-active-set tests
def check_feasibility(P: Program, aset: Sequence[int], check_rank=True) -> bool:
    if check_rank and not is_full_rank(P.A, aset):
        return False
    A = numpy.block([[P.A, -P.F], [numpy.zeros((P.q, P.n)), P.A_t]])
    b = numpy.vstack([P.b, P.b_t])
    return solve_lp(numpy.zeros(P.n + P.t), A, b, aset) is not None


def check_optimality(P: Program, aset: Sequence[int]) -> bool:
    """Truthiness of the reference's check_optimality: the 'max t' LP over (x, theta, lambda, s, t) is solved."""
    aset = list(aset)
    n, t, m, k = P.n, P.t, P.m, len(aset)
    if not P.is_qp and k != n:  # mplp_program.py:472-473
        return False
    ka = k - P.n_eq
    inact = [i for i in range(m) if i not in set(aset)]
    ni = m - k
    nv = n + t + m + 1
    ox, ot, ol, os_, oq = 0, n, n + t, n + t + k, n + t + m
    Qm = P.Q if P.is_qp else numpy.zeros((n, n))
    rows, rhs = [], []

    def add(block, r):
        if block.shape[0]:
            rows.append(block)
            rhs.append(numpy.asarray(r, dtype=float).reshape(-1, 1))
    # equalities: stationarity, active rows, inactive rows with slack
    B = numpy.zeros((n, nv)); B[:, ox:ot] = Qm; B[:, ot:ol] = P.H; B[:, ol:os_] = P.A[aset].T
    add(B, -P.c)
    B = numpy.zeros((k, nv)); B[:, ox:ot] = P.A[aset]; B[:, ot:ol] = -P.F[aset]
    add(B, P.b[aset])
    B = numpy.zeros((ni, nv)); B[:, ox:ot] = P.A[inact]; B[:, ot:ol] = -P.F[inact]; B[:, os_:oq] = numpy.eye(ni)
    add(B, P.b[inact])
    # inequalities
    lam_cols = numpy.arange(ol + P.n_eq, ol + k)
    B = numpy.zeros((ka, nv)); B[numpy.arange(ka), lam_cols] = -1.0; B[:, oq] = 1.0      # t <= lambda (activated only)
    add(B, numpy.zeros(ka))
    B = numpy.zeros((ni, nv)); B[:, os_:oq] = -numpy.eye(ni); B[:, oq] = 1.0             # t <= s
    add(B, numpy.zeros(ni))
    B = numpy.zeros((1, nv)); B[0, oq] = -1.0                                            # t >= 0
    add(B, numpy.zeros(1))
    B = numpy.zeros((ka, nv)); B[numpy.arange(ka), lam_cols] = -1.0                      # lambda >= 0
    add(B, numpy.zeros(ka))
    B = numpy.zeros((ni, nv)); B[:, os_:oq] = -numpy.eye(ni)                             # s >= 0
    add(B, numpy.zeros(ni))
    B = numpy.zeros((P.q, nv)); B[:, ot:ol] = P.A_t                                      # theta in Theta
    add(B, P.b_t)
    A = numpy.vstack(rows)
    b = numpy.vstack(rhs)
    cost = numpy.zeros(nv); cost[oq] = -1.0
    n_equalities = n + m if k > 0 else m  # the reference's lp_active_limit quirk (mpqp_program.py:299-304)
    return solve_lp(cost, A, b, range(n_equalities)) is not None


def optimal_control_law(P: Program, aset: Sequence[int]):
    aset = list(aset)
    if not P.is_qp:
        pinv = numpy.linalg.pinv(P.A[aset])
        return pinv @ P.F[aset], pinv @ P.b[aset], -pinv.T @ P.H, -pinv.T @ P.c
    k = len(aset)
    Aa = P.A[aset]
    K = numpy.block([[Aa, numpy.zeros((k, k))], [P.Q, Aa.T]])
    consts = numpy.linalg.solve(K, numpy.vstack([P.b[aset], -P.c]))
    mats = numpy.linalg.solve(K, numpy.vstack([P.F[aset], -P.H]))
    return mats[:P.n], consts[:P.n], mats[P.n:], consts[P.n:]


def gen_region(P: Program, aset: Sequence[int]) -> Optional[Dict]:
    """gen_cr_from_active_set / _1d: returns a dict with the CriticalRegion fields, or None."""
    aset = list(aset)
    ne = P.n_eq
    active = aset[ne:]
    inactive = [i for i in range(P.m) if i not in set(aset)]
    Ax, bx, Al, bl = optimal_control_law(P, aset)
    lamA, lamb = -Al[ne:], bl[ne:]
    inA = P.A[inactive] @ Ax - P.F[inactive]
    inb = P.b[inactive] - P.A[inactive] @ bx
    E = numpy.vstack([lamA, inA, P.A_t])
    f = numpy.vstack([lamb, inb, P.b_t])
    kept = nonzero_rows(E)
    E, f = scale_rows(E[kept], f[kept])
    nl, ni = lamA.shape[0], inA.shape[0]
    if P.t == 1:
        lo, hi = float('-inf'), float('inf')
        for i in range(E.shape[0]):
            if E[i] > 0:
                hi = min(hi, f[i][0] / E[i][0])
            else:
                lo = max(lo, f[i][0] / E[i][0])
        if not (lo + 1e-8 <= hi):
            return None
        good = [kept[i] for i in range(E.shape[0]) if lo <= f[i] / E[i] <= hi]
        Eo, fo = numpy.array([[1], [-1]]), numpy.array([[hi], [-lo]])
    else:
        r = chebyshev_radius(E, f)
        if r is None or not (r > 1e-8):
            return None
        good_pos = [i for i in range(E.shape[0]) if solve_lp(None, E, f, [i]) is not None]
        good = [kept[i] for i in good_pos]
        Eo, fo = drop_duplicate_rows(E[good_pos], f[good_pos])
    lam_kept = [i for i in good if i < nl]
    reg_kept = [i - nl for i in good if nl <= i < nl + ni]
    om_kept = [i - nl - ni for i in good if i >= nl + ni]
    return dict(A=Ax, b=bx, C=Al, d=bl, E=Eo, f=fo, active_set=aset, omega_set=om_kept,
                lambda_set=[active[i] for i in lam_kept], regular_set=[reg_kept, [inactive[i] for i in reg_kept]])


def region_radius(region: Dict) -> Optional[float]:
    return chebyshev_radius(numpy.asarray(region['E'], dtype=float), numpy.asarray(region['f'], dtype=float))


# ------------------------------------------------------------------ enumeration bookkeeping
class CombinationTester:
    def __init__(self):
        self.combos = set()

    def check(self, aset) -> bool:
        s = set(aset)
        if not s:
            return True
        return not any(s.issuperset(c) for c in self.combos)

    def add_combo(self, aset):
        self.combos.add(tuple(aset))


def children_of(aset, m, tester: Optional[CombinationTester]):
    start = 0 if len(aset) == 0 else aset[-1] + 1
    return [[*aset, i] for i in range(start, m) if tester is None or tester.check([*aset, i])]


def evaluate_candidate(P: Program, aset) -> int:
    """status byte: 1 full rank, 2 feasible, 4 check_optimality truthy, 8 region built (tests/golden convention)"""
    st = 1 if is_full_rank(P.A, aset) else 0
    if not check_feasibility(P, aset):
        return st
    st |= 2
    if check_optimality(P, aset):
        st |= 4
        if gen_region(P, aset) is not None:
            st |= 8
    return st


_POOL_PROGRAM = None


def _pool_init(P):
    global _POOL_PROGRAM
    _POOL_PROGRAM = P
    for k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[k] = '1'  # the reference pins BLAS threads to 1 (src/ppopt/__init__.py:1-7)


def _pool_eval(aset):
    P = _POOL_PROGRAM
    st = 1 if is_full_rank(P.A, aset) else 0
    region = None
    if check_feasibility(P, aset):
        st |= 2
        if check_optimality(P, aset):
            st |= 4
            region = gen_region(P, aset)
            if region is not None:
                st |= 8
    return st, region


def evaluate_many(P: Program, asets, cores: int = 1):
    """[(status, region-or-None)] for a list of candidates, on `cores` worker processes (fork), order preserved."""
    if cores <= 1 or len(asets) < 2 * cores:
        _pool_init(P)
        return [_pool_eval(a) for a in asets]
    ctx = multiprocessing.get_context('fork')
    with ctx.Pool(cores, initializer=_pool_init, initargs=(P,)) as pool:
        return pool.map(_pool_eval, asets, chunksize=max(1, len(asets) // (cores * 8)))


def solve(P: Program, max_levels: Optional[int] = None, cores: int = 1, record=None):
    """The serial reference's level loop (mpqp_combinatorial.py:10-72); `cores` > 1 only parallelises the per-level
    candidate evaluation (as the reference's parallel twin does) without changing the serial pruning semantics."""
    tester = CombinationTester()
    regions = []
    max_depth = max(P.n, P.t) - P.n_eq
    depth = max_depth if max_levels is None else min(max_depth, max_levels)
    to_check = children_of(P.equality_indices, P.m, tester)
    for i in range(depth):
        if not P.is_qp:
            to_check = [c for c in to_check if not (c[-1] >= len(c) + P.m - P.n)]
        if not to_check:
            break
        outs = evaluate_many(P, to_check, cores)
        feasible = []
        for c, (st, reg) in zip(to_check, outs):
            if st & 2:
                feasible.append(c)
                if reg is not None:
                    regions.append(reg)
            else:
                tester.add_combo(c)
        if record is not None:
            record.append((to_check, numpy.array([o[0] for o in outs], dtype=numpy.uint8)))
        nxt = []
        if i + 1 != max_depth and not (max_levels is not None and i + 1 == depth):
            for c in feasible:
                nxt.extend(children_of(c, P.m, tester))
        to_check = nxt
    if max_levels is None or max_levels >= max_depth:
        eq = P.equality_indices
        if check_feasibility(P, eq) and check_optimality(P, eq):
            reg = gen_region(P, eq)
            if reg is not None:
                r = region_radius(reg)
                if r is not None and r > 1e-8:
                    regions.append(reg)
    return regions
