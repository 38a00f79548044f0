"""Golden vectors for the mpMIQP enumeration.

TEST INFRASTRUCTURE.  Build container only:  python oracle/gen_mpmiqp_golden.py

The UNMODIFIED reference, under the LP shim, does everything the enumeration algorithm does after the binary tree
(/root/reference/src/ppopt/mp_solvers/mpmiqp_enumeration.py:40-64): MPMIQP_Program.generate_substituted_problem for every
binary combination (mpmiqp_program.py:71-116, including the sub-problem constructor's presolve) and
solve_mpqp(sub_problem, mpqp_algorithm.combinatorial).  The tree itself (MITree / check_bin_feasibility) needs a MILP
solver and the reference only binds Gurobi (solver.py:278-281), which is not in this image; for a LEAF the MILP
feasibility question is the LP feasibility of the substituted constraint set, which the reference's own
sub_problem.check_feasibility answers - that is what selects the feasible combinations here.  The mixed-integer programs
are built with post_process=False (their constructor presolve is MILP based, mpmilp_program.py:75-148).
Stored per program: the feasible combinations in the tree's order, the presolved arrays of every sub-problem, and the
regions of every sub-problem in the reference's order.
"""
import itertools
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import gen_golden as gg  # noqa: E402  (loads the reference under the shim)
from ppopt.mp_solvers.solve_mpqp import mpqp_algorithm, solve_mpqp  # noqa: E402
from ppopt.mpmiqp_program import MPMIQP_Program  # noqa: E402


def simple_mpmiqp():   # /root/reference/tests/test_fixtures.py:207-220
    A = numpy.array([[0, 1, 1], [1, 0, 0], [-1, 0, 0], [1, -1, 0], [1, 0, -1]], float)
    b = numpy.array([1, 0, 0, 0, 0], float).reshape(-1, 1)
    F = numpy.array([0, 1, 0, 0, 0], float).reshape(-1, 1)
    c = numpy.array([-3, 0, 0], float).reshape(-1, 1)
    H = numpy.zeros((F.shape[1], A.shape[1])).T
    A_t = numpy.array([1, 1], float).reshape(-1, 1)
    b_t = numpy.array([2, 2], float).reshape(-1, 1)
    return MPMIQP_Program(A, b, c, H, numpy.eye(3), A_t, b_t, F, binary_indices=[1, 2], post_process=False)


def market_mpmiqp():   # /root/reference/tests/test_fixtures.py:246-262
    A = numpy.array(
        [[1, 1, 0, 0, 0], [0, 0, 1, 1, 0], [-1, 0, -1, 0, 0], [0, -1, 0, -1, -500], [-1, 0, 0, 0, 0], [0, -1, 0, 0, 0],
         [0, 0, -1, 0, 0], [0, 0, 0, -1, 0], [0, 0, 0, 0, -1], [0, 0, 0, 0, 1]], float)
    b = numpy.array([350, 600, 0, 0, 0, 0, 0, 0, 0, 1], float).reshape(-1, 1)
    F = numpy.array([[0, 0], [0, 0], [-1, 0], [0, -1], [0, 0], [0, 0], [0, 0], [0, 0], [0, 0], [0, 0]], float)
    A_t = numpy.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0], [0.0, -1.0]], float)
    b_t = numpy.array([[1000.0], [1000.0], [0.0], [0.0]], float)
    H = numpy.zeros([5, 2])
    Q = numpy.diag([153, 162, 162, 126, 1]).astype(float)
    c = numpy.array([25, 25, 25, 25, 7.6e6], float).reshape(-1, 1)
    return MPMIQP_Program(A, b, c, H, Q, A_t, b_t, F, binary_indices=[4], post_process=False)


def random_mpmiqp(seed=7, n_cont=5, n_bin=4, t=2, m=10):
    """a dense random mpMIQP: continuous box + random rows that the binaries shift (16 combinations, most feasible)"""
    rng = numpy.random.default_rng(seed)
    n = n_cont + n_bin
    A = numpy.vstack([rng.normal(size=(m, n)), numpy.hstack([numpy.eye(n_cont), numpy.zeros((n_cont, n_bin))]),
                      numpy.hstack([-numpy.eye(n_cont), numpy.zeros((n_cont, n_bin))])])
    A[:m, n_cont:] *= 0.5
    b = numpy.vstack([rng.uniform(1.0, 3.0, size=(m, 1)), 2.0 * numpy.ones((2 * n_cont, 1))])
    F = numpy.vstack([rng.normal(size=(m, t)) * 0.5, numpy.zeros((2 * n_cont, t))])
    R = rng.normal(size=(n, n))
    Q = R.T @ R + numpy.eye(n)
    c = rng.normal(size=(n, 1))
    H = rng.normal(size=(n, t)) * 0.3
    A_t = numpy.vstack([numpy.eye(t), -numpy.eye(t)])
    b_t = numpy.ones((2 * t, 1))
    return MPMIQP_Program(A, b, c, H, Q, A_t, b_t, F, binary_indices=list(range(n_cont, n)), post_process=False)


PROGRAMS = {'simple_mpmiqp': simple_mpmiqp, 'market_mpmiqp': market_mpmiqp, 'random_mpmiqp_5_4': random_mpmiqp}


def generate(name):
    prog = PROGRAMS[name]()
    nb = len(prog.binary_indices)
    out = {'binary_indices': numpy.array(prog.binary_indices, dtype=numpy.int32),
           'cont_indices': numpy.array(prog.cont_indices, dtype=numpy.int32)}
    combos = []
    n_regions = 0
    A_c, A_b = prog.A[:, prog.cont_indices], prog.A[:, prog.binary_indices]
    pure = [i for i in range(prog.A.shape[0]) if numpy.allclose(A_c[i], 0) and numpy.allclose(prog.F[i], 0)]
    for y in itertools.product((0, 1), repeat=nb):
        # rows without continuous variables and parameters are dropped by generate_substituted_problem (:85-93): the tree
        # is what enforces them (check_bin_feasibility); for a full assignment that is plain arithmetic
        lhs = A_b[pure] @ numpy.array(y, float).reshape(-1, 1) - prog.b[pure]
        viol = [i for k, i in enumerate(pure) if (abs(lhs[k, 0]) > 1e-9 if i in prog.equality_indices else lhs[k, 0] > 1e-9)]
        if viol:
            print(f'   {y}: violates the binary-only rows {viol}')
            continue
        try:
            sub = prog.generate_substituted_problem(list(y))
            feasible = sub.check_feasibility(list(sub.equality_indices)) is not None
        except Exception as e:   # noqa: BLE001
            print(f'   {y}: substituted problem rejected by the constructor ({type(e).__name__})')
            feasible = False
        if not feasible:
            continue
        k = len(combos)
        combos.append(list(y))
        sol = solve_mpqp(sub, mpqp_algorithm.combinatorial)
        for key in ('A', 'b', 'c', 'H', 'Q', 'A_t', 'b_t', 'F'):
            out[f's{k}_{key}'] = numpy.asarray(getattr(sub, key), dtype=float)
        out[f's{k}_n_eq'] = numpy.int64(len(sub.equality_indices))
        regs = {}
        gg.pack_regions(list(sol.critical_regions), regs)
        for key, v in regs.items():
            out[f's{k}_{key}'] = v
        n_regions += len(sol.critical_regions)
        print(f'   {y}: {sub.A.shape[0]} x {sub.A.shape[1]} after presolve, {len(sol.critical_regions)} regions')
    out['combinations'] = numpy.array(combos, dtype=numpy.int32).reshape(len(combos), nb)
    dst = os.path.join(ROOT, 'tests', 'golden', 'mpmiqp')
    os.makedirs(dst, exist_ok=True)
    numpy.savez_compressed(os.path.join(dst, name + '.npz'), **out)
    print(f'[{name}] {len(combos)} of {2 ** nb} combinations feasible, {n_regions} regions', flush=True)


if __name__ == '__main__':
    for nm in (sys.argv[1:] or PROGRAMS):
        generate(nm)
