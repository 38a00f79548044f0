"""The real drop-in path: a genuine ``ppopt.MPQP_Program`` built by the UNMODIFIED reference, ``ppopt_b200.install()``,
then the reference's own ``ppopt.mp_solvers.solve_mpqp.solve_mpqp(prog, mpqp_algorithm.combinatorial)``
(/root/reference/src/ppopt/mp_solvers/solve_mpqp.py:52,70-71).  The Solution and its regions must be ppopt's OWN classes
and equal the golden regions the unmodified reference produced for the same program.

The reference travels to the GPU box as the git-ignored ``baseline/_ref`` install and is imported
under the LP shim of oracle/ref_shim (cvxopt/GLPK are not in this image); the solve itself never touches an LP solver -
it runs in libppgpu.so."""
import os
import sys

import numpy
import pytest

from conftest import GOLDEN, ROOT
from parity import REL_TOL, golden_regions, rel_err, rows_match_as_sets

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, 'oracle'))


def _reference():
    import ref_harness
    if not ref_harness.available():
        pytest.skip('reference not present (baseline/_ref missing)')
    return ref_harness.load()


@pytest.mark.parametrize('name', ['factory_mpqp', 'transport_mplp', 'mpc_n3', 'rand_6_3_12_s1', 'doc_portfolio'])
def test_install_runs_the_reference_api_on_the_gpu(name):
    ppopt = _reference()
    import problems
    import ppopt_b200
    from ppopt.critical_region import CriticalRegion
    from ppopt.mp_solvers import mpqp_combinatorial
    from ppopt.mp_solvers.solve_mpqp import mpqp_algorithm, solve_mpqp
    from ppopt.mplp_program import MPLP_Program
    from ppopt.mpqp_program import MPQP_Program
    from ppopt.solution import Solution
    d = problems.CONFIGS[name]()
    kw = {'post_process': d['post_process']} if 'post_process' in d else {}
    if d['kind'] == 'qp':
        prog = MPQP_Program(d['A'], d['b'], d['c'], d['H'], d['Q'], d['A_t'], d['b_t'], d['F'],
                            equality_indices=list(d['equality_indices']), **kw)
    else:
        prog = MPLP_Program(d['A'], d['b'], d['c'], d['H'], d['A_t'], d['b_t'], d['F'],
                            equality_indices=list(d['equality_indices']), **kw)
    g = numpy.load(os.path.join(GOLDEN, name + '.npz'))
    assert numpy.array_equal(prog.A, g['A'])   # the reference's own presolve ran
    original = ppopt_b200.install(ppopt)
    try:
        assert mpqp_combinatorial.solve is not original
        sol = solve_mpqp(prog, mpqp_algorithm.combinatorial)
    finally:
        mpqp_combinatorial.solve = original
    assert type(sol) is Solution and sol.program is prog
    assert getattr(sol, 'gpu_launches', 0) > 0, 'the solve did not run in libppgpu.so'
    ref = golden_regions(g)
    assert [list(r.active_set) for r in sol.critical_regions] == [r['active_set'].tolist() for r in ref]
    for a, b in zip(sol.critical_regions, ref):
        assert type(a) is CriticalRegion
        for fld in 'AbCd':
            assert rel_err(getattr(a, fld), b[fld]) <= REL_TOL, (name, a.active_set, fld)
        if prog.num_t() == 1:
            assert a.E.dtype.kind == 'i' and rel_err(a.f, b['f']) <= REL_TOL
        else:
            u1, u2 = rows_match_as_sets(a.E, a.f, b['E'], b['f'])
            assert not u1 and not u2, (name, a.active_set)
        assert a.omega_set == b['omega_set'].tolist() and a.lambda_set == b['lambda_set'].tolist() or prog.num_t() == 1
    # the objects behave like the reference's: evaluate / is_inside come from ppopt.critical_region
    r0 = sol.critical_regions[0]
    theta = numpy.linalg.lstsq(numpy.asarray(r0.E, dtype=float), numpy.asarray(r0.f, dtype=float) - 1e-3, rcond=None)[0]
    assert r0.evaluate(theta).shape == (prog.num_x(), 1)
    # point lookup through the reference's Solution API
    got = sol.get_region(theta)
    assert got is None or type(got) is CriticalRegion
