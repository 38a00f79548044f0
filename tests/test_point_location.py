"""Batched point location: oracle restatement against vectors produced by the reference's own
upop.PointLocation / Solution.get_region (tests/golden/pointloc/*.npz, oracle/gen_pointloc_golden.py), and the K7 kernel
against the same vectors through the C ABI."""
import os
import sys

import numpy
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
PL_DIR = os.path.join(GOLDEN, 'pointloc')
NAMES = sorted(f[:-4] for f in os.listdir(PL_DIR) if f.endswith('.npz'))


@pytest.mark.parametrize('name', NAMES)
def test_oracle_matches_reference_point_location(name):
    import point_location_oracle as plo
    g = numpy.load(os.path.join(PL_DIR, name + '.npz'))
    regions = plo.load_regions(g)
    tol = float(g['point_location_tolerance'])
    assert tol == 1e-5   # solution.py:23
    differ = 0
    for p, th in enumerate(g['thetas']):
        ju, js = plo.locate_upop(regions, th), plo.locate_solution(regions, th, tol)
        assert ju == g['idx_upop'][p] and js == g['idx_solution'][p]
        differ += ju != js
        x = plo.evaluate(regions, js, th)
        if js < 0:
            assert x is None and numpy.all(numpy.isnan(g['x_solution'][p]))
        else:
            assert numpy.array_equal(x.ravel(), g['x_solution'][p])
    assert differ >= 50   # the cloud holds points between the two acceptance rules (3e-6 outside a facet)


@pytest.mark.parametrize('name', NAMES)
def test_oracle_matches_reference_overlap_rule(name):
    """lowest objective among the containing regions, ties to the later one - bit-identical numpy operations, so even the
    rounding-level ties of points that sit in two regions reproduce"""
    import point_location_oracle as plo
    g, m = numpy.load(os.path.join(PL_DIR, name + '.npz')), numpy.load(os.path.join(GOLDEN, name + '.npz'))
    prog = {'Q': m['Q'] if 'Q' in m else None, 'H': m['H'], 'c': m['c']}
    regions = plo.load_regions(g)
    for p, th in enumerate(g['thetas']):
        assert plo.locate_overlap(regions, prog, th, 1e-5) == g['idx_solution_overlap'][p]
        assert plo.locate_overlap(regions, prog, th, None) == g['idx_upop_overlap'][p]


def _solution_from_golden(g, program=None, overlapping=False):
    from ppopt_b200.critical_region import CriticalRegion
    from ppopt_b200.solution import Solution
    regs = [CriticalRegion(g[f'r{i}_A'], g[f'r{i}_b'], None, None, numpy.asarray(g[f'r{i}_E'], dtype=float), g[f'r{i}_f'], [])
            for i in range(int(g['n_regions']))]
    return Solution(program, regs, is_overlapping=overlapping)


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_gpu_point_location_matches_reference(name):
    from ppopt_b200 import PointLocation
    g = numpy.load(os.path.join(PL_DIR, name + '.npz'))
    pl = PointLocation(_solution_from_golden(g))
    thetas = g['thetas']
    assert numpy.array_equal(pl.locate_batch(thetas), g['idx_upop'])                 # E theta <= f
    idx, x = pl.evaluate_batch(thetas, tol=float(g['point_location_tolerance']))     # E theta - f < 1e-5
    assert numpy.array_equal(idx, g['idx_solution'])
    found = idx >= 0
    assert numpy.all(numpy.isnan(x[~found]))
    scale = numpy.maximum(1.0, numpy.abs(g['x_solution'][found]))
    assert numpy.max(numpy.abs(x[found] - g['x_solution'][found]) / scale) <= 1e-12   # fma chain vs BLAS gemv
    # the reference's single-point interface (upop/point_location.py:91-133)
    for p in (0, 1, 2, len(thetas) - 1):
        th = thetas[p].reshape(-1, 1)
        assert pl.locate(th) == g['idx_upop'][p] and pl.is_inside(th) == (g['idx_upop'][p] >= 0)
        xe = pl.evaluate(th)
        assert (xe is None) == (g['idx_upop'][p] < 0)
    assert pl.locate_batch(numpy.zeros((0, thetas.shape[1]))).shape == (0,)


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_gpu_point_location_overlap_rule(name):
    """overlapping solutions: the reported region must contain the point and carry the reference's objective value; the
    index itself must match wherever a single region contains the point (where two do, their objectives agree to
    ~1e-14 and the reference's own choice is rounding noise - measured in oracle/gen_pointloc_golden.py's cloud)"""
    import point_location_oracle as plo
    from ppopt_b200 import PointLocation
    from ppopt_b200.mplp_program import load_presolved
    g = numpy.load(os.path.join(PL_DIR, name + '.npz'))
    program = load_presolved(os.path.join(GOLDEN, name + '.npz'))
    prog = {'Q': getattr(program, 'Q', None), 'H': program.H, 'c': program.c}
    regions = plo.load_regions(g)
    pl = PointLocation(_solution_from_golden(g, program, overlapping=True))
    thetas = g['thetas']
    for tol, key in ((None, 'idx_upop_overlap'), (1e-5, 'idx_solution_overlap')):
        idx, x = pl.evaluate_batch(thetas, tol=tol)
        ref = g[key]
        assert numpy.array_equal(idx >= 0, ref >= 0)
        ties = 0
        for p in numpy.nonzero(idx != ref)[0]:
            th = thetas[p].reshape(-1, 1)
            A, b, E, f = regions[idx[p]]
            assert numpy.all(E @ th <= f) if tol is None else numpy.all(E @ th - f < tol)
            o_gpu = plo.objective(prog, A @ th + b, th)
            o_ref = plo.objective(prog, regions[ref[p]][0] @ th + regions[ref[p]][1], th)
            assert abs(o_gpu - o_ref) <= 1e-9 * max(1.0, abs(o_ref))
            ties += 1
        multi = sum(1 for th in thetas if plo.locate_overlap(regions, prog, th, tol, return_gap=True)[1] != numpy.inf)
        assert ties <= multi
        for p in numpy.nonzero(idx >= 0)[0][:50]:
            A, b = regions[idx[p]][0], regions[idx[p]][1]
            want = (A @ thetas[p].reshape(-1, 1) + b).ravel()
            assert numpy.max(numpy.abs(x[p] - want) / numpy.maximum(1.0, numpy.abs(want))) <= 1e-12


@pytest.mark.gpu
def test_gpu_point_location_on_an_engine_solution():
    """end to end: enumerate on the GPU, then locate on the GPU; the regions equal the reference's to 1e-8, the cloud's
    closest points are 3e-6 from a facet, so the located indices must be the reference's"""
    from ppopt_b200 import PointLocation, mpqp_algorithm, solve_mpqp
    from ppopt_b200.mplp_program import load_presolved
    g = numpy.load(os.path.join(PL_DIR, 'factory_mpqp.npz'))
    sol = solve_mpqp(load_presolved(os.path.join(GOLDEN, 'factory_mpqp.npz')), mpqp_algorithm.combinatorial)
    assert sol.is_overlapping   # what solve_mpqp leaves on every solution (solve_mpqp.py:105-112)
    pl = PointLocation(sol)     # factory: no point of the cloud sits in two regions, so the indices are unambiguous
    assert numpy.array_equal(pl.locate_batch(g['thetas']), g['idx_upop_overlap'])
    assert numpy.array_equal(pl.locate_batch(g['thetas'], tol=1e-5), g['idx_solution_overlap'])
    sol.is_overlapping = False
    pl = PointLocation(sol)
    assert numpy.array_equal(pl.locate_batch(g['thetas']), g['idx_upop'])
    assert numpy.array_equal(pl.locate_batch(g['thetas'], tol=1e-5), g['idx_solution'])


def test_point_location_needs_the_gpu_library():
    """no CPU fallback: constructing the locator without a CUDA device fails loudly"""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from ppopt_b200 import PointLocation
    g = numpy.load(os.path.join(PL_DIR, 'factory_mpqp.npz'))
    with pytest.raises(RuntimeError):
        PointLocation(_solution_from_golden(g))
