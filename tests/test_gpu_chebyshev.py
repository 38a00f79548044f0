"""Batched Chebyshev balls (ppgpu_chebyshev_batch, csrc/k34c_compact.cu::cheb_batch_kernel) against the radius the reference
computes: chebyshev_ball (utils/chebyshev_ball.py:10-63) is  max r : E theta + |E_i| r <= f, r >= 0  - restated here with
HiGHS (the LP backend of the golden run) on the regions of the golden files, plus hand-made polytopes with known radii."""
import os

import numpy
import pytest
from scipy.optimize import linprog

from conftest import GOLDEN
from parity import golden_regions

pytestmark = pytest.mark.gpu


def _radius_highs(E, f):
    nrm = numpy.linalg.norm(E, axis=1, keepdims=True)
    c = numpy.zeros(E.shape[1] + 1); c[-1] = -1.0
    res = linprog(c, A_ub=numpy.hstack([E, nrm]), b_ub=f.ravel(), bounds=[(None, None)] * E.shape[1] + [(None, None)], method='highs')
    return -res.fun if res.status == 0 else (numpy.inf if res.status == 3 else -numpy.inf)


def test_known_radii():
    from ppopt_b200.chebyshev import chebyshev_radii
    box = (numpy.vstack([numpy.eye(3), -numpy.eye(3)]), numpy.array([1, 2, 3, 1, 2, 3.0]))           # radius 1
    simplex = (numpy.array([[-1, 0], [0, -1], [1, 1.0]]), numpy.array([0, 0, 1.0]))                    # 1 / (2 + sqrt 2)
    flat = (numpy.array([[1, 0], [-1, 0], [0, 1], [0, -1.0]]), numpy.array([1, 1, 0, 0.0]))            # radius 0
    empty = (numpy.array([[1, 0], [-1, 0.0]]), numpy.array([-1, -1.0]))                                # empty: radius -1
    rad = chebyshev_radii([box]) .tolist() + chebyshev_radii([simplex, flat, empty]).tolist()
    assert abs(rad[0] - 1.0) < 1e-12 and abs(rad[1] - 1.0 / (2.0 + numpy.sqrt(2.0))) < 1e-12
    assert abs(rad[2]) < 1e-12 and abs(rad[3] + 1.0) < 1e-12


@pytest.mark.parametrize('name', ['rand_6_3_12_s1', 'ctrl_alloc_n2', 'mpc_n7', 'synthetic_30_6_40_s0'])
def test_region_radii_match_the_lp_backend(name):
    from ppopt_b200.chebyshev import chebyshev_radii, full_dimensional
    from ppopt_b200.critical_region import CriticalRegion
    regs = [r for r in golden_regions(numpy.load(os.path.join(GOLDEN, name + '.npz'))) if r['E'].shape[1] > 1]
    polys = [(numpy.asarray(r['E'], float), numpy.asarray(r['f'], float)) for r in regs]
    rad = chebyshev_radii(polys)
    want = numpy.array([_radius_highs(E, f) for E, f in polys])
    assert numpy.max(numpy.abs(rad - want) / numpy.maximum(1.0, numpy.abs(want))) < 1e-8
    # every region of a solution is full dimensional (that is how the reference kept it)
    objs = [CriticalRegion(None, None, None, None, E, f, []) for E, f in polys]
    assert all(full_dimensional(objs)) and objs[0].is_full_dimension()


def test_program_warnings():
    from ppopt_b200.mplp_program import load_presolved
    prog = load_presolved(os.path.join(GOLDEN, 'factory_mpqp.npz'))
    assert prog.warnings() == []
    prog.b = prog.b.copy(); prog.b[:] = -1e3    # no point satisfies the constraints any more
    w = prog.warnings()
    assert any('not feasible' in x for x in w)
