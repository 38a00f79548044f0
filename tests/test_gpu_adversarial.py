"""Adversarial numerics: decisions that sit inside a tolerance band must be escalated on
the GPU and reported, never taken silently.

Rank.  A program with a NEARLY dependent pair of rows: row m is row 3 plus relative noise delta.  The reference decides
rank with numpy.linalg.matrix_rank (SVD; constraint_utilities.py:222-236): for delta down to ~1e-14 the pair is still of
full rank for numpy, while a pivoted-QR ratio test at 1e-11 alone would call it deficient from 1e-11 downwards.  The
engine flags every such candidate PPG_ST_BORDER and re-decides it by singular values (one-sided Jacobi,
csrc/k1_rank.cu::k1_svd_recheck_kernel): the rank bit must equal numpy's on every candidate, for every noise level.
(The program bypasses the constructor's presolve, which would remove the duplicated row: this is a test of the kernels.)"""
import os

import numpy
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('delta', [1e-4, 1e-7, 1e-9, 1e-11, 1e-12, 1e-13, 0.0])
def test_rank_decisions_follow_the_singular_values(delta):
    import itertools
    from ppopt_b200 import engine
    from ppopt_b200._lib import ST_BORDER, ST_RANK
    from ppopt_b200.mplp_program import MPQP_Program
    g = numpy.load(os.path.join(GOLDEN, 'rand_5_3_10_s2.npz'))
    rng = numpy.random.default_rng(3)
    A, b, F = g['A'].copy(), g['b'].copy(), g['F'].copy()
    noise = rng.normal(size=A.shape[1])
    A = numpy.vstack([A, A[3] * (1.0 + delta * noise)])
    b = numpy.vstack([b, b[3:4] + 1.0])
    F = numpy.vstack([F, F[3]])
    prog = MPQP_Program(A, b, g['c'], g['H'], g['Q'], g['A_t'], g['b_t'], F, equality_indices=[], presolved=True)
    eng = engine.Engine(engine.program_arrays(prog))
    m = A.shape[0]
    n_flagged = 0
    for k in (2, 3):
        cands = [list(c) for c in itertools.combinations(range(m), k)]
        st = eng.level_eval(eng.masks_from_lists(cands), k, stages=1).cpu().numpy()
        want = numpy.array([numpy.linalg.matrix_rank(A[c]) == k for c in cands])
        got = (st & ST_RANK) != 0
        bad = numpy.nonzero(want != got)[0]
        assert bad.size == 0, f'delta {delta}: rank differs from numpy on {[cands[i] for i in bad[:5]]}'
        n_flagged += int(numpy.sum((st & ST_BORDER) != 0))
        pair = numpy.array([3 in c and m - 1 in c for c in cands])
        if 1e-14 < delta <= 1e-9:
            assert numpy.all((st[pair] & ST_BORDER) != 0), 'a borderline decision was taken without a flag'
    assert n_flagged == eng.counters()['border']
    eng.close()
