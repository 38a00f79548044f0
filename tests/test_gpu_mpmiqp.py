"""GPU parity of the mpMIQP enumeration (ppopt_b200.mp_solvers.mpmiqp_enumeration) against
tests/golden/mpmiqp/*.npz: the UNMODIFIED reference's generate_substituted_problem + solve_mpqp(combinatorial) for every
feasible binary combination (oracle/gen_mpmiqp_golden.py).  The program object below stands for the reference's
MPMIQP_Program: it hands out the sub-problems the reference built (post-presolve arrays), so what is tested is everything
this package adds - screening the 2^nb assignments, solving the sub-problems concurrently on CUDA streams, annotating and
merging the regions in the reference's order."""
import os

import numpy
import pytest

from conftest import GOLDEN
from parity import REL_TOL, rel_err, rows_match_as_sets

pytestmark = pytest.mark.gpu
MI = os.path.join(GOLDEN, 'mpmiqp')
NAMES = sorted(f[:-4] for f in os.listdir(MI) if f.endswith('.npz')) if os.path.isdir(MI) else []


class StoredMixedIntegerProgram:
    def __init__(self, g):
        from ppopt_b200.mplp_program import MPQP_Program
        self._cls = MPQP_Program
        self.g = g
        self.binary_indices = g['binary_indices'].tolist()
        self.cont_indices = g['cont_indices'].tolist()
        self.index = {tuple(c): k for k, c in enumerate(g['combinations'].tolist())}
        n = len(self.binary_indices) + len(self.cont_indices)
        # the binary-only rows of the original program are not part of the fixtures: no assignment is excluded by them here
        self.A, self.F, self.b = numpy.zeros((0, n)), numpy.zeros((0, 1)), numpy.zeros((0, 1))
        self.equality_indices = []

    def generate_substituted_problem(self, y):
        if tuple(y) not in self.index:
            raise ValueError('infeasible assignment (not in the reference\'s list)')
        k, g = self.index[tuple(y)], self.g
        return self._cls(g[f's{k}_A'], g[f's{k}_b'], g[f's{k}_c'], g[f's{k}_H'], g[f's{k}_Q'], g[f's{k}_A_t'], g[f's{k}_b_t'],
                         g[f's{k}_F'], equality_indices=list(range(int(g[f's{k}_n_eq']))), presolved=True)


def _ref_regions(g, k):
    keys = ('active_set', 'A', 'b', 'C', 'd', 'E', 'f')
    return [{key: g[f's{k}_r{i}_{key}'] for key in keys} for i in range(int(g[f's{k}_n_regions']))]


@pytest.mark.parametrize('name', NAMES)
@pytest.mark.parametrize('streams', [1, 4])
def test_enumeration_matches_reference(name, streams):
    from ppopt_b200.mp_solvers.solve_mpmiqp import solve_mpmiqp
    g = numpy.load(os.path.join(MI, name + '.npz'))
    prog = StoredMixedIntegerProgram(g)
    combos = g['combinations'].tolist()
    # given the tree's leaves, and screening all 2^nb assignments itself (infeasible ones raise in this stand-in program)
    for given in (combos, None):
        from ppopt_b200.mp_solvers import mpmiqp_enumeration
        sol = mpmiqp_enumeration.solve_mpmiqp_enumeration(prog, feasible_combinations=given, streams=streams)
        assert sol.feasible_combinations == combos
        assert sol.is_overlapping
        want = [(c, r) for k, c in enumerate(combos) for r in _ref_regions(g, k)]
        assert len(sol.critical_regions) == len(want)
        for a, (c, b) in zip(sol.critical_regions, want):
            assert list(a.active_set) == b['active_set'].tolist() and a.y_fixation == c
            assert a.y_indices == prog.binary_indices and a.x_indices == prog.cont_indices
            for fld in 'AbCd':
                assert rel_err(getattr(a, fld), b[fld]) <= REL_TOL, (name, c, a.active_set, fld)
            u1, u2 = rows_match_as_sets(a.E, a.f, b['E'], b['f'])
            assert not u1 and not u2, (name, c, a.active_set)
    full = sol.critical_regions[0].evaluate(numpy.zeros((a.A.shape[1], 1)))
    assert full.shape == (len(prog.binary_indices) + len(prog.cont_indices), 1)
    assert solve_mpmiqp(prog, feasible_combinations=combos).feasible_combinations == combos
