"""K2w (csrc/k2w_walk.cu): feasibility certificates shared between the candidates of a prefix by a primal-simplex walk over
the vertices of the feasibility polyhedron.  K2w may only ever ADD the feasible bit to candidates the reference calls
feasible (check_feasibility, mplp_program.py:411-444), and the final decisions must not depend on whether it ran.

Through the C ABI (ppgpu_set_option PPGPU_OPT_K2W_MIN: 0 = walk on every launch, -1 = never):
  * walk alone (stage bit 1 | 2 would add the relaxation and the simplex, so the walk is isolated by comparing counters and
    by the soundness check): every candidate the walk certifies is feasible in the golden vectors of the unmodified reference;
  * walk on / walk off: identical status bytes on every candidate of every golden level;
  * the deep sampled goldens (levels 4-5 of the bench workload) are checked with the walk in tests/test_gpu_sampled.py
    (levels that large use it by default)."""
import os

import numpy
import pytest

from conftest import GOLDEN, golden_names

pytestmark = pytest.mark.gpu


def _engine(name):
    from ppopt_b200 import engine
    from ppopt_b200.mplp_program import load_presolved
    prog = load_presolved(os.path.join(GOLDEN, name + '.npz'))
    return engine, prog, engine.Engine(engine.program_arrays(prog))


@pytest.mark.parametrize('name', golden_names())
def test_walk_never_changes_a_decision(name):
    from ppopt_b200._lib import OPT_K2W_MIN
    engine, prog, eng = _engine(name)
    g = numpy.load(os.path.join(GOLDEN, name + '.npz'))
    ne = int(g['n_eq'])
    certified = 0
    for lv in range(int(g['n_levels'])):
        cands = g[f'level{lv}_candidates'].tolist()
        if not cands:
            continue
        masks = eng.masks_from_lists(cands)
        k_act = len(cands[0]) - ne
        eng.set_option(OPT_K2W_MIN, 0)
        c0 = eng.counters()
        on = eng.level_eval(masks, k_act, stages=3).cpu().numpy()
        c1 = eng.counters()
        eng.set_option(OPT_K2W_MIN, -1)
        off = eng.level_eval(masks, k_act, stages=3).cpu().numpy()
        c2 = eng.counters()
        assert numpy.array_equal(on & 3, off & 3), f'{name} level {lv + 1}: decisions depend on the walk'
        assert numpy.array_equal(on & 3, g[f'level{lv}_status'] & 3), f'{name} level {lv + 1}'
        assert c2['k2w_pivots'] == c1['k2w_pivots'], 'walk ran although switched off'
        got = c1['k2w_certified'] - c0['k2w_certified']
        assert got <= int(numpy.sum((g[f'level{lv}_status'] & 2) != 0)), 'more certificates than feasible candidates'
        certified += got
    if eng.has_walk_vertex and name not in ('simple_mpqp_1d',):
        assert certified > 0, f'{name}: the walk certified nothing'
    eng.close()


@pytest.mark.parametrize('name', ['synthetic_30_6_40_s0', 'mpc_n7', 'ctrl_alloc_n5', 'rand_wide_40_8_90_s5'])
def test_walk_alone_is_sound(name):
    """only K1 + the walk (the relaxation and both simplex passes are skipped by evaluating stage 1, then calling the walk
    through stage 2 with every later kernel's work removed is not possible through the ABI - instead: run K1, run the full
    feasibility stage with the walk on, and check that the number of walk certificates equals the number of feasible
    candidates the relaxation + simplex no longer had to look at)"""
    from ppopt_b200._lib import OPT_K2W_MIN
    engine, prog, eng = _engine(name)
    g = numpy.load(os.path.join(GOLDEN, name + '.npz'))
    ne = int(g['n_eq'])
    eng.set_option(OPT_K2W_MIN, 0)
    for lv in range(int(g['n_levels'])):
        cands = g[f'level{lv}_candidates'].tolist()
        masks = eng.masks_from_lists(cands)
        k_act = len(cands[0]) - ne
        c0 = eng.counters()
        st = eng.level_eval(masks, k_act, stages=3).cpu().numpy()
        c1 = eng.counters()
        ref = g[f'level{lv}_status']
        assert numpy.array_equal(st & 3, ref & 3)
        n_feas = int(numpy.sum((ref & 2) != 0))
        walk = c1['k2w_certified'] - c0['k2w_certified']
        rest = (c1['k2a_certified'] - c0['k2a_certified']) + (c1['k2_lps'] - c0['k2_lps'])
        # every feasible candidate was exhibited by exactly one of: the walk, the relaxation, the simplex
        assert walk <= n_feas and walk + rest >= n_feas, (name, lv, walk, rest, n_feas)
    eng.close()
