"""Witness inheritance (ppgpu_level_eval_w, csrc/k6_children.cu::inherit_kernel): the vertex walk leaves, for every
candidate it certifies, the mask of ALL rows active at the certifying vertex; a candidate of the next level one of whose
parents has a witness that also holds the added row is certified by that same vertex before any LP work is spent on it.

The reference solves one LP per candidate (check_feasibility, mplp_program.py:411-444), so the only admissible effect is
on WHICH kernel exhibits the feasible point - never on a decision:
  * with and without witnesses, every status byte of every level is identical (and equal to the golden vectors of the
    unmodified reference where a golden level exists) - all golden programs, the walk forced on at every level;
  * every witness contains its candidate;
  * at the benchmarked depth (levels 1..5 of the 100x30x6 program, 74.5 M candidates) the complete status arrays of the two
    paths are equal byte for byte; the path without witnesses is the one tests/test_gpu_sampled.py pins to verdicts of the
    unmodified reference;
  * engine.solve() with PPGPU_INHERIT=0 / 1: same digest, same regions."""
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


def _levels(engine, eng, depth, witnesses):
    """level loop through the C ABI; returns [(masks, status, witness)]"""
    import torch
    from ppopt_b200._lib import ST_FEAS, WITNESS_SLOTS
    masks, parent, out = eng.root_level(), None, []
    for lvl in range(depth):
        n = masks.shape[0]
        if n == 0:
            break
        wit = torch.zeros((n, WITNESS_SLOTS, eng.W), dtype=torch.int64, device=eng.tdev) if witnesses else None
        st = eng.level_eval(masks, lvl + 1, witness=wit, parent=parent)
        out.append((masks, st, wit))
        if lvl + 1 == depth:
            break
        feas = eng.select(st, ST_FEAS, ST_FEAS)
        keep = {}
        masks = eng.children(masks, feas, lvl + 1, keep=keep)
        parent = engine.ParentLevel(keep['feas_masks'], keep['ws'], keep['nf'], wit[feas].contiguous()) if (witnesses and keep) else None
    return out


@pytest.mark.parametrize('name', golden_names())
def test_inheritance_never_changes_a_decision(name):
    import torch
    from ppopt_b200._lib import OPT_K2W_MIN
    engine, prog, eng = _engine(name)
    if not eng.has_walk_vertex:
        eng.close()
        pytest.skip('feasibility polyhedron without a vertex: no walk, no witnesses')
    g = numpy.load(os.path.join(GOLDEN, name + '.npz'))
    depth = min(int(g['n_levels']), eng.max_depth)
    eng.set_option(OPT_K2W_MIN, 0)          # the walk (the source of witnesses) on every level, however small
    c0 = eng.counters()
    on = _levels(engine, eng, depth, True)
    c1 = eng.counters()
    off = _levels(engine, eng, depth, False)
    c2 = eng.counters()
    assert c2['inherited'] == c1['inherited'], 'inheritance ran without a parent level'
    assert len(on) == len(off)
    for lv, ((m1, s1, w1), (m0, s0, _)) in enumerate(zip(on, off)):
        assert torch.equal(m1, m0), f'{name} level {lv + 1}: different candidates'
        assert torch.equal(s1 & 15, s0 & 15), f'{name} level {lv + 1}: decisions depend on the witnesses'
        if f'level{lv}_status' in g and len(g[f'level{lv}_status']) == s1.shape[0]:
            assert numpy.array_equal(s1.cpu().numpy() & 3, g[f'level{lv}_status'] & 3), f'{name} level {lv + 1}'
        # a witness holds its candidate
        for sl in range(w1.shape[1]):
            ws = w1[:, sl]
            has = (ws != 0).any(dim=1)
            assert bool(((m1 & ~ws) == 0).all(dim=1)[has].all()), f'{name} level {lv + 1}: witness does not contain its candidate'
            # ... and the first slot belongs to a candidate certified feasible (the second may also sit on a candidate the walk
            # gave up on earlier: it is only ever used if the relaxation / the simplex call that candidate feasible)
            if sl == 0:
                assert bool(((s1[has] & 2) != 0).all())
    eng.close()
    if name in ('synthetic_30_6_40_s0', 'rand_wide_40_8_90_s5'):
        assert c1['inherited'] - c0['inherited'] > 0, 'nothing was inherited'


def test_inheritance_at_the_benchmarked_depth():
    """levels 1..5 of the bench program with the default thresholds: byte-identical status arrays, most of level 5 inherited"""
    import torch
    engine, prog, eng = _engine('synthetic_30_6_40_s0')
    c0 = eng.counters()
    on = _levels(engine, eng, 5, True)
    c1 = eng.counters()
    sums = [(int(s.shape[0]), int((s & 2).ne(0).sum()), engine._checksum(s)) for _, s, _ in on]
    n5, inherited, piv_on = on[4][0].shape[0], c1['inherited'] - c0['inherited'], c1['k2w_pivots'] - c0['k2w_pivots']
    del on
    torch.cuda.empty_cache()
    off = _levels(engine, eng, 5, False)
    c2 = eng.counters()
    assert sums == [(int(s.shape[0]), int((s & 2).ne(0).sum()), engine._checksum(s)) for _, s, _ in off]
    piv_off = c2['k2w_pivots'] - c1['k2w_pivots']
    assert n5 == 70_595_661 or n5 > 7e7
    assert inherited > 0.5 * n5, (inherited, n5)
    assert piv_on < 0.8 * piv_off, (piv_on, piv_off)
    eng.close()


def test_solve_with_and_without_inheritance(monkeypatch):
    engine, prog, eng = _engine('synthetic_30_6_40_s0')
    monkeypatch.setenv('PPGPU_INHERIT', '1')
    a = engine.solve(prog, max_levels=4, engine=eng, digest=True)
    monkeypatch.setenv('PPGPU_INHERIT', '0')
    b = engine.solve(prog, max_levels=4, engine=eng, digest=True)
    assert a.digest == b.digest
    assert [r.active_set for r in a.critical_regions] == [r.active_set for r in b.critical_regions]
    assert a.engine_counters['inherited'] > 0
    eng.close()
