"""Pins the CPU oracle (oracle/ppopt_oracle.py) and the CPU checker of the batched algorithms (oracle/twin.cpp) to the
golden vectors that the UNMODIFIED reference produced in the build container (oracle/gen_golden.py)."""
import os
import sys

import numpy
import pytest

from conftest import GOLDEN, ROOT, golden_names
from parity import REL_TOL, check_status_bits, golden_regions, rel_err, rows_match_as_sets

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import ppopt_oracle as oracle  # noqa: E402
import twin_binding  # noqa: E402

FULL = ['factory_mpqp', 'transport_mplp', 'simple_mpqp_1d', 'simple_mplp', 'portfolio_analog', 'doc_portfolio', 'mpc_n3',
        'ctrl_alloc_n1', 'rand_lp_4_2_8_s3']
SAMPLED = ['mpc_n5', 'rand_6_3_12_s1', 'rand_5_3_10_s2', 'ctrl_alloc_n2', 'synthetic_30_6_40_s0', 'mpc_n10']


@pytest.mark.parametrize('name', FULL)
def test_oracle_full_solve_matches_reference(name):
    path = os.path.join(GOLDEN, name + '.npz')
    g = numpy.load(path)
    P = oracle.Program.from_npz(path)
    record = []
    regions = oracle.solve(P, record=record)
    assert len(record) == int(g['n_levels'])
    for lv, (cands, status) in enumerate(record):
        assert [list(c) for c in cands] == g[f'level{lv}_candidates'].tolist()
        assert numpy.array_equal(status, g[f'level{lv}_status']), f'{name} level {lv + 1}'
    ref = golden_regions(g)
    assert [r['active_set'] for r in regions] == [r['active_set'].tolist() for r in ref]
    for a, b in zip(regions, ref):
        for fld in 'AbCdEf':
            assert rel_err(a[fld], b[fld]) <= 1e-12, (name, a['active_set'], fld)
        assert a['omega_set'] == b['omega_set'].tolist()
        assert a['lambda_set'] == b['lambda_set'].tolist()
        assert a['regular_set'] == [b['regular_pos'].tolist(), b['regular_idx'].tolist()]


@pytest.mark.parametrize('name', SAMPLED)
def test_oracle_sampled_candidates_match_reference(name):
    path = os.path.join(GOLDEN, name + '.npz')
    g = numpy.load(path)
    P = oracle.Program.from_npz(path)
    rng = numpy.random.default_rng(7)
    for lv in range(int(g['n_levels'])):
        cands, status = g[f'level{lv}_candidates'], g[f'level{lv}_status']
        pick = set(rng.choice(len(cands), size=min(25, len(cands)), replace=False).tolist())
        pick |= set(numpy.nonzero(status & 8)[0][:6].tolist()) | set(numpy.nonzero(~status & 2)[0][:6].tolist())
        for i in sorted(pick):
            assert oracle.evaluate_candidate(P, cands[i].tolist()) == int(status[i]), (name, lv, cands[i])


@pytest.mark.parametrize('name', golden_names())
def test_twin_status_matches_reference(name):
    """rank / feasible / region decisions of the batched algorithm (CPU statement) == the reference's, every candidate"""
    path = os.path.join(GOLDEN, name + '.npz')
    g = numpy.load(path)
    tw = twin_binding.Twin.from_npz(path)
    for lv in range(int(g['n_levels'])):
        cands, ref = g[f'level{lv}_candidates'], g[f'level{lv}_status']
        if len(cands) > 12000:  # keep the CPU suite short; the GPU suite covers every candidate
            sel = numpy.random.default_rng(lv).choice(len(cands), 12000, replace=False)
            cands, ref = cands[sel], ref[sel]
        st = tw.eval(tw.masks(cands.tolist()))
        tolerated = check_status_bits(st, ref, f'{name} level {lv + 1}')
        assert len(tolerated) <= 0.001 * len(st) + (26 if name == 'ctrl_alloc_n5' else 0)
        assert not numpy.any(st & 32)


@pytest.mark.parametrize('name', [n for n in golden_names() if n != 'ctrl_alloc_n5'])
def test_twin_regions_match_reference(name):
    path = os.path.join(GOLDEN, name + '.npz')
    g = numpy.load(path)
    tw = twin_binding.Twin.from_npz(path)
    n, t = tw.n, tw.t
    for r in golden_regions(g):
        rc, laws, rows, flags, info = tw.emit(tw.masks([r['active_set'].tolist()])[0])
        assert rc == 1
        assert rel_err(laws[:n, 1:], r['A']) <= REL_TOL and rel_err(laws[:n, :1], r['b']) <= REL_TOL
        assert rel_err(laws[n:, 1:], r['C']) <= REL_TOL and rel_err(laws[n:, :1], r['d']) <= REL_TOL
        if t == 1:
            assert rel_err(numpy.array([[info[3]], [-info[2]]]), r['f']) <= REL_TOL
        else:
            sel = [i for i in range(len(flags)) if (flags[i] & 3) == 3 and not (flags[i] & 4)]
            u1, u2 = rows_match_as_sets(rows[sel, 1:], rows[sel, :1], r['E'], r['f'])
            assert not u1 and not u2, (name, r['active_set'])


def test_weakly_redundant_row_ambiguity_is_confined():
    """ctrl_alloc_n5: the reference backend's own keep/drop decisions overlap for margins in [-2.3e-8, 5e-11]
    (DESIGN.md); every row on which the twin and the reference disagree must lie in that band."""
    path = os.path.join(GOLDEN, 'ctrl_alloc_n5.npz')
    g = numpy.load(path)
    tw = twin_binding.Twin.from_npz(path)
    m, ne = tw.m, tw.n_eq
    worst = 0.0
    for r in golden_regions(g):
        aset = r['active_set'].tolist()
        rc, laws, rows, flags, info, mg = tw.emit(tw.masks([aset])[0], margins=True)
        ka, ninact = len(aset) - ne, m - len(aset)
        ref_kept = set([i for i, a in enumerate(aset[ne:]) if a in r['lambda_set'].tolist()]
                       + [ka + int(p) for p in r['regular_pos']] + [ka + ninact + int(o) for o in r['omega_set']])
        for i in range(len(flags)):
            if flags[i] & 1:
                mine_kept = mg[i] >= -1e-9
                if mine_kept != (i in ref_kept):
                    worst = max(worst, abs(mg[i]))
    assert worst < 5e-8
