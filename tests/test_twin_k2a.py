"""CPU checks of the feasibility-certificate algorithm (K2a) and of the K2a -> K2 hand-over, on the sequential restatement
in oracle/twin.cpp, against the reference's golden status bytes (no GPU needed):
  * soundness: a candidate the relaxation certifies is feasible for the reference (bit 1 of the golden status);
  * the residuals handed over for an uncertified candidate are exact, and the feasibility LP seen from that point decides
    exactly like the LP seen from the origin (it is the same LP, translated)."""
import os
import sys

import numpy
import pytest

from conftest import GOLDEN, ROOT, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))

CAP = 20000   # candidates per level looked at (seeded sample above that)


def _levels(g, tw, max_rows=8):
    n_eq = int(g['n_eq'])
    for lv in range(int(g['n_levels'])):
        c, st = g[f'level{lv}_candidates'], g[f'level{lv}_status']
        if len(c) == 0 or c.shape[1] - n_eq > max_rows:
            continue   # the register-resident certificate kernel covers 1..8 activated rows, the prefix form 1..18
        if len(c) > CAP:
            sel = numpy.random.default_rng(lv).choice(len(c), CAP, replace=False)
            c, st = c[sel], st[sel]
        yield c, st


@pytest.mark.parametrize('prefix', [False, True], ids=['per_candidate', 'prefix_form'])
@pytest.mark.parametrize('name', golden_names())
def test_certificates_are_sound_and_handover_is_exact(name, prefix):
    from twin_binding import Twin
    g = numpy.load(os.path.join(GOLDEN, name + '.npz'))
    tw = Twin.from_npz(os.path.join(GOLDEN, name + '.npz'))
    if tw.R0 > 128:
        pytest.skip('outside the envelope of the register-resident kernel (R0 <= 128)')
    n_eq = int(g['n_eq'])
    checked_lp = 0
    for c, st in _levels(g, tw, 18 if prefix else 8):
        flags, steps, resid = tw.k2a(tw.masks(c.tolist()), prefix=prefix)
        rank_ok, feasible = (st & 1) != 0, (st & 2) != 0
        assert not numpy.any((flags == 1) & rank_ok & ~feasible), 'a certified candidate is infeasible for the reference'
        assert numpy.all(steps <= 192)
        # hand-over: a sample of the uncertified candidates, feasible and infeasible ones
        todo = numpy.nonzero((flags == 0) & rank_ok)[0]
        for i in todo[:60]:
            rows = [int(a) - n_eq for a in c[i][n_eq:]]
            # (prefix form: the prefix rows are held by the numeric projector W, not re-solved - on ill-conditioned prefixes
            # their residual is cond x eps; the residuals are still the exact ones of the point, which is all K2 needs)
            if not prefix:
                assert numpy.max(numpy.abs(resid[i][rows])) <= 1e-6 * max(1.0, numpy.max(numpy.abs(resid[i]))), \
                    'the last iterate does not satisfy the active rows'
            cold, _ = tw.feas_from(rows)
            assert cold == bool(feasible[i])
            if numpy.isnan(resid[i][0]):
                continue   # prefix form, ill-conditioned prefix: no warm hand-over, the simplex starts cold
            warm, _ = tw.feas_from(rows, -resid[i])
            assert warm == cold
            checked_lp += 1
    if name in ('mpc_n5', 'mpc_n7', 'rand_6_3_12_s1', 'rand_5_3_10_s2'):
        assert checked_lp > 0   # these programs do leave candidates to the simplex


@pytest.mark.parametrize('prefix', [False, True], ids=['per_candidate', 'prefix_form'])
def test_certification_rate_on_the_headline_program(prefix):
    """synthetic 100 x 30 x 6 program: at levels 1-2 everything that is feasible is certified (the GPU counters say the
    same: k2a certified == tried at these levels), at level 3 at least 97 %"""
    from twin_binding import Twin
    name = 'synthetic_30_6_40_s0'
    g = numpy.load(os.path.join(GOLDEN, name + '.npz'))
    tw = Twin.from_npz(os.path.join(GOLDEN, name + '.npz'))
    for lv, (c, st) in enumerate(_levels(g, tw)):
        flags, steps, _ = tw.k2a(tw.masks(c.tolist()), prefix=prefix)
        ok = (st & 1) != 0
        if lv < 2:
            assert numpy.array_equal(flags[ok] == 1, ((st & 2) != 0)[ok])
        else:
            assert (flags[ok] == 1).sum() >= 0.97 * ((st & 2) != 0)[ok].sum()
        assert 5 < steps[ok].mean() < 45


@pytest.mark.parametrize('name', golden_names())
def test_walk_start_vertex_is_a_feasible_vertex(name):
    """host_math.hpp::build_walk_dictionary (the start of every K2w walk): the nonbasic rows form a nonsingular system, the
    dictionary equals the one numpy derives from that basis, and the vertex is primal feasible"""
    from twin_binding import Twin
    tw = Twin.from_npz(os.path.join(GOLDEN, name + '.npz'))
    wd = tw.walk_dict()
    if wd is None:
        pytest.skip('feasibility polyhedron without a vertex: the walk is off for this program')
    D0, bvar, nvar, T0 = wd
    h, G = T0[:, 0], T0[:, 1:-1]
    assert sorted(bvar.tolist() + nvar.tolist()) == list(range(G.shape[0]))
    M = G @ numpy.linalg.inv(G[nvar])
    beta = (h - M @ h[nvar])[bvar]
    scale = max(1.0, float(numpy.max(numpy.abs(h))))
    assert numpy.max(numpy.abs(D0[:, 1:] + M[bvar])) <= 1e-9 * max(1.0, float(numpy.max(numpy.abs(M))))
    assert numpy.max(numpy.abs(D0[:, 0] - numpy.maximum(beta, 0.0))) <= 1e-9 * scale
    assert beta.min() >= -1e-9 * scale


@pytest.mark.parametrize('name', golden_names())
def test_walk_certificates_are_sound(name):
    """CPU restatement of K2w (oracle/twin.cpp::k2w_walk_level = csrc/k2w_walk.cu): on every golden level no candidate the
    walk certifies is infeasible for the unmodified reference, and on programs whose feasible sets dominate it certifies
    most of them with about one pivot per candidate"""
    from twin_binding import Twin
    path = os.path.join(GOLDEN, name + '.npz')
    g = numpy.load(path)
    tw = Twin.from_npz(path)
    if tw.walk_dict() is None:
        pytest.skip('no start vertex')
    tot = cert = feas = piv = 0
    n_eq = int(g['n_eq'])
    for lv in range(int(g['n_levels'])):
        c, st = g[f'level{lv}_candidates'], g[f'level{lv}_status']
        if len(c) == 0 or c.shape[1] - n_eq > 32 or len(c) > 200000:
            continue          # (whole levels, in order: the walk shares certificates between neighbours)
        cert_f, p = tw.k2w(tw.masks(c.tolist()))
        ok = (st & 3) == 3          # full rank and feasible for the reference
        # (the kernel only looks at candidates that passed the rank screen: the LP of a rank-deficient set may well be
        # feasible, the reference reports such sets infeasible without solving it, mplp_program.py:433-435)
        bad = cert_f.astype(bool) & ((st & 1) != 0) & ((st & 2) == 0)
        assert not numpy.any(bad), f'{name}: the walk certified an infeasible candidate'
        tot += len(c); cert += int(numpy.sum(cert_f.astype(bool) & ok)); feas += int(ok.sum()); piv += p
    assert tot > 0
    if name == 'synthetic_30_6_40_s0':
        assert cert >= 0.99 * feas and piv <= 2.0 * tot, (cert, feas, piv, tot)


@pytest.mark.parametrize('name', golden_names())
def test_witnesses_are_feasible_vertices_and_inheritance_is_sound(name):
    """What the next level inherits (ppgpu_level_eval_w, k6_children.cu::inherit_kernel), restated on the CPU: the witness of
    a certified candidate is the active-row set of the certifying vertex (slot 0) or of a later vertex of the walk that holds
    the candidate as well (slot 1).  Checked with numpy, independently of the walk:
    the witness contains its candidate, its rows determine a point z, and z satisfies every row of the feasibility system.
    And the inheritance rule against the UNMODIFIED reference's verdicts: every golden candidate of the next level that
    lies inside the witness of one of its parents (and passes the rank screen) is feasible for the reference."""
    from twin_binding import Twin
    path = os.path.join(GOLDEN, name + '.npz')
    g = numpy.load(path)
    tw = Twin.from_npz(path)
    wd = tw.walk_dict()
    if wd is None:
        pytest.skip('no start vertex')
    T0 = wd[3]
    h, G = T0[:, 0], T0[:, 1:-1]
    n_eq = int(g['n_eq'])
    scale = max(1.0, float(numpy.max(numpy.abs(h))))
    checked = inherited = 0
    for lv in range(int(g['n_levels']) - 1):
        c, c2, st2 = g[f'level{lv}_candidates'], g[f'level{lv + 1}_candidates'], g[f'level{lv + 1}_status']
        st = g[f'level{lv}_status']
        if len(c) == 0 or len(c2) == 0 or c.shape[1] - n_eq > 32 or c.shape[1] - n_eq < 1 or len(c) > 20000:
            continue
        masks = tw.masks(c.tolist())
        cert, _, wit2 = tw.k2w_witness(masks, slots=2)
        wit = numpy.ascontiguousarray(wit2[:, 0])
        m64 = numpy.ascontiguousarray(masks).view(numpy.uint64).reshape(len(c), -1)
        has = cert.astype(bool)
        assert numpy.all((wit2[~has, 0] == 0))   # (slot 1 may also sit on a candidate the walk gave up on: still a vertex that holds it)
        bits = numpy.unpackbits(wit.view(numpy.uint8), axis=1, bitorder='little')[:, :G.shape[0]].astype(bool)
        seen = set()
        for sl in range(2):
            ws = numpy.ascontiguousarray(wit2[:, sl])
            hs = (ws != 0).any(axis=1)
            assert numpy.all((m64[hs] & ~ws[hs, :m64.shape[1]]) == 0), f'{name}: a witness does not contain its candidate'
            bs = numpy.unpackbits(ws.view(numpy.uint8), axis=1, bitorder='little')[:, :G.shape[0]].astype(bool)
            for i in numpy.nonzero(hs)[0][:400]:
                key = ws[i].tobytes()
                if key in seen:
                    continue
                seen.add(key)
                rows = numpy.nonzero(bs[i])[0]
                assert len(rows) == G.shape[1], f'{name}: a witness is not a basis'
                z = numpy.linalg.solve(G[rows], h[rows])
                assert numpy.max(G @ z - h) <= 1e-7 * scale, f'{name}: the witness vertex violates a row'
                checked += 1
        bits2 = numpy.unpackbits(numpy.ascontiguousarray(wit2[:, 1]).view(numpy.uint8), axis=1, bitorder='little')[:, :G.shape[0]].astype(bool)
        # inheritance: child = parent + one row, inside the parent's witness
        by_set = {tuple(r): i for i, r in enumerate(c.tolist())}
        for child, s2 in zip(c2.tolist(), st2.tolist()):
            for drop in range(n_eq, len(child)):
                par = tuple(child[:drop] + child[drop + 1:])
                i = by_set.get(par)
                if i is None or (st[i] & 3) != 3:     # the engine only looks at the witnesses of FEASIBLE parents
                    continue
                if bits[i, child[drop] - n_eq] or bits2[i, child[drop] - n_eq]:
                    inherited += 1
                    if s2 & 1:
                        assert s2 & 2, f'{name}: {child} inherits a certificate but the reference calls it infeasible'
                    break
    if name == 'synthetic_30_6_40_s0':
        assert checked > 0 and inherited > 1000
