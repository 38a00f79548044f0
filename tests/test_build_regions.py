"""Host assembly of CriticalRegion objects from K5's buffers (engine.build_regions) against an independent per-region
statement kept in this test: values, shapes, contiguity and the Python types of the index lists
(field meaning: /root/reference/src/ppopt/utils/mpqp_utils.py:181-195).  Runs without a GPU."""
import types

import numpy
import pytest


def _plain(eng, cr_cls, active_sets, k_act, laws, rows, flags, info):
    n, t, ne, m = eng.n, eng.t, eng.n_eq, eng.m
    out = []
    n_inact = eng.mi - k_act
    for si, aset in enumerate(active_sets):
        if info[si, 0] != 1.0:
            out.append(None)
            continue
        law, fl = laws[si], flags[si]
        kept = numpy.nonzero((fl & 3) == 3)[0]
        if t == 1:
            E = numpy.array([[1], [-1]])
            f = numpy.array([[info[si, 3]], [-info[si, 2]]])
        else:
            nd = numpy.nonzero(((fl & 3) == 3) & ((fl & 4) == 0))[0]
            E = numpy.ascontiguousarray(rows[si][nd, 1:])
            f = numpy.ascontiguousarray(rows[si][nd, :1])
        active = aset[ne:]
        inactive = [i for i in range(m) if i not in set(aset)]
        lam = [active[i] for i in kept if i < k_act]
        reg_pos = [int(i - k_act) for i in kept if k_act <= i < k_act + n_inact]
        omega = [int(i - k_act - n_inact) for i in kept if i >= k_act + n_inact]
        out.append(cr_cls(numpy.ascontiguousarray(law[:n, 1:]), numpy.ascontiguousarray(law[:n, :1]),
                          numpy.ascontiguousarray(law[n:, 1:]), numpy.ascontiguousarray(law[n:, :1]), E, f, list(aset),
                          omega, lam, [reg_pos, [inactive[p] for p in reg_pos]]))
    return out


@pytest.mark.parametrize('n,t,ne,m,q,k_act', [(30, 6, 0, 100, 12, 5), (9, 2, 6, 18, 4, 3), (4, 1, 2, 12, 2, 1), (6, 3, 0, 12, 6, 6),
                                                (5, 2, 1, 9, 4, 0)])
def test_region_assembly_equals_plain_statement(n, t, ne, m, q, k_act):
    from ppopt_b200 import engine
    from ppopt_b200.critical_region import CriticalRegion
    mi = m - ne
    eng = types.SimpleNamespace(n=n, t=t, n_eq=ne, m=m, mi=mi)
    R0, k, N = mi + q, ne + k_act, 37
    rng = numpy.random.default_rng(n * 100 + t)
    laws, rows = rng.standard_normal((N, n + k, t + 1)), rng.standard_normal((N, R0, t + 1))
    info = numpy.ones((N, 4))
    info[:, 2], info[:, 3] = -rng.random(N), rng.random(N)
    info[rng.random(N) < 0.2, 0] = 0.0
    info[3, 0] = -1.0
    flags = rng.choice(numpy.array([0, 1, 2, 3, 7], dtype=numpy.int32), size=(N, R0))
    flags[5] = 0
    asets = [list(range(ne)) + sorted((ne + rng.choice(mi, k_act, replace=False)).tolist()) for _ in range(N)]
    got = engine.build_regions(eng, CriticalRegion, asets, k_act, laws, rows, flags, info)
    want = _plain(eng, CriticalRegion, asets, k_act, laws, rows, flags, info)
    assert len(got) == len(want) == N
    for a, b in zip(got, want):
        assert (a is None) == (b is None)
        if a is None:
            continue
        for fld in 'AbCdEf':
            x, y = getattr(a, fld), getattr(b, fld)
            assert x.shape == y.shape and x.dtype == y.dtype and numpy.array_equal(x, y) and x.flags.c_contiguous, fld
        assert a.active_set == b.active_set and a.omega_set == b.omega_set and a.lambda_set == b.lambda_set
        assert a.regular_set == b.regular_set
        for lst in (a.omega_set, a.lambda_set, a.regular_set[0], a.regular_set[1]):
            assert all(type(v) is int for v in lst)
    assert engine.build_regions(eng, CriticalRegion, [], k_act, laws[:0], rows[:0], flags[:0], info[:0]) == []
