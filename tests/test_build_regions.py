"""Host assembly of CriticalRegion objects from K5's buffers (engine.build_regions), checked against what the UNMODIFIED
reference produced (tests/golden/*.npz): the buffers come from the CPU checker's emit (oracle/twin.cpp::emit_region, same
layout as ppgpu_regions_emit), the expectations - matrices, kept-index lists omega_set / lambda_set / regular_set
(/root/reference/src/ppopt/utils/mpqp_utils.py:181-195), dtypes - from the golden regions.  Runs without a GPU."""
import os
import sys
import types

import numpy
import pytest

from conftest import GOLDEN, ROOT
from parity import REL_TOL, golden_regions, index_lists_match, rel_err, rows_match_as_sets

sys.path.insert(0, os.path.join(ROOT, 'oracle'))

NAMES = ['factory_mpqp', 'transport_mplp', 'simple_mplp', 'doc_portfolio', 'portfolio_analog', 'mpc_n5', 'mpc_n10',
         'ctrl_alloc_n1', 'rand_6_3_12_s1', 'rand_5_3_10_s2', 'rand_lp_4_2_8_s3', 'synthetic_30_6_40_s0']


@pytest.mark.parametrize('name', NAMES)
def test_region_assembly_reproduces_the_reference_regions(name):
    from twin_binding import Twin
    from ppopt_b200 import engine
    from ppopt_b200.critical_region import CriticalRegion
    path = os.path.join(GOLDEN, name + '.npz')
    g = numpy.load(path)
    tw = Twin.from_npz(path)
    eng = types.SimpleNamespace(n=tw.n, t=tw.t, n_eq=tw.n_eq, m=tw.m, mi=tw.m - tw.n_eq)
    ref = golden_regions(g)[:120]
    by_k = {}
    for r in ref:
        by_k.setdefault(len(r['active_set']) - tw.n_eq, []).append(r)
    n_checked = 0
    for k_act, regs in by_k.items():
        bufs = [tw.emit(tw.masks([r['active_set'].tolist()])[0], margins=True) for r in regs]
        assert all(b[0] == 1 for b in bufs)
        laws = numpy.stack([b[1] for b in bufs])
        rows = numpy.stack([b[2] for b in bufs])
        flags = numpy.stack([b[3] for b in bufs])
        info = numpy.stack([b[4] for b in bufs])
        asets = [r['active_set'].tolist() for r in regs]
        got = engine.build_regions(eng, CriticalRegion, asets, k_act, laws, rows, flags, info)
        assert len(got) == len(regs)
        for a, r, b in zip(got, regs, bufs):
            assert a is not None and a.active_set == r['active_set'].tolist()
            for fld in 'AbCd':
                x = getattr(a, fld)
                assert x.shape == r[fld].shape and x.dtype == numpy.float64 and x.flags.c_contiguous, fld
                assert rel_err(x, r[fld]) <= REL_TOL, (name, a.active_set, fld)
            if tw.t == 1:
                assert a.E.dtype.kind == 'i' and a.E.tolist() == r['E'].tolist() and rel_err(a.f, r['f']) <= REL_TOL
            else:
                assert a.E.shape[1] == r['E'].shape[1] and a.f.shape[1] == 1 and a.E.flags.c_contiguous
            for lst in (a.omega_set, a.lambda_set, a.regular_set[0], a.regular_set[1]):
                assert all(type(v) is int for v in lst)
            bad = index_lists_match(a, r, rows=b[2], flags=b[3], margins=b[5], k_act=k_act, n_eq=tw.n_eq, m=tw.m, one_d=tw.t == 1)
            assert not bad, (name, a.active_set, bad)
            if tw.t > 1:
                u1, u2 = rows_match_as_sets(a.E, a.f, r['E'], r['f'])
                assert not u1 and not u2, (name, a.active_set)
            n_checked += 1
    assert n_checked == len(ref)


def test_region_assembly_skips_non_regions_and_handles_empty_input():
    from ppopt_b200 import engine
    from ppopt_b200.critical_region import CriticalRegion
    eng = types.SimpleNamespace(n=3, t=2, n_eq=0, m=6, mi=6)
    laws, rows = numpy.zeros((3, 5, 3)), numpy.zeros((3, 8, 3))
    flags, info = numpy.zeros((3, 8), dtype=numpy.int32), numpy.zeros((3, 4))
    info[1, 0] = -1.0   # singular KKT system
    assert engine.build_regions(eng, CriticalRegion, [[0, 1]] * 3, 2, laws, rows, flags, info) == [None, None, None]
    assert engine.build_regions(eng, CriticalRegion, [], 2, laws[:0], rows[:0], flags[:0], info[:0]) == []


@pytest.mark.parametrize('name', ['synthetic_30_6_40_s0', 'mpc_n5', 'transport_mplp', 'doc_portfolio'])
def test_tensor_buffers_assemble_like_numpy_buffers(name):
    """engine.build_regions has a whole-array numpy statement (host buffers) and a whole-array torch statement (device
    buffers: the solver's case) of the kept-index arithmetic and of the half-space gather; fed the same buffers - as CPU
    tensors here - they must build identical regions"""
    import torch
    from twin_binding import Twin
    from ppopt_b200 import engine
    from ppopt_b200.critical_region import CriticalRegion
    path = os.path.join(GOLDEN, name + '.npz')
    g = numpy.load(path)
    tw = Twin.from_npz(path)
    eng = types.SimpleNamespace(n=tw.n, t=tw.t, n_eq=tw.n_eq, m=tw.m, mi=tw.m - tw.n_eq)
    by_k = {}
    for r in golden_regions(g)[:200]:
        by_k.setdefault(len(r['active_set']) - tw.n_eq, []).append(r)
    for k_act, regs in by_k.items():
        bufs = [tw.emit(tw.masks([r['active_set'].tolist()])[0], margins=True) for r in regs]
        laws, rows, flags, info = (numpy.stack([b[i] for b in bufs]) for i in (1, 2, 3, 4))
        asets = [r['active_set'].tolist() for r in regs]
        a = engine.build_regions(eng, CriticalRegion, asets, k_act, laws, rows, flags, info)
        b = engine.build_regions(eng, CriticalRegion, asets, k_act, *[torch.from_numpy(x.copy()) for x in (laws, rows, flags, info)])
        assert len(a) == len(b) == len(regs)
        for x, y in zip(a, b):
            assert (x is None) == (y is None)
            if x is None:
                continue
            for fld in 'AbCdEf':
                assert numpy.array_equal(getattr(x, fld), getattr(y, fld)), (name, x.active_set, fld)
            assert (x.active_set, x.omega_set, x.lambda_set, x.regular_set) == (y.active_set, y.omega_set, y.lambda_set, y.regular_set)
            assert all(type(v) is int for lst in (y.omega_set, y.lambda_set, y.regular_set[0], y.regular_set[1]) for v in lst)
