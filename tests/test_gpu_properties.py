"""GPU tests beyond the golden vectors: size-independent properties at the headline size, K6 against the reference's
set arithmetic under arbitrary feasibility patterns, edge cases, slice/whole equivalence."""
import os
import sys

import numpy
import pytest

from conftest import GOLDEN, ROOT
from parity import masks_to_lists

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(ROOT, 'oracle'))


def _engine(name):
    from ppopt_b200 import engine
    from ppopt_b200.mplp_program import load_presolved
    prog = load_presolved(os.path.join(GOLDEN, name + '.npz'))
    return engine, prog, engine.Engine(engine.program_arrays(prog))


def _lex_sorted(lists):
    return all(a < b for a, b in zip(lists, lists[1:]))


@pytest.mark.parametrize('name', ['factory_mpqp', 'synthetic_30_6_40_s0', 'transport_mplp'])
def test_k6_matches_reference_set_arithmetic_for_arbitrary_feasibility(name):
    """next-level generation + superset pruning against CombinationTester/generate_children_sets
    (solver_utils.py:29-55,154-166) when 'feasible' is an arbitrary pseudo-random predicate (W = 1 and W = 2 masks,
    mpQP and mpLP cardinality filter)."""
    import torch
    from ppopt_b200.mp_solvers.solver_utils import CombinationTester, generate_children_sets
    engine, prog, eng = _engine(name)
    is_lp = type(prog).__name__ == 'MPLP_Program'
    m, n, ne = prog.num_constraints(), prog.num_x(), len(prog.equality_indices)
    rng = numpy.random.default_rng(5)
    tester = CombinationTester()
    to_check = generate_children_sets(prog.equality_indices, m, tester)
    masks = eng.root_level()
    for level in range(1, 4):
        if is_lp:
            to_check = [c for c in to_check if not (c[-1] >= len(c) + m - n)]
        assert masks_to_lists(masks.cpu().numpy(), ne) == to_check, f'level {level}'
        if len(to_check) > 3000:  # thin the level (keeps the reference-side python loops short)
            keep_p = 3000.0 / len(to_check)
        else:
            keep_p = 0.8
        feas = rng.random(len(to_check)) < keep_p
        status = torch.from_numpy(numpy.where(feas, 3, 1).astype(numpy.uint8)).to(eng.tdev)
        feasible = []
        for c, ok in zip(to_check, feas):
            if ok:
                feasible.append(c)
            else:
                tester.add_combo(c)
        nxt = []
        for c in feasible:
            nxt.extend(generate_children_sets(c, m, tester))
        idx = eng.select(status, 2, 2)
        assert idx.cpu().tolist() == numpy.nonzero(feas)[0].tolist()
        masks = eng.children(masks, idx, level)
        to_check = nxt
        if not to_check:
            assert masks.shape[0] == 0
            break
    eng.close()


def test_headline_size_properties():
    """synthetic 100x30x6, levels 1..4 (3.94e6 candidates): sortedness, cardinality, idempotence, agreement with the CPU
    checker and with the reference-algorithm oracle on random samples, closure of the pruning rule."""
    import torch
    import ppopt_oracle as oracle
    import twin_binding
    engine, prog, eng = _engine('synthetic_30_6_40_s0')
    path = os.path.join(GOLDEN, 'synthetic_30_6_40_s0.npz')
    sol = engine.solve(prog, max_levels=4, engine=eng, collect_status=True)
    sizes = [s['candidates'] for s in sol.level_stats]
    assert sizes[:2] == [100, 4950] and sizes[2] == 158760 and sizes[3] == 3776565
    tw = twin_binding.Twin.from_npz(path)
    P = oracle.Program.from_npz(path)
    rng = numpy.random.default_rng(11)
    feas_prev = None
    for lv, (masks, status) in enumerate(sol.level_status):
        m64 = masks.view(numpy.uint64).reshape(len(masks), -1)
        pc = numpy.unpackbits(m64.view(numpy.uint8), axis=1).sum(axis=1)
        assert numpy.all(pc == lv + 1)
        # lexicographic order of index lists == descending order of the bit-reversed key (checked on a sample of pairs)
        pick = numpy.sort(rng.choice(len(masks) - 1, size=min(4000, len(masks) - 1), replace=False))
        a = masks_to_lists(masks[pick], 0)
        b = masks_to_lists(masks[pick + 1], 0)
        assert all(x < y for x, y in zip(a, b)), f'level {lv + 1} not sorted'
        # idempotence: re-evaluating the level reproduces the status bytes
        again = eng.level_eval(torch.from_numpy(masks).to(eng.tdev), lv + 1).cpu().numpy()
        assert numpy.array_equal(again & 7, status & 7)
        # CPU checker on a sample (every decision bit), oracle (reference algorithm, HiGHS) on a smaller one
        pick = rng.choice(len(masks), size=min(3000, len(masks)), replace=False)
        assert numpy.array_equal(tw.eval(masks[pick]) & 11, status[pick] & 11), f'level {lv + 1} vs CPU checker'
        for i in rng.choice(len(masks), size=40, replace=False):
            aset = masks_to_lists(masks[i:i + 1], 0)[0]
            assert oracle.evaluate_candidate(P, aset) & 11 == int(status[i]) & 11, (lv + 1, aset)
        # pruning closure: every candidate's (k-1)-subsets are feasible members of the previous level
        if feas_prev is not None:
            for i in rng.choice(len(masks), size=300, replace=False):
                aset = masks_to_lists(masks[i:i + 1], 0)[0]
                for j in range(len(aset)):
                    assert tuple(aset[:j] + aset[j + 1:]) in feas_prev
        if lv < 3:
            feas_prev = set(map(tuple, masks_to_lists(masks[(status & 2) != 0], 0)))
    assert not any(sol.engine_counters[k] for k in ('numeric', 'border'))
    eng.close()


def test_k6_split_between_ranks_equals_single_pass():
    """the multi-GPU path counts the children of disjoint parent ranges on different ranks and sums the per-parent
    results (engine._children_sharded): any split must reproduce the single-pass child array"""
    import torch

    engine, prog, eng = _engine('synthetic_30_6_40_s0')
    g = numpy.load(os.path.join(GOLDEN, 'synthetic_30_6_40_s0.npz'))
    masks = eng.masks_from_lists(g['level1_candidates'].tolist())
    status = eng.level_eval(masks, 2)
    rng = numpy.random.default_rng(3)
    keep = torch.from_numpy(rng.random(masks.shape[0]) < 0.9).to(status.device)
    feas_idx = torch.nonzero(((status & 2) != 0) & keep).flatten().contiguous()
    whole = eng.children(masks, feas_idx, 2)
    from ppopt_b200 import _lib, sharding
    import ctypes
    nf = feas_idx.shape[0]
    for world in (2, 3):
        feas_masks = eng.empty((nf, eng.W), torch.int64)
        ws_bytes = eng.lib.ppgpu_scan_workspace_bytes(nf)
        ws = eng.empty(((ws_bytes + 7) // 8,), torch.int64)
        _lib.check(eng.lib.ppgpu_children_prepare(eng.h, masks.data_ptr(), feas_idx.data_ptr(), nf, feas_masks.data_ptr(),
                                                  ws.data_ptr(), ws_bytes, eng._stream()), 'prepare')
        survive_sum = torch.zeros((nf, eng.W), dtype=torch.int64, device=eng.tdev)
        counts_sum = torch.zeros((nf + 1,), dtype=torch.int64, device=eng.tdev)
        for rank in range(world):
            survive = torch.zeros_like(survive_sum)
            counts = torch.zeros_like(counts_sum)
            bounds = [(nf * rank // world, nf * (rank + 1) // world)] if nf < 65536 * world else sharding.chunks(nf, rank, world)
            for lo, hi in bounds:
                _lib.check(eng.lib.ppgpu_children_count_range(eng.h, feas_masks.data_ptr(), nf, 2, survive.data_ptr(),
                                                              counts.data_ptr(), lo, hi, ws.data_ptr(), ws_bytes,
                                                              eng._stream()), 'count_range')
            survive_sum += survive
            counts_sum += counts
        tot = ctypes.c_int64(0)
        _lib.check(eng.lib.ppgpu_children_scan(eng.h, counts_sum.data_ptr(), nf, ctypes.byref(tot), ws.data_ptr(), ws_bytes,
                                               eng._stream()), 'scan')
        assert tot.value == whole.shape[0]
        out = eng.empty((tot.value, eng.W), torch.int64)
        _lib.check(eng.lib.ppgpu_children_write(eng.h, feas_masks.data_ptr(), survive_sum.data_ptr(), counts_sum.data_ptr(),
                                                nf, out.data_ptr(), eng._stream()), 'write')
        assert torch.equal(out, whole)
    eng.close()


def test_slice_evaluation_equals_whole_level():
    """the multi-GPU path evaluates [lo, hi) slices of a level: statuses must not depend on the slicing"""
    import torch
    engine, prog, eng = _engine('mpc_n5')
    g = numpy.load(os.path.join(GOLDEN, 'mpc_n5.npz'))
    cands = g['level3_candidates'].tolist()
    masks = eng.masks_from_lists(cands)
    whole = eng.level_eval(masks, 4).cpu().numpy()
    parts = torch.zeros(len(cands), dtype=torch.uint8, device=eng.tdev)
    cuts = [0, 1, 17, 500, 501, len(cands)]
    for lo, hi in zip(cuts, cuts[1:]):
        eng.level_eval(masks, 4, parts, 7, lo, hi)
    assert numpy.array_equal(parts.cpu().numpy(), whole)
    assert numpy.array_equal(whole & 3, g['level3_status'] & 3)
    eng.close()


def test_edge_cases():
    import torch
    engine, prog, eng = _engine('factory_mpqp')
    # empty level
    empty = eng.empty((0, eng.W), torch.int64)
    assert eng.level_eval(empty, 1).shape[0] == 0
    assert eng.select(torch.zeros(0, dtype=torch.uint8, device=eng.tdev), 2, 2).shape[0] == 0
    assert eng.children(empty, eng.empty((0,), torch.int64), 1).shape[0] == 0
    # the empty active set is full rank by definition (constraint_utilities.py:231-232) and feasible
    st = eng.level_eval(torch.zeros((1, eng.W), dtype=torch.int64, device=eng.tdev), 0).cpu().numpy()
    assert st[0] & 3 == 3
    # more active rows than variables can never be full rank
    st = eng.level_eval(eng.masks_from_lists([[0, 1, 2, 3, 4]]), 5, stages=1).cpu().numpy()
    assert st[0] & 1 == 0
    eng.close()
    # reference test tests/other_tests/test_mpqp_utils.py:5-12 on the same fixture: [] and [0] feasible
    from ppopt_b200.mp_solvers.mpqp_combinatorial import check_child_feasibility
    from ppopt_b200.mp_solvers.solver_utils import CombinationTester
    t = CombinationTester()
    out = check_child_feasibility(prog, [[], [0], [2, 3], [0, 1, 2, 3, 4]], t)
    assert out == [[], [0], [2, 3]] and t.combos == {(0, 1, 2, 3, 4)}


def test_status_bytes_stay_on_device_and_counters_move():
    engine, prog, eng = _engine('mpc_n3')
    sol = engine.solve(prog, engine=eng)
    c = sol.engine_counters
    assert c['k2_lps'] > 0 and c['k2_pivots'] > c['k2_lps'] and c['k4_lps'] > 0 and sol.gpu_launches > 10
    assert len(sol.critical_regions) == 7
    eng.close()


@pytest.mark.parametrize('name', ['mpc_n5', 'rand_6_3_12_s1', 'transport_mplp', 'portfolio_analog', 'ctrl_alloc_n2'])
def test_relaxation_certificates_never_change_a_decision(name):
    """K2a only certifies feasibility (exactly re-verified point); with it disabled (stage bit 8) the simplex alone must
    produce the same status bytes on every golden candidate"""
    engine, prog, eng = _engine(name)
    g = numpy.load(os.path.join(GOLDEN, name + '.npz'))
    ne = int(g['n_eq'])
    for lv in range(int(g['n_levels'])):
        cands = g[f'level{lv}_candidates'].tolist()
        masks = eng.masks_from_lists(cands)
        k_act = len(cands[0]) - ne
        with_k2a = eng.level_eval(masks, k_act, stages=3).cpu().numpy()
        without = eng.level_eval(masks, k_act, stages=3 | 8).cpu().numpy()
        assert numpy.array_equal(with_k2a & 3, without & 3), f'{name} level {lv + 1}'
        assert numpy.array_equal(with_k2a & 3, g[f'level{lv}_status'] & 3)
    c = eng.counters()
    assert c['k2a_tried'] > 0 and c['k2a_certified'] <= c['k2a_tried']
    eng.close()


@pytest.mark.parametrize('name', ['ctrl_alloc_n5', 'ctrl_alloc_n2', 'rand_6_3_12_s1'])
def test_redundancy_flags_match_cpu_checker(name):
    """K5's per-row keep/drop decisions (redundancy LPs with one row forced to equality) against the CPU checker, whose
    LP margins are verified against HiGHS in tests/test_oracle.py.  Thin regions (radius ~1e-5, theta ranges ~1e2) make
    this the most demanding accuracy test of the register simplex: a row may only differ if its margin sits within 2e-8
    of the keep threshold."""
    import torch
    import twin_binding
    engine, prog, eng = _engine(name)
    path = os.path.join(GOLDEN, name + '.npz')
    g = numpy.load(path)
    tw = twin_binding.Twin.from_npz(path)
    ne = int(g['n_eq'])
    by_k = {}
    for i in range(int(g['n_regions'])):
        a = g[f'r{i}_active_set'].tolist()
        by_k.setdefault(len(a) - ne, []).append(a)
    n_rows = n_diff = 0
    for k_act, asets in by_k.items():
        masks = eng.masks_from_lists(asets)
        status = torch.full((len(asets),), 7, dtype=torch.uint8, device=eng.tdev)
        sel = torch.arange(len(asets), dtype=torch.int64, device=eng.tdev)
        laws, rows, flags, info = [x.cpu().numpy() for x in eng.emit(masks, sel, k_act, status)]
        for si, a in enumerate(asets):
            rc, tl, tr, tf, ti, mg = tw.emit(tw.masks([a])[0], margins=True)
            assert (info[si, 0] == 1.0) == (rc == 1)
            if rc != 1 or eng.t == 1:
                continue
            assert numpy.array_equal(flags[si] & 1, tf & 1)
            for r in numpy.nonzero(tf & 1)[0]:
                n_rows += 1
                if (flags[si][r] & 2) != (tf[r] & 2):
                    n_diff += 1
                    assert abs(mg[r] + 1e-9) < 2e-8, (name, a, int(r), float(mg[r]))
    assert n_diff <= 0.01 * n_rows
    eng.close()


def test_constructor_presolve_matches_reference():
    """MPQP_Program / MPLP_Program built from RAW data (constructor presolve, redundancy LPs batched on
    the GPU) must end with exactly the arrays the reference constructor produced, for every golden program"""
    from conftest import golden_names
    from ppopt_b200 import MPLP_Program, MPQP_Program
    for name in golden_names():
        g = numpy.load(os.path.join(GOLDEN, name + '.npz'))
        kw = dict(equality_indices=g['raw_equality_indices'].tolist(), post_process=bool(g['raw_post_process']))
        if str(g['kind']) == 'qp':
            p = MPQP_Program(g['raw_A'], g['raw_b'], g['raw_c'], g['raw_H'], g['raw_Q'], g['raw_A_t'], g['raw_b_t'], g['raw_F'], **kw)
        else:
            p = MPLP_Program(g['raw_A'], g['raw_b'], g['raw_c'], g['raw_H'], g['raw_A_t'], g['raw_b_t'], g['raw_F'], **kw)
        assert len(p.equality_indices) == int(g['n_eq']), name
        for fld in ('A', 'b', 'F', 'A_t', 'b_t'):
            assert numpy.array_equal(getattr(p, fld), g[fld]), (name, fld, getattr(p, fld).shape, g[fld].shape)


def test_raw_to_solution_end_to_end():
    """raw data -> constructor presolve -> solve_mpqp: the whole user path without any reference code"""
    import problems
    from ppopt_b200 import MPQP_Program, mpqp_algorithm, solve_mpqp
    d = problems.mpc_double_integrator(5)
    prog = MPQP_Program(d['A'], d['b'], d['c'], d['H'], d['Q'], d['A_t'], d['b_t'], d['F'], equality_indices=d['equality_indices'])
    sol = solve_mpqp(prog, mpqp_algorithm.combinatorial)
    g = numpy.load(os.path.join(GOLDEN, 'mpc_n5.npz'))
    assert [list(r.active_set) for r in sol.critical_regions] == [g[f'r{i}_active_set'].tolist() for i in range(int(g['n_regions']))]


def test_wide_templates_emission_matches_cpu_checker():
    """rand_wide_40_8_90_s5 (n'=40 > 32, t=8 > 6, 186 rows > 128) drives the widest kernel instantiations (K1 NP=64, K2
    8 warps x 64 columns, K3/K4/K5 8 rows per lane x 16 columns).  No region exists at levels 1-2, so K5 is checked on
    arbitrary feasible candidates: laws and normalised rows against the CPU checker."""
    import torch
    import twin_binding
    engine, prog, eng = _engine('rand_wide_40_8_90_s5')
    path = os.path.join(GOLDEN, 'rand_wide_40_8_90_s5.npz')
    g = numpy.load(path)
    tw = twin_binding.Twin.from_npz(path)
    cands, st = g['level1_candidates'], g['level1_status']
    pick = numpy.nonzero(st & 2)[0][::701][:16]
    asets = [cands[i].tolist() for i in pick]
    masks = eng.masks_from_lists(asets)
    status = torch.full((len(asets),), 7, dtype=torch.uint8, device=eng.tdev)
    sel = torch.arange(len(asets), dtype=torch.int64, device=eng.tdev)
    laws, rows, flags, info = [x.cpu().numpy() for x in eng.emit(masks, sel, 2, status)]
    for si, a in enumerate(asets):
        rc, tl, tr, tf, ti = tw.emit(tw.masks([a])[0])
        assert (info[si, 0] == 1.0) == (rc == 1)
        assert numpy.allclose(laws[si], tl, rtol=1e-8, atol=1e-8 * max(1.0, numpy.abs(tl).max()))
        assert numpy.array_equal(flags[si] & 1, tf & 1)
        assert numpy.allclose(rows[si], tr, rtol=1e-8, atol=1e-8 * max(1.0, numpy.abs(tr).max()))
    eng.close()
