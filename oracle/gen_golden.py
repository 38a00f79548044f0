"""Generates tests/golden/*.npz by running the UNMODIFIED reference (under the LP shim).

TEST INFRASTRUCTURE.  Run in the build container only:  python oracle/gen_golden.py [names...]

For each configuration in oracle/problems.py::CONFIGS it
  1. builds the reference's MPQP_Program / MPLP_Program (so the reference's own presolve runs),
  2. replays the level loop of mpqp_combinatorial.solve
     (/root/reference/src/ppopt/mp_solvers/mpqp_combinatorial.py:31-70) calling ONLY the
     reference's own functions, recording per candidate:
        bit0 is_full_rank, bit1 check_feasibility, bit2 check_optimality truthy, bit3 region built,
  3. for programs small enough, also calls the reference's solve() and asserts the replay found the
     same regions in the same order,
  4. stores the post-presolve program arrays, the per-level candidate lists + status bytes and all
     region matrices in one .npz.
Depth caps (levels evaluated) are applied to programs the reference cannot finish.
"""
import os
import sys
import time
import warnings

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')

import ref_harness  # noqa: E402

ppopt = ref_harness.load()
from ppopt.mplp_program import MPLP_Program  # noqa: E402
from ppopt.mpqp_program import MPQP_Program  # noqa: E402
from ppopt.mp_solvers import mpqp_combinatorial  # noqa: E402
from ppopt.mp_solvers.solver_utils import CombinationTester, generate_children_sets  # noqa: E402
from ppopt.utils.constraint_utilities import is_full_rank  # noqa: E402
from ppopt.utils.mpqp_utils import gen_cr_from_active_set  # noqa: E402

import problems  # noqa: E402

import multiprocess  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')

# name -> (max levels to replay or None for all, also run reference solve() and compare)
PLAN = {
    'factory_mpqp': (None, True),
    'transport_mplp': (None, True),
    'simple_mpqp_1d': (None, True),
    'simple_mplp': (None, True),
    'portfolio_analog': (None, True),
    'doc_portfolio': (None, True),
    'mpc_n3': (None, True),
    'mpc_n5': (None, True),
    'mpc_n7': (None, False),
    'mpc_n10': (6, False),
    'ctrl_alloc_n1': (None, True),
    'ctrl_alloc_n2': (None, True),
    'ctrl_alloc_n5': (4, False),
    'rand_6_3_12_s1': (None, True),
    'rand_5_3_10_s2': (None, True),
    'rand_lp_4_2_8_s3': (None, True),
    'synthetic_30_6_40_s0': (3, False),
    'rand_wide_40_8_90_s5': (2, False),
}


# programs whose next-level list after the cap is too slow to generate with the reference's CombinationTester
# (single-threaded, 20 min for mpc_n10 level 6 -> 7): no frontier_count is stored for them
NO_FRONTIER = {'mpc_n10'}


def build_reference_program(d):
    kw = {}
    if 'post_process' in d:
        kw['post_process'] = d['post_process']
    if d['kind'] == 'qp':
        return MPQP_Program(d['A'], d['b'], d['c'], d['H'], d['Q'], d['A_t'], d['b_t'], d['F'],
                            equality_indices=list(d['equality_indices']), **kw)
    return MPLP_Program(d['A'], d['b'], d['c'], d['H'], d['A_t'], d['b_t'], d['F'],
                        equality_indices=list(d['equality_indices']), **kw)


_PROG = None


def _eval_candidate(active_set):
    """status bits for one candidate, using the reference's functions only."""
    p = _PROG
    st = 0
    if is_full_rank(p.A, active_set):
        st |= 1
    if not p.check_feasibility(active_set):
        return st, None
    st |= 2
    region = None
    if p.check_optimality(active_set):
        st |= 4
        region = gen_cr_from_active_set(p, active_set)
        if region is not None:
            st |= 8
    return st, region


def replay(program, max_levels, procs, skip_frontier=False):
    global _PROG
    _PROG = program
    murder = CombinationTester()
    n_eq = len(program.equality_indices)
    max_depth = max(program.num_x(), program.num_t()) - n_eq
    to_check = generate_children_sets(program.equality_indices, program.num_constraints(), murder)
    levels = []
    regions = []
    pool = multiprocess.Pool(procs) if procs > 1 else None
    depth = max_depth if max_levels is None else min(max_depth, max_levels)
    t_levels = []
    for i in range(depth):
        t0 = time.time()
        if type(program) is MPLP_Program:
            cond = lambda child: child[-1] >= len(child) + program.num_constraints() - program.num_x()
            to_check = [c for c in to_check if not cond(c)]
        if pool is not None and len(to_check) > 64:
            outs = pool.map(_eval_candidate, to_check, chunksize=max(1, len(to_check) // (procs * 8)))
        else:
            outs = [_eval_candidate(c) for c in to_check]
        status = numpy.array([o[0] for o in outs], dtype=numpy.uint8)
        feasible = []
        for c, (st, reg) in zip(to_check, outs):
            if st & 2:
                feasible.append(c)
                if reg is not None:
                    regions.append(reg)
            else:
                murder.add_combo(c)
        levels.append((numpy.array(to_check, dtype=numpy.int32).reshape(len(to_check), -1), status))
        future = []
        if i + 1 != max_depth and not (skip_frontier and i + 1 == depth):
            for c in feasible:
                future.extend(generate_children_sets(c, program.num_constraints(), murder))
        t_levels.append(time.time() - t0)
        print(f'   level {i + 1}: {len(to_check)} candidates, {len(feasible)} feasible, '
              f'{int((status & 8 != 0).sum())} regions, {t_levels[-1]:.1f}s', flush=True)
        to_check = future
        if not to_check:
            break
    if pool is not None:
        pool.close()
        pool.join()
    base_status = 0
    complete = max_levels is None or max_levels >= max_depth
    eq = list(program.equality_indices)
    if is_full_rank(program.A, eq):
        base_status |= 1
    if program.check_feasibility(eq):
        base_status |= 2
        if program.check_optimality(eq):
            base_status |= 4
            region = gen_cr_from_active_set(program, eq)
            if region is not None and region.is_full_dimension():
                base_status |= 8
                if complete:
                    regions.append(region)
    return levels, regions, base_status, numpy.array(t_levels), (to_check if not complete else [])


def pack_regions(regions, out):
    out['n_regions'] = numpy.int64(len(regions))
    for i, r in enumerate(regions):
        out[f'r{i}_active_set'] = numpy.array(r.active_set, dtype=numpy.int32)
        out[f'r{i}_A'] = numpy.asarray(r.A)
        out[f'r{i}_b'] = numpy.asarray(r.b)
        out[f'r{i}_C'] = numpy.asarray(r.C)
        out[f'r{i}_d'] = numpy.asarray(r.d)
        out[f'r{i}_E'] = numpy.asarray(r.E)
        out[f'r{i}_f'] = numpy.asarray(r.f)
        out[f'r{i}_omega_set'] = numpy.array(r.omega_set, dtype=numpy.int32)
        out[f'r{i}_lambda_set'] = numpy.array(r.lambda_set, dtype=numpy.int32)
        out[f'r{i}_regular_pos'] = numpy.array(r.regular_set[0], dtype=numpy.int32)
        out[f'r{i}_regular_idx'] = numpy.array(r.regular_set[1], dtype=numpy.int32)


def generate(name, procs):
    max_levels, run_solve = PLAN[name]
    raw = problems.CONFIGS[name]()
    t0 = time.time()
    prog = build_reference_program(raw)
    print(f'[{name}] presolve {time.time() - t0:.2f}s: n={prog.num_x()} t={prog.num_t()} m={prog.num_constraints()} '
          f'q={prog.A_t.shape[0]} n_eq={len(prog.equality_indices)}', flush=True)
    out = {'kind': numpy.array(raw['kind']), 'n_eq': numpy.int64(len(prog.equality_indices))}
    for k in ('A', 'b', 'c', 'H', 'A_t', 'b_t', 'F'):
        out[k] = numpy.array(getattr(prog, k))
        out['raw_' + k] = numpy.array(raw[k], dtype=float)
    if raw['kind'] == 'qp':
        out['Q'] = numpy.array(prog.Q, dtype=float)
        out['raw_Q'] = numpy.array(raw['Q'], dtype=float)
    out['raw_equality_indices'] = numpy.array(raw['equality_indices'], dtype=numpy.int32)
    out['raw_post_process'] = numpy.bool_(raw.get('post_process', True))
    t0 = time.time()
    levels, regions, base_status, t_levels, frontier = replay(prog, max_levels, procs, name in NO_FRONTIER)
    t_replay = time.time() - t0
    out['n_levels'] = numpy.int64(len(levels))
    out['level_cap'] = numpy.int64(-1 if max_levels is None else max_levels)
    out['base_status'] = numpy.uint8(base_status)
    out['replay_seconds'] = t_levels
    out['replay_procs'] = numpy.int64(procs)
    for i, (cands, status) in enumerate(levels):
        out[f'level{i}_candidates'] = cands
        out[f'level{i}_status'] = status
    if len(frontier):
        out['frontier_count'] = numpy.int64(len(frontier))
    pack_regions(regions, out)
    if run_solve:
        t0 = time.time()
        sol = mpqp_combinatorial.solve(prog)
        out['reference_solve_seconds'] = numpy.float64(time.time() - t0)
        a = [list(r.active_set) for r in sol.critical_regions]
        bb = [list(r.active_set) for r in regions]
        assert a == bb, (name, a, bb)
        for r1, r2 in zip(sol.critical_regions, regions):
            for fld in 'AbCdEf':
                assert numpy.array_equal(getattr(r1, fld), getattr(r2, fld)), (name, fld)
    n_c = sum(len(l[1]) for l in levels)
    print(f'[{name}] {n_c} candidates, {len(regions)} regions, replay {t_replay:.1f}s '
          f'(procs={procs}) base_status={base_status}', flush=True)
    numpy.savez_compressed(os.path.join(OUT, name + '.npz'), **out)


if __name__ == '__main__':
    names = sys.argv[1:] or list(PLAN)
    procs = int(os.environ.get('GOLDEN_PROCS', '8'))
    os.makedirs(OUT, exist_ok=True)
    for nm in names:
        generate(nm, procs)
