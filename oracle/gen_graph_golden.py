"""Golden vectors for the combinatorial connected-graph algorithm.

TEST INFRASTRUCTURE.  Build container only:  python oracle/gen_graph_golden.py [names...]

The reference's driver, mp_solvers/mpqp_combi_graph.py:69-145, seeds itself with ``program.sample_theta_space(1)``, which
needs a QP solver that is not in this image.  So the driver loop is replayed here around the UNMODIFIED reference's own
functions - is_full_rank, feasability_check (:48-66), gen_cr_from_active_set, sorted_tuple / remove_i / add_i - from an
explicit seed (the first region of the program's combinatorial golden; the reference accepts explicit seeds through
combinatorial_graph_initialization, :10-29).  Stored: every visited active set with its three decisions, and all regions.
"""
import os
import sys
import time

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import gen_golden as gg  # noqa: E402  (loads the reference under the shim)
import problems  # noqa: E402
from ppopt.mplp_program import MPLP_Program  # noqa: E402
from ppopt.mp_solvers.mpqp_combi_graph import add_i, feasability_check, remove_i, sorted_tuple  # noqa: E402
from ppopt.utils.constraint_utilities import is_full_rank  # noqa: E402
from ppopt.utils.mpqp_utils import gen_cr_from_active_set  # noqa: E402

NAMES = ['factory_mpqp', 'mpc_n3', 'mpc_n5', 'ctrl_alloc_n1', 'rand_6_3_12_s1', 'rand_5_3_10_s2']


def replay(program, seed):
    E, S = {sorted_tuple(seed)}, {sorted_tuple(seed)}
    trace, regions = {}, []
    eqs = set(program.equality_indices)
    while S:
        A = S.pop()
        rank_ok = bool(is_full_rank(program.A, list(A)))
        nonempty = bool(rank_ok and feasability_check(program, A))
        region = None
        if nonempty and not (type(program) is MPLP_Program and len(A) != program.num_x()):
            region = gen_cr_from_active_set(program, list(A))
            if region is not None:
                regions.append(region)
        trace[A] = (rank_ok, nonempty, region is not None)
        if (not rank_ok) or nonempty:
            for i in A:
                if i not in eqs:
                    t = remove_i(A, i)
                    if t not in E:
                        S.add(t); E.add(t)
        if nonempty:
            for i in range(program.num_constraints()):
                if i not in A:
                    t = add_i(A, i)
                    if t not in E:
                        S.add(t); E.add(t)
    return trace, regions


def generate(name):
    prog = gg.build_reference_program(problems.CONFIGS[name]())
    g = numpy.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
    seed = g['r0_active_set'].tolist()
    t0 = time.time()
    trace, regions = replay(prog, seed)
    keys = sorted(trace, key=lambda a: (len(a), a))
    out = {'seed': numpy.array(seed, dtype=numpy.int32), 'n_visited': numpy.int64(len(keys))}
    width = max(len(k) for k in keys)
    vis = numpy.full((len(keys), width), -1, dtype=numpy.int32)
    for i, k in enumerate(keys):
        vis[i, :len(k)] = k
    out['visited'] = vis
    out['decisions'] = numpy.array([[int(x) for x in trace[k]] for k in keys], dtype=numpy.uint8)
    regions.sort(key=lambda r: (len(r.active_set), list(r.active_set)))
    gg.pack_regions(regions, out)
    dst = os.path.join(ROOT, 'tests', 'golden', 'graph')
    os.makedirs(dst, exist_ok=True)
    numpy.savez_compressed(os.path.join(dst, name + '.npz'), **out)
    print(f'[{name}] visited {len(keys)}, full rank {int(out["decisions"][:, 0].sum())}, non-empty '
          f'{int(out["decisions"][:, 1].sum())}, regions {len(regions)}, {time.time() - t0:.1f}s', flush=True)


if __name__ == '__main__':
    for nm in (sys.argv[1:] or NAMES):
        generate(nm)
