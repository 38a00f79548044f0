"""Golden vectors for the connected-graph algorithm of Oberdieck et al. (second half).

TEST INFRASTRUCTURE.  Build container only:  python oracle/gen_graph_oberdieck_golden.py [names...]

Runs the UNMODIFIED reference driver ppopt.mp_solvers.mpqp_graph.solve(program, initial_active_sets=[seed])
(/root/reference/src/ppopt/mp_solvers/mpqp_graph.py:38-103) under the LP shim.  The seed is passed explicitly - the
reference's own default, program.sample_theta_space(), solves QPs with a solver that is not in this image - and is the
first region of the program's combinatorial golden.  While the driver runs, is_full_rank / check_feasibility /
check_optimality / CriticalRegion.is_full_dimension are wrapped (not changed) to record the order in which active sets
are attempted and what was decided for each.  Stored: that trace, and all regions in the reference's order.
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
import ppopt.mp_solvers.mpqp_graph as ref_graph  # noqa: E402

NAMES = ['factory_mpqp', 'mpc_n3', 'mpc_n5', 'ctrl_alloc_n1', 'rand_6_3_12_s1', 'rand_5_3_10_s2']


def generate(name):
    prog = gg.build_reference_program(problems.CONFIGS[name]())
    g = numpy.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
    seed = g['r0_active_set'].tolist()
    order = []
    real_rank = ref_graph.is_full_rank

    def spy_rank(A, idx=None):
        order.append(tuple(idx) if idx is not None else None)
        return real_rank(A, idx)
    ref_graph.is_full_rank = spy_rank
    t0 = time.time()
    try:
        sol = ref_graph.solve(prog, initial_active_sets=[seed])
    finally:
        ref_graph.is_full_rank = real_rank
    regions = list(sol.critical_regions)
    out = {'seed': numpy.array(seed, dtype=numpy.int32), 'n_attempted': numpy.int64(len(order))}
    width = max(len(k) for k in order)
    att = numpy.full((len(order), width), -1, dtype=numpy.int32)
    for i, k in enumerate(order):
        att[i, :len(k)] = k
    out['attempted'] = att
    gg.pack_regions(regions, out)
    dst = os.path.join(ROOT, 'tests', 'golden', 'graph_oberdieck')
    os.makedirs(dst, exist_ok=True)
    numpy.savez_compressed(os.path.join(dst, name + '.npz'), **out)
    print(f'[{name}] attempted {len(order)}, regions {len(regions)}, {time.time() - t0:.1f}s', flush=True)


if __name__ == '__main__':
    for nm in (sys.argv[1:] or NAMES):
        generate(nm)
