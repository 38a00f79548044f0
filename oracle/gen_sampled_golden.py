"""Sampled golden vectors for enumeration depths the reference cannot finish.

TEST INFRASTRUCTURE.  Build container only:  python oracle/gen_sampled_golden.py <name> [more names]

Input:  gpurun_out/sample_<name>.npz written on the GPU box by scripts/sample_levels.py - a random sample of the
        candidates of the engine's own deep level arrays plus every candidate the engine called a region there.
Output: tests/golden/sampled/<name>.npz - for each of those candidates the verdict of the UNMODIFIED reference
        (oracle/gen_golden.py::_eval_candidate: is_full_rank / check_feasibility / check_optimality /
        gen_cr_from_active_set, same status bits as the level goldens) and all matrices of every region it builds.
The GPU test (tests/test_gpu_sampled.py) re-evaluates exactly these candidates through the C ABI, checks that each one is
a member of the level the engine enumerates, and compares status bits and region matrices.
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
import multiprocess  # noqa: E402

import problems  # noqa: E402


def _lists(masks, n_eq):
    m = numpy.ascontiguousarray(masks).view(numpy.uint64)
    m = m.reshape(m.shape[0], -1)
    bits = numpy.unpackbits(m.view(numpy.uint8).reshape(m.shape[0], -1), axis=1, bitorder='little')
    eq = list(range(n_eq))
    return [eq + (numpy.nonzero(r)[0] + n_eq).tolist() for r in bits]


def pack_regions_compact(regions, out):
    """thousands of regions: one concatenated array per field (+ row offsets) instead of eleven zip members per region;
    tests/parity.py::golden_regions reads both layouts"""
    def cat(get, dtype=numpy.float64):
        parts = [numpy.asarray(get(r), dtype=dtype).reshape(-1) for r in regions]
        off = numpy.zeros(len(parts) + 1, dtype=numpy.int64)
        off[1:] = numpy.cumsum([len(x) for x in parts])
        return (numpy.concatenate(parts) if parts else numpy.zeros(0, dtype=dtype)), off
    out['n_regions'] = numpy.int64(len(regions))
    out['packed'] = numpy.bool_(True)
    for key, get, dt in (('active_set', lambda r: r.active_set, numpy.int32), ('A', lambda r: r.A, numpy.float64),
                         ('b', lambda r: r.b, numpy.float64), ('C', lambda r: r.C, numpy.float64),
                         ('d', lambda r: r.d, numpy.float64), ('E', lambda r: r.E, numpy.float64),
                         ('f', lambda r: r.f, numpy.float64), ('omega_set', lambda r: r.omega_set, numpy.int32),
                         ('lambda_set', lambda r: r.lambda_set, numpy.int32),
                         ('regular_pos', lambda r: r.regular_set[0], numpy.int32),
                         ('regular_idx', lambda r: r.regular_set[1], numpy.int32)):
        out['pk_' + key], out['pk_' + key + '_off'] = cat(get, dt)
    out['pk_E_int'] = numpy.array([numpy.asarray(r.E).dtype.kind == 'i' for r in regions], dtype=numpy.bool_)


def generate(name, procs):
    src = numpy.load(os.path.join(ROOT, 'gpurun_out', f'sample_{name}.npz'))
    prog = gg.build_reference_program(problems.CONFIGS[name]())
    n_eq = len(prog.equality_indices)
    assert n_eq == int(src['n_eq'])
    gg._PROG = prog
    pool = multiprocess.Pool(procs)
    out = {'levels': src['levels'], 'n_eq': numpy.int64(n_eq)}
    regions = []
    region_level = []
    for lvl in src['levels'].tolist():
        cands = _lists(src[f'level{lvl}_masks'], n_eq)
        t0 = time.time()
        res = pool.map(gg._eval_candidate, cands, chunksize=max(1, len(cands) // (procs * 16)))
        status = numpy.array([r[0] for r in res], dtype=numpy.uint8)
        for st, reg in res:
            if reg is not None:
                regions.append(reg)
                region_level.append(lvl)
        out[f'level{lvl}_candidates'] = numpy.array(cands, dtype=numpy.int32)
        out[f'level{lvl}_status'] = status
        out[f'level{lvl}_pos'] = src[f'level{lvl}_pos']
        out[f'level{lvl}_size'] = src[f'level{lvl}_size']
        gpu = src[f'level{lvl}_status']
        diff = int(((gpu & 11) != (status & 11)).sum())
        print(f'[{name}] level {lvl + 1}: {len(cands)} sampled of {int(src[f"level{lvl}_size"])}, feasible '
              f'{int((status & 2 != 0).sum())}, regions {int((status & 8 != 0).sum())}, '
              f'{time.time() - t0:.1f}s; engine-vs-reference status differences at sampling time: {diff}', flush=True)
    pool.close()
    pool.join()
    pack_regions_compact(regions, out)
    out['region_level'] = numpy.array(region_level, dtype=numpy.int64)
    dst = os.path.join(ROOT, 'tests', 'golden', 'sampled')
    os.makedirs(dst, exist_ok=True)
    numpy.savez_compressed(os.path.join(dst, name + '.npz'), **out)


if __name__ == '__main__':
    procs = int(os.environ.get('GOLDEN_PROCS', '8'))
    for nm in sys.argv[1:]:
        generate(nm, procs)
