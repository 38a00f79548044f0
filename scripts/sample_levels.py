"""Draws a random sample of candidates (plus every candidate the engine calls a region) from the GPU engine's OWN level
arrays at depths the reference cannot enumerate, so that the unmodified reference can judge them on the CPU afterwards
(oracle/gen_sampled_golden.py -> tests/golden/sampled/*.npz).  Run on the GPU box:

    python scripts/sample_levels.py synthetic_30_6_40_s0 5 20000 [first level to sample]  ->  gpurun_out/sample_<name>.npz

For every level beyond the fixture's own cap the file holds: the sampled candidates (inequality-row bitmasks), their
position in the level, the engine's status byte at sampling time (informational; the test re-evaluates), the level size.
"""
import os
import sys

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ppopt_b200 import engine  # noqa: E402
from ppopt_b200._lib import ST_FEAS, ST_OPT  # noqa: E402
from ppopt_b200.mplp_program import load_presolved  # noqa: E402


def main(name, depth, n_sample, first=None, seed=2026):
    path = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
    g = numpy.load(path)
    cap = int(g['level_cap']) if first is None else first - 1
    eng = engine.Engine(engine.program_arrays(load_presolved(path)))
    gen = torch.Generator(device='cpu').manual_seed(seed)
    out = {'name': numpy.array(name), 'n_eq': numpy.int64(eng.n_eq), 'words': numpy.int64(eng.W), 'golden_cap': numpy.int64(cap)}
    masks = eng.root_level()
    levels = []
    for lvl in range(min(depth, eng.max_depth)):
        n = masks.shape[0]
        if n == 0:
            break
        k_act = lvl + 1
        status = eng.level_eval(masks, k_act)
        opt_idx = eng.select(status, ST_OPT, ST_OPT)
        if opt_idx.shape[0]:
            eng.emit(masks, opt_idx, k_act, status)   # sets the region bit
        if cap >= 0 and lvl >= cap:
            pick = torch.randperm(n, generator=gen)[:min(n_sample, n)].sort().values.to(eng.tdev)
            keep = torch.unique(torch.cat([pick, opt_idx]))   # sorted
            out[f'level{lvl}_pos'] = keep.cpu().numpy()
            out[f'level{lvl}_masks'] = masks[keep].cpu().numpy()
            out[f'level{lvl}_status'] = status[keep].cpu().numpy()
            out[f'level{lvl}_size'] = numpy.int64(n)
            levels.append(lvl)
            st = out[f'level{lvl}_status']
            print(f'level {lvl + 1}: {n} candidates, sampled {len(st)} (regions {int((st & 8 != 0).sum())}, '
                  f'infeasible {int((st & 2 == 0).sum())})', flush=True)
        feas_idx = eng.select(status, ST_FEAS, ST_FEAS)
        if lvl + 1 < min(depth, eng.max_depth):
            masks = eng.children(masks, feas_idx, k_act)
    out['levels'] = numpy.array(levels, dtype=numpy.int64)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    dst = os.path.join(ROOT, 'gpurun_out', f'sample_{name}.npz')
    numpy.savez_compressed(dst, **out)
    print('wrote', dst)
    eng.close()


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else None)
