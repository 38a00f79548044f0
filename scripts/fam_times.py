"""per-kernel-family milliseconds for one capped solve of a golden program:  python scripts/fam_times.py [name] [levels] [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ppopt_b200 import engine  # noqa: E402
from ppopt_b200.mplp_program import load_presolved  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'synthetic_30_6_40_s0'
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
prog = load_presolved(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
eng = engine.Engine(engine.program_arrays(prog))
for rep in range(reps):
    eng.counters(reset=True)
    eng.profile(True)
    eng.profile_read(reset=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sol = engine.solve(prog, max_levels=levels, engine=eng, materialize=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    fam = eng.profile_read(reset=True)
    print(f'rep {rep}: {sol.total_candidates} candidates, {len(sol.critical_regions)} regions, {dt * 1e3:.1f} ms, '
          f'{sol.total_candidates / dt:.3e} cand/s')
    print('   ', {k: round(v['ms'], 2) for k, v in fam.items() if v['launches']})
    c = sol.engine_counters
    print('   ', {k: c[k] for k in c if c[k]})
    print('   ', [(s['candidates'], s['feasible'], s['optimal'], s['regions']) for s in sol.level_stats])
eng.close()
