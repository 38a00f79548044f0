import sys, time, os
sys.path.insert(0, '.')
import numpy, torch
from ppopt_b200 import engine
from ppopt_b200.mplp_program import load_presolved
for name, cap in (('synthetic_30_6_40_s0', 3), ('synthetic_30_6_40_s0', 4), ('mpc_n7', None), ('mpc_n10', 6), ('ctrl_alloc_n5', 5)):
    prog = load_presolved(f'tests/golden/{name}.npz')
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sol = engine.solve(prog, max_levels=cap)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(name, cap, 'candidates', sol.total_candidates, 'regions', len(sol.critical_regions), f'{dt:.3f}s', f'{sol.total_candidates/dt:.3e} cand/s')
    for s in sol.level_stats: print('   ', s)
    print('   ', sol.engine_counters)
print('fp64 peak TFLOP/s', engine.measure_fp64_peak())
