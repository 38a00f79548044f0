"""Small end-to-end solves for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
sys.path.insert(0, '.')
from ppopt_b200 import engine
from ppopt_b200.mplp_program import load_presolved
for name, cap in (('factory_mpqp', None), ('transport_mplp', None), ('portfolio_analog', None), ('mpc_n3', None), ('rand_6_3_12_s1', 3),
                  ('synthetic_30_6_40_s0', 2), ('rand_wide_40_8_90_s5', 1)):
    sol = engine.solve(load_presolved(f'tests/golden/{name}.npz'), max_levels=cap)
    print(name, sol.total_candidates, len(sol.critical_regions), flush=True)
