import sys; sys.path.insert(0, '.')
import torch
from ppopt_b200 import engine
from ppopt_b200.mplp_program import load_presolved
name, L = sys.argv[1], int(sys.argv[2])
prog = load_presolved(f'tests/golden/{name}.npz')
eng = engine.Engine(engine.program_arrays(prog))
prev = eng.counters()
masks = eng.root_level()
for lvl in range(L):
    st = eng.level_eval(masks, lvl + 1)
    torch.cuda.synchronize()
    c = eng.counters()
    d = {k: c[k] - prev[k] for k in c}
    prev = c
    n = masks.shape[0]
    print(f'level {lvl+1}: n={n} k2a tried={d["k2a_tried"]} certified={d["k2a_certified"]} steps/lp={d["k2a_steps"]/max(1,d["k2a_tried"]):.1f} | simplex lps={d["k2_lps"]} pivots/lp={d["k2_pivots"]/max(1,d["k2_lps"]):.1f} | k4 lps={d["k4_lps"]}')
    feas = eng.select(st, 2, 2)
    masks = eng.children(masks, feas, lvl + 1)
