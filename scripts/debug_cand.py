"""evaluate explicit active sets on the GPU and print status + K5 info:  python scripts/debug_cand.py name "a,b,c;d,e,f" """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import numpy, torch
from ppopt_b200 import engine
from ppopt_b200.mplp_program import load_presolved
import twin_binding
name = sys.argv[1]
sets = [[int(x) for x in s.split(',')] for s in sys.argv[2].split(';')]
path = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
eng = engine.Engine(engine.program_arrays(load_presolved(path)))
tw = twin_binding.Twin.from_npz(path)
masks = eng.masks_from_lists(sets)
k_act = len(sets[0]) - eng.n_eq
st = eng.level_eval(masks, k_act)
print('status after level_eval', st.cpu().tolist())
sel = torch.arange(len(sets), device=eng.tdev)
laws, rows, flags, info = eng.emit(masks, sel, k_act, st)
print('status after emit', st.cpu().tolist())
print('info', info.cpu().numpy())
tst, aux = tw.eval(tw.masks(sets), aux=True)
print('twin', tst.tolist(), aux.tolist())
for i, s in enumerate(sets):
    rc, l2, r2, f2, i2, mg = tw.emit(tw.masks([s])[0], margins=True)
    print('twin emit', rc, i2)
    print('max row diff gpu-vs-twin', numpy.abs(rows[i].cpu().numpy() - r2).max())
print(eng.counters())
