"""One GPU playing rank r of a G-way split of the LAST level of the bench workload, with witness inheritance: per-family kernel
times of the rank's packed share against 1/G of the whole level.   python scripts/shard_probe2.py G [ranks...]
(K2w partitioning knobs are read once per process: PPGPU_K2W_GROUPS / PPGPU_K2W_SPLIT / PPGPU_K2W_ITEM)"""
import os, sys
sys.path.insert(0, '.')
import torch
from ppopt_b200 import engine, sharding, _lib
from ppopt_b200.mplp_program import load_presolved
G = int(sys.argv[1]); ranks = [int(x) for x in sys.argv[2:]] or [0]
L = 5
prog = load_presolved('tests/golden/synthetic_30_6_40_s0.npz')
eng = engine.Engine(engine.program_arrays(prog))
masks, parent = eng.root_level(), None
for lvl in range(L - 1):
    n = masks.shape[0]
    wit = torch.zeros((n, _lib.WITNESS_SLOTS, eng.W), dtype=torch.int64, device=eng.tdev) if n >= engine.WITNESS_MIN_LEVEL else None
    st = eng.level_eval(masks, lvl + 1, witness=wit, parent=parent)
    feas = eng.select(st, 2, 2)
    keep = {}
    nxt = eng.children(masks, feas, lvl + 1, keep=keep)
    parent = engine.ParentLevel(keep['feas_masks'], keep['ws'], keep['nf'], wit[feas].contiguous()) if wit is not None else None
    masks = nxt
n = masks.shape[0]
def run(ranges):
    packed = torch.cat([masks[lo:hi] for lo, hi in ranges]) if len(ranges) > 1 else masks[ranges[0][0]:ranges[0][1]]
    eng.profile(True); eng.profile_read(reset=True)
    eng.level_eval(packed.contiguous(), L, None, 7, parent=parent)
    torch.cuda.synchronize()
    p = eng.profile_read(reset=True); eng.profile(False)
    return {k: round(v['ms'], 1) for k, v in p.items() if v['launches']}
run([(0, n // 64)])
whole = run([(0, n)])
print('env', {k: v for k, v in os.environ.items() if k.startswith('PPGPU_K2W')}, 'whole level', whole, '-> ideal k2w per rank %.1f' % (whole['k2w_walk'] / G))
for per_rank in [int(x) for x in os.environ.get('PROBE_CHUNKS', '16').split(',')]:
    sharding.CHUNKS_PER_RANK = per_rank
    rows = []
    for r in ranks:
        ch = sharding.chunks(n, r, G)
        rows.append(run(ch))
    print(f'chunks per rank {per_rank} ({len(ch)} x {ch[0][1] - ch[0][0]}):', ' | '.join(
        f"r{r} k2w {x['k2w_walk']} k34 {x['k34_kkt_cheb']} sum {round(sum(x.values()), 1)}" for r, x in zip(ranks, rows)))
