"""One GPU playing rank 0 of a G-way split: time of its chunks of the last level against 1/G of the whole level."""
import sys, time
sys.path.insert(0, '.')
import torch
from ppopt_b200 import engine, sharding
from ppopt_b200.mplp_program import load_presolved
G = int(sys.argv[1]); L = int(sys.argv[2]) if len(sys.argv) > 2 else 5
prog = load_presolved('tests/golden/synthetic_30_6_40_s0.npz')
eng = engine.Engine(engine.program_arrays(prog))
masks = eng.root_level()
for lvl in range(L - 1):
    st = eng.level_eval(masks, lvl + 1)
    masks = eng.children(masks, eng.select(st, 2, 2), lvl + 1)
n = masks.shape[0]
def timed(ranges):
    status = torch.zeros((n,), dtype=torch.uint8, device=eng.tdev)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for lo, hi in ranges:
        eng.level_eval(masks, L, status, 7, lo, hi)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
timed([(0, n // 64)])
whole = timed([(0, n)])
print(f'n={n} whole level {whole:.1f} ms -> ideal per rank {whole / G:.1f} ms')
for per_rank, min_chunk in ((32, 16384), (16, 65536), (8, 65536), (4, 65536)):
    sharding.CHUNKS_PER_RANK, sharding.MIN_CHUNK = per_rank, min_chunk
    for r in (0, G // 2, G - 1):
        ch = sharding.chunks(n, r, G)
        ms = timed(ch)
        print(f'per_rank {per_rank:2d} min_chunk {min_chunk:8d} rank {r}: {len(ch):2d} chunks of {ch[0][1] - ch[0][0]:8d} -> {ms:.1f} ms')
