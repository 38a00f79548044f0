"""Throughput of the batched point location (K7) next to the CPU restatement of the reference's per-point loop."""
import os, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import numpy, torch
import point_location_oracle as plo
from ppopt_b200 import PointLocation
from ppopt_b200.critical_region import CriticalRegion
from ppopt_b200.solution import Solution
name = sys.argv[1] if len(sys.argv) > 1 else 'rand_6_3_12_s1'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
g = numpy.load(f'tests/golden/pointloc/{name}.npz')
regs = [CriticalRegion(g[f'r{i}_A'], g[f'r{i}_b'], None, None, numpy.asarray(g[f'r{i}_E'], dtype=float), g[f'r{i}_f'], [])
        for i in range(int(g['n_regions']))]
pl = PointLocation(Solution(None, regs))
lo, hi = g['thetas'].min(0), g['thetas'].max(0)
thetas = numpy.random.default_rng(0).uniform(lo, hi, size=(N, len(lo)))
d = torch.from_numpy(thetas).cuda()
pl._run(d[:1000], None, True); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); reg, x = pl._run(d, None, True); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
t0 = time.perf_counter(); idx, xx = pl.evaluate_batch(thetas); t1 = time.perf_counter()
regions = plo.load_regions(g)
n_cpu = 300
t2 = time.perf_counter()
ref = [plo.locate_upop(regions, th) for th in thetas[:n_cpu]]
t3 = time.perf_counter()
assert ref == idx[:n_cpu].tolist()
print(f'{name}: {len(regs)} regions, {int(pl.region_constraints[-1])} half-spaces, t={pl.t}, n={pl.n_x}; {N} points: '
      f'device {ms:.2f} ms = {N / ms * 1e3:.3e} points/s, host arrays in/out {1e3 * (t1 - t0):.1f} ms = {N / (t1 - t0):.3e} points/s; '
      f'located {(idx >= 0).mean():.2f}; CPU restatement {n_cpu / (t3 - t2):.3e} points/s (numpy, 1 core)')
