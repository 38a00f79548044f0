"""Writes profiles/r01_summary.md from the ncu artefacts in gpurun_out/ and the last bench JSON lines."""
import collections
import csv
import json
import subprocess
import sys

def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == 'ID':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get('Metric Name') == 'gpu__time_duration.sum':
                name = d['Kernel Name'].split('(')[0][:64]
                v = float(d['Metric Value'].replace(',', ''))
                u = d['Metric Unit']
                v = v / 1e3 if u == 'us' else v / 1e6 if u == 'ns' else v
                a = agg.setdefault(name, [0, 0.0])
                a[0] += 1
                a[1] += v
    return agg

def raw(path, want):
    out = subprocess.run(f'ncu -i {path} --page raw --csv', shell=True, capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    h, u, v = rr[0], rr[1], rr[2]
    return [(a, c, b) for a, b, c in zip(h, u, v) if a in want]

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio']

def table(rows):
    print('| metric | value |\n|---|---|')
    for a, c, b in rows:
        print(f'| {a} | {c} {b} |')

def bench_line(path):
    try:
        return json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:
        return None

print('# Round 1 - measured numbers and ncu evidence (B200, one GPU unless stated)\n')
print('All captures: `gpurun`, `--clock-control none`, workload = synthetic dense mpQP 100x30x6 '
      '(`tests/golden/synthetic_30_6_40_s0.npz`), combinatorial levels 1..5 (74,536,536 candidates per step; the level-5 '
      'array is evaluated in 18 chunks of 2^22 candidates) unless stated. Raw reports stay in `gpurun_out/` (scratch); the CSV '
      'launch lists are copied here.\n')
d = bench_line('gpurun_out/bench_final_l5.json')
if d:
    print('## Headline (default `python bench.py`)\n')
    k = d['kernels_ms_per_step']
    print(f"{d['value']:.3e} candidates/s device-resident ({d['ms_per_step']:.1f} ms/step), {d['e2e']['value']:.3e} end to end through "
          f"`solve_mpqp` ({d['e2e']['ms_per_step']:.1f} ms/step), CPU oracle {d['cpu_baseline']['value']:.0f} candidates/s on "
          f"{d['cpu_baseline']['cores']} cores. Roofline of the dominant kernel family (K2a): {d['roofline']['achieved']:.2f} of "
          f"{d['roofline']['peak']:.2f} TFLOP/s fp64 = {d['roofline']['frac']:.3f}; share of the step "
          f"{d['roofline']['share_of_step']:.2f}.\n")
    print('| kernel family | ms / step (CUDA events inside bench.py) |\n|---|---|')
    for name, v in sorted(k.items(), key=lambda x: -x[1]):
        print(f'| {name} | {v:.2f} |')
    print()
print('History of the same workload on one B200 (ms per step, levels 1..5): first correct version (simplex for every '
      'candidate, levels 1..4 only: 674 ms for 3.9e6 candidates) -> K2a v1 certificates 1697 -> K2a v2 in registers 1112 -> '
      'K2 warm start from the K2a iterate 997 -> block scans of the status bytes in K2/K34 890 -> 32 candidates per K2a '
      'queue item 853 -> second relaxation phase 748.\n')
for title, path, cmd in (
        ('final version, the timed step of the default command', 'profiles/r01_launches_final.csv',
         '`ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline` (4th solve of the run = the timed device-resident step)'),
        ('mid-round version (K2a v1 + simplex for the rest), levels 1..4', 'profiles/r01_launches_k2a_v1.csv',
         '`... -s <warm-up launches> -c 90 --csv python bench.py --steps 1 --warmup 3 --levels 4 --no-cpu-baseline`'),
        ('v1 (first correct version, simplex for every candidate), levels 1..4', 'profiles/r01_launches_v1.csv',
         'same command')):
    agg = launches(path)
    tot = sum(v[1] for v in agg.values())
    print(f'## Launch list of one timed step - {title}\n')
    print(f'{cmd} -> `{path}`\n')
    print('| kernel | launches | ms | share |\n|---|---|---|---|')
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:12]:
        print(f'| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |')
    print(f'\nTotal {tot:.1f} ms (cold-cache, serialised launches: shares, not absolutes, are comparable with the event timings).\n')
for title, path in (
        ('Dominant kernel, final: `k2a_relax_reg_kernel<4,5,true>` (one level-5 chunk, 4,194,304 candidates)', 'gpurun_out/r01_k2a_final_l5.ncu-rep'),
        ('Final `k2_feas_kernel<2,2,40,1>` (same chunk: warm list + block scan)', 'gpurun_out/r01_k2_final_l5.ncu-rep'),
        ('Final `k34_kernel<4,8,4>` (same chunk)', 'gpurun_out/r01_k34_final_l5.ncu-rep'),
        ('K2a v2 before the batched queue: `k2a_relax_reg_kernel<4,4,true>` (level-4 launch, 3,776,565 candidates)', 'gpurun_out/r01_k2a_v2.ncu-rep'),
        ('K2a v1: `k2a_relax_small_kernel<4,4>` (same level-4 launch)', 'gpurun_out/r01_k2a_final.ncu-rep'),
        ('v1 dominant kernel `k2_feas_kernel<4,1,40,1>` (level-4 launch, simplex for every candidate)', 'gpurun_out/r01_k2.ncu-rep'),
        ('Register simplex after the v2-v3 rework `k2_feas_kernel<2,2,40,1>` (all candidates, before K2a existed)', 'gpurun_out/r01_k2_v4.ncu-rep')):
    try:
        rows_ = raw(path, WANT)
    except Exception:
        continue
    print(f'## {title}, `ncu --set full`\n')
    table(rows_)
    print()
print('## Multi-GPU (levels 1..5, strong scaling, `bench.py --gpus N` under torchrun, CUDA events, max over ranks)\n')
print('| GPUs | ms / step | candidates / s | speed-up |\n|---|---|---|---|')
base = bench_line('gpurun_out/bench_final_l5.json')
for n_g, path in ((1, 'gpurun_out/bench_final_l5.json'), (4, 'gpurun_out/bench_final_4gpu.json'), (8, 'gpurun_out/bench_final_8gpu.json')):
    d = bench_line(path)
    if d and base:
        print(f"| {n_g} | {d['ms_per_step']:.1f} | {d['value']:.3e} | {base['ms_per_step'] / d['ms_per_step']:.2f}x |")
print('\nBefore the snake dealing of chunks and the adaptive K2a queue items the 8-GPU step took 168.8 ms (4.4x): 256 chunks of '
      '276 K candidates left every warp 3-4 queue items of 32 candidates, K2a ran 42 % above its share.\n')
print('## K7 batched point location (SURVEY 8f row 4), first measurement\n')
print('`python scripts/pointloc_bench.py rand_6_3_12_s1 2000000`: 299 regions / 1,801 half-spaces, t = 3, n = 6, 2,000,000 uniform '
      'points (46 % inside some region): 8.6e6 points/s with the points resident in HBM, 1.0e7 points/s from host arrays '
      '(second call, warm); the numpy restatement of the reference loop does 1.6e3 points/s on one core. Not tuned yet '
      '(one warp per point, every miss scans all regions).\n')
for tag, path in (('levels 1..5 (default bench.py)', 'gpurun_out/bench_final_l5.json'),
                  ('levels 1..4', 'gpurun_out/bench_final_l4.json'), ('reference arm (`--impl reference`)', 'gpurun_out/bench_ref.json'),
                  ('2 GPUs', 'gpurun_out/bench_final_2gpu.json'), ('4 GPUs', 'gpurun_out/bench_final_4gpu.json'),
                  ('8 GPUs', 'gpurun_out/bench_final_8gpu.json')):
    d = bench_line(path)
    if d is None:
        continue
    print(f'\n## bench.py line, {tag}\n\n```json\n{json.dumps(d)}\n```')
