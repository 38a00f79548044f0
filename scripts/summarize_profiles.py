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

print('# Round 1 - measured numbers and ncu evidence (B200, one GPU unless stated)\n')
print('All captures: `gpurun`, `--clock-control none`, workload = synthetic dense mpQP 100x30x6 '
      '(`tests/golden/synthetic_30_6_40_s0.npz`), combinatorial levels 1..4 (3,940,375 candidates per step) unless stated. '
      'Raw reports stay in `gpurun_out/` (scratch); the CSV launch lists are copied here.\n')
for title, path in (('final version (K2a certificates + simplex for the rest)', 'profiles/r01_launches_final.csv'),
                    ('v1 (first correct version, simplex for every candidate)', 'profiles/r01_launches_v1.csv')):
    agg = launches(path)
    tot = sum(v[1] for v in agg.values())
    print(f'## Launch list of one timed step - {title}\n')
    print(f'`ncu --metrics gpu__time_duration.sum --clock-control none -s <warm-up launches> -c 90 --csv python bench.py --steps 1 --warmup 3 --levels 4 --no-cpu-baseline` -> `{path}`\n')
    print('| kernel | launches | ms | share |\n|---|---|---|---|')
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:10]:
        print(f'| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |')
    print(f'\nTotal {tot:.1f} ms.\n')
print('## Dominant kernel, final version: `k2a_relax_small_kernel<4,4>` (level-4 launch, 3,776,565 candidates), `ncu --set full`\n')
table(raw('gpurun_out/r01_k2a_final.ncu-rep', WANT))
print('\n## v1 dominant kernel `k2_feas_kernel<4,1,40,1>` (same launch, `ncu --set full`)\n')
table(raw('gpurun_out/r01_k2.ncu-rep', WANT))
print('\n## Register simplex after the v2-v3 rework `k2_feas_kernel<2,2,40,1>` (all candidates, before K2a existed)\n')
table(raw('gpurun_out/r01_k2_v4.ncu-rep', WANT))
for tag, path in (('levels 1..5 (default bench.py, 78,385,935 candidates per step)', 'gpurun_out/bench_final_l5.json'),
                  ('levels 1..4', 'gpurun_out/bench_final_l4.json'), ('reference arm (`--impl reference`)', 'gpurun_out/bench_ref.json')):
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:
        continue
    print(f'\n## bench.py line, {tag}\n\n```json\n{json.dumps(d)}\n```')
