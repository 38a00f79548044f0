import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith('{'):
        continue
    d = json.loads(line)
    print('value %.4g %s | ms/step %.1f | e2e %.4g | k2 frac %.3f (%.2f TF) | launches %d' % (
        d['value'], d['unit'], d['ms_per_step'], d.get('e2e', {}).get('value', 0), d.get('roofline', {}).get('frac', 0) or 0,
        d.get('roofline', {}).get('achieved', 0) or 0, d.get('gpu_launches', 0)))
    print('kernels', {k: round(v, 2) for k, v in d.get('kernels_ms_per_step', {}).items()})
    if 'cpu_baseline' in d:
        print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
