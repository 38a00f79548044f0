"""Dump one kernel's SASS from libppgpu.so compactly:  python scripts/sass_loop.py <mangled substring> [lo hi]"""
import re, subprocess, sys
name = sys.argv[1]
txt = subprocess.run(['cuobjdump', '-sass', 'ppopt_b200/libppgpu.so'], capture_output=True, text=True).stdout
rows, on = [], False
for line in txt.splitlines():
    if 'Function :' in line:
        if on: break
        on = name in line
        continue
    if on:
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
        if m: rows.append((int(m.group(1), 16), m.group(2).strip()))
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
print('total', len(rows))
for a, t in rows:
    if lo <= a <= hi: print(f'{a:04x} {t}')
