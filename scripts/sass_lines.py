"""Per-source-line instruction / stall-sample shares of one kernel:  python scripts/sass_lines.py x.ncu-rep obj.o 'mangled substring'
(ncu's source page is SASS only in CSV form; the line table comes from nvdisasm -g on the same object file)"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj, key = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iex, ismp = hdr.index('Instructions Executed'), hdr.index('# Samples')
ins = [(int(r[iex] or 0), int(r[ismp] or 0)) for r in rows[2:] if len(r) > iex]
tmp = tempfile.mkdtemp()
subprocess.run(f'cd {tmp} && cuobjdump -xelf all {os.path.abspath(obj)} > /dev/null', shell=True)
cub = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], None, False
for l in dis:
    if l.startswith('//--------------------- .text.'):
        on = key in l
        continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): lines.append(cur)
print(len(ins), 'profiled instructions,', len(lines), 'disassembled')
n = min(len(ins), len(lines))
agg = collections.defaultdict(lambda: [0, 0])
for (ex, sm), ln in zip(ins[:n], lines[:n]):
    agg[ln][0] += ex; agg[ln][1] += sm
te, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
src = {}
for ln, (ex, sm) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
    if ln is None: continue
    f = os.path.join(os.path.dirname(os.path.abspath(obj)), '..', ln[0])
    if ln[0] not in src:
        try: src[ln[0]] = open(f).read().splitlines()
        except OSError: src[ln[0]] = []
    text = src[ln[0]][ln[1] - 1].strip()[:90] if len(src[ln[0]]) >= ln[1] else ''
    print(f'{ln[0]}:{ln[1]:4d} {100.0*ex/te:5.1f}% instr {100.0*sm/ts:5.1f}% samples | {text}')
