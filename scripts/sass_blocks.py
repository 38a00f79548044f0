"""Basic-block view of an ncu source page:  python scripts/sass_blocks.py x.ncu-rep [min_share]
groups consecutive SASS instructions with the same execution count, prints share of all warp instructions, stall samples
and the mnemonic mix of every block above min_share (default 1 %)"""
import csv, subprocess, sys, io, collections
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
ins = [(r[isrc].strip(), int(r[iex] or 0), int(r[ismp] or 0)) for r in rows[2:] if len(r) > iex]
tot = sum(x[1] for x in ins); tots = sum(x[2] for x in ins)
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print(f'{len(ins)} SASS instructions, {tot:.4g} warp instructions executed, {tots} samples')
blocks, start = [], 0
for i in range(1, len(ins) + 1):
    if i == len(ins) or ins[i][1] != ins[start][1]:
        blocks.append((start, i)); start = i
for a, b in blocks:
    n = sum(x[1] for x in ins[a:b]); s = sum(x[2] for x in ins[a:b])
    if 100.0 * n / tot < minshare and 100.0 * s / tots < minshare: continue
    mix = collections.Counter(x[0].split()[1 if x[0].startswith('@') else 0].split('.')[0] for x in ins[a:b])
    print(f'[{a:5d},{b:5d}) len {b-a:4d} x {ins[a][1]:.3g} = {100.0*n/tot:5.1f}% instr {100.0*s/tots:5.1f}% samples  ' +
          ' '.join(f'{k}:{v}' for k, v in mix.most_common(9)))
