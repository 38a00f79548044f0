"""python scripts/traffic_from_rep.py <x.ncu-rep> <out.json> <candidates_per_launch> <note>
DRAM bytes (read + write) and the shared-memory wavefront utilisation of the one kernel launch in an `ncu --set full`
report -> the JSON bench.py reads for roofline.traffic (profiles/r02_<family>_traffic.json)."""
import csv
import io
import json
import subprocess
import sys

rep, out, ncand, note = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[-1]
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def get(name):
    i = hdr.index(name)
    return float(vals[i].replace(',', '')), units[i]


rd, ru = get('dram__bytes_read.sum')
wr, wu = get('dram__bytes_write.sum')
rd, wr = rd * scale[ru], wr * scale[wu]
d = {'dram_bytes_per_launch': int(rd + wr), 'dram_read': int(rd), 'dram_write': int(wr), 'candidates_per_launch': ncand,
     'kernel': vals[hdr.index('Kernel Name')][:120], 'duration_ms_under_ncu': get('gpu__time_duration.sum')[0], 'source': note}
try:
    d['smem_wavefronts_pct_of_peak'] = get('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed')[0]
except ValueError:
    pass
json.dump(d, open(out, 'w'))
print(d)
