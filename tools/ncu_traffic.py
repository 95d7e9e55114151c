"""DRAM traffic of one benchmark step from an `ncu --set full` capture of exactly that step's launches:
python tools/ncu_traffic.py <rep> <workload> <command string>  ->  one JSON object (merged into profiles/roofline_traffic.json)"""
import csv
import io
import json
import subprocess
import sys
import time

rep, wl, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ir, iw, ik, it = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('Kernel Name'), hdr.index('gpu__time_duration.sum')
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
tot_r = tot_w = 0.0
kern = {}
for r in rows[2:]:
    rd, wr = float(r[ir]) * scale[units[ir]], float(r[iw]) * scale[units[iw]]
    tot_r += rd; tot_w += wr
    k = r[ik].split('(')[0]
    e = kern.setdefault(k, {'launches': 0, 'dram_bytes': 0.0, 'time_' + units[it]: 0.0})
    e['launches'] += 1; e['dram_bytes'] += rd + wr; e['time_' + units[it]] += float(r[it])
print(json.dumps({wl: {'dram_bytes_per_step': tot_r + tot_w, 'dram_bytes_read': tot_r, 'dram_bytes_write': tot_w, 'launches': len(rows) - 2,
                       'by_kernel': kern, 'source': 'ncu --set full --clock-control none, one step of `%s`, %s' % (cmd, time.strftime('%Y-%m-%d'))}}))
