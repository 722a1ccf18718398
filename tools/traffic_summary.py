"""Per-kernel time and DRAM traffic of an ncu launch list taken with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv.
Usage: python tools/traffic_summary.py gpurun_out/traffic.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
st = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
h = rows[st]
iN, iM, iU, iV = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Unit'), h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[st + 1:]:
    if len(r) < len(h):
        continue
    name = re.sub(r'\(anonymous namespace\)::|<unnamed>::|void |\(.*', '', r[iN])
    v = float(r[iV].replace(',', ''))
    u = r[iU]
    a = agg.setdefault(name, {'n': 0, 'us': 0.0, 'rd': 0.0, 'wr': 0.0})
    if r[iM] == 'gpu__time_duration.sum':
        a['us'] += v * {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(u, 1)
        a['n'] += 1
    else:
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
        a['rd' if 'read' in r[iM] else 'wr'] += v
tot = dict(us=0, rd=0, wr=0, n=0)
print("%-30s %4s %10s %10s %10s" % ("kernel", "n", "us", "read MB", "write MB"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
    print("%-30s %4d %10.1f %10.1f %10.1f" % (k[:30], a['n'], a['us'], a['rd'] / 1e6, a['wr'] / 1e6))
    for q in tot:
        tot[q] += a[q]
print("%-30s %4d %10.1f %10.1f %10.1f" % ("total", tot['n'], tot['us'], tot['rd'] / 1e6, tot['wr'] / 1e6))
