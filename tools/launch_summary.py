"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, total time and share per kernel.
Usage: python tools/launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[start]
iname, ival, iunit = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) < len(hdr):
        continue
    v = float(r[ival].replace(',', ''))
    v = {'ns': v / 1e3, 'us': v, 'usecond': v, 'ms': v * 1e3, 'msecond': v * 1e3, 'nsecond': v / 1e3}.get(r[iunit], v)
    name = re.sub(r'\(anonymous namespace\)::|<unnamed>::|void ', '', r[iname])
    name = re.sub(r'\(.*', '', name)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("%-50s %8s %12s %8s %7s" % ("kernel", "launches", "total us", "avg us", "share"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-50s %8d %12.1f %8.1f %6.1f%%" % (k[:50], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
print("%-50s %8d %12.1f" % ("total", sum(a[0] for a in agg.values()), tot))
