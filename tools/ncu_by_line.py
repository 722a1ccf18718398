"""Joins an ncu SASS source page (per-instruction samples / executed counts) with nvdisasm -g line
info and prints the hottest CUDA source lines.  Usage:
  ncu -i rep --page source --csv > src.csv ; nvdisasm -g -c batch.cubin > batch.sass
  python tools/ncu_by_line.py src.csv batch.sass <mangled kernel substring> [file.cu]"""
import csv, re, sys, collections
srccsv, sass, kern = sys.argv[1:4]
cu = sys.argv[4] if len(sys.argv) > 4 else None
rows = list(csv.reader(open(srccsv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
ci, si = hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
inst = []
for r in rows[hi + 1:]:
    if len(r) <= ci or not r[0].strip():
        continue
    try:
        stalls = {hdr[i]: int(r[i] or 0) for i in stall_cols}
        inst.append((r[1], int(r[ci] or 0), int(r[si] or 0), stalls))
    except ValueError:
        pass
# line info from nvdisasm
lines = open(sass).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and kern in l and l.rstrip().endswith(':'))
cur = None
seq = []
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('//-------'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        seq.append(cur)
print('ncu instructions', len(inst), 'nvdisasm instructions', len(seq))
n = min(len(inst), len(seq))
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for k in range(n):
    a = agg[seq[k]]
    a[0] += inst[k][1]; a[1] += inst[k][2]; a[2].update(inst[k][3])
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
text = {}
if cu:
    for i, l in enumerate(open(cu).read().split('\n')):
        text[i + 1] = l.strip()
print('total warp instructions %d, samples %d' % (ti, ts))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[5]) if len(sys.argv) > 5 else 35]:
    top = ', '.join('%s %d%%' % (k.replace('stall_', ''), 100 * v / max(a[1], 1)) for k, v in a[2].most_common(3))
    src = text.get(key[1], '') if key and cu and key[0].endswith(cu.split('/')[-1]) else ''
    print('%5.1f%% instr %5.1f%% samples  %s:%s  [%s]  %s' % (100 * a[0] / ti, 100 * a[1] / ts, key[0] if key else '?', key[1] if key else '?', top, src[:90]))
