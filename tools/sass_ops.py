"""Per kernel of libmassb200.so: how many SASS instructions of the Blackwell-specific kinds it holds (tcgen05 MMA =
UTC*MMA, TMEM loads = LDTM, TMA-engine bulk copies = UBLKCP / UTMALDG, mbarrier waits = SYNCS, packed fp32 FMA =
FFMA2, cp.async = LDGSTS).  Usage: python tools/sass_ops.py [library] > profiles/rNN_sass_ops.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mass_b200", "csrc", "libmassb200.so")
text = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
KINDS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG",
         "SYNCS", "FFMA2", "LDGSTS", "MATCH", "REDUX"]
per = collections.OrderedDict()
cur = None
for line in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
        cur = per.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        for k in KINDS:
            if op.startswith(k):
                cur[k] += 1
print("SASS of %s (sm_100a), instructions of Blackwell-specific kinds per kernel" % os.path.basename(so))
print("%-44s %s" % ("kernel", " ".join("%8s" % k for k in KINDS if any(c[k] for c in per.values()))))
used = [k for k in KINDS if any(c[k] for c in per.values())]
for name, c in per.items():
    if any(c[k] for k in used):
        print("%-44s %s" % (name[:44], " ".join("%8d" % c[k] for k in used)))
