"""Reads a bench.py JSON line on stdin and prints ms per step + the stage timers on one short line."""
import json
import sys

lines = [l for l in sys.stdin.read().splitlines() if l.startswith("{")]
if not lines:
    print("no JSON line")
    sys.exit(0)
d = json.loads(lines[-1])
r = d.get("roofline") or {}
print("%.4f ms  %s  frac %.4f  launches %s" % (d["ms_per_step"], {k: round(v, 4) for k, v in (r.get("stage_ms") or {}).items()},
                                             r.get("frac", 0.0), d.get("gpu_launches")))
