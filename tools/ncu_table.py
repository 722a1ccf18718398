"""Prints one line of key metrics per profiled launch of an `ncu --page raw --csv` export.
Usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_table.py raw.csv"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"),
        ("dram__bytes_write.sum", "wrMB"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__inst_executed.sum", "Minst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("launch__registers_per_thread", "regs"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit"), ("lts__t_sector_hit_rate.pct", "l2hit"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar")]
idx = [(hdr.index(c), n) for c, n in cols if c in hdr]
units = rows[1]
print(" ".join("%9s" % n for _, n in idx))
for r in rows[2:]:
    out = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void |\(.*", "", v)[:28]
            out.append("%-28s" % v)
            continue
        x = float(v.replace(",", ""))
        u = units[i]
        if n == "us":
            x = {"ns": x / 1e3, "us": x, "ms": x * 1e3}.get(u, x)
        if n in ("rdMB", "wrMB"):
            x = {"byte": x / 1e6, "Kbyte": x / 1e3, "Mbyte": x, "Gbyte": x * 1e3}.get(u, x)
        if n == "Minst":
            x /= 1e6
        out.append("%9.1f" % x)
    print(" ".join(out))
