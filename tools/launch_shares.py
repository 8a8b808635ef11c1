#!/usr/bin/env python
"""Per-kernel totals and shares from an ncu launch list (--metrics gpu__time_duration.sum --csv).
usage: tools/launch_shares.py launches.csv [skip]"""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = rows[0]
ki, vi, ui, idi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
tot = collections.OrderedDict()
for r in rows[1:]:
    if int(r[idi]) < skip:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v
    name = r[ki].split("(")[0][-60:]
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += v
total = sum(t[1] for t in tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:60s} {n:8d} {us:12.1f} {us / n:10.1f} {100 * us / total:6.1f}%")
