#!/usr/bin/env python
"""Like ncu_lines.py but aggregated per source line over all inlined copies, sorted by executed
thread-instructions.  usage: ncu_lines_agg.py rep kernel_regex [top]"""
import csv, subprocess, sys, io, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        inst = int(d.get("Instructions Executed", "0") or 0); samp = int(d.get("# Samples", "0") or 0)
        thr = int(d.get("Thread Instructions Executed", "0") or 0)
    except ValueError:
        continue
    a = agg[(fname, line)]
    a[0] += samp; a[1] += inst; a[2] += thr; a[3] = r[1].strip()[:100]
ti = sum(a[1] for a in agg.values()); ts = sum(a[0] for a in agg.values())
print(f"total warp inst {ti} samples {ts}")
for (f, l), (s, i, t, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*i/max(1,ti):5.1f}% inst {100*s/max(1,ts):5.1f}% smp thr/inst {t/max(1,i):5.1f}  {f}:{l}  {src}")
