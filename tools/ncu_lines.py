#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an .ncu-rep: instructions executed, stall samples.
usage: ncu_lines.py rep kernel_regex [top]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname = None; hdr = None; out = []
tot_inst = tot_samp = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name":
        continue
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
    if inst or samp:
        out.append((samp, inst, thr, fname, line, r[1].strip()[:90]))
        tot_inst += inst; tot_samp += samp
print(f"total inst {tot_inst}  samples {tot_samp}")
for samp, inst, thr, f, l, src in sorted(out, reverse=True)[:top]:
    print(f"{100*samp/max(1,tot_samp):5.1f}% smp {100*inst/max(1,tot_inst):5.1f}% inst  thr/inst {thr/max(1,inst):5.1f}  {f}:{l}  {src}")
