#!/usr/bin/env python
"""Where a kernel touches local memory: LDL / STL instructions of one function by source line.
usage: nvdisasm --print-line-info x.cubin > all.sass; tools/sass_locals.py all.sass <substring of the mangled name>"""
import re
import sys
from collections import Counter

path, want = sys.argv[1], sys.argv[2]
inside, cur = False, None
c, total = Counter(), 0
for l in open(path):
    if l.startswith(".text."):
        inside = want in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m[1].split("/")[-1], int(m[2]), "inlined" if "inlined" in m[3] else "")
        continue
    if re.search(r"^\s+/\*[0-9a-f]+\*/", l):
        total += 1
        k = re.search(r"\b(LDL|STL)(\.\w+)*\b", l)
        if k:
            c[(cur[0], cur[1], k[1])] += 1
print(f"{total} instructions, {sum(v for (f,n,k),v in c.items() if k=='LDL')} LDL, {sum(v for (f,n,k),v in c.items() if k=='STL')} STL")
for (f, n, k), v in sorted(c.items()):
    print(f"  {f}:{n}  {k} x{v}")
