#!/usr/bin/env python
"""Per-source-line executed warp instructions / stall samples from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py rep [top]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
for i, r in enumerate(rows):
    if "Instructions Executed" in r:
        hdr = r; start = i + 1; break
iinst, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
print(hdr[:6])
isrc = hdr.index("Source")
agg = defaultdict(lambda: [0, 0, 0])
cur = "?"
tot = 0
for r in rows[start:]:
    if len(r) <= iinst:
        continue
    try:
        n = int(r[iinst] or 0); s = int(r[isamp] or 0)
    except ValueError:
        continue
    key = r[isrc].strip()[:110]
    agg[key][0] += n; agg[key][1] += s; agg[key][2] += 1
    tot += n
for k, (n, s, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n / max(tot, 1) * 100:5.2f}% inst {s:7d} smp {c:4d} rows | {k}")
