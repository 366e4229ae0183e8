#!/usr/bin/env python
"""Per-device-function instruction / stall-sample attribution from an .ncu-rep (source page), splitting the kernel at CALL targets."""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, isamp, iinst, ithr = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
recs = []
for r in rows[2:]:
    if len(r) <= ithr or not r[ia].startswith("0x"):
        continue
    recs.append((int(r[ia], 16), r[isrc].strip(), int(r[isamp] or 0), int(r[iinst] or 0), int(r[ithr] or 0)))
base = recs[0][0]
targets = set()
for a, s, *_ in recs:
    m = re.search(r"CALL\.REL\.NOINC\s+(0x[0-9a-f]+)", s)
    if m:
        targets.add(int(m.group(1), 16) - base)
bounds = sorted(t for t in targets)
tot_i = sum(r[3] for r in recs); tot_s = sum(r[2] for r in recs)
def seg(lo, hi):
    rs = [r for r in recs if lo <= r[0] - base < hi]
    return len(rs), sum(r[3] for r in rs), sum(r[2] for r in rs), sum(r[4] for r in rs)
edges = [0] + bounds + [recs[-1][0] - base + 16]
print(f"total warp-inst {tot_i:.3e}  samples {tot_s}")
for lo, hi in zip(edges[:-1], edges[1:]):
    n, i, s, th = seg(lo, hi)
    if i == 0:
        continue
    print(f"  fn@{lo:#8x} {n:5d} SASS  {i / tot_i * 100:5.1f}% inst  {s / max(1, tot_s) * 100:5.1f}% samples  {th / max(1, i):5.1f} thr/inst")
