#!/usr/bin/env python
"""Basic-block view of a kernel from an .ncu-rep source page: consecutive SASS rows with the same executed count, with the
CUDA source line of the first row.  usage: ncu_blocks.py rep lo hi [min_share_pct]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
lo = int(sys.argv[2], 16); hi = int(sys.argv[3], 16)
minp = float(sys.argv[4]) if len(sys.argv) > 4 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, iinst, ithr, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
recs = []
base = None
for r in rows[2:]:
    if len(r) <= ithr or not r[ia].startswith("0x"):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    recs.append((a - base, r[isrc].strip(), int(r[iinst] or 0), int(r[ithr] or 0), int(r[isamp] or 0)))
tot = sum(r[2] for r in recs)
blocks = []
cur = None
for a, s, n, th, sm in recs:
    if not (lo <= a < hi):
        continue
    if cur and cur["n"] == n:
        cur["len"] += 1; cur["thr"] += th; cur["smp"] += sm; cur["ops"].append(s.split()[0] if s else "?")
    else:
        cur = {"a": a, "n": n, "len": 1, "thr": th, "smp": sm, "ops": [s.split()[0] if s else "?"], "first": s}
        blocks.append(cur)
print(f"kernel total {tot:.3e}; range total {sum(b['n'] * b['len'] for b in blocks):.3e}")
for b in blocks:
    share = b["n"] * b["len"] / tot * 100
    if share >= minp:
        print(f"  {b['a']:#7x} len {b['len']:4d} x {b['n']:9d} = {share:5.2f}%  thr/inst {b['thr'] / max(1, b['n'] * b['len']):5.1f}  smp {b['smp']:6d}  {b['first'][:70]}")
