#!/usr/bin/env python
"""Opcode mix (executed warp instructions) of a kernel from an .ncu-rep source page; optional address range.
usage: ncu_opmix.py rep [lo hi]   (hex offsets relative to the kernel start)"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ia, isrc, iinst, ithr = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
base = None
mix = defaultdict(lambda: [0, 0])
tot = 0
for r in rows[2:]:
    if len(r) <= ithr or not r[ia].startswith("0x"):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    if not (lo <= a - base < hi):
        continue
    s = r[isrc].strip()
    s = re.sub(r"^@!?U?P\d+\s+", "", s)
    op = s.split()[0].split(".")[0] if s else "?"
    n = int(r[iinst] or 0)
    mix[op][0] += n; mix[op][1] += int(r[ithr] or 0)
    tot += n
print(f"range {lo:#x}..{hi:#x}: {tot:.3e} warp instructions")
for op, (n, th) in sorted(mix.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"  {op:12s} {n / tot * 100:6.2f}%   {th / max(n, 1):5.1f} thr/inst")
