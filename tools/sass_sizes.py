#!/usr/bin/env python
"""Code-footprint map of a kernel: sizes of the local subroutines (noinline device functions) inside one kernel's SASS,
derived from CALL targets.  usage: sass_sizes.py [kernel-substring]"""
import re, subprocess, sys
so = "smplsim_b200/libsmplsim_b200.so"
pat = sys.argv[1] if len(sys.argv) > 1 else "k_step3I4WCfgILi24ELi75ELi24ELi64ELi32"
names = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn = [l.split()[2] for l in names.splitlines() if "Function :" in l and pat in l][0]
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, so], capture_output=True, text=True).stdout
addrs = [int(m.group(1), 16) for m in re.finditer(r"/\*([0-9a-f]{4,6})\*/\s+\S", sass)]
end = max(addrs) + 16
targets = sorted({int(m.group(1), 16) for m in re.finditer(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", sass)} | {0})
calls = {}
for m in re.finditer(r"/\*([0-9a-f]{4,6})\*/\s+CALL\.REL\.NOINC (0x[0-9a-f]+)", sass):
    calls.setdefault(int(m.group(2), 16), []).append(int(m.group(1), 16))
print(fn, "total", end // 1024, "KB")
for a, b in zip(targets, targets[1:] + [end]):
    callers = sorted({max(t for t in targets if t <= c) for c in calls.get(a, [])})
    print(f"  0x{a:05x}  {(b - a) / 1024:6.1f} KB   called from {[hex(c) for c in callers]}")
