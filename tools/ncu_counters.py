#!/usr/bin/env python
"""profiles/k_step5_counters.json from an `ncu --set full` capture of k_step5 (what bench.py's roofline.secondary reads).
usage: ncu_counters.py rep num_envs out.json"""
import csv
import json
import subprocess
import sys

rep, n, outp = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, vals = rows[0], rows[2]
g = {h: vals[i] for i, h in enumerate(hdr)}
f = lambda k: float(g[k].replace(",", ""))  # noqa: E731
# FP32 operations from the source page: thread-level executed counts of FFMA (2 flops), FMUL, FADD (+ MUFU)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
sh = srows[1]
i_src, i_thr = sh.index("Source"), sh.index("Thread Instructions Executed")
ffma = fmul = fadd = 0.0
import re
for r in srows[2:]:
    if len(r) <= i_thr:
        continue
    op = re.sub(r"^@!?U?P\d+\s+", "", r[i_src].strip()).split(".")[0].split()[0:1]
    if not op:
        continue
    t = float(r[i_thr] or 0)
    if op[0] == "FFMA": ffma += t
    elif op[0] == "FMUL": fmul += t
    elif op[0] == "FADD": fadd += t
d = {
    "source": rep.split("/")[-1] + " (ncu --set full --clock-control none, one k_step5 launch, cfg2)",
    "kernel": g["Kernel Name"], "num_envs": n, "time_ms_under_ncu": f("gpu__time_duration.sum") / 1e6 if g["gpu__time_duration.sum"] and f("gpu__time_duration.sum") > 1e4 else f("gpu__time_duration.sum"),
    "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "threads_per_inst": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "warps_per_sm": f("sm__warps_active.avg.per_cycle_active"),
    "warp_inst_per_env_step": f("smsp__inst_executed.sum") / n,
    "fp32_flops_per_env_step": (2 * ffma + fmul + fadd) / n,
    "dram_bytes_per_launch": f("dram__bytes_read.sum") * (1e6 if "Mbyte" in rows[1][hdr.index("dram__bytes_read.sum")] else 1e3 if "Kbyte" in rows[1][hdr.index("dram__bytes_read.sum")] else 1)
                             + f("dram__bytes_write.sum") * (1e6 if "Mbyte" in rows[1][hdr.index("dram__bytes_write.sum")] else 1e3 if "Kbyte" in rows[1][hdr.index("dram__bytes_write.sum")] else 1),
    "registers_per_thread": f("launch__registers_per_thread"), "shared_mem_per_block_bytes": f("launch__shared_mem_per_block_dynamic") * 1e3,
    "grid": f("launch__grid_size"), "block": f("launch__block_size"),
}
json.dump(d, open(outp, "w"), indent=1)
print(json.dumps(d, indent=1))
