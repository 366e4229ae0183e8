#!/usr/bin/env python
"""Throughput of the CPU oracle port vs thread count (sizing the cpu_baseline / --impl reference arm honestly)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
print("logical", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "host_cores()", bench.host_cores(), flush=True)
for p in ("/sys/fs/cgroup/cpu.max", "/proc/loadavg"):
    try: print(p, open(p).read().strip())
    except Exception as e: print(p, "n/a")
for th in [int(x) for x in (sys.argv[1:] or ["8", "16", "32", "64", "128"])]:
    co = bench.CpuOracle(threads=th)
    co.run(2)
    n = 40
    t = co.run(n)
    print(f"threads {th:4d}: {co.nenv * n / t:9.0f} env-steps/s  ({co.nenv * n / t / th:6.0f} per thread)", flush=True)
