#!/usr/bin/env python
"""Per-substep trace of the constraint solver on the captured regression case (tests/golden/regress_default_step_case1.npz):
builds a -DSMPLSIM_TRACE copy of the library (smplsim_b200/libsmplsim_trace.so), loads it through SMPLSIM_SO and prints, for
env 0, every solver iteration (adopt / line search / final), the line-search derivative and step, and max|qacc| per substep.
Run under gpurun:  python tools/solver_trace.py"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
SO = os.path.join(ROOT, "smplsim_b200", "libsmplsim_trace.so")
os.environ["SMPLSIM_SO"] = SO
from smplsim_b200 import _lib  # noqa: E402

if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(s) for s in _lib.sources()):
    subprocess.check_call([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + _lib.NVCC_FLAGS + ["-DSMPLSIM_TRACE", "-o", SO, "smplsim_capi.cu"],
                          cwd=os.path.join(ROOT, "smplsim_b200", "csrc"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from smplsim_b200.batched import HumanoidBatchB200  # noqa: E402
from smplsim_b200.cfg import make_cfg  # noqa: E402

d = np.load(os.path.join(ROOT, "tests", "golden", "regress_default_step_case1.npz"))
env = HumanoidBatchB200(make_cfg(env="speed"), num_envs=1, seed=0)
env.reset()
env.task_change_step.fill_(10 ** 6)
L = _lib.lib()
buf = (C.c_float * 8192)()
L.smplsim_debug_trace(buf, 8192)                      # drop the reset trace
env.step(torch.as_tensor(d["action"], device="cuda:0")[None])
n = L.smplsim_debug_trace(buf, 8192)
v = np.array(buf[:n]); i = 0
while i < n:
    k = int(v[i]); r = v[i + 1:i + 1 + k]; i += k + 1
    tag = int(r[0])
    if tag == 0:
        print("SUB s=%d iters=%d nrows=%d max|qacc|=%.4g z=%.5f" % (r[1], r[2], r[3], r[4], r[5]))
    elif tag == 1:
        print("   it=%d %s max|qstar|=%.4g nslots=%d" % (r[1], {0: "-", 1: "fin", 2: "adopt", 3: "ls"}[int(r[2])], r[3], r[4]))
    else:
        print("      ls f0=%.5g al=%.5g g1=%.5g g2=%.5g" % tuple(r[1:5]))
print("max|qvel| after the step", env.qvel.abs().max().item(), "(oracle: 3.6229)")
