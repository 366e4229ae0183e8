"""per-substep solver trace of the regression case (needs the -DSMPLSIM_TRACE build: SMPLSIM_SO=smplsim_b200/libsmplsim_trace.so)"""
import sys, os, ctypes as C, torch, numpy as np
sys.path.insert(0, '/root/repo')
from smplsim_b200.batched import HumanoidBatchB200
from smplsim_b200.cfg import make_cfg
from smplsim_b200 import _lib
d = np.load('/root/repo/tests/golden/regress_default_step_case1.npz')
e1 = HumanoidBatchB200(make_cfg(env="speed"), num_envs=1, seed=0)
e1.reset()
e1.task_target[0, 0] = float(d['task_target'][0]); e1.task_change_step.fill_(10**6)
L = _lib.lib()
buf = (C.c_float * 8192)()
L.smplsim_debug_trace(buf, 8192)   # drop the reset trace
a = torch.as_tensor(d['action'], device="cuda:0")[None]
e1.step(a)
n = L.smplsim_debug_trace(buf, 8192)
v = np.array(buf[:n]); i = 0
while i < n:
    k = int(v[i]); r = v[i + 1:i + 1 + k]; i += k + 1
    tag = int(r[0])
    if tag == 0: print("SUB s=%d iters=%d nrows=%d max|qacc|=%.4g z=%.5f" % (r[1], r[2], r[3], r[4], r[5]))
    elif tag == 1: print("   it=%d %s max|qstar|=%.4g nslots=%d" % (r[1], {0: '-', 1: 'fin', 2: 'adopt', 3: 'ls'}[int(r[2])], r[3], r[4]))
    else: print("      ls f0=%.5g al=%.5g g1=%.5g g2=%.5g" % tuple(r[1:5]))
print("WARMSET", os.environ.get("SMPLSIM_WARMSET"), "max|qvel|", e1.qvel.abs().max().item())
