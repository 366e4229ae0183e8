#!/usr/bin/env python
"""Tiny workload for compute-sanitizer (memcheck / racecheck): reset + steps of speed (Default) and getup (Fall init) envs.
usage: sanitize_probe.py [speed|getup|both] [n_envs] [steps]"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smplsim_b200.batched import HumanoidBatchB200
from smplsim_b200.cfg import make_cfg
which = sys.argv[1] if len(sys.argv) > 1 else "both"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 18
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
for task in (("speed", "getup") if which == "both" else (which,)):
    env = HumanoidBatchB200(make_cfg(env=task), num_envs=n, seed=2)
    env.reset()
    g = torch.Generator(device="cuda:0"); g.manual_seed(0)
    for t in range(steps):
        env.step(torch.clamp(torch.randn(n, env.num_actions, generator=g, device="cuda:0") * 0.3, -1, 1))
        env.reset_done()
    torch.cuda.synchronize()
    print(task, "ok", float(env.qvel.abs().max()))
