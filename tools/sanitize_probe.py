#!/usr/bin/env python
"""Tiny workload for compute-sanitizer (memcheck / racecheck): reset + steps of speed (Default) and getup (Fall init) envs;
"shapes": speed with three body shapes interleaved (smplsim_create_shapes), "selfcol": getup with geom-geom rows.
usage: sanitize_probe.py [speed|getup|both|shapes|selfcol|all] [n_envs] [steps]"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smplsim_b200.batched import HumanoidBatchB200
from smplsim_b200.cfg import make_cfg
which = sys.argv[1] if len(sys.argv) > 1 else "both"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 18
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cases = {"both": ("speed", "getup"), "all": ("speed", "getup", "shapes", "selfcol")}.get(which, (which,))
for task in cases:
    kw, ov, tk = {}, {}, task
    if task == "shapes":
        from smplsim_b200 import model as M
        from smplsim_b200.abi import model_from_cfg
        tk = "speed"
        c0 = make_cfg(env=tk); m0 = model_from_cfg(c0); base = M.load_parsed("smpl")
        mk = lambda **v: M.build_model(M.shape_variant(base, **v), timestep=m0.timestep, contact_bodies=list(c0.env.contact_bodies), control_mode=c0.env.control_mode)
        kw = dict(models=[m0, mk(leg=1.15, girth=1.1), mk(leg=0.9, arm=1.1, density=0.9)], env_model=[(i * 5) % 3 for i in range(n)])
    if task == "selfcol":
        tk, ov = "getup", {"env.self_collision": True}
    env = HumanoidBatchB200(make_cfg(env=tk, overrides=ov), num_envs=n, seed=2, **kw)
    env.reset()
    g = torch.Generator(device="cuda:0"); g.manual_seed(0)
    for t in range(steps):
        env.step(torch.clamp(torch.randn(n, env.num_actions, generator=g, device="cuda:0") * 0.3, -1, 1))
        env.reset_done()
    torch.cuda.synchronize()
    print(task, "ok", float(env.qvel.abs().max()))
