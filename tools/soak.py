#!/usr/bin/env python
"""Soak test: long random-action rollouts on every task; reports non-finite envs and the largest |qvel| seen.
usage: soak.py [n_envs] [steps]"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smplsim_b200.batched import HumanoidBatchB200
from smplsim_b200.cfg import make_cfg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
g = torch.Generator(device="cuda:0"); g.manual_seed(1)
for task, robot, sigma in [("speed", "smpl_humanoid", 0.0821), ("speed", "smpl_humanoid", 0.5), ("reach", "smpl_humanoid", 0.3),
                           ("getup", "smpl_humanoid", 0.3), ("getup", "smplx_humanoid", 0.2)]:
    ne = n if robot == "smpl_humanoid" else max(n // 4, 64)
    env = HumanoidBatchB200(make_cfg(env=task, robot=robot), num_envs=ne, seed=3)
    env.reset()
    worst, bad, nres, itmax, st8, st16, stbad = 0.0, 0, 0, 0, 0, 0, 0
    for t in range(steps):
        a = torch.clamp(torch.randn(ne, env.num_actions, generator=g, device="cuda:0") * sigma, -1, 1)
        env.step(a)
        qv = env.qvel
        fin = torch.isfinite(qv).all(dim=1) & torch.isfinite(env.obs_buf).all(dim=1)
        bad += int((~fin).sum())
        worst = max(worst, float(torch.nan_to_num(qv.abs(), nan=0.0, posinf=0.0).max()))
        itmax = max(itmax, int(env.solver_iter.max()))
        sb = env.status
        st8 += int(((sb & 8) != 0).sum()); st16 += int(((sb & 16) != 0).sum()); stbad += int(((sb & 7) != 0).sum())
        nres += int(env.reset_buf.sum())
        env.reset_done()
    print(f"{task:6s} {robot:15s} sigma {sigma:.3f} envs {ne} steps {steps}: non-finite env-steps {bad}, max|qvel| {worst:.1f}, resets {nres}, max solver iters (last substep) {itmax}, env-steps with rows dropped {st8}, at the iteration cap {st16}, mj_check resets {stbad}", flush=True)
