#!/usr/bin/env python
"""Tiny workload for ncu: N envs of the cfg2 workload, warm-up with random actions, then a few timed steps."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smplsim_b200.batched import HumanoidBatchB200  # noqa: E402
from smplsim_b200.cfg import make_cfg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 30
OV = {"env.self_collision": True} if os.environ.get("SMPLSIM_SELFCOL") else {}
env = HumanoidBatchB200(make_cfg(env="speed", overrides=OV), num_envs=n, with_aux=False)
env.reset()
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
acts = torch.clamp(torch.randn(warm + 8, n, env.num_actions, generator=g, device="cuda:0") * 0.0821, -1, 1)
for i in range(warm):
    env.step(acts[i]); env.reset_done()
torch.cuda.synchronize(); t0 = time.time()
for i in range(8):
    env.step(acts[warm + i]); env.reset_done()
torch.cuda.synchronize(); dt = (time.time() - t0) / 8
env2 = HumanoidBatchB200(make_cfg(env="speed", overrides=OV), num_envs=min(n, 1024))
env2.reset()
its = []
for i in range(warm + 8):
    env2.step(acts[i][: env2.num_envs]); its.append(env2.solver_iter.float().mean().item()); env2.reset_done()
print(f"kernel v{env.kernel_version} N={n}: {dt * 1e3:.3f} ms/step -> {n / dt / 1e6:.3f} M env-steps/s ; mean extra solves (last substep) {sum(its[-8:]) / 8:.2f}")
