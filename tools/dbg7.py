import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from smplsim_b200.batched import HumanoidBatchB200
from smplsim_b200.cfg import make_cfg
n = 64
g = torch.Generator(device="cuda:0"); g.manual_seed(1)
acts = torch.clamp(torch.randn(12, n, 69, generator=g, device="cuda:0") * 0.0821, -1, 1)
envA = HumanoidBatchB200(make_cfg(env="speed"), num_envs=n, seed=0)
envA.reset()
for t in range(5): envA.step(acts[t])
mask = torch.zeros(n, dtype=torch.uint8, device="cuda:0"); mask[3] = 1; mask[40] = 1
envA.reset(mask)
qa = envA.qpos.clone()
envA.step(acts[5])
envB = HumanoidBatchB200(make_cfg(env="speed"), num_envs=n, seed=0)
envB.reset()
envB.step(acts[5])
for j in (3, 40):
    print("env", j, "after masked reset qpos z/quat", qa[j, 2:7].tolist(), "| step diff vs fresh:", (envA.qpos[j] - envB.qpos[j]).abs().max().item(), "max|qvel|", envA.qvel[j].abs().max().item(), envB.qvel[j].abs().max().item())
print("unreset env 0 qvel max", envA.qvel[0].abs().max().item())
