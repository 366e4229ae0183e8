import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from smplsim_b200.batched import HumanoidBatchB200
from smplsim_b200.cfg import make_cfg
n = 256
cfg2 = make_cfg(env="getup", robot="smplx_humanoid")
env2 = HumanoidBatchB200(cfg2, num_envs=n, seed=5)
obs = env2.reset()
print("smplx getup reset: nan envs", int((~torch.isfinite(obs).all(dim=1)).sum()), "qpos nan", int((~torch.isfinite(env2.qpos).all(dim=1)).sum()), "kernel", env2.kernel_version)
env = HumanoidBatchB200(make_cfg(env="speed"), num_envs=1024, seed=0)
env.reset()
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
nreset = 0
for t in range(70):
    a = torch.clamp(torch.randn(1024, 69, generator=g, device="cuda:0") * 0.0821, -1, 1)
    obs, rew, term, trunc = env.step(a)
    bad = (~torch.isfinite(obs).all(dim=1))
    if bad.any() or rew.min() < 0 or rew.max() > 1 or (env.progress_buf > 301).any():
        print("step", t, "nonfinite envs", int(bad.sum()), "rew range", rew.min().item(), rew.max().item(), "progress max", env.progress_buf.max().item())
        break
    nreset += int(env.reset_buf.sum()); env.reset_done()
print("rollout done t", t, "nreset", nreset, "progress max", env.progress_buf.max().item())
