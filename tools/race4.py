import sys, torch
sys.path.insert(0, '/root/repo')
from smplsim_b200.batched import HumanoidBatchB200
from smplsim_b200.cfg import make_cfg
env = HumanoidBatchB200(make_cfg(env="speed"), num_envs=64, seed=7)
env.reset()
a = torch.zeros(64, 69, device="cuda:0"); a[:, 3] = 0.2
env.step(a); env.step(a)
torch.cuda.synchronize()
print("done")
