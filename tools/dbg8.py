import sys, os, torch, numpy as np
sys.path.insert(0, '/root/repo')
from smplsim_b200.batched import HumanoidBatchB200
from smplsim_b200.cfg import make_cfg
d = np.load('/root/repo/tests/golden/regress_default_step_case1.npz')
e1 = HumanoidBatchB200(make_cfg(env="speed"), num_envs=1, seed=0)
e1.reset()
e1.task_target[0, 0] = float(d['task_target'][0]); e1.task_change_step.fill_(10**6)
a = torch.as_tensor(d['action'], device="cuda:0")[None]
e1.step(a)
print(os.environ.get("SMPLSIM_KERNEL"), os.environ.get("SMPLSIM_WARMSET"), "kernel", e1.kernel_version, "max|qvel|", e1.qvel.abs().max().item(), "z", e1.qpos[0, 2].item(), "iters", e1.solver_iter.item())
