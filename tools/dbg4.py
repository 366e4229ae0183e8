import os, sys, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from smplsim_b200.batched import HumanoidBatchB200
from smplsim_b200.cfg import make_cfg
cfg = make_cfg(env="speed")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
env = HumanoidBatchB200(cfg, num_envs=n, seed=7)
env.reset(); env.task_target[:, 0] = 1.5
a = torch.zeros(n, 69, device="cuda:0"); a[:, 3] = 0.2
for s in range(4):
    env.step(a)
    q = env.qpos.cpu().numpy(); o = env.obs_buf.cpu().numpy(); it = env.solver_iter.cpu().numpy()
    for nm in ("ctrl", "qacc", "qvel", "qpos_fwd", "qvel_fwd", "qacc_warm", "xpos", "body_linvel"):
        t_ = getattr(env, nm).cpu().numpy().reshape(n, -1); d_ = np.abs(t_ - t_[0:1]).max(axis=1)
        print("   ", nm, "max diff", d_.max(), "n", int((d_ > 0).sum()))
    dq = np.abs(q - q[0:1]).max(axis=1); do = np.abs(o - o[0:1]).max(axis=1)
    bad = np.nonzero(dq > 0)[0]
    print("step", s, "n differing", len(bad), "max dq", dq.max(), "max dobs", do.max(), "first bad", bad[:16], "iters uniq", np.unique(it))
    if len(bad):
        j = bad[0]; dd = np.abs(q[j] - q[0]); print("  worst qpos idx", np.argsort(-dd)[:6], dd[np.argsort(-dd)[:6]])
