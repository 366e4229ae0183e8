import sys, os, torch, numpy as np
sys.path.insert(0, '/root/repo')
from smplsim_b200.batched import HumanoidBatchB200
from smplsim_b200.cfg import make_cfg
n = 1024
env = HumanoidBatchB200(make_cfg(env="speed"), num_envs=n, seed=0)
env.reset()
g = torch.Generator(device="cuda:0"); g.manual_seed(0)
names = ("qpos", "qvel", "qpos_fwd", "qvel_fwd", "qacc_warm", "task_target", "task_change_step", "progress_buf", "recovery", "rng_counter")
for t in range(80):
    a = torch.clamp(torch.randn(n, 69, generator=g, device="cuda:0") * 0.0821, -1, 1)
    prev = {k: getattr(env, k).clone() for k in names}
    obs, rew, term, trunc = env.step(a)
    bad = ((~torch.isfinite(obs).all(dim=1)) | (env.qvel.abs().max(dim=1).values > 100)).nonzero().flatten()
    if len(bad):
        j = int(bad[0])
        print("step", t, "bad envs", bad.tolist()[:8], "iters", env.solver_iter[j].item(), "mask", hex(int(env.contact_mask[j].item())))
        out = {k: prev[k][j].cpu().numpy() for k in names}; out["action"] = a[j].cpu().numpy()
        out["after_qpos"] = env.qpos[j].cpu().numpy(); out["after_qvel"] = env.qvel[j].cpu().numpy()
        os.makedirs("gpurun_out", exist_ok=True)
        np.savez("gpurun_out/nan_case.npz", **out)
        # replay the same env alone, substep by substep through env.step with 1-substep cfg is not identical; use mj_step-free replay: single env batch
        e1 = HumanoidBatchB200(make_cfg(env="speed"), num_envs=1, seed=0)
        for k in names: getattr(e1, k).copy_(prev[k][j:j+1])
        o1, _, _, _ = e1.step(a[j:j+1])
        print("replay alone finite:", bool(torch.isfinite(o1).all()), "qvel max", e1.qvel.abs().max().item())
        break
    env.reset_done()
print("done", t)
