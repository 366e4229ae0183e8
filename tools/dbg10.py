import sys, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle import oracle as orc
from util_states import make_models, rollout_states
from smplsim_b200.cfg import make_cfg
from smplsim_b200.batched import HumanoidBatchB200
def _t(x): return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32, device="cuda:0")
mode = "uhc_pd"
cfg, om = make_models(env="speed", control_mode=mode)
m = om.model; n = 48
q, v, w = rollout_states(make_models(control_mode="uhc_pd")[1], n, seed=21)
rng = np.random.default_rng(4)
act = np.clip(rng.normal(size=(n, m.nu)) * 0.4, -1, 1)
qs = q.copy(); qs[:, 7:] += rng.normal(size=(n, m.nu)) * 0.003
cfg1 = make_cfg(env="speed", overrides={"env.control_mode": mode, "env.control_frequency_inv": 1})
env = HumanoidBatchB200(cfg1, n)
env.set_state(_t(q), _t(v)); env.qpos_fwd.copy_(_t(qs)); env.qvel_fwd.copy_(_t(v * 0.9)); env.qacc_warm.copy_(_t(w))
env.task_change_step.fill_(10000)
env.step(_t(act))
st = env.status.cpu().numpy()
print("status", st)
for i in np.nonzero(st)[0]:
    e = orc.OracleEnv(om)
    e.qpos[:] = qs[i]; e.qvel[:] = v[i] * 0.9; e.forward()
    e.qpos[:] = q[i]; e.qvel[:] = v[i]
    tau = e.compute_torque(act[i]); e.ctrl[:] = tau; e.warn = 0
    print(i, "max|v|", np.abs(v[i]).max(), "max|tau|", np.abs(tau).max())
    e.mj_step()
    print("  oracle warn", e.warn, "max|qacc|", np.abs(e.qacc).max(), "iters", orc.lib().orc_solver_iter(e.ptr), "ncon", orc.lib().orc_ncon(e.ptr))
