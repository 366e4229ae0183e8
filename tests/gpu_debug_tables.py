#!/usr/bin/env python
"""Print CUDA-vs-oracle error tables (development aid next to the parity tests -- it uses the oracle, so it lives under tests/;
run under gpurun: python tests/gpu_debug_tables.py)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import oracle as orc  # noqa: E402
from smplsim_b200.batched import HumanoidBatchB200  # noqa: E402
from util_states import airborne_states, make_models, relerr, rollout_states  # noqa: E402


def T(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32, device="cuda:0")


def main():
    cfg, om = make_models(control_mode="torque")
    m = om.model
    env = HumanoidBatchB200(cfg, num_envs=16)
    print("smem bytes/env", env.smem_bytes_per_env(), "kernel version", env.kernel_version, "schedule steps", env.schedule_steps)
    q, v = airborne_states(m, 16, seed=3)
    xp, xq = env.kinematics(T(q))
    e = orc.OracleEnv(om)
    errs = []
    for i in range(16):
        e.qpos[:] = q[i]; e.kinematics()
        errs.append(np.abs(xp[i].cpu().numpy() - e.xpos).max())
    print("FK max err", max(errs))
    ctrl = np.random.default_rng(5).uniform(-50, 50, (16, m.nu))
    env.set_state(T(q), T(v)); env.mj_step(T(ctrl), 1)
    torch.cuda.synchronize()
    ga, gv, gq = env.qacc.cpu().numpy(), env.qvel.cpu().numpy(), env.qpos.cpu().numpy()
    for i in range(4):
        e.qpos[:] = q[i]; e.qvel[:] = v[i]; e.qacc_warm[:] = 0; e.ctrl[:] = ctrl[i]; e.mj_step()
        print(f"airborne {i}: qacc {relerr(ga[i], e.qacc):.2e} qvel {relerr(gv[i], e.qvel):.2e} qpos {relerr(gq[i], e.qpos):.2e}")
        if relerr(ga[i], e.qacc) > 1e-2:
            print("  gpu qacc[:12]", ga[i][:12]); print("  ref qacc[:12]", e.qacc[:12])
    for mode in ("torque", "uhc_pd"):
        cfg, om = make_models(control_mode=mode)
        n = 48
        q, v, w = rollout_states(make_models(control_mode="uhc_pd")[1], n, seed=11)
        env = HumanoidBatchB200(cfg, num_envs=n)
        ctrl = np.random.default_rng(2).uniform(-80, 80, (n, m.nu))
        env.set_state(T(q), T(v)); env.qacc_warm.copy_(T(w)); env.mj_step(T(ctrl), 1)
        torch.cuda.synchronize()
        ga, gv, gq = env.qacc.cpu().numpy(), env.qvel.cpu().numpy(), env.qpos.cpu().numpy()
        gm = env.contact_mask.cpu().numpy(); it = env.solver_iter.cpu().numpy()
        e = orc.OracleEnv(om)
        worst = 0
        for i in range(n):
            e.qpos[:] = q[i]; e.qvel[:] = v[i]; e.qacc_warm[:] = w[i]; e.ctrl[:] = ctrl[i]; e.mj_step()
            r = (relerr(ga[i], e.qacc), relerr(gv[i], e.qvel), relerr(gq[i], e.qpos))
            worst = max(worst, r[1])
            if i < 12 or r[1] > 1e-4:
                print(f"{mode} contact {i}: ncon {e.ncon} it gpu/ref {it[i]}/{e.solver_iter} qacc {r[0]:.2e} qvel {r[1]:.2e} qpos {r[2]:.2e} "
                      f"mask {'==' if int(gm[i]) == e.contact_mask else '!='}")
        print(mode, "worst qvel relerr", worst)
    # env step
    for task in ("speed", "getup"):
        cfg, om = make_models(env=task, seed=3)
        n = 8
        env = HumanoidBatchB200(cfg, num_envs=n, seed=3)
        o0 = env.reset().cpu().numpy().copy()
        oes = [orc.OracleEnv(om, env_id=i) for i in range(n)]
        for i, e in enumerate(oes):
            o = e.reset()
            print(task, "reset obs err", i, np.abs(o - o0[i]).max(), "target", env.task_target[i, 0].item(), e.target[0], env.task_change_step[i].item(), e.change_step)
        rng = np.random.default_rng(9)
        for t in range(3):
            act = np.clip(rng.normal(size=(n, m.nu)) * 0.1, -1, 1)
            obs, rew, term, trunc = [x.cpu().numpy() for x in env.step(T(act))]
            for i, e in enumerate(oes[:4]):
                o, r, te, tr = e.step(act[i])
                print(task, f"step {t} env {i}: obs err {np.abs(o - obs[i]).max():.2e} rew {rew[i]:.5f}/{r:.5f} term {term[i]}/{te} iters {env.solver_iter[i].item()}")
    # timing
    cfg, om = make_models(env="speed")
    for n in (4096, 16384):
        env = HumanoidBatchB200(cfg, num_envs=n)
        env.reset()
        a = torch.zeros(n, m.nu, device="cuda:0")
        for _ in range(3):
            env.step(a); env.reset_done()
        torch.cuda.synchronize(); t0 = time.time()
        g = torch.Generator(device="cuda:0"); g.manual_seed(0)
        for _ in range(40):
            env.step(torch.clamp(torch.randn(n, m.nu, generator=g, device="cuda:0") * 0.0821, -1, 1)); env.reset_done()
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(10):
            env.step(a); env.reset_done()
        torch.cuda.synchronize(); dt = (time.time() - t0) / 10
        print(f"N={n}: {dt * 1e3:.2f} ms/step -> {n / dt / 1e6:.3f} M env-steps/s")


if __name__ == "__main__":
    main()
