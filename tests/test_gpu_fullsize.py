"""BASELINE.json configs at their full per-GPU sizes, checked through size-independent properties (no oracle at this scale):
determinism / lane independence, closed-form free fall, yaw + translation invariance of the proprioception, finite rollouts."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from smplsim_b200.cfg import make_cfg  # noqa: E402

pytestmark = pytest.mark.gpu


def _env(cfg, n, **kw):
    from smplsim_b200.batched import HumanoidBatchB200
    return HumanoidBatchB200(cfg, num_envs=n, device="cuda:0", **kw)


def _yaw_quat(ang):
    return torch.stack([torch.cos(ang / 2), torch.zeros_like(ang), torch.zeros_like(ang), torch.sin(ang / 2)], -1)


def _qmul(a, b):
    w1, x1, y1, z1 = a.unbind(-1); w2, x2, y2, z2 = b.unbind(-1)
    return torch.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                        w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], -1)


def test_cfg3_16384_reach_obs_v2_heading_invariance():
    """config 3 (16 384 envs, env=reach, self_obs_v=2): the self observation is invariant under a yaw rotation and an xy
    translation of the whole scene (the reference's own 'heading invariance check', humanoid_env.py:497-504)."""
    n = 16384
    cfg = make_cfg(env="reach", overrides={"env.self_obs_v": 2, "robot.create_vel_sensors": True})
    env = _env(cfg, n, seed=1)
    m = env.model
    g = torch.Generator(device="cuda:0"); g.manual_seed(0)
    q = torch.zeros(n, m.nq, device="cuda:0")
    q[:, 2] = 0.9
    quat = torch.randn(n, 4, generator=g, device="cuda:0"); q[:, 3:7] = quat / quat.norm(dim=1, keepdim=True)
    q[:, 7:] = (torch.rand(n, m.nu, generator=g, device="cuda:0") - 0.5) * 2.0
    lin = torch.randn(n, m.nbody, 3, generator=g, device="cuda:0"); ang = torch.randn(n, m.nbody, 3, generator=g, device="cuda:0")
    xp, xq = env.kinematics(q)
    o1 = env.self_obs(2, xp, xq, linvel=lin, angvel=ang)
    yaw = torch.rand(n, generator=g, device="cuda:0") * 6.2831853
    yq = _yaw_quat(yaw)
    c, s = torch.cos(yaw), torch.sin(yaw)
    R = torch.zeros(n, 3, 3, device="cuda:0"); R[:, 0, 0] = c; R[:, 0, 1] = -s; R[:, 1, 0] = s; R[:, 1, 1] = c; R[:, 2, 2] = 1
    q2 = q.clone()
    q2[:, 3:7] = _qmul(yq, q[:, 3:7])
    q2[:, 0] = 7.0; q2[:, 1] = -3.0
    xp2, xq2 = env.kinematics(q2)
    o2 = env.self_obs(2, xp2, xq2, linvel=torch.einsum("nij,nbj->nbi", R, lin), angvel=torch.einsum("nij,nbj->nbi", R, ang))
    assert o1.shape == (n, 358)
    assert (o1 - o2).abs().max().item() < 2e-4      # fp32: absolute xy offsets of 7 m, velocities O(3)


def test_cfg4_8192_motion_feed_mocap_rollout():
    """config 4 per-GPU shard (65 536 / 8 = 8 192 envs, obs v2): synthetic motion table -> gather -> MoCap reset -> steps; envs
    that share (clip, time) evolve bit-identically (lane / CTA independence at full size)."""
    from smplsim_b200.motion_lib import MotionLibB200, synthetic_tables
    n = 8192
    cfg = make_cfg(env="speed", overrides={"env.self_obs_v": 2, "robot.create_vel_sensors": True})
    env = _env(cfg, n, seed=4, with_aux=False)
    lib = MotionLibB200(env, synthetic_tables(env, num_clips=16, frames=90))
    ids = (torch.arange(n, device="cuda:0") % 16).to(torch.int32)
    times = ((torch.arange(n, device="cuda:0") % 16) * 0.1).float()
    st = lib.get_motion_state_intervaled(ids, times)
    env.reset(None, init_mode=2, qpos0=st["qpos"], qvel0=st["qvel"])
    env.task_target[:, 0] = 1.0
    env.task_change_step.fill_(10_000)
    a = torch.zeros(n, 69, device="cuda:0")
    for _ in range(3):
        env.step(a)
    assert torch.isfinite(env.obs_buf).all()
    q = env.qpos.view(n // 16, 16, -1)
    assert torch.equal(q, q[0:1].expand_as(q))


def test_cfg5_4096_smplx_getup_free_fall_and_rollout():
    """config 5 per-GPU shard (32 768 / 8 = 4 096 SMPL-X envs): closed-form free fall, then a Fall-init getup rollout."""
    n = 4096
    cfg = make_cfg(env="getup", robot="smplx_humanoid", overrides={"env.control_mode": "torque"})
    env = _env(cfg, n, seed=5, with_aux=False)
    m = env.model
    q = torch.zeros(n, m.nq, device="cuda:0"); q[:, 2] = 6.0; q[:, 3] = 1.0
    env.set_state(q, torch.zeros(n, m.nv, device="cuda:0"))
    k, h = 20, m.timestep
    env.mj_step(torch.zeros(n, m.nu, device="cuda:0"), k)
    assert (env.qvel[:, 2] + 9.81 * k * h).abs().max().item() < 1e-5
    assert (env.qpos[:, 2] - (6.0 - 9.81 * h * h * k * (k + 1) / 2)).abs().max().item() < 1e-5
    cfg2 = make_cfg(env="getup", robot="smplx_humanoid")
    env2 = _env(cfg2, n, seed=5, with_aux=False)
    obs = env2.reset()
    assert obs.shape == (n, 626) and torch.isfinite(obs).all()
    env2.step(torch.zeros(n, m.nu, device="cuda:0"))
    assert torch.isfinite(env2.obs_buf).all() and (env2.recovery == 59).all()


def test_cfg2_4096_long_rollout_statistics():
    """config 2: 4 096 envs, 60 steps with the bench's action distribution and in-stream autoreset: everything stays finite,
    episodes end (fall -> terminate -> reset), rewards stay in [0, 1]."""
    n = 4096
    env = _env(make_cfg(env="speed"), n, seed=0, with_aux=False)
    env.reset()
    g = torch.Generator(device="cuda:0"); g.manual_seed(0)
    nreset = 0
    for _ in range(70):
        a = torch.clamp(torch.randn(n, 69, generator=g, device="cuda:0") * 0.0821, -1, 1)
        obs, rew, term, trunc = env.step(a)
        nreset += int(env.reset_buf.sum().item())
        env.reset_done()
        assert torch.isfinite(obs).all() and rew.min().item() >= 0.0 and rew.max().item() <= 1.0
    assert nreset > 0
    assert (env.progress_buf <= 301).all()
