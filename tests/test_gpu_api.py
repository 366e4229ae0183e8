"""GPU tests of the drop-in surfaces: single-env gym protocol (B1), GymVectEnv protocol (B2), motion-lib gather (a12)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from conftest import GOLDEN  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from smplsim_b200.cfg import make_cfg  # noqa: E402

pytestmark = pytest.mark.gpu


def test_single_env_protocol_matches_oracle_getup():
    """BASELINE config 1 (env=getup, num_envs=1) through the B1 classes: shapes, info, rollout vs the oracle."""
    from smplsim_b200.envs import HumanoidGetup, make_env
    cfg = make_cfg(env="getup", seed=0)
    env = make_env(cfg)
    assert isinstance(env, HumanoidGetup)
    assert env.observation_space.shape == (290,) and env.action_space.shape == (69,)
    assert len(env.actuator_names) == 69 and env.actuator_names[0] == "L_Hip_x"
    obs, info = env.reset()
    assert obs.dtype == np.float32 and obs.shape == (290,) and np.array_equal(info["critic_state"], obs)
    om = orc.OracleModel.from_cfg(cfg, seed=env._b.seed)
    e = orc.OracleEnv(om, env_id=0)
    o = e.reset()
    assert np.abs(o - obs).max() < 2e-3
    rng = np.random.default_rng(1)
    for t in range(3):
        a = np.clip(rng.normal(size=69) * 0.1, -1, 1).astype(np.float32)
        obs, r, died, timed_out, info = env.step(a)
        o, rr, te, tr = e.step(a.astype(np.float64))
        assert isinstance(r, float) and isinstance(died, bool) and isinstance(timed_out, bool)
        assert np.abs(o - obs).max() < 5e-3 and abs(r - rr) < 5e-3 and died == te and timed_out == tr
        assert env.render() is None
    d = env.mj_data
    assert d.qpos.shape == (76,) and d.xpos.shape == (25, 3) and d.sensordata.shape == (144,)
    assert set(d.contact.geom1.tolist()) <= {0}
    mm = env.mj_model                                  # the MjModel reads of humanoid_env.py:262-289
    assert mm.nbody == 25 and mm.nv == 75 and mm.body(1).name == "Pelvis" and abs(mm.opt.timestep - 1 / 450) < 1e-12
    assert np.allclose(mm.joint("L_Hip_x").range, [-np.pi, np.pi]) and abs(mm.body_mass.sum() - 71.81) < 0.01


def test_gym_vect_env_autoreset_and_final_observation():
    from smplsim_b200.batched import GymVectEnvB200, HumanoidBatchB200
    cfg = make_cfg(env="speed", overrides={"env.episode_length": 3})
    n = 32
    venv = GymVectEnvB200(HumanoidBatchB200(cfg, num_envs=n, seed=2))
    obs, info = venv.reset()
    assert obs.shape == (n, 292) and venv.num_envs == n
    a = torch.zeros(n, 69, device="cuda:0")
    for t in range(4):
        obs, rew, term, trunc, info = venv.step(a)
        assert not trunc.any() or t == 3
    assert trunc.all()                                   # cur_t = 4 > episode_length = 3 (strict, quirk Q7)
    assert "final_observation" in info and info["final_observation"].shape == (n, 292)
    assert (venv._env.progress_buf == 0).all()           # envs were reset in-stream
    assert abs(obs[0, 0].item() - 0.94) < 1e-6           # returned obs is the post-reset one


def test_motion_gather_matches_reference_frame_index():
    """frame index rule of MotionLibBase.get_motion_state_intervaled (golden from the reference source) + row gather."""
    from smplsim_b200.batched import HumanoidBatchB200
    from smplsim_b200.motion_lib import MotionLibB200, TABLE_KEYS
    g = np.load(os.path.join(GOLDEN, "frame_blend.npz"))
    env = HumanoidBatchB200(make_cfg(env="speed"), num_envs=1)
    m = env.model
    nf = g["num_frames"].astype(np.int32)
    K = nf.shape[0]
    starts = np.concatenate([[0], np.cumsum(nf)[:-1]]).astype(np.int32)
    total = int(nf.sum())
    rng = np.random.default_rng(0)
    widths = dict(qpos=m.nq, qvel=m.nv, xpos=3 * m.nbody, xquat=4 * m.nbody, body_vel=3 * m.nbody, body_ang_vel=3 * m.nbody, dof_pos=m.nu, dof_vel=m.nu)
    tabs = {k: rng.normal(size=(total, widths[k])).astype(np.float32) for k in TABLE_KEYS}
    tabs.update(motion_num_frames=nf, motion_dt=g["dt"].astype(np.float32), motion_lengths=g["motion_len"].astype(np.float32), length_starts=starts)
    lib = MotionLibB200(env, tabs)
    ids = torch.arange(K, dtype=torch.int32)
    out = lib.get_motion_state_intervaled(ids, torch.as_tensor(g["time"], dtype=torch.float32))
    fr = out["frame_idx"].cpu().numpy()
    ref = g["frame_idx"] + starts
    # fp32 vs fp64 time arithmetic may differ by one frame exactly at frame boundaries
    bad = np.nonzero(fr != ref)[0]
    for i in bad:
        tt = g["time"][i] / g["dt"][i]
        assert abs(tt - round(tt)) < 1e-3 and abs(fr[i] - ref[i]) == 1, (i, fr[i], ref[i], tt)
    assert len(bad) <= 4
    assert np.array_equal(out["qpos"].cpu().numpy(), tabs["qpos"][fr])
    assert np.array_equal(out["xquat"].cpu().numpy().reshape(K, -1), tabs["xquat"][fr])
    assert np.array_equal(out["root_pos"].cpu().numpy(), tabs["xpos"][fr][:, :3])


def test_mocap_state_init_from_synthetic_motion_table():
    """Config-4 plumbing: synthetic clip table -> gather -> StateInit.MoCap reset -> one step stays finite."""
    from smplsim_b200.batched import HumanoidBatchB200
    from smplsim_b200.motion_lib import MotionLibB200, synthetic_tables
    cfg = make_cfg(env="speed", overrides={"env.self_obs_v": 2, "robot.create_vel_sensors": True})
    n = 64
    env = HumanoidBatchB200(cfg, num_envs=n)
    lib = MotionLibB200(env, synthetic_tables(env, num_clips=8, frames=60))
    ids = torch.arange(n, dtype=torch.int32) % 8
    st = lib.get_motion_state_intervaled(ids, torch.rand(n) * 1.5)
    obs = env.reset(None, init_mode=2, qpos0=st["qpos"], qvel0=st["qvel"])
    assert torch.allclose(env.qpos[:, 7:], st["dof_pos"], atol=1e-6) and torch.isfinite(obs).all()
    xp = env.xpos
    assert torch.allclose(xp, st["xpos"], atol=1e-4)
    env.step(torch.zeros(n, 69, device="cuda:0"))
    assert torch.isfinite(env.obs_buf).all() and torch.isfinite(env.qpos).all()


def test_smplx_env_step_matches_oracle():
    """BASELINE config 5 model (52 bodies, 159 dofs): getup env step vs the oracle."""
    from smplsim_b200.batched import HumanoidBatchB200
    cfg = make_cfg(env="getup", robot="smplx_humanoid", seed=5)
    n = 8
    env = HumanoidBatchB200(cfg, num_envs=n, seed=5)
    assert env.num_obs == 626 and env.num_actions == 153
    obs0 = env.reset().cpu().numpy().copy()
    om = orc.OracleModel.from_cfg(cfg, seed=env.seed)
    rng = np.random.default_rng(3)
    act = np.clip(rng.normal(size=(n, 153)) * 0.1, -1, 1)
    obs, rew, term, trunc = [x.cpu().numpy() for x in env.step(torch.as_tensor(act, dtype=torch.float32))]
    for i in range(n):
        e = orc.OracleEnv(om, env_id=i)
        o0 = e.reset()
        assert np.abs(o0 - obs0[i]).max() < 5e-3
        o, r, te, tr = e.step(act[i])
        assert np.abs(o - obs[i]).max() < 1e-3 and abs(r - rew[i]) < 1e-3      # one env step = 15 substeps from the Fall-init pose


def test_gae_matches_reference_golden():
    """estimate_advantages CUDA reverse scan vs the reference's Python loop (tests/golden/gae.npz)."""
    from smplsim_b200.learning import estimate_advantages
    g = np.load(os.path.join(GOLDEN, "gae.npz"))
    t = lambda k: torch.as_tensor(g[k], dtype=torch.float32, device="cuda:0")  # noqa: E731
    # the reference scans ONE flat batch: the last sample of a column that is truncated but not dead bootstraps from the next
    # stored sample, i.e. the first value of the following column (0 after the very last one)
    nxt = torch.cat([t("values")[0, 1:], torch.zeros(1, device="cuda:0")])
    adv, ret = estimate_advantages(t("rewards"), t("not_done"), t("not_dead"), t("values"), 0.99, 0.95, next_value=nxt)
    assert np.abs(adv.cpu().numpy() - g["advantages"]).max() < 2e-5
    assert np.abs(ret.cpu().numpy() - g["returns"]).max() < 2e-5


def test_batched_sampler_shapes_and_masks():
    from smplsim_b200.batched import HumanoidBatchB200
    from smplsim_b200.learning import BatchedSampler, estimate_advantages
    cfg = make_cfg(env="speed", overrides={"env.episode_length": 5})
    n, T = 64, 8
    env = HumanoidBatchB200(cfg, num_envs=n, seed=3)
    torch.manual_seed(0)
    W = torch.randn(env.num_obs, env.num_actions, device="cuda:0") * 0.01
    sampler = BatchedSampler(env, lambda o: o @ W)
    b = sampler.sample(T)
    assert b["states"].shape == (T, n, 292) and b["actions"].shape == (T, n, 69) and b["rewards"].shape == (T, n)
    assert b["states"].abs().max() <= 5.0            # observations clipped; the RAW action is recorded (agent.py:81-93), see test_sampler_cpu.py
    nd = b["not_done"].cpu().numpy()
    assert (nd[5] == 0).all() and (nd[:5] == 1).all()            # cur_t = 6 > 5 at the 6th step: every env truncates together
    assert torch.equal(b["states"][6, :, 0], torch.full((n,), 0.94, device="cuda:0"))   # next state after the in-stream reset
    adv, ret = estimate_advantages(b["rewards"], b["not_done"], b["not_dead"], torch.zeros(T, n, device="cuda:0"), 0.99, 0.95)
    assert torch.isfinite(adv).all() and abs(adv.mean().item()) < 1e-4 and abs(adv.std().item() - 1) < 1e-3


def test_two_ppo_epochs_getup_plumbing():
    """BASELINE config 1 as a plumbing check: env=getup, two PPO epochs (sample -> GAE -> update) complete through the
    batched surface with obs / action shapes (290,) / (69,) and finite losses; the policy parameters move."""
    from smplsim_b200.batched import HumanoidBatchB200
    from smplsim_b200.learning import BatchedSampler
    from smplsim_b200.ppo import PolicyGaussian, PPOLearner, Value
    env = HumanoidBatchB200(make_cfg(env="getup"), num_envs=64, seed=1)
    assert env.num_obs == 290 and env.num_actions == 69
    torch.manual_seed(0)
    policy = PolicyGaussian(290, 69, [64, 32]).to("cuda:0")
    value = Value(290, [64, 32]).to("cuda:0")
    learner = PPOLearner(policy, value, opt_num_epochs=2)
    w0 = policy.action_mean.weight.detach().clone()

    def act(o):
        policy.eval()
        return policy.select_action(o)

    sampler = BatchedSampler(env, act)
    for _ in range(2):
        info = learner.update(sampler.sample(6))
        assert all(np.isfinite(v) for v in info.values()), info
    assert int(policy.norm.n) == 2 * 2 * 6 * 64 and not torch.equal(w0, policy.action_mean.weight)


def test_amass_loader_tables_feed_motion_lib_and_cuda_fk():
    """AMASS-style clip -> amass_tables -> MotionLibB200: the loader's numpy FK agrees with the CUDA kinematics kernel on the same
    qpos (xpos 1e-5, xquat up to sign), and a MoCap reset from the gathered state reproduces it."""
    from smplsim_b200.batched import HumanoidBatchB200
    from smplsim_b200.motion_lib import MotionLibB200
    from smplsim_b200.motion_loader import amass_tables
    g = np.load(os.path.join(GOLDEN, "motion_fk.npz"))
    env = HumanoidBatchB200(make_cfg(env="speed"), num_envs=16, seed=0)
    tb = amass_tables(env.model, {"clip": {"pose_aa": g["pose_aa"], "trans": g["trans"], "fps": int(g["fps"])}}, fix_height="geom")
    xp, xq = env.kinematics(torch.as_tensor(tb["qpos"], device="cuda:0"))
    assert np.abs(xp.cpu().numpy().reshape(-1, 72) - tb["xpos"]).max() < 2e-5
    a, b = xq.cpu().numpy().reshape(-1, 24, 4), tb["xquat"].reshape(-1, 24, 4)
    assert np.minimum(np.abs(a - b).max(-1), np.abs(a + b).max(-1)).max() < 2e-5
    lib = MotionLibB200(env, tb)
    ids = torch.zeros(16, dtype=torch.int32, device="cuda:0")
    tm = torch.linspace(0, float(tb["motion_lengths"][0]), 16, device="cuda:0")
    st = lib.get_motion_state_intervaled(ids, tm)
    env.reset(init_mode=2, qpos0=st["qpos"], qvel0=st["qvel"])
    assert torch.allclose(env.qpos[:, :3], st["qpos"][:, :3], atol=1e-6) and torch.isfinite(env.obs_buf).all()
    assert (env.qpos[:, 2] > 0.5).all()
