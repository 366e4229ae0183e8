"""Boundary B1: the single-env classes behind the reference's own sampling loop.

tests/golden/b1_sample_worker.npz was recorded by the REFERENCE's Agent.sample_worker + Memory / TrajBatch
(smpl_sim/agents/agent.py:64-109, run unchanged from /root/reference by tests/golden/make_golden_b1.py) driving
smplsim_b200.envs.HumanoidGetup.  Here the same episodes are replayed through the gym protocol with the recorded raw actions:
on the host emulator build (bit-for-bit: same kernels, same env code) and on the GPU (fp32 tolerance, Fall init = 45 substeps of
random actions before every episode)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from conftest import GOLDEN  # noqa: E402
from b1_script import EmuGetup, make_b1_cfg  # noqa: E402


def _close(a, b, atol):
    """Pose part of the getup observation (root height, body positions, 6-d rotations: the first 214 numbers) within atol; the
    joint-velocity block (tens of rad/s while thrashing on the floor, chaotic between two fp32 roundings) within 50 atol."""
    d = np.abs(a - b)
    return d[:214].max() <= atol and d[214:].max() <= 50 * atol


def _replay(env, g, atol):
    T = g["states"].shape[0]
    t = 0
    nep = 0
    while t < T:
        obs, info = env.reset()
        nep += 1
        assert obs.dtype == np.float32 and np.array_equal(info["critic_state"], obs)
        for _ in range(10000):
            state = np.clip(obs, -5.0, 5.0)
            assert _close(state, g["states"][t], atol), (t, np.abs(state - g["states"][t]).max())
            a = np.clip(g["actions"][t], -1.0, 1.0)                     # Agent.preprocess_actions (agent.py:153-161)
            obs, r, died, timed_out, info = env.step(a)
            assert isinstance(r, float) and isinstance(died, bool) and isinstance(timed_out, bool)
            assert abs(r - g["rewards"][t]) <= atol
            assert _close(np.clip(obs, -5.0, 5.0), g["next_states"][t], atol)
            assert int(not (died or timed_out)) == int(g["not_done"][t]) and int(not died) == int(g["not_dead"][t]), t
            t += 1
            if died or timed_out or t >= T:
                break
    return nep


def test_reference_sample_worker_trajectory_replays_on_emulator():
    g = np.load(os.path.join(GOLDEN, "b1_sample_worker.npz"))
    assert (g["not_done"] == 0).sum() >= 3 and np.abs(g["actions"]).max() > 1.0      # several episodes; actions beyond the clip
    nep = _replay(EmuGetup(make_b1_cfg()), g, 0.0)
    assert nep >= 3


@pytest.mark.gpu
def test_reference_sample_worker_trajectory_replays_on_gpu():
    from smplsim_b200.envs import HumanoidGetup
    g = np.load(os.path.join(GOLDEN, "b1_sample_worker.npz"))
    nep = _replay(HumanoidGetup(make_b1_cfg()), g, 5e-2)        # getup: Fall init + up to 7 steps on the floor in fp32 (emulator vs GPU rounding)
    assert nep >= 3


def test_reset_seed_reseeds_the_device_stream():
    env = EmuGetup(make_b1_cfg())
    o1, _ = env.reset(seed=11)
    o2, _ = env.reset(seed=12)
    o3, _ = env.reset(seed=11)
    assert np.array_equal(o1, o3) and not np.array_equal(o1, o2)
