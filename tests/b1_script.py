"""Shared by tests/golden/make_golden_b1.py (the reference's Agent.sample_worker driving the B1 env on the emulator) and
tests/test_b1_protocol.py (replay on the emulator and on the GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
from smplsim_b200.cfg import make_cfg  # noqa: E402
from smplsim_b200.envs import HumanoidGetup  # noqa: E402


def make_b1_cfg():
    return make_cfg(env="getup", seed=3, overrides={"env.episode_length": 6, "env.recovery_steps": 3})


def policy_matrix(obs_dim, act_dim):
    rng = np.random.default_rng(5)
    return (rng.normal(size=(obs_dim, act_dim)) * 0.05).astype(np.float32)


class EmuGetup(HumanoidGetup):
    """HumanoidGetup with its num_envs=1 handle bound to the emulator build (test infrastructure)."""

    def _make_batch(self, cfg, device):
        import emu_env
        return emu_env.EmuBatch(cfg, 1, seed=int(cfg.get("seed", 0)))

    def __init__(self, cfg):
        super().__init__(cfg, device="cpu")
