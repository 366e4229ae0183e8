#!/usr/bin/env python
"""Generate tests/golden/b1_sample_worker.npz: the REFERENCE's own sampling loop -- Agent.sample_worker with Memory / TrajBatch
(smpl_sim/agents/agent.py:64-109) -- run UNCHANGED against this repo's single-env class HumanoidGetup (boundary B1,
smplsim_b200/envs.py), the env backed by the host SIMT emulator build of the kernels (tests/emu; build container only, no GPU).
The recorded trajectory (clipped observations, raw policy actions, rewards, not_done / not_dead, episode boundaries over several
Fall-init resets) is what tests/test_b1_protocol.py replays: on the emulator (`-m "not gpu"`, bit-for-bit) and on the GPU."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "emu"))
from b1_script import make_b1_cfg, EmuGetup, policy_matrix  # noqa: E402  (before the stubs: smplsim_b200.envs must not see a mock gymnasium)
import make_golden as MG  # noqa: E402,F401  (installs the third-party stubs, puts /root/reference on sys.path)

AG = MG._import_with_stubs(lambda: __import__("smpl_sim.agents.agent", fromlist=["x"]))
from smpl_sim.learning.trajbatch import TrajBatch  # noqa: E402


class _Logger:
    def __init__(self):
        self.num_steps = 0; self.num_episodes = 0

    def start_episode(self, env): self.num_episodes += 1
    def step(self, env, reward, info): self.num_steps += 1
    def end_episode(self, env): pass
    def end_sampling(self): pass


def main():
    cfg = make_b1_cfg()
    env = EmuGetup(cfg)
    W = policy_matrix(env.get_obs_size(), env.get_action_size())

    class _Policy:
        type = "gaussian"

        def select_action(self, x, mean_action):
            a = torch.tanh(x @ torch.as_tensor(W, dtype=x.dtype)) * 0.15
            a[..., 3] = 1.3; a[..., 40] = -1.2                                           # two components beyond the [-1, 1] clip on purpose
            return a

    ag = types.SimpleNamespace(env=env, policy_net=_Policy(), logger_rl_cls=_Logger, mean_action=True, noise_rate=1.0, dtype=torch.float32,
                               np_dtype=np.float32, headless=True, clip_obs=True, obs_low=-5.0, obs_high=5.0, clip_actions=True,
                               actions_low=env.action_space.low, actions_high=env.action_space.high)
    for name in ("seed_worker", "pre_sample", "push_memory", "preprocess_obs", "preprocess_actions", "sample_worker"):
        setattr(ag, name, types.MethodType(getattr(AG.Agent, name), ag))
    memory, logger = ag.sample_worker(0, None, 26)
    tb = TrajBatch([memory])
    out = {k: np.asarray(getattr(tb, k)) for k in ("states", "actions", "not_done", "not_dead", "next_states", "rewards")}
    np.savez_compressed(os.path.join(HERE, "b1_sample_worker.npz"), **out)
    print("b1_sample_worker.npz:", {k: v.shape for k, v in out.items()}, "episodes", logger.num_episodes, "not_done zeros", int((out["not_done"] == 0).sum()))


if __name__ == "__main__":
    main()
