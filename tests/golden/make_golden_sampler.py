#!/usr/bin/env python
"""Generate tests/golden/sampler.npz by running the REFERENCE's own Agent.sample_worker (smpl_sim/agents/agent.py:64-109) with its
Memory / TrajBatch (smpl_sim/learning/memory.py, trajbatch.py) against a scripted env and a scripted deterministic policy
(build container only).  The unbound methods run on a namespace carrying exactly the attributes they read -- the Agent ctor
(spawns loggers / networks) is not used.  The script (tests/sampler_script.py) defines, per env index k: episode lengths,
death / time-out flags, observations with components beyond the +-5 clip, rewards; the policy is a fixed linear map with gain
3 so that actions saturate the [-1, 1] clip.  For every env the reference worker records its Memory; the first T samples of
each are the golden (the worker finishes the episode that crosses T, the batched sampler stops at T)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
import make_golden as MG  # noqa: E402,F401  (installs the third-party stubs, puts /root/reference on sys.path)
from sampler_script import D, A, K, T, W, ScriptEnv  # noqa: E402

AG = MG._import_with_stubs(lambda: __import__("smpl_sim.agents.agent", fromlist=["x"]))
from smpl_sim.learning.memory import Memory  # noqa: E402,F401
from smpl_sim.learning.trajbatch import TrajBatch  # noqa: E402


class _Logger:
    def __init__(self):
        self.num_steps = 0

    def start_episode(self, env): pass
    def step(self, env, reward, info): self.num_steps += 1
    def end_episode(self, env): pass
    def end_sampling(self): pass


class _Policy:
    type = "gaussian"

    def select_action(self, x, mean_action):
        return (x @ torch.as_tensor(W.T, dtype=x.dtype)) * 3.0


def main():
    out = {}
    for k in range(K):
        env = ScriptEnv(k)
        ag = types.SimpleNamespace(env=env, policy_net=_Policy(), logger_rl_cls=_Logger, mean_action=True, noise_rate=1.0, dtype=torch.float32,
                                   np_dtype=np.float32, headless=True, clip_obs=True, obs_low=-5.0, obs_high=5.0, clip_actions=True,
                                   actions_low=-np.ones(A, np.float32), actions_high=np.ones(A, np.float32))
        for name in ("seed_worker", "pre_sample", "push_memory", "preprocess_obs", "preprocess_actions", "sample_worker"):
            setattr(ag, name, types.MethodType(getattr(AG.Agent, name), ag))
        memory, logger = ag.sample_worker(0, None, T)
        tb = TrajBatch([memory])
        assert tb.states.shape[0] >= T
        for key in ("states", "actions", "not_done", "not_dead", "next_states", "rewards"):
            out[f"{key}_{k}"] = np.asarray(getattr(tb, key))[:T]
        out[f"env_actions_{k}"] = np.asarray(env.actions_seen)[:T]
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **out)
    print("sampler.npz:", {k: v.shape for k, v in out.items() if k.endswith("_0")})


if __name__ == "__main__":
    main()
