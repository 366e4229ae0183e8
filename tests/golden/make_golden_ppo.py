#!/usr/bin/env python
"""Generate tests/golden/ppo_update.npz by running the REFERENCE's own learner code on a tiny problem (build container only):

  RunningNorm                       smpl_sim/learning/running_norm.py
  MLP / PolicyGaussian / Value      smpl_sim/learning/mlp.py, policy_gaussian.py, critic.py
  estimate_advantages, get_optimizer  smpl_sim/learning/learning_utils.py:188-218
  AgentPG.update_value, AgentPPO.update_policy / ppo_loss / clip_policy_grad   smpl_sim/agents/agent_pg.py:21-28, agent_ppo.py:20-107

The agent classes are not instantiated (their ctor spawns envs / loggers); the unbound methods run on a namespace that carries
exactly the attributes they read.  T = 48 samples of one env (flat batch), 3 optimisation epochs."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402,F401  (installs the third-party stubs, puts /root/reference on sys.path)

PG = MG._import_with_stubs(lambda: __import__("smpl_sim.learning.policy_gaussian", fromlist=["x"]))
from smpl_sim.learning.mlp import MLP  # noqa: E402
from smpl_sim.learning.critic import Value  # noqa: E402
LU = MG._import_with_stubs(lambda: __import__("smpl_sim.learning.learning_utils", fromlist=["x"]))
APG = MG._import_with_stubs(lambda: __import__("smpl_sim.agents.agent_pg", fromlist=["x"]))
APPO = MG._import_with_stubs(lambda: __import__("smpl_sim.agents.agent_ppo", fromlist=["x"]))


def sd_np(sd, prefix):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    S, A, T = 7, 3, 48
    units, act = [16, 8], "silu"
    cfg = types.SimpleNamespace(learning=types.SimpleNamespace(mlp=types.SimpleNamespace(units=units, activation=act), fix_std=True, log_std=-2.5))
    policy = PG.PolicyGaussian(cfg, action_dim=A, state_dim=S)
    value = Value(MLP(S, units, act))
    out = {}
    out.update(sd_np(policy.state_dict(), "p0."))
    out.update(sd_np(value.state_dict(), "v0."))
    states = torch.randn(T, S) * 2.0 + 0.5
    actions = torch.randn(T, A) * 0.1
    rewards = torch.rand(T, 1)
    not_done = torch.ones(T, 1); not_done[[11, 29, T - 1]] = 0.0
    not_dead = torch.ones(T, 1); not_dead[29] = 0.0
    exps = torch.ones(T)
    ag = types.SimpleNamespace(policy_net=policy, value_net=value, update_modules=[policy, value], trans_value=lambda x: x,
                               optimizer_policy=LU.get_optimizer(policy, 5e-5, 0.0, "adam"), optimizer_value=LU.get_optimizer(value, 3e-4, 0.0, "adam"),
                               opt_num_epochs=3, value_opt_niter=1, use_mini_batch=False, clip_epsilon=0.2, policy_grad_clip=[(policy, 25)],
                               device=torch.device("cpu"), dtype=torch.float32, gamma=0.99, tau=0.95)
    for name in ("update_value",):
        setattr(ag, name, types.MethodType(getattr(APG.AgentPG, name), ag))
    for name in ("update_policy", "ppo_loss", "clip_policy_grad"):
        setattr(ag, name, types.MethodType(getattr(APPO.AgentPPO, name), ag))
    # AgentPG.update_params body (agent_pg.py:41-60) on tensors
    LU.to_train(*ag.update_modules)
    with LU.to_test(*ag.update_modules):
        with torch.no_grad():
            values = value(states)
    adv, ret = LU.estimate_advantages(rewards, not_done, not_dead, values, ag.gamma, ag.tau)
    ag.update_policy(states, states, actions, ret, adv, exps)
    out.update(sd_np(policy.state_dict(), "p1."))
    out.update(sd_np(value.state_dict(), "v1."))
    out.update(states=states.numpy(), actions=actions.numpy(), rewards=rewards.numpy(), not_done=not_done.numpy(), not_dead=not_dead.numpy(),
               values=values.numpy(), advantages=adv.numpy(), returns=ret.numpy(), units=np.array(units))
    np.savez_compressed(os.path.join(HERE, "ppo_update.npz"), **out)
    print("ppo_update.npz: norm.n", int(policy.norm.n), "|adv| max", float(adv.abs().max()), "keys", len(out))


if __name__ == "__main__":
    main()
