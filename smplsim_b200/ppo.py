"""PPO update plumbing right after the hot path, data-parallel over ranks.

What the reference does per epoch (agents/agent_pg.py:41-60, agents/agent_ppo.py:20-107): values of the whole batch, GAE
(``estimate_advantages``), then ``opt_num_epochs`` (10) full-batch iterations of [one critic step, one clipped-surrogate policy
step with a gradient-norm clip of 25].  Here the batch is the device-resident ``[T, N, .]`` rollout of ``BatchedSampler``; with
``torch.distributed`` initialised each rank holds its env shard and the three places where the reference's single process sees
the whole batch are reduced over ranks: advantage moments (``estimate_advantages``), ``RunningNorm`` batch moments, gradients.

The modules keep the reference's parameter names (``norm.*``, ``net.affine_layers.i.*``, ``action_mean.*``, ``action_log_std``,
``value_head.*``), so a checkpoint written by ``AgentHumanoid.get_nn_weights`` (agents/agent_humanoid.py:117-121) loads as is.
PyTorch is plumbing here: the networks are plain ``nn.Linear`` stacks (library GEMMs), none of this is on the simulated path."""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn

from .dist import allreduce_mean_grads
from .learning import estimate_advantages

_ACT = {"tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid, "gelu": nn.functional.gelu, "silu": nn.functional.silu}


def _dist_on():
    return torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1


class RunningNorm(nn.Module):
    """``y = clip((x - mean) / (std + 1e-8), +-clip)`` with running moments (learning/running_norm.py:6-42): in training mode every
    forward first merges the batch's biased variance / mean into the running ones (Chan's parallel update, weights n/(n+m)); the
    reference therefore updates it during the PPO epochs and not while sampling (quirk kept).  Under ``torch.distributed`` the
    batch moments are those of the union of all ranks' batches, so every rank holds identical statistics."""

    def __init__(self, dim: int, demean: bool = True, destd: bool = True, clip: float = 5.0):
        super().__init__()
        self.dim, self.demean, self.destd, self.clip = dim, demean, destd, clip
        self.register_buffer("n", torch.tensor(0, dtype=torch.long))
        self.register_buffer("mean", torch.zeros(dim))
        self.register_buffer("var", torch.zeros(dim))
        self.register_buffer("std", torch.zeros(dim))

    @torch.no_grad()
    def update(self, x: torch.Tensor):
        m = x.shape[0]
        if _dist_on():     # batch moments over all ranks; float64 one-pass sums so that the result matches var_mean of the joint batch
            xd = x.double()
            s = torch.cat([xd.sum(0), (xd * xd).sum(0), xd.new_tensor([float(m)])])
            torch.distributed.all_reduce(s)
            m = int(round(float(s[-1])))
            mean_d = s[: self.dim] / m
            var_x = (s[self.dim: 2 * self.dim] / m - mean_d * mean_d).clamp_min(0.0).to(x.dtype)
            mean_x = mean_d.to(x.dtype)
        else:
            var_x, mean_x = torch.var_mean(x, dim=0, unbiased=False)
        w = self.n.to(x.dtype) / (m + self.n).to(x.dtype)
        self.var.copy_(w * self.var + (1 - w) * var_x + w * (1 - w) * (mean_x - self.mean).pow(2))
        self.mean.copy_(w * self.mean + (1 - w) * mean_x)
        self.std.copy_(torch.sqrt(self.var))
        self.n += m

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            self.update(x)
        if self.n > 0:
            if self.demean:
                x = x - self.mean
            if self.destd:
                x = x / (self.std + 1e-8)
            if self.clip:
                x = torch.clamp(x, -self.clip, self.clip)
        return x


class MLP(nn.Module):
    """``activation(Linear(.))`` stack, activation after every layer incl. the last (learning/mlp.py:32-58)."""

    def __init__(self, input_dim: int, hidden_dims: Sequence[int] = (128, 128), activation: str = "tanh"):
        super().__init__()
        self.activation = _ACT[activation]
        self.out_dim = hidden_dims[-1]
        self.affine_layers = nn.ModuleList()
        last = input_dim
        for nh in hidden_dims:
            self.affine_layers.append(nn.Linear(last, nh))
            last = nh

    def forward(self, x):
        for affine in self.affine_layers:
            x = self.activation(affine(x))
        return x


class PolicyGaussian(nn.Module):
    """Diagonal-Gaussian actor (learning/policy_gaussian.py:15-43): RunningNorm -> MLP -> Linear mean head (weights x0.1, bias 0),
    state-independent log-std (``learning.log_std`` = -2.5, fixed when ``learning.fix_std``)."""

    def __init__(self, state_dim: int, action_dim: int, units: Sequence[int], activation: str = "silu", log_std: float = -2.5,
                 fix_std: bool = True):
        super().__init__()
        self.type = "gaussian"
        self.norm = RunningNorm(state_dim)
        self.net = MLP(state_dim, units, activation)
        self.action_mean = nn.Linear(self.net.out_dim, action_dim)
        with torch.no_grad():
            self.action_mean.weight.mul_(0.1)
            self.action_mean.bias.mul_(0.0)
        self.action_log_std = nn.Parameter(torch.ones(1, action_dim) * log_std, requires_grad=not fix_std)

    def forward(self, x):
        mean = self.action_mean(self.net(self.norm(x)))
        return mean, self.action_log_std.expand_as(mean)

    def select_action(self, x, mean_action: bool = False, generator: Optional[torch.Generator] = None):
        mean, log_std = self.forward(x)
        if mean_action:
            return mean
        return mean + torch.exp(log_std) * torch.randn(mean.shape, device=mean.device, dtype=mean.dtype, generator=generator)

    def get_log_prob(self, x, action):
        mean, log_std = self.forward(x)
        var = torch.exp(2.0 * log_std)
        lp = -((action - mean) ** 2) / (2.0 * var) - log_std - 0.5 * math.log(2.0 * math.pi)
        return lp.sum(1, keepdim=True)


class Value(nn.Module):
    """Critic: MLP -> Linear(1) (weights x0.1, bias 0), no input normalisation (learning/critic.py:5-19)."""

    def __init__(self, state_dim: int, units: Sequence[int], activation: str = "silu"):
        super().__init__()
        self.net = MLP(state_dim, units, activation)
        self.value_head = nn.Linear(self.net.out_dim, 1)
        with torch.no_grad():
            self.value_head.weight.mul_(0.1)
            self.value_head.bias.mul_(0.0)

    def forward(self, x):
        return self.value_head(self.net(x))


def ppo_loss(policy: PolicyGaussian, states, actions, advantages, fixed_log_probs, ind, clip_epsilon: float):
    """Clipped surrogate over the rows ``ind`` (agents/agent_ppo.py:96-107)."""
    log_probs = policy.get_log_prob(states[ind], actions[ind])
    ratio = torch.exp(log_probs - fixed_log_probs[ind])
    adv = advantages[ind]
    return -torch.min(ratio * adv, torch.clamp(ratio, 1.0 - clip_epsilon, 1.0 + clip_epsilon) * adv).mean()


class PPOLearner:
    """``AgentPPO.update_params`` (agent_pg.py:41-60 + agent_ppo.py:20-94, the default ``use_mini_batch=False`` branch) on device
    tensors.  ``update(batch)`` takes the dict ``BatchedSampler.sample`` returns (``[T, N, .]``)."""

    def __init__(self, policy: PolicyGaussian, value: Value, gamma: float = 0.99, tau: float = 0.95, clip_epsilon: float = 0.2,
                 opt_num_epochs: int = 10, value_opt_niter: int = 1, policy_lr: float = 5e-5, value_lr: float = 3e-4,
                 policy_grad_clip: Optional[float] = 25.0, weight_decay: float = 0.0, value_weight_decay: Optional[float] = None):
        self.policy, self.value = policy, value
        self.gamma, self.tau, self.clip_epsilon = gamma, tau, clip_epsilon
        self.opt_num_epochs, self.value_opt_niter, self.policy_grad_clip = opt_num_epochs, value_opt_niter, policy_grad_clip
        # get_optimizer (learning/learning_utils.py:188-190)
        self.optimizer_policy = torch.optim.Adam(policy.parameters(), eps=1e-8, lr=policy_lr, weight_decay=weight_decay)
        self.optimizer_value = torch.optim.Adam(value.parameters(), eps=1e-8, lr=value_lr,
                                                weight_decay=weight_decay if value_weight_decay is None else value_weight_decay)

    def _step(self, loss, net, opt, clip=None):
        opt.zero_grad()
        loss.backward()
        if _dist_on():
            allreduce_mean_grads(net.parameters())
        if clip is not None:
            torch.nn.utils.clip_grad_norm_(net.parameters(), clip)
        opt.step()

    def update_value(self, critic_states, returns):
        for _ in range(self.value_opt_niter):
            loss = (self.value(critic_states) - returns).pow(2).mean()
            self._step(loss, self.value, self.optimizer_value)
        return float(loss.detach())

    def update(self, batch: Dict[str, torch.Tensor], advantages_fn=None) -> Dict[str, float]:
        T, N = batch["rewards"].shape
        states = batch["states"].reshape(T * N, -1)
        actions = batch["actions"].reshape(T * N, -1)
        critic_states = batch.get("critic_states", batch["states"]).reshape(T * N, -1)
        exps = batch.get("exps")
        self.policy.train(); self.value.train()
        with torch.no_grad():
            self.value.eval()
            values = self.value(critic_states).reshape(T, N)
            last_v = self.value(batch["last_obs"]).reshape(N) if "last_obs" in batch else None
            self.value.train()
        fn = advantages_fn or estimate_advantages
        adv, ret = fn(batch["rewards"], batch["not_done"], batch["not_dead"], values, self.gamma, self.tau, next_value=last_v)
        adv, ret = adv.reshape(T * N, 1), ret.reshape(T * N, 1)
        ind = torch.arange(T * N, device=states.device) if exps is None else exps.reshape(-1).nonzero(as_tuple=False).squeeze(1)
        with torch.no_grad():
            self.policy.eval()
            fixed_log_probs = self.policy.get_log_prob(states, actions)
            self.policy.train()
        info = {}
        for _ in range(self.opt_num_epochs):
            info["value_loss"] = self.update_value(critic_states, ret)
            loss = ppo_loss(self.policy, states, actions, adv, fixed_log_probs, ind, self.clip_epsilon)
            self._step(loss, self.policy, self.optimizer_policy, self.policy_grad_clip)
            info["policy_loss"] = float(loss.detach())
        info["mean_reward"] = float(batch["rewards"].mean())
        return info


def build_from_cfg(cfg, state_dim: int, action_dim: int, device="cuda:0"):
    """``AgentHumanoid.setup_policy / setup_value / setup_optimizer`` + the ``AgentPPO`` ctor arguments of
    agents/agent_humanoid.py:60-80 from ``cfg.learning`` (data/cfg/learning/*.yaml keys)."""
    L = cfg.learning
    g = (lambda k, d: L.get(k, d)) if hasattr(L, "get") else (lambda k, d: getattr(L, k, d))
    mlp = g("mlp", None)
    units = list(mlp["units"] if isinstance(mlp, dict) else mlp.units)
    act = mlp["activation"] if isinstance(mlp, dict) else mlp.activation
    policy = PolicyGaussian(state_dim, action_dim, units, act, float(g("log_std", -2.5)), bool(g("fix_std", True))).to(device)
    value = Value(state_dim, units, act).to(device)
    learner = PPOLearner(policy, value, gamma=float(g("gamma", 0.99)), tau=float(g("tau", 0.95)), clip_epsilon=float(g("clip_epsilon", 0.2)),
                         opt_num_epochs=int(g("opt_num_epochs", 10)), policy_lr=float(g("policy_lr", 5e-5)), value_lr=float(g("value_lr", 3e-4)),
                         policy_grad_clip=g("policy_grad_clip", 25), weight_decay=float(g("policy_weightdecay", 0.0)),
                         value_weight_decay=float(g("value_weightdecay", 0.0)))
    return policy, value, learner
