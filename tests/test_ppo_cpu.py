"""PPO update plumbing (smplsim_b200/ppo.py) against the reference's own learner code: the fixture
tests/golden/ppo_update.npz was produced by running RunningNorm / PolicyGaussian / Value / estimate_advantages /
AgentPG.update_value / AgentPPO.update_policy from /root/reference on a tiny problem (tests/golden/make_golden_ppo.py)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from conftest import GOLDEN  # noqa: E402
from smplsim_b200.ppo import PolicyGaussian, PPOLearner, RunningNorm, Value  # noqa: E402


def _gae_torch(rewards, not_done, not_dead, values, gamma, tau, next_value=None):
    """learning_utils.py:198-218 per column ([T,N]); the reference normalises with torch's unbiased std."""
    T, N = rewards.shape
    adv = torch.zeros(T, N)
    prev_v = torch.zeros(N) if next_value is None else next_value
    prev_a = torch.zeros(N)
    for t in reversed(range(T)):
        delta = rewards[t] + gamma * prev_v * not_dead[t] - values[t]
        adv[t] = delta + gamma * tau * prev_a * not_done[t]
        prev_v = values[t]
        prev_a = adv[t]
    ret = values + adv
    return (adv - adv.mean()) / adv.std(), ret


def _load(g, prefix, module):
    sd = {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}
    module.load_state_dict(sd)           # same parameter / buffer names as the reference modules


def test_ppo_update_matches_reference():
    g = np.load(os.path.join(GOLDEN, "ppo_update.npz"))
    units = [int(u) for u in g["units"]]
    S, A = g["states"].shape[1], g["actions"].shape[1]
    policy, value = PolicyGaussian(S, A, units, "silu", -2.5, True), Value(S, units, "silu")
    _load(g, "p0.", policy); _load(g, "v0.", value)
    learner = PPOLearner(policy, value, gamma=0.99, tau=0.95, clip_epsilon=0.2, opt_num_epochs=3, policy_lr=5e-5, value_lr=3e-4, policy_grad_clip=25)
    T = g["states"].shape[0]
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    batch = dict(states=t("states").reshape(T, 1, S), actions=t("actions").reshape(T, 1, A), rewards=t("rewards").reshape(T, 1),
                 not_done=t("not_done").reshape(T, 1), not_dead=t("not_dead").reshape(T, 1))
    seen = {}

    def adv_fn(*a, **k):
        adv, ret = _gae_torch(*a, **k)
        seen["adv"], seen["ret"] = adv, ret
        return adv, ret

    learner.update(batch, advantages_fn=adv_fn)
    assert np.abs(seen["adv"].numpy() - g["advantages"].reshape(T, 1)).max() < 1e-5
    assert np.abs(seen["ret"].numpy() - g["returns"].reshape(T, 1)).max() < 1e-5
    for prefix, mod in (("p1.", policy), ("v1.", value)):
        for k, v in mod.state_dict().items():
            ref = g[prefix + k]
            assert np.abs(v.numpy() - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), (prefix + k, np.abs(v.numpy() - ref).max())
    assert int(policy.norm.n) == 3 * T          # the quirk: RunningNorm only moves during the optimisation epochs


def test_running_norm_distributed_moments_equal_single_process():
    """Two gloo ranks, each with half of the batch, end with the statistics one process gets from the whole batch."""
    import torch.multiprocessing as mp
    x = torch.randn(64, 5) * 3 + 1
    ref = RunningNorm(5); ref.train(); ref(x[:40]); ref(x[40:])
    port = 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_rn_worker, args=(r, 2, port, x, q)) for r in range(2)]
    [p.start() for p in ps]
    outs = [q.get(timeout=120) for _ in ps]
    [p.join(60) for p in ps]
    for mean, var, n in outs:
        assert n == 64 and np.abs(mean - ref.mean.numpy()).max() < 1e-5 and np.abs(var - ref.var.numpy()).max() < 1e-4


def _rn_worker(rank, world, port, x, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rn = RunningNorm(5); rn.train()
    rn(x[:40][rank::2]); rn(x[40:][rank::2])
    q.put((rn.mean.numpy().copy(), rn.var.numpy().copy(), int(rn.n)))
    dist.destroy_process_group()
