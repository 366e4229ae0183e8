#!/usr/bin/env python
"""PPO on the B200 stepper -- the loop of smpl_sim/run.py (AgentHumanoid.optimize_policy: sample -> update_params -> log) with
the batched pieces of this repo: HumanoidBatchB200 (env shard on this GPU), BatchedSampler (device-resident rollout),
smplsim_gae (advantages), PPOLearner (AgentPPO.update_params semantics, gradients / moments all-reduced over ranks).

    python examples/train_ppo.py --env speed --num-envs 4096 --horizon 16 --epochs 5
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/train_ppo.py --env getup

``--horizon x --num-envs`` plays the role of ``learning.min_batch_size`` (51 200 in the reference yaml = 4096 envs x 12.5 steps)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from smplsim_b200.batched import HumanoidBatchB200  # noqa: E402
from smplsim_b200.cfg import make_cfg  # noqa: E402
from smplsim_b200.dist import max_over_ranks, rank_seed  # noqa: E402
from smplsim_b200.learning import BatchedSampler  # noqa: E402
from smplsim_b200.ppo import PolicyGaussian, PPOLearner, Value  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="speed", choices=["speed", "reach", "getup"])
    ap.add_argument("--num-envs", type=int, default=4096, help="envs on THIS GPU")
    ap.add_argument("--horizon", type=int, default=16)
    ap.add_argument("--epochs", type=int, default=5)
    ap.add_argument("--units", type=int, nargs="+", default=[2048, 1536, 1024, 1024, 512, 512])   # data/cfg/learning/simple_mlp.yaml
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(dev))
    torch.manual_seed(a.seed)                         # identical initial weights on every rank
    env = HumanoidBatchB200(make_cfg(env=a.env), num_envs=a.num_envs, device=dev, seed=a.seed, rank=rank, with_aux=False)
    policy = PolicyGaussian(env.num_obs, env.num_actions, a.units, "silu", -2.5, True).to(dev)
    value = Value(env.num_obs, a.units, "silu").to(dev)
    learner = PPOLearner(policy, value)               # gamma .99, tau .95, clip .2, 10 epochs, lr 5e-5 / 3e-4, grad clip 25
    gen = torch.Generator(device=dev); gen.manual_seed(rank_seed(a.seed, rank))

    def act(obs):
        policy.eval()                                 # to_test during sampling: RunningNorm frozen (agents/agent.py:123)
        return policy.select_action(obs, generator=gen)

    sampler = BatchedSampler(env, act)
    for ep in range(a.epochs):
        t0 = time.time()
        batch = sampler.sample(a.horizon)
        torch.cuda.synchronize(); t1 = time.time()
        info = learner.update(batch)
        torch.cuda.synchronize(); t2 = time.time()
        ts, tu = max_over_ranks(t1 - t0, dev), max_over_ranks(t2 - t1, dev)
        if rank == 0:
            n = a.horizon * a.num_envs * world
            print(f"epoch {ep}: {n} samples  sample {ts:.2f}s ({n / ts / 1e3:.0f}k env-steps/s)  update {tu:.2f}s  "
                  f"reward {info['mean_reward']:.4f}  policy_loss {info['policy_loss']:.4f}  value_loss {info['value_loss']:.4f}", flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
