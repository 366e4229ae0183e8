"""Scripted env / policy shared by tests/golden/make_golden_sampler.py (drives the REFERENCE's Agent.sample_worker) and
tests/test_sampler_cpu.py (drives smplsim_b200.learning.BatchedSampler): observations, rewards and episode ends are closed-form
functions of (env index k, episode e, step t)."""
import numpy as np

D, A, K, T = 6, 4, 3, 14
W = (np.arange(A * D, dtype=np.float64).reshape(A, D) % 7 - 3.0) * 0.11      # policy: a = 3 W s  (saturates the action clip)


def ep_len(k, e):
    return 3 + (2 * k + 3 * e) % 5


def died(k, e):
    return (k + e) % 2 == 0          # otherwise the episode times out


def obs_of(k, e, t):
    base = np.array([k + 1.0, e - 1.5, t * 0.7, -6.0 + t, 7.5 - 2.0 * k, 0.3 * (t + 1) * (e + 1)])   # components beyond +-5 exist
    return (base * (1.0 if (k + t) % 2 == 0 else -1.0)).astype(np.float32)


def rew_of(k, e, t):
    return 0.1 * t + k + 0.01 * e


class ScriptEnv:
    """Single-env gym protocol (what Agent.sample_worker drives)."""

    def __init__(self, k):
        self.k, self.e, self.t = k, -1, 0
        self.actions_seen = []
        self.np_random = np.random.default_rng(0)

    def reset(self):
        self.e += 1; self.t = 0
        o = obs_of(self.k, self.e, 0)
        return o, {"critic_state": o}

    def step(self, a):
        self.actions_seen.append(np.asarray(a, dtype=np.float32).copy())
        r = rew_of(self.k, self.e, self.t)
        self.t += 1
        end = self.t >= ep_len(self.k, self.e)
        d = bool(end and died(self.k, self.e))
        to = bool(end and not died(self.k, self.e))
        o = obs_of(self.k, self.e, self.t)
        return o, r, d, to, {"critic_state": o}

    def render(self):
        pass
