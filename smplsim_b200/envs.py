"""Single-env gymnasium protocol (boundary B1): what ``smpl_sim/run.py``'s loop talks to
(``Agent.sample_worker``, smpl_sim/agents/agent.py:64-109) -- the API of ``BaseEnv`` / ``HumanoidEnv`` /
``HumanoidTask`` (smpl_sim/envs/base_env.py:64-110, humanoid_env.py:148-184,455-469, humanoid_task.py) backed by a
``num_envs=1`` CUDA handle.  numpy in, numpy out; there is no CPU fallback.

A CUDA context does not survive ``fork()`` (agent.py:130-133 forks its samplers), so use ``num_threads=1`` with these
classes, or the batched surface (``smplsim_b200.batched``) for real throughput.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Optional

import numpy as np
import torch

from .batched import HumanoidBatchB200

try:                                     # gymnasium is optional in this image
    import gymnasium as _gym
    _Box = _gym.spaces.Box
    _EnvBase = _gym.Env
except Exception:                        # pragma: no cover - minimal stand-ins with the attributes the agents read
    _gym = None

    class _Box:                          # noqa: D401
        def __init__(self, low, high, dtype=np.float32):
            self.low, self.high, self.dtype = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype), dtype
            self.shape = self.low.shape

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def sample(self):
            return np.random.uniform(np.maximum(self.low, -1e3), np.minimum(self.high, 1e3)).astype(self.dtype)

    class _EnvBase:
        pass


class _Contacts:
    """``mj_data.contact.geom1 / geom2`` of the last forward pass (floor contacts only; floor = geom 0)."""

    def __init__(self, mask: int):
        g2 = [g for g in range(1, 64) if (mask >> g) & 1]
        self.geom2 = np.asarray(g2, dtype=np.int32)
        self.geom1 = np.zeros_like(self.geom2)


class HumanoidEnvB200(_EnvBase):
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 30}
    TASK = None

    def __init__(self, cfg: Any, device: str = "cuda:0"):
        self.cfg = cfg
        if self.TASK is not None and cfg.env.get("task", self.TASK) != self.TASK:
            raise ValueError(f"{type(self).__name__} expects cfg.env.task == {self.TASK!r}")
        e = cfg.env
        # BaseEnv.__init__ (base_env.py:23-49)
        self.clip_actions = e.clip_actions
        self.render_mode = e.render_mode
        assert self.render_mode is None or self.render_mode in self.metadata["render_modes"]
        self.headless = cfg.headless
        self.sim_timestep_inv = e.sim_timestep_inv
        self.sim_timestep = 1.0 / self.sim_timestep_inv
        self.control_freq_inv = e.control_frequency_inv
        self.cur_t = 0
        self.dt = self.sim_timestep * self.control_freq_inv
        self.paused = False
        self.disable_reset = False
        self.viewer = None
        self.renderer = None
        self.dtype = np.float32
        self._b = self._make_batch(cfg, device)
        m = self._b.model
        self.control_mode = e.control_mode
        self.max_episode_length = e.episode_length
        self.self_obs_v = e.self_obs_v
        self.humanoid_type = cfg.robot.humanoid_type
        self.body_names_orig = list(m.body_names)
        self.mj_body_names = ["world"] + list(m.body_names)
        self.num_rigid_bodies = m.nbody
        self.dof_names = self.body_names_orig[1:]
        self.actuator_names = list(m.joint_names)
        self.dof_size = m.nu
        self.qpos_lim, self.qvel_lim = m.nq, m.nv
        self.contact_bodies = list(e.contact_bodies)
        self.contact_bodies_ids = [m.geom_names.index(n) + 1 for n in self.contact_bodies]
        self.floor_idx = 0
        self.jkp, self.jkd, self.torque_lim = m.act_kp.copy(), m.act_kd.copy(), m.act_torque_lim.copy()
        self._pd_action_scale, self._pd_action_offset = m.act_scale.copy(), m.act_offset.copy()
        self.observation_space = _Box(-np.inf * np.ones(self.get_obs_size(), dtype=np.float32),
                                      np.inf * np.ones(self.get_obs_size(), dtype=np.float32), dtype=np.float32)
        lim = np.ones(self.get_action_size(), dtype=np.float32) if self.clip_actions else np.inf * np.ones(self.get_action_size(), dtype=np.float32)
        self.action_space = _Box(-lim, lim, dtype=np.float32)
        self.np_random = np.random.default_rng(int(cfg.get("seed", 0)))
        self.reward_info = {}

    def _make_batch(self, cfg, device):
        """The num_envs=1 handle behind this env (CUDA; there is no CPU path)."""
        return HumanoidBatchB200(cfg, num_envs=1, device=device, seed=int(cfg.get("seed", 0)))

    # ------------------------------------------------------------------ sizes
    def get_action_size(self):
        return self._b.num_actions

    def get_obs_size(self):
        return self._b.num_obs

    # ------------------------------------------------------------------ gym API
    def seed(self, seed: Optional[int] = None):
        """BaseEnv.seed / gym reset(seed=): reseeds np_random AND the device-side task / Fall-init stream (Philox counter jump:
        the key stays (cfg.seed, env id), the per-env counter restarts from a value derived from `seed`)."""
        self.np_random = np.random.default_rng(seed)
        if seed is not None:
            self._b.rng_counter.fill_(int((int(seed) * 2654435761) & 0x7FFFFFFF))

    def reset(self, seed=None, options=None):
        if seed is not None:
            self.seed(seed)
        obs = self._b.reset().cpu().numpy()[0].astype(self.dtype)
        self.cur_t = 0
        return obs, {"critic_state": obs}

    def step(self, action):
        a = torch.as_tensor(np.asarray(action, dtype=np.float32)[None, : self.dof_size])
        obs, rew, term, trunc = self._b.step(a)
        obs = obs.cpu().numpy()[0].astype(self.dtype)
        self.cur_t = int(self._b.progress_buf[0].item())
        died, timed_out = bool(term[0].item()), bool(trunc[0].item())
        if self.disable_reset:
            died, timed_out = False, False
        info = dict(self.reward_info)
        info["critic_state"] = obs
        return obs, float(rew[0].item()), died, timed_out, info

    def render(self):        # called every step by worker 0 even when headless (agent.py:96-97): cheap no-op
        return None

    def close(self):
        pass

    # ------------------------------------------------------------------ mj_data-style views (examples poke env.mj_data)
    @property
    def mj_data(self):
        b = self._b
        sens = np.concatenate([b.body_linvel[0].cpu().numpy().ravel(), b.body_angvel[0].cpu().numpy().ravel()])
        z3, z4 = np.zeros((1, 3)), np.array([[1.0, 0, 0, 0]])
        return SimpleNamespace(qpos=b.qpos[0].cpu().numpy().astype(np.float64), qvel=b.qvel[0].cpu().numpy().astype(np.float64),
                               ctrl=b.ctrl[0].cpu().numpy().astype(np.float64), qacc=b.qacc[0].cpu().numpy().astype(np.float64),
                               xpos=np.concatenate([z3, b.xpos[0].cpu().numpy()]), xquat=np.concatenate([z4, b.xquat[0].cpu().numpy()]),
                               sensordata=sens, contact=_Contacts(int(b.contact_mask[0].item())))

    @property
    def mj_model(self):
        """The MjModel attributes the reference reads (humanoid_env.py:262-289,325-370; utils/mujoco_utils.py): sizes, names, joint
        ranges, timestep.  Body / geom 0 is the world / floor, as in MuJoCo."""
        m = self._b.model
        rng = np.asarray(m.dof_range)
        names = ["world"] + list(m.body_names)
        joints = {n: SimpleNamespace(name=n, range=rng[6 + i].copy(), qposadr=np.array([7 + i]), dofadr=np.array([6 + i]))
                  for i, n in enumerate(m.joint_names)}
        return SimpleNamespace(nbody=m.nbody + 1, nq=m.nq, nv=m.nv, nu=m.nu, ngeom=len(m.geom_names) + 1, opt=SimpleNamespace(timestep=self.sim_timestep),
                               body=lambda i: SimpleNamespace(name=names[i], id=i), joint=lambda n: joints[n],
                               geom=lambda i: SimpleNamespace(name=(["floor"] + list(m.geom_names))[i], id=i),
                               body_mass=np.concatenate([[0.0], np.asarray(m.body_mass)]))

    def get_qpos(self):
        return self._b.qpos[0].cpu().numpy().astype(np.float64)

    def get_qvel(self):
        return self._b.qvel[0].cpu().numpy().astype(np.float64)

    def get_body_xpos(self):
        return self._b.xpos[0].cpu().numpy().astype(np.float64)

    def get_body_xquat(self):
        return self._b.xquat[0].cpu().numpy().astype(np.float64)

    def get_body_xpos_by_id(self, body_id):
        return self.get_body_xpos()[body_id]

    def get_root_pos(self):
        return self.get_body_xpos()[0].copy()

    def compute_torque(self, ctrl):
        """Torque applied in the last substep for the last action (the reference recomputes it per substep)."""
        return self._b.ctrl[0].cpu().numpy().astype(np.float64)


class HumanoidSpeed(HumanoidEnvB200):
    TASK = "HumanoidSpeed"


class HumanoidReach(HumanoidEnvB200):
    TASK = "HumanoidReach"


class HumanoidGetup(HumanoidEnvB200):
    TASK = "HumanoidGetup"


def make_env(cfg: Any, device: str = "cuda:0"):
    """``eval(cfg.env.task)(cfg)`` of AgentHumanoid.setup_env (agents/agent_humanoid.py:91-92)."""
    return {"HumanoidEnv": HumanoidEnvB200, "HumanoidSpeed": HumanoidSpeed, "HumanoidReach": HumanoidReach,
            "HumanoidGetup": HumanoidGetup}[cfg.env.task](cfg, device=device)
