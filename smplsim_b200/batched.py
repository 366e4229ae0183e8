"""Batched tensor protocol (boundary B2): the reference's own GPU-sim surface
(smpl_sim/envs/nv/base_task.py:97-124,294-313, smpl_sim/envs/nv/humanoid.py:2127-2143) backed by
libsmplsim_b200.so instead of Isaac Gym, with MuJoCo-env semantics (smpl_sim/envs/humanoid_env.py).

PyTorch owns every tensor (device memory, streams); the C ABI receives raw pointers per call.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Any, Optional

import numpy as np
import torch

from . import _lib
from .abi import SmplsimAuxC, SmplsimStateC, env_cfg_from, model_from_cfg
from .model import ModelDesc


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class HumanoidBatchB200:
    """N independent SMPL humanoid envs stepped in lockstep on one GPU.

    Attributes mirroring the reference's tensor API: ``num_envs, num_obs, num_actions, device, obs_buf,
    rew_buf, reset_buf, progress_buf, extras`` (+ ``terminate_buf, truncate_buf``).  State lives in the
    SoA tensors ``qpos[N,nq], qvel[N,nv]`` etc.
    """

    def __init__(self, cfg: Any, num_envs: Optional[int] = None, device: str = "cuda:0", seed: Optional[int] = None,
                 model: Optional[ModelDesc] = None, rank: int = 0, with_aux: bool = True,
                 models: Optional[Sequence[ModelDesc]] = None, env_model: Optional[Sequence[int]] = None):
        """``models`` + ``env_model`` ([num_envs] indices into models): per-env body shapes in one batch -- what one SMPL_Robot per
        env process gives the reference (humanoid_env.py:219-250).  Same tree / joints / geom layout in every model."""
        self._require_device(device)
        self.cfg = cfg
        e = cfg.env
        self.num_envs = int(num_envs if num_envs is not None else (e.get("num_envs", 1) if hasattr(e, "get") else 1))
        self.device = torch.device(device)
        self.models = list(models) if models is not None else None
        if self.models is not None:
            if env_model is None or len(env_model) != self.num_envs:
                raise ValueError("models= needs env_model with one entry per env")
            model = self.models[0]
        self.model = model if model is not None else model_from_cfg(cfg)
        seed = int(cfg.get("seed", 0) if seed is None and hasattr(cfg, "get") else (seed or 0))
        self.seed = (seed + 0x9E3779B97F4A7C15 * rank) & 0xFFFFFFFFFFFFFFFF      # per-rank Philox key (env sharding)
        self.envcfg = env_cfg_from(cfg, self.model, seed=self.seed)
        self._cmodel = self.model.c_struct()
        self._h = C.c_void_p()
        L = self._L()
        dev_index = self._device_index()
        if self.models is None:
            self._check(L.smplsim_create(C.addressof(self._cmodel), C.addressof(self.envcfg), self.num_envs, dev_index, C.byref(self._h)))
            self.env_model = None
        else:
            self._cmodels_each = [m.c_struct() for m in self.models]          # keeps the arrays the structs point at alive
            arr = (type(self._cmodel) * len(self.models))(*self._cmodels_each)
            em = np.ascontiguousarray(np.asarray(env_model, dtype=np.int32))
            self._check(L.smplsim_create_shapes(C.addressof(arr), len(self.models), C.c_void_p(em.ctypes.data), C.addressof(self.envcfg),
                                                self.num_envs, dev_index, C.byref(self._h)))
            self.env_model = em
        self.num_obs = L.smplsim_obs_dim(self._h)
        self.num_actions = self.model.nu
        self.dt = float(self.model.timestep * self.envcfg.nsubsteps)
        self.max_episode_length = int(e.episode_length)
        N, m, dv = self.num_envs, self.model, self.device
        f32 = dict(dtype=torch.float32, device=dv)
        i32 = dict(dtype=torch.int32, device=dv)
        self.qpos = torch.zeros(N, m.nq, **f32)
        self.qpos[:, 3] = 1.0
        self.qvel = torch.zeros(N, m.nv, **f32)
        self.qpos_fwd = self.qpos.clone()
        self.qvel_fwd = torch.zeros(N, m.nv, **f32)
        self.qacc_warm = torch.zeros(N, m.nv, **f32)
        self.task_target = torch.zeros(N, 4, **f32)
        self.task_change_step = torch.zeros(N, **i32)
        self.progress_buf = torch.zeros(N, **i32)
        self.recovery = torch.zeros(N, **i32)
        self.rng_counter = torch.zeros(N, **i32)
        self.pid_integral = torch.zeros(N, m.nu, **f32)              # SimplePID state (simple_pid only; survives reset, like the reference object)
        self.pid_last_error = torch.full((N, m.nu), float('nan'), **f32)   # NaN = None
        self.obs_buf = torch.zeros(N, self.num_obs, **f32)
        self.rew_buf = torch.zeros(N, **f32)
        self.terminate_buf = torch.zeros(N, dtype=torch.uint8, device=dv)
        self.truncate_buf = torch.zeros(N, dtype=torch.uint8, device=dv)
        self.reset_buf = torch.ones(N, dtype=torch.uint8, device=dv)
        # side outputs of the last forward pass
        self.xpos = torch.zeros(N, m.nbody, 3, **f32)
        self.xquat = torch.zeros(N, m.nbody, 4, **f32)
        self.body_linvel = torch.zeros(N, m.nbody, 3, **f32)
        self.body_angvel = torch.zeros(N, m.nbody, 3, **f32)
        self.contact_mask = torch.zeros(N, dtype=torch.int64, device=dv)
        self.qacc = torch.zeros(N, m.nv, **f32)
        self.ctrl = torch.zeros(N, m.nu, **f32)
        self.solver_iter = torch.zeros(N, **i32)
        self.status = torch.zeros(N, dtype=torch.uint8, device=dv)   # mj_warning bits of the last call (1 qpos, 2 qvel, 4 qacc)
        self.extras = {"terminate": self.terminate_buf, "sim_warning": self.status}
        self._state = SmplsimStateC(*[t.data_ptr() for t in (self.qpos, self.qvel, self.qpos_fwd, self.qvel_fwd, self.qacc_warm,
                                                             self.task_target, self.task_change_step, self.progress_buf,
                                                             self.recovery, self.rng_counter, self.pid_integral, self.pid_last_error)])
        self._aux = SmplsimAuxC(*[t.data_ptr() for t in (self.xpos, self.xquat, self.body_linvel, self.body_angvel,
                                                         self.contact_mask, self.qacc, self.ctrl, self.solver_iter, self.status)])
        if not with_aux:        # throughput runs: skip the side outputs (NULL pointers), keep the 1-byte status
            self._aux = SmplsimAuxC(status=self.status.data_ptr())
        self.gpu_launches = 0

    def __del__(self):
        try:
            if self._h:
                self._L().smplsim_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ binding to libsmplsim_b200.so (CUDA only)
    def _require_device(self, device):
        if not torch.cuda.is_available():
            raise RuntimeError("HumanoidBatchB200 needs a CUDA device (there is no CPU fallback for the stepper)")

    def _L(self):
        return _lib.lib()

    def _check(self, rc):
        if rc != 0:
            raise _lib.SmplsimError(f"libsmplsim_b200 error {rc}: {self._L().smplsim_last_error().decode()}")

    def _device_index(self):
        return self.device.index if self.device.index is not None else torch.cuda.current_device()

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ reference tensor API
    def reset(self, env_ids: Optional[torch.Tensor] = None, init_mode: int = -1, qpos0: Optional[torch.Tensor] = None,
              qvel0: Optional[torch.Tensor] = None):
        """HumanoidTask.reset for ``env_ids`` (None = all; an index tensor or a uint8/bool mask)."""
        mask = None
        if env_ids is not None:
            if env_ids.dtype in (torch.bool, torch.uint8) and env_ids.numel() == self.num_envs:
                mask = env_ids.to(torch.uint8).contiguous()
            else:
                if env_ids.numel() == 0:
                    return self.obs_buf
                mask = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
                mask[env_ids.long()] = 1
        if qpos0 is not None:
            qpos0 = qpos0.to(self.device, torch.float32).contiguous()
            qvel0 = qvel0.to(self.device, torch.float32).contiguous()
        self._check(self._L().smplsim_reset(self._h, C.byref(self._state), _ptr(mask), int(init_mode), _ptr(qpos0), _ptr(qvel0),
                                            _ptr(self.obs_buf), C.byref(self._aux), self._stream()))
        self.gpu_launches += 1
        self._keep = (mask, qpos0, qvel0)
        return self.obs_buf

    def reset_done(self):
        """Reset every env whose reset_buf is set (in-stream, no host sync)."""
        return self.reset(self.reset_buf)

    def step(self, actions: torch.Tensor):
        """BaseEnv.step for all envs: obs_buf, rew_buf, terminate/truncate/reset_buf, progress_buf updated in place."""
        a = actions.to(self.device, torch.float32)
        if bool(self.cfg.env.clip_actions):
            a = torch.clamp(a, -1.0, 1.0)      # action_space Box(-1,1) (humanoid_env.py:180-184; agent.py:153-161)
        a = a.contiguous()
        assert a.shape == (self.num_envs, self.num_actions)
        self._check(self._L().smplsim_step(self._h, C.byref(self._state), _ptr(a), _ptr(self.obs_buf), _ptr(self.rew_buf),
                                           _ptr(self.terminate_buf), _ptr(self.truncate_buf), C.byref(self._aux), self._stream()))
        self.gpu_launches += 1
        torch.bitwise_or(self.terminate_buf, self.truncate_buf, out=self.reset_buf)
        self._keep_a = a
        return self.obs_buf, self.rew_buf, self.terminate_buf, self.truncate_buf

    # ------------------------------------------------------------------ raw physics / kinematics (parity + tools)
    def mj_step(self, ctrl: torch.Tensor, nsub: int = 1):
        """data.ctrl[:] = ctrl; mujoco.mj_step x nsub."""
        c = ctrl.to(self.device, torch.float32).contiguous()
        self._check(self._L().smplsim_mj_step(self._h, C.byref(self._state), _ptr(c), int(nsub), C.byref(self._aux), self._stream()))
        self.gpu_launches += 1
        self._keep_a = c

    def kinematics(self, qpos: torch.Tensor):
        q = qpos.to(self.device, torch.float32).contiguous()
        n = q.shape[0]
        xpos = torch.empty(n, self.model.nbody, 3, dtype=torch.float32, device=self.device)
        xquat = torch.empty(n, self.model.nbody, 4, dtype=torch.float32, device=self.device)
        self._check(self._L().smplsim_kinematics(self._h, _ptr(q), _ptr(xpos), _ptr(xquat), n, self._stream()))
        self.gpu_launches += 1
        return xpos, xquat

    def self_obs(self, version: int, xpos, xquat, qvel=None, linvel=None, angvel=None):
        def prep(t):
            return None if t is None else t.to(self.device, torch.float32).contiguous()
        xpos, xquat, qvel, linvel, angvel = map(prep, (xpos, xquat, qvel, linvel, angvel))
        n, nb = xpos.shape[0], self.model.nbody
        dim = (1 if self.envcfg.root_height_obs else 0) + 3 * (nb - 1) + 6 * nb + (6 + self.model.nu if version == 1 else 6 * nb)
        obs = torch.empty(n, dim, dtype=torch.float32, device=self.device)
        self._check(self._L().smplsim_self_obs(self._h, version, _ptr(qvel), _ptr(xpos), _ptr(xquat), _ptr(linvel), _ptr(angvel),
                                               _ptr(obs), n, self._stream()))
        self.gpu_launches += 1
        return obs

    def set_state(self, qpos: torch.Tensor, qvel: torch.Tensor):
        """Write qpos/qvel (and mark them as the last-forward state, as mj_forward would)."""
        self.qpos.copy_(qpos.to(self.device, torch.float32))
        self.qvel.copy_(qvel.to(self.device, torch.float32))
        self.qpos_fwd.copy_(self.qpos)
        self.qvel_fwd.copy_(self.qvel)

    def smem_bytes_per_env(self):
        return self._L().smplsim_smem_bytes_per_env(self._h)

    @property
    def kernel_version(self) -> int:
        """5: lane-chain kernels (csrc/lane_kernels.cuh)."""
        return self._L().smplsim_kernel_version(self._h)

    @property
    def schedule_steps(self) -> int:
        return self._L().smplsim_schedule_steps(self._h)


class GymVectEnvB200:
    """``GymVectEnv`` (smpl_sim/envs/nv/gymwrapper.py:7-65) over a HumanoidBatchB200: gymnasium-vector style
    ``reset`` / ``step`` with autoreset and ``info["final_observation"]``."""

    def __init__(self, env: HumanoidBatchB200, clip_observations: float = float(np.inf)):
        self._env = env
        self._clip_obs = clip_observations
        self.single_observation_shape = (env.num_obs,)
        self.single_action_shape = (env.num_actions,)

    @property
    def num_envs(self) -> int:
        return self._env.num_envs

    def _get_clipped_obs(self):
        return torch.clamp(self._env.obs_buf, -self._clip_obs, self._clip_obs)

    def reset(self, seed=None, options=None):
        self._env.reset(None)
        return self._get_clipped_obs().clone(), {}

    def close(self):
        pass

    def step(self, actions):
        env = self._env
        env.step(actions)
        obs = self._get_clipped_obs().clone()
        reward = env.rew_buf.clone()
        terminated = env.terminate_buf.bool()
        truncated = env.truncate_buf.bool()
        done = env.reset_buf.clone()
        env.reset(done)                         # masked, in-stream; no host sync
        info = {"terminate": terminated, "final_observation": obs, "_final_observation": done.bool()}
        return self._get_clipped_obs().clone(), reward, terminated, truncated, info
