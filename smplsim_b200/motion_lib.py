"""Reference-pose feed: device-resident motion tables + ``get_motion_state_intervaled``.

Replaces ``MotionLibBase.get_motion_state_intervaled`` / ``_calc_frame_blend`` (smpl_sim/smpllib/motion_lib_base.py:313-354,
448-458): the flat per-frame tables (``gts, grs, gvs, gavs, dof_pos, dvs, qpos, qvel`` + ``length_starts``, built in
``load_motions`` :173-198) live on the GPU and one CUDA kernel does the nearest-frame gather for all envs.  The AMASS
loader itself (SMPL pkl files, ``smplx``) is a "next" row (f3); ``synthetic_tables`` builds the BASELINE config-4 stand-in
(K clips x F frames of sinusoidal joint angles, tables from this repo's own FK with finite-difference + gaussian velocities,
following smpl_sim/smpllib/torch_smpl_humanoid_batch.py:118-228).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np
import torch

from . import _lib
from .batched import HumanoidBatchB200

TABLE_KEYS = ("qpos", "qvel", "xpos", "xquat", "body_vel", "body_ang_vel", "dof_pos", "dof_vel")
OPTIONAL_KEYS = ("motion_aa",)          # per-frame axis-angle pose of the source clip (motion_lib_base.py:340-354 returns it)


def _gauss1d(x, sigma):
    from scipy.ndimage import gaussian_filter1d
    return gaussian_filter1d(x, sigma, axis=0, mode="nearest")


def synthetic_tables(env: HumanoidBatchB200, num_clips: int = 64, frames: int = 300, fps: float = 30.0, seed: int = 4) -> Dict[str, np.ndarray]:
    """Config-4 synthetic clip tables: 3 sinusoids per joint, root on a 1 m/s straight path at fixed height."""
    m = env.model
    rng = np.random.default_rng(seed)
    K, F, nu = num_clips, frames, m.nu
    t = np.arange(F) / fps
    amp = rng.uniform(0, 0.4, (K, 3, nu))
    frq = rng.uniform(0.2, 1.5, (K, 3, nu))
    ph = rng.uniform(0, 2 * np.pi, (K, 3, nu))
    dof = (amp[:, :, None, :] * np.sin(2 * np.pi * frq[:, :, None, :] * t[None, None, :, None] + ph[:, :, None, :])).sum(1) / 3.0   # [K,F,nu]
    qpos = np.zeros((K, F, m.nq))
    qpos[..., 0] = t[None, :] * 1.0
    qpos[..., 2] = 0.94
    qpos[..., 3:7] = 0.5
    qpos[..., 7:] = dof
    flat_q = qpos.reshape(K * F, m.nq)
    xpos, xquat = env.kinematics(torch.as_tensor(flat_q, dtype=torch.float32))
    xpos = xpos.cpu().numpy().reshape(K, F, m.nbody, 3)
    xquat = xquat.cpu().numpy().reshape(K, F, m.nbody, 4)
    dt = 1.0 / fps
    body_vel = np.stack([_gauss1d(np.gradient(xpos[k], dt, axis=0), 2) for k in range(K)])
    dof_vel = np.stack([_gauss1d(np.gradient(dof[k], dt, axis=0), 2) for k in range(K)])
    # angular velocity from consecutive quaternions (world frame), gaussian filtered
    def qmul(a, b):
        w1, x1, y1, z1 = np.moveaxis(a, -1, 0); w2, x2, y2, z2 = np.moveaxis(b, -1, 0)
        return np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                         w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], -1)
    qn = np.concatenate([xquat[:, 1:], xquat[:, -1:]], 1)
    qc = xquat * np.array([1, -1, -1, -1.0])
    dq = qmul(qn, qc)
    dq *= np.sign(dq[..., :1] + 1e-12)
    ang = 2 * np.arctan2(np.linalg.norm(dq[..., 1:], axis=-1), dq[..., 0])
    ax = dq[..., 1:] / np.maximum(np.linalg.norm(dq[..., 1:], axis=-1, keepdims=True), 1e-9)
    body_ang_vel = np.stack([_gauss1d((ax * ang[..., None] / dt)[k], 2) for k in range(K)])
    qvel = np.zeros((K, F, m.nv))
    qvel[..., 0:3] = body_vel[:, :, 0]
    qvel[..., 3:6] = body_ang_vel[:, :, 0]
    qvel[..., 6:] = dof_vel
    f32 = lambda a: np.ascontiguousarray(a.reshape(K * F, -1), dtype=np.float32)  # noqa: E731
    return dict(qpos=f32(qpos), qvel=f32(qvel), xpos=f32(xpos), xquat=f32(xquat), body_vel=f32(body_vel), body_ang_vel=f32(body_ang_vel),
                dof_pos=f32(dof), dof_vel=f32(dof_vel), motion_num_frames=np.full(K, F, np.int32), motion_dt=np.full(K, dt, np.float32),
                motion_lengths=np.full(K, dt * (F - 1), np.float32), length_starts=(np.arange(K) * F).astype(np.int32),
                motion_aa=np.zeros((K * F, 3 * m.nbody), np.float32), motion_bodies=np.zeros((K, 17), np.float32),
                motion_fps=np.full(K, fps, np.float32), motion_keys=np.array([f"synthetic_{k:03d}" for k in range(K)]))


class MotionLibB200:
    """GPU-resident motion tables with the reference's query API."""

    def __init__(self, env: HumanoidBatchB200, tables: Dict[str, np.ndarray]):
        self.env = env
        dv = env.device
        self.tables = {k: torch.as_tensor(tables[k], dtype=torch.float32, device=dv).contiguous() for k in TABLE_KEYS}
        self._motion_num_frames = torch.as_tensor(tables["motion_num_frames"], dtype=torch.int32, device=dv)
        self._motion_dt = torch.as_tensor(tables["motion_dt"], dtype=torch.float32, device=dv)
        self._motion_lengths = torch.as_tensor(tables["motion_lengths"], dtype=torch.float32, device=dv)
        self.length_starts = torch.as_tensor(tables["length_starts"], dtype=torch.int32, device=dv)
        self._keys = TABLE_KEYS + tuple(k for k in OPTIONAL_KEYS if k in tables)
        for k in self._keys[len(TABLE_KEYS):]:
            self.tables[k] = torch.as_tensor(tables[k], dtype=torch.float32, device=dv).contiguous()
        self.widths = (C.c_int32 * len(self._keys))(*[self.tables[k].shape[1] for k in self._keys])
        self._tabs = (C.c_void_p * len(self._keys))(*[self.tables[k].data_ptr() for k in self._keys])
        K = int(self._motion_lengths.shape[0])
        # per-motion shape parameters (motion_lib_base.py:_motion_bodies, 17 = 16 betas + gender) and library bookkeeping
        self._motion_bodies = torch.as_tensor(tables["motion_bodies"], dtype=torch.float32, device=dv) if "motion_bodies" in tables else torch.zeros(K, 17, device=dv)
        self._motion_fps = torch.as_tensor(tables["motion_fps"], dtype=torch.float32, device=dv) if "motion_fps" in tables else 1.0 / self._motion_dt
        self._motion_data_keys = np.asarray(tables["motion_keys"]) if "motion_keys" in tables else np.array([f"clip_{i:04d}" for i in range(K)])
        self._num_unique_motions = K
        # sampling weights (PMCP), motion_lib_base.py:102-116,231-270: uniform until evaluation reports failed sequences
        self._sampling_prob = torch.full((K,), 1.0 / K, dtype=torch.float64, device=dv)
        self._sampling_batch_prob = self._sampling_prob.clone()
        self._termination_history = torch.zeros(K, dtype=torch.float64, device=dv)
        self.curr_failed_keys = []
        self._gen = torch.Generator(device=dv)
        self._gen.manual_seed(int(getattr(env, "seed", 0)) & 0x7FFFFFFF)

    def num_motions(self):
        return int(self._motion_lengths.shape[0])

    def get_motion_length(self, motion_ids=None):
        return self._motion_lengths if motion_ids is None else self._motion_lengths[motion_ids.long()]

    def get_motion_state_intervaled(self, motion_ids: torch.Tensor, motion_times: torch.Tensor, offset=None):
        dv = self.env.device
        ids = motion_ids.to(dv, torch.int32).contiguous()
        tm = motion_times.to(dv, torch.float32).contiguous()
        n = ids.shape[0]
        outs = {k: torch.empty(n, self.tables[k].shape[1], dtype=torch.float32, device=dv) for k in self._keys}
        optr = (C.c_void_p * len(self._keys))(*[outs[k].data_ptr() for k in self._keys])
        frame = torch.empty(n, dtype=torch.int32, device=dv)
        _lib.check(_lib.lib().smplsim_motion_gather(self.env._h, ids.data_ptr(), tm.data_ptr(), n, self._motion_lengths.data_ptr(),
                                                    self._motion_num_frames.data_ptr(), self._motion_dt.data_ptr(), self.length_starts.data_ptr(),
                                                    len(self._keys), self._tabs, self.widths, optr, frame.data_ptr(),
                                                    C.c_void_p(torch.cuda.current_stream(dv).cuda_stream)))
        self.env.gpu_launches += 1
        nb = self.env.model.nbody
        xpos = outs["xpos"].view(n, nb, 3)
        if offset is not None:
            xpos = xpos + offset[:, None, :]
        xquat, bv, bav = outs["xquat"].view(n, nb, 4), outs["body_vel"].view(n, nb, 3), outs["body_ang_vel"].view(n, nb, 3)
        return dict(root_pos=xpos[:, 0], root_rot=xquat[:, 0], dof_pos=outs["dof_pos"], root_vel=bv[:, 0], root_ang_vel=bav[:, 0],
                    dof_vel=outs["dof_vel"], xpos=xpos, xquat=xquat, body_vel=bv, body_ang_vel=bav, qpos=outs["qpos"], qvel=outs["qvel"],
                    motion_aa=outs.get("motion_aa"), motion_bodies=self._motion_bodies[ids.long()], frame_idx=frame)

    # ------------------------------------------------------------------ sampling (motion_lib_base.py:231-312), on the device
    def _uniform(self):
        K = self._num_unique_motions
        self._sampling_prob = torch.full((K,), 1.0 / K, dtype=torch.float64, device=self._sampling_prob.device)

    def _key_indexes(self, failed_keys):
        all_keys = self._motion_data_keys.tolist()
        return [all_keys.index(k) for k in failed_keys]

    def update_hard_sampling_weight(self, failed_keys):
        """Auto PMCP, hard: train only on the failed sequences (motion_lib_base.py:231-243)."""
        if len(failed_keys) > 0:
            idx = self._key_indexes(failed_keys)
            self._sampling_prob.zero_()
            self._sampling_prob[idx] = 1.0 / len(idx)
        else:
            self._uniform()
        self._sampling_batch_prob = self._sampling_prob / self._sampling_prob.sum()

    def update_soft_sampling_weight(self, failed_keys):
        """Auto PMCP, soft: weights proportional to the termination history (motion_lib_base.py:245-262)."""
        if len(failed_keys) > 0:
            self.curr_failed_keys = list(failed_keys)
            self._termination_history[self._key_indexes(failed_keys)] += 1
            self.update_sampling_prob(self._termination_history)
        else:
            self._uniform()
        self._sampling_batch_prob = self._sampling_prob / self._sampling_prob.sum()

    def update_sampling_prob(self, termination_history):
        th = torch.as_tensor(termination_history, dtype=torch.float64, device=self._sampling_prob.device)
        if th.numel() != self._sampling_prob.numel():
            return False
        self._sampling_prob = th / th.sum()
        self._termination_history = th.clone()
        self._sampling_batch_prob = self._sampling_prob.clone()
        return True

    def set_termination_history(self, termination_history):
        self._termination_history = torch.as_tensor(termination_history["termination_history"], dtype=torch.float64, device=self._sampling_prob.device)
        self.curr_failed_keys = termination_history["failed_keys"]
        self.update_sampling_prob(self._termination_history)

    def sample_motions(self, n: int = 1):
        """np.random.choice(p=_sampling_batch_prob, replace=True) of the reference (:272-274) as a device multinomial (its own
        Philox stream: distributional parity, the reference draws from the global numpy RNG)."""
        return torch.multinomial(self._sampling_batch_prob.float(), int(n), replacement=True, generator=self._gen).to(torch.int32)

    def sample_time(self, motion_ids, truncate_time=None):
        phase = torch.rand(motion_ids.shape, generator=self._gen, device=self._motion_lengths.device)
        motion_len = self._motion_lengths[motion_ids.long()].clone()
        if truncate_time is not None:
            assert truncate_time >= 0.0
            motion_len -= truncate_time
        return phase * motion_len

    def sample_time_interval(self, motion_ids, truncate_time=None):
        curr_fps = 1.0 / 30.0
        t = self.sample_time(motion_ids, truncate_time)
        return (t / curr_fps).long().float() * curr_fps

    def get_motion_num_steps(self, motion_ids=None):
        nf = self._motion_num_frames if motion_ids is None else self._motion_num_frames[motion_ids.long()]
        fps = self._motion_fps if motion_ids is None else self._motion_fps[motion_ids.long()]
        return (nf.float() * 30.0 / fps).to(torch.int32)
