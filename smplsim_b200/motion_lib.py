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
                motion_lengths=np.full(K, dt * (F - 1), np.float32), length_starts=(np.arange(K) * F).astype(np.int32))


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
        self.widths = (C.c_int32 * len(TABLE_KEYS))(*[self.tables[k].shape[1] for k in TABLE_KEYS])
        self._tabs = (C.c_void_p * len(TABLE_KEYS))(*[self.tables[k].data_ptr() for k in TABLE_KEYS])

    def num_motions(self):
        return int(self._motion_lengths.shape[0])

    def get_motion_length(self, motion_ids=None):
        return self._motion_lengths if motion_ids is None else self._motion_lengths[motion_ids.long()]

    def get_motion_state_intervaled(self, motion_ids: torch.Tensor, motion_times: torch.Tensor, offset=None):
        dv = self.env.device
        ids = motion_ids.to(dv, torch.int32).contiguous()
        tm = motion_times.to(dv, torch.float32).contiguous()
        n = ids.shape[0]
        outs = {k: torch.empty(n, self.tables[k].shape[1], dtype=torch.float32, device=dv) for k in TABLE_KEYS}
        optr = (C.c_void_p * len(TABLE_KEYS))(*[outs[k].data_ptr() for k in TABLE_KEYS])
        frame = torch.empty(n, dtype=torch.int32, device=dv)
        _lib.check(_lib.lib().smplsim_motion_gather(self.env._h, ids.data_ptr(), tm.data_ptr(), n, self._motion_lengths.data_ptr(),
                                                    self._motion_num_frames.data_ptr(), self._motion_dt.data_ptr(), self.length_starts.data_ptr(),
                                                    len(TABLE_KEYS), self._tabs, self.widths, optr, frame.data_ptr(),
                                                    C.c_void_p(torch.cuda.current_stream(dv).cuda_stream)))
        self.env.gpu_launches += 1
        nb = self.env.model.nbody
        xpos = outs["xpos"].view(n, nb, 3)
        if offset is not None:
            xpos = xpos + offset[:, None, :]
        xquat, bv, bav = outs["xquat"].view(n, nb, 4), outs["body_vel"].view(n, nb, 3), outs["body_ang_vel"].view(n, nb, 3)
        return dict(root_pos=xpos[:, 0], root_rot=xquat[:, 0], dof_pos=outs["dof_pos"], root_vel=bv[:, 0], root_ang_vel=bav[:, 0],
                    dof_vel=outs["dof_vel"], xpos=xpos, xquat=xquat, body_vel=bv, body_ang_vel=bav, qpos=outs["qpos"], qvel=outs["qvel"],
                    frame_idx=frame)
