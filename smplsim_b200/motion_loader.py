"""AMASS-style motion files -> the flat per-frame tables ``MotionLibB200`` serves.

Restates the data path of ``MotionLibSMPL.load_motion_with_skeleton`` (smpl_sim/smpllib/motion_lib_smpl.py:94-155) +
``Humanoid_Batch.fk_batch`` (smpl_sim/smpllib/torch_smpl_humanoid_batch.py:118-228) + the table concatenation of
``MotionLibBase.load_motions`` (smpl_sim/smpllib/motion_lib_base.py:127-208):

  file format  joblib / pickle dict ``{key: {"pose_aa": [F, 72 | 156], "trans": [F, 3] (or "trans_orig"), "fps": 30, ...}}``
  joints       SMPL order -> MuJoCo body order (``smpl_2_mujoco``), axis-angle -> rotation matrices
  root         ``trans + offset[Pelvis]`` (``count_offset``), root quaternion = pose of joint 0
  dof_pos      intrinsic XYZ euler angles of the 23 body joints (= the three hinges _x,_y,_z of each body), with the reference's
               (imperfect, frames 1..F-2 only) continuity fix
  FK           parent-ordered sweep over the XML's body offsets (rounded to 5 decimals like ``update_model``)
  velocities   forward differences (last frame repeated) + gaussian sigma = 2 (mode nearest) for body linear / angular velocities;
               dof velocities are raw forward differences; ``qvel = [root lin vel, R_root^T w_root, dof_vel]``

Differences, stated: the reference derives the offsets from SMPL betas (needs the SMPL model files, absent here) -- this loader
uses the offsets of the model it is given (the shipped XML = neutral zero-beta body), and ``fix_trans_height`` (mesh vertices)
is replaced by ``fix_height="geom"`` (lowest collision-geom point over the first 30 frames) or ``"no_fix"``."""
from __future__ import annotations

import pickle
from typing import Dict, Iterable, Optional, Union

import numpy as np

SMPL_BONE_ORDER_NAMES = ["Pelvis", "L_Hip", "R_Hip", "Torso", "L_Knee", "R_Knee", "Spine", "L_Ankle", "R_Ankle", "Chest", "L_Toe", "R_Toe",
                         "Neck", "L_Thorax", "R_Thorax", "Head", "L_Shoulder", "R_Shoulder", "L_Elbow", "R_Elbow", "L_Wrist", "R_Wrist",
                         "L_Hand", "R_Hand"]       # the SMPL kinematic-tree order AMASS poses come in (smpl_joint_names.py:19-44)
_FINGERS = [f + str(i) for f in ("Index", "Middle", "Pinky", "Ring", "Thumb") for i in (1, 2, 3)]
SMPLH_BONE_ORDER_NAMES = SMPL_BONE_ORDER_NAMES[:22] + ["L_" + f for f in _FINGERS] + ["R_" + f for f in _FINGERS]   # smpl_joint_names.py:46-99 (52)


def bone_order_for(model):
    """Pose-file joint order of the humanoid class: 24-body SMPL, or the 52-body SMPL-H / SMPL-X body (Humanoid_Batch uses
    SMPLH_BONE_ORDER_NAMES for both, torch_smpl_humanoid_batch.py:47-71)."""
    names = list(model.body_names)
    for order in (SMPL_BONE_ORDER_NAMES, SMPLH_BONE_ORDER_NAMES):
        if sorted(names) == sorted(order):
            return order
    raise NotImplementedError("fk_motion: the model is neither the 24-body SMPL nor the 52-body SMPL-H/X humanoid")


def aa_to_quat(aa: np.ndarray) -> np.ndarray:
    ang = np.linalg.norm(aa, axis=-1, keepdims=True)
    half = 0.5 * ang
    small = ang < 1e-6
    k = np.where(small, 0.5 - ang * ang / 48.0, np.sin(half) / np.where(small, 1.0, ang))
    return np.concatenate([np.cos(half), aa * k], axis=-1)


def quat_to_mat(q: np.ndarray) -> np.ndarray:
    w, x, y, z = np.moveaxis(q, -1, 0)
    s = 2.0 / (q * q).sum(-1)
    m = np.stack([1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w),
                  s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w),
                  s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)], -1)
    return m.reshape(q.shape[:-1] + (3, 3))


def mat_to_quat(m: np.ndarray) -> np.ndarray:
    """Rotation matrix -> unit quaternion wxyz with w >= 0 (best-conditioned branch per element)."""
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = [m[..., i, j] for i in range(3) for j in range(3)]
    q2 = np.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], -1)
    qa = np.sqrt(np.maximum(q2, 0.0))
    cand = np.stack([np.stack([q2[..., 0], m21 - m12, m02 - m20, m10 - m01], -1),
                     np.stack([m21 - m12, q2[..., 1], m10 + m01, m02 + m20], -1),
                     np.stack([m02 - m20, m10 + m01, q2[..., 2], m12 + m21], -1),
                     np.stack([m10 - m01, m20 + m02, m21 + m12, q2[..., 3]], -1)], -2)
    cand = cand / (2.0 * np.maximum(qa[..., None], 0.1))
    best = np.argmax(q2, axis=-1)
    q = np.take_along_axis(cand, best[..., None, None], axis=-2)[..., 0, :]
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    return q * np.where(q[..., :1] < 0, -1.0, 1.0)


def mat_to_euler_xyz(m: np.ndarray) -> np.ndarray:
    """Intrinsic XYZ angles (a, b, c) with R = Rx(a) Ry(b) Rz(c) -- pytorch3d ``matrix_to_euler_angles(., "XYZ")``."""
    b = np.arcsin(np.clip(m[..., 0, 2], -1.0, 1.0))
    a = np.arctan2(-m[..., 1, 2], m[..., 2, 2])
    c = np.arctan2(-m[..., 0, 1], m[..., 0, 0])
    return np.stack([a, b, c], -1)


def fix_continuous_dof(dof: np.ndarray) -> np.ndarray:
    """The reference's euler-flip repair (pytorch3d_transforms.py:749-771): frames 1..F-2, at most two passes per frame."""
    dof = dof.copy()
    T = dof.shape[0] - 1
    for t in range(1, T):
        diff = dof[t] - dof[t - 1]
        times = 0
        while np.abs(diff).max() >= 3:
            ch = np.abs(diff).sum(-1) >= 3
            d = dof[t][ch].copy()
            d[:, 0] = np.pi + d[:, 0]; d[:, 1] = np.pi - d[:, 1]; d[:, 2] = np.pi + d[:, 2]
            d[d > np.pi] -= 2 * np.pi
            d[d < -np.pi] += 2 * np.pi
            dof[t][ch] = d
            diff = dof[t] - dof[t - 1]
            times += 1
            if times > 1:
                break
    return dof


def _gauss(x):
    from scipy.ndimage import gaussian_filter1d
    return gaussian_filter1d(x, 2, axis=0, mode="nearest")


def _qmul(a, b):
    w1, x1, y1, z1 = np.moveaxis(a, -1, 0); w2, x2, y2, z2 = np.moveaxis(b, -1, 0)
    return np.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], -1)


def fk_motion(model, pose_aa: np.ndarray, trans: np.ndarray, fps: float, filter_vel: bool = True) -> Dict[str, np.ndarray]:
    """One clip: ``pose_aa [F, J, 3]`` (SMPL joint order, J = 24; SMPL-H order, J = 52), ``trans [F, 3]`` -> per-frame state in MuJoCo
    body order (fk_batch, return_full)."""
    names = list(model.body_names)
    order = bone_order_for(model)
    if pose_aa.shape[1] != len(order):
        raise ValueError(f"pose_aa has {pose_aa.shape[1]} joints, the model needs {len(order)}")
    s2m = [order.index(n) for n in names]
    F = pose_aa.shape[0]
    dt = 1.0 / float(fps)
    pose_quat = aa_to_quat(pose_aa.astype(np.float64))
    mats = quat_to_mat(pose_quat)[:, s2m]                              # [F, 24, 3, 3] local rotations, MuJoCo order
    off = np.round(np.asarray(model.body_pos, np.float64), 5)
    parent = list(model.body_parent)
    root_pos = trans.astype(np.float64) + off[0][None]
    gpos = np.zeros((F, len(names), 3)); gmat = np.zeros((F, len(names), 3, 3))
    for i, p in enumerate(parent):
        if p < 0:
            gpos[:, i], gmat[:, i] = root_pos, mats[:, 0]
        else:
            gpos[:, i] = np.einsum("fij,j->fi", gmat[:, p], off[i]) + gpos[:, p]
            gmat[:, i] = gmat[:, p] @ mats[:, i]
    gquat = mat_to_quat(gmat)
    vel = (gpos[1:] - gpos[:-1]) / dt
    vel = np.concatenate([vel, vel[-1:]], 0)
    dq = np.zeros_like(gquat); dq[..., 0] = 1.0
    d = _qmul(gquat[1:], gquat[:-1] * np.array([1.0, -1.0, -1.0, -1.0]))
    d = d * np.where(d[..., :1] < 0, -1.0, 1.0)
    dq[:-1] = d / np.linalg.norm(d, axis=-1, keepdims=True)
    ang = np.arccos(np.clip(2.0 * dq[..., 0] ** 2 - 1.0, -1.0, 1.0))
    axis = dq[..., 1:] / np.maximum(np.linalg.norm(dq[..., 1:], axis=-1, keepdims=True), 1e-10)
    angvel = axis * ang[..., None] / dt
    if filter_vel:
        vel, angvel = _gauss(vel), _gauss(angvel)
    dof_pos = fix_continuous_dof(mat_to_euler_xyz(mats)[:, 1:])        # [F, J-1, 3]
    dof_vel = (dof_pos[1:] - dof_pos[:-1]) / dt
    dof_vel = np.concatenate([dof_vel, dof_vel[-1:]], 0)
    qpos = np.concatenate([root_pos, pose_quat[:, 0], dof_pos.reshape(F, -1)], -1)
    local_w = np.einsum("fji,fj->fi", gmat[:, 0], angvel[:, 0])
    qvel = np.concatenate([vel[:, 0], local_w, dof_vel.reshape(F, -1)], -1)
    return dict(qpos=qpos, qvel=qvel, xpos=gpos, xquat=gquat, body_vel=vel, body_ang_vel=angvel, dof_pos=dof_pos.reshape(F, -1),
                dof_vel=dof_vel.reshape(F, -1))


def _lowest_geom_z(model, xpos, xquat) -> float:
    """Lowest point of the collision geoms over the given frames (capsule / sphere: centre - radius along z; box: lowest corner)."""
    R = quat_to_mat(xquat)
    lo = np.inf
    gb, gt = np.asarray(model.geom_body), np.asarray(model.geom_type)
    for g in range(len(gb)):
        b = int(gb[g])
        Rg = R[:, b] @ np.asarray(model.geom_mat[g]).reshape(3, 3)
        c = xpos[:, b] + np.einsum("fij,j->fi", R[:, b], np.asarray(model.geom_pos[g]))
        size = np.asarray(model.geom_size[g])
        if int(gt[g]) == 6:                                            # mjGEOM_BOX
            z = c[:, 2] - (np.abs(Rg[:, 2, :]) * size[None]).sum(-1)
        else:                                                          # capsule (half length size[1] along local z) / sphere
            hl = size[1] if int(gt[g]) == 3 else 0.0
            z = c[:, 2] - np.abs(Rg[:, 2, 2]) * hl - size[0]
        lo = min(lo, float(z.min()))
    return lo


def load_motion_file(path: str) -> Dict[str, dict]:
    try:
        import joblib
        return joblib.load(path)
    except ImportError:      # pragma: no cover
        with open(path, "rb") as f:
            return pickle.load(f)


def amass_tables(model, motions: Union[str, Dict[str, dict], Iterable[dict]], fix_height: str = "geom", randomize_heading: bool = False,
                 max_length: int = -1, seed: int = 0) -> Dict[str, np.ndarray]:
    """All clips -> the concatenated fp32 tables + ``length_starts / motion_lengths / motion_dt / motion_num_frames`` (motion_lib_base.py:
    173-198); the result feeds ``MotionLibB200(env, tables)``."""
    if isinstance(motions, str):
        motions = load_motion_file(motions)
    clips = list(motions.values()) if isinstance(motions, dict) else list(motions)
    rng = np.random.default_rng(seed)
    acc = {k: [] for k in ("qpos", "qvel", "xpos", "xquat", "body_vel", "body_ang_vel", "dof_pos", "dof_vel")}
    nfr, dts = [], []
    for clip in clips:
        fps = float(clip.get("fps", 30))
        pose = np.asarray(clip["pose_aa"], np.float64)
        trans = np.asarray(clip["trans"] if "trans" in clip else clip["trans_orig"], np.float64).copy()
        F = pose.shape[0]
        if max_length != -1 and F >= max_length:
            s = int(rng.integers(0, F - max_length + 1))
            pose, trans = pose[s:s + max_length], trans[s:s + max_length]
        nj = len(bone_order_for(model))
        if pose.shape[1] == 156 and nj == 24:                          # SMPL-H clip on the SMPL body: keep the 22 body joints, zero the hands (:128-129)
            pose = np.concatenate([pose[:, :66], np.zeros((pose.shape[0], 6))], 1)
        if pose.shape[1] == 72 and nj == 52:                           # SMPL clip on the SMPL-H/X body: body joints, open hands
            pose = np.concatenate([pose[:, :66], np.zeros((pose.shape[0], 90))], 1)
        if pose.shape[1] != 3 * nj:
            raise ValueError(f"pose_aa must be [F,72] or [F,156], got {pose.shape}")
        pose = pose.reshape(-1, nj, 3)
        if randomize_heading:                                          # :138-144
            yaw = np.pi * (2 * rng.random() - 1.0)
            qh = np.array([np.cos(yaw / 2), 0.0, 0.0, np.sin(yaw / 2)])
            q0 = _qmul(np.broadcast_to(qh, (pose.shape[0], 4)), aa_to_quat(pose[:, 0]))
            q0 = q0 * np.where(q0[:, :1] < 0, -1.0, 1.0)
            half = np.arccos(np.clip(q0[:, 0], -1, 1)); sn = np.maximum(np.sin(half), 1e-12)
            pose[:, 0] = q0[:, 1:] / sn[:, None] * (2 * half)[:, None]
            Rh = quat_to_mat(qh)
            trans = trans @ Rh.T
        st = fk_motion(model, pose, trans, fps)
        if fix_height == "geom":
            k = min(30, st["xpos"].shape[0])
            dz = _lowest_geom_z(model, st["xpos"][:k], st["xquat"][:k])
            st["qpos"][:, 2] -= dz; st["xpos"][..., 2] -= dz
        elif fix_height != "no_fix":
            raise NotImplementedError("fix_height: 'geom' | 'no_fix' (the reference's mesh-vertex modes need the SMPL model files)")
        for k in acc:
            acc[k].append(st[k].reshape(st[k].shape[0], -1))
        nfr.append(st["qpos"].shape[0]); dts.append(1.0 / fps)
    nfr = np.asarray(nfr, np.int32); dts = np.asarray(dts, np.float32)
    out = {k: np.ascontiguousarray(np.concatenate(v, 0), dtype=np.float32) for k, v in acc.items()}
    starts = np.concatenate([[0], np.cumsum(nfr)[:-1]]).astype(np.int32)
    out.update(motion_num_frames=nfr, motion_dt=dts, motion_lengths=(dts * (nfr - 1)).astype(np.float32), length_starts=starts)
    return out
