#!/usr/bin/env python
"""Generate tests/golden/motion_fk.npz with the REFERENCE's Humanoid_Batch.fk_batch (smpl_sim/smpllib/torch_smpl_humanoid_batch.py:118-228:
SMPL axis-angle pose -> MuJoCo-ordered qpos/qvel, global body transforms, finite-difference + gaussian velocities).  The class
ctor needs the SMPL model files, so the unbound methods run on a namespace carrying what they read: `_offsets` (the shipped XML's
body offsets, rounded to 5 decimals like update_model does, :113), `_parents`, `smpl_2_mujoco`, `dt`, `filter_vel`."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402

HB = MG._import_with_stubs(lambda: __import__("smpl_sim.smpllib.torch_smpl_humanoid_batch", fromlist=["x"]))
from smpl_sim.smpllib.smpl_joint_names import SMPL_BONE_ORDER_NAMES, SMPL_MUJOCO_NAMES  # noqa: E402
from smplsim_b200.model import load_model  # noqa: E402


class _Dict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def main():
    HB.EasyDict = _Dict
    m = load_model("smpl")
    assert list(m.body_names) == list(SMPL_MUJOCO_NAMES)
    rng = np.random.default_rng(8)
    F, fps = 48, 30
    # smooth random motion: per-joint sinusoids in axis-angle (SMPL joint order), walking root
    t = np.arange(F) / fps
    amp = rng.uniform(0.0, 0.5, (24, 3)); frq = rng.uniform(0.3, 1.5, (24, 3)); ph = rng.uniform(0, 6.28, (24, 3))
    pose_aa = amp[None] * np.sin(2 * np.pi * frq[None] * t[:, None, None] + ph[None])
    pose_aa[:, 0] = np.array([1.2, 1.2, 1.2])[None] + 0.2 * np.sin(2 * np.pi * 0.5 * t)[:, None]   # root near the upright-start rotation
    pose_aa[:, 16, 2] += 2.0 * np.sin(2 * np.pi * 0.7 * t)          # one large shoulder swing to exercise the euler unwrap
    trans = np.stack([1.0 * t, 0.1 * np.sin(t), 0.9 + 0.02 * np.sin(3 * t)], axis=1)
    ns = types.SimpleNamespace(_offsets=torch.from_numpy(np.round(m.body_pos[None].astype(np.float32), decimals=5)),
                               _parents=[-1, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 17, 11, 19, 20, 21, 22],
                               smpl_2_mujoco=[SMPL_BONE_ORDER_NAMES.index(q) for q in SMPL_MUJOCO_NAMES], dt=1.0 / fps, filter_vel=True)
    for name in ("forward_kinematics_batch",):
        setattr(ns, name, types.MethodType(getattr(HB.Humanoid_Batch, name), ns))
    ns._compute_velocity = HB.Humanoid_Batch._compute_velocity
    ns._compute_angular_velocity = HB.Humanoid_Batch._compute_angular_velocity
    out = HB.Humanoid_Batch.fk_batch(ns, torch.from_numpy(pose_aa[None]).float(), torch.from_numpy(trans[None]).float(), return_full=True, count_offset=True)
    g = {k: out[k][0].numpy() for k in ("global_translation", "global_rotation", "global_velocity", "global_angular_velocity", "dof_pos", "dof_vels", "qpos", "qvel")}
    g.update(pose_aa=pose_aa.reshape(F, 72).astype(np.float32), trans=trans.astype(np.float32), fps=fps)
    np.savez_compressed(os.path.join(HERE, "motion_fk.npz"), **g)
    print("motion_fk.npz:", {k: v.shape for k, v in g.items() if hasattr(v, "shape")}, "max|dof_pos|", np.abs(g["dof_pos"]).max())


def main_smplh():
    """Same through the SMPL-H / SMPL-X branch of Humanoid_Batch (52 joints, :47-71)."""
    from smpl_sim.smpllib.smpl_joint_names import SMPLH_BONE_ORDER_NAMES, SMPLH_MUJOCO_NAMES
    HB.EasyDict = _Dict
    m = load_model("smplx")
    assert list(m.body_names) == list(SMPLH_MUJOCO_NAMES)
    rng = np.random.default_rng(9)
    F, fps = 24, 30
    t = np.arange(F) / fps
    amp = rng.uniform(0.0, 0.4, (52, 3)); frq = rng.uniform(0.3, 1.5, (52, 3)); ph = rng.uniform(0, 6.28, (52, 3))
    pose_aa = amp[None] * np.sin(2 * np.pi * frq[None] * t[:, None, None] + ph[None])
    pose_aa[:, 0] = np.array([1.2, 1.2, 1.2])[None]
    trans = np.stack([0.5 * t, 0 * t, 0.9 + 0 * t], axis=1)
    parents = [-1, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 17, 18, 19, 17, 21, 22, 17, 24, 25, 17, 27, 28, 17, 30, 31, 11, 33, 34,
               35, 36, 37, 38, 36, 40, 41, 36, 43, 44, 36, 46, 47, 36, 49, 50]
    ns = types.SimpleNamespace(_offsets=torch.from_numpy(np.round(m.body_pos[None].astype(np.float32), decimals=5)), _parents=parents,
                               smpl_2_mujoco=[SMPLH_BONE_ORDER_NAMES.index(q) for q in SMPLH_MUJOCO_NAMES], dt=1.0 / fps, filter_vel=True)
    ns.forward_kinematics_batch = types.MethodType(HB.Humanoid_Batch.forward_kinematics_batch, ns)
    ns._compute_velocity = HB.Humanoid_Batch._compute_velocity
    ns._compute_angular_velocity = HB.Humanoid_Batch._compute_angular_velocity
    out = HB.Humanoid_Batch.fk_batch(ns, torch.from_numpy(pose_aa[None]).float(), torch.from_numpy(trans[None]).float(), return_full=True, count_offset=True)
    g = {k: out[k][0].numpy() for k in ("global_translation", "global_rotation", "dof_pos", "qpos", "qvel")}
    g.update(pose_aa=pose_aa.reshape(F, 156).astype(np.float32), trans=trans.astype(np.float32), fps=fps)
    np.savez_compressed(os.path.join(HERE, "motion_fk_smplh.npz"), **g)
    print("motion_fk_smplh.npz:", {k: v.shape for k, v in g.items() if hasattr(v, "shape")})


if __name__ == "__main__" and "--sampling" not in sys.argv:
    main()
    main_smplh()


def gen_sampling():
    """PMCP weight rules of MotionLibBase (motion_lib_base.py:225-270) run unbound on a namespace -> motion_sampling.npz."""
    import ast
    import types
    # importing the module needs joblib / smplx / torch-geometry stubs (make_golden.py replaces it by a mock), so -- as for
    # _calc_frame_blend -- the methods under test are executed from the reference SOURCE TEXT
    src = open(os.path.join(MG.REF, "smpl_sim/smpllib/motion_lib_base.py")).read()
    want = ("update_hard_sampling_weight", "update_soft_sampling_weight", "update_sampling_prob", "set_termination_history")
    fns = [n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name in want]
    env = {"np": np, "print": lambda *a, **k: None}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "motion_lib_base.py", "exec"), env)
    K = 6
    ns = types.SimpleNamespace(_motion_data_keys=np.array([f"k{i}" for i in range(K)]), _num_unique_motions=K,
                               _sampling_prob=np.ones(K) / K, _termination_history=np.zeros(K), curr_failed_keys=[])
    for name in want:
        setattr(ns, name, types.MethodType(env[name], ns))
    out = {}
    ns.update_hard_sampling_weight(["k1", "k4"]); out["hard"] = ns._sampling_prob.copy()
    ns.update_hard_sampling_weight([]); out["hard_empty"] = ns._sampling_prob.copy()
    ns.update_soft_sampling_weight(["k0", "k2"]); ns.update_soft_sampling_weight(["k2", "k5"])
    out["soft2"] = ns._sampling_prob.copy(); out["soft2_hist"] = ns._termination_history.copy()
    hist = np.array([1.0, 0.0, 3.0, 2.0, 0.0, 4.0])
    ns.set_termination_history(dict(termination_history=hist.copy(), failed_keys=["k3"]))
    out["restore"] = ns._sampling_prob.copy(); out["restore_hist"] = hist
    np.savez_compressed(os.path.join(HERE, "motion_sampling.npz"), **out)
    print("motion_sampling.npz", {k: v.round(3).tolist() for k, v in out.items()})


if __name__ == "__main__" and "--sampling" in sys.argv:
    gen_sampling()
