#!/usr/bin/env python
"""G4: pin the physics restatement against the REAL MuJoCo -- for the day a `mujoco` wheel is importable next to
/root/reference (it is not in this image: "parity unpinned").  Writes tests/golden/mj_step_g4.npz; tests/test_oracle_physics.py::
test_oracle_matches_mujoco_g4 picks it up automatically (and is skipped while the file is absent).

    pip install mujoco            # wherever that is possible
    python tests/golden/make_golden_mujoco.py [/root/reference]

Per state: qpos, qvel, qacc_warmstart, ctrl -> after one mj_step at h = 1/450: qpos', qvel', qacc, the floor-contact geom set,
plus the model constants the restatement derives itself (body_mass, body_invweight0, dof_invweight0, efc_R / efc_aref of the first
state with contacts) -- each "VERIFY" item of SURVEY App. A maps to one of these arrays."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("SMPLSIM_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main():
    import mujoco                                            # the real one
    from util_states import airborne_states, make_models, rollout_states
    out = {}
    for name, xml in (("smpl", os.path.join(REF, "smpl_sim/data/assets/mjcf/smpl_humanoid.xml")), ("smplx", os.path.join(REF, "smpl_humanoid.xml"))):
        mm = mujoco.MjModel.from_xml_path(xml)
        mm.opt.timestep = 1.0 / 450.0                        # base_env.py:142
        md = mujoco.MjData(mm)
        cfg, om = make_models(robot=f"{name}_humanoid", control_mode="torque")
        m = om.model
        assert (mm.nq, mm.nv, mm.nu) == (m.nq, m.nv, m.nu)
        n = 200 if name == "smpl" else 40
        q1, v1, w1 = rollout_states(make_models(robot=f"{name}_humanoid")[1], n, seed=31)
        q2, v2 = airborne_states(m, n // 4, seed=32)
        q = np.concatenate([q1, q2]); v = np.concatenate([v1, v2]); w = np.concatenate([w1, np.zeros_like(v2)])
        rng = np.random.default_rng(33)
        ctrl = rng.uniform(-50, 50, (q.shape[0], m.nu))
        res = dict(qpos=q, qvel=v, qacc_warm=w, ctrl=ctrl, qpos1=[], qvel1=[], qacc=[], floor_geoms=[], ncon_self=[])
        floor = mujoco.mj_name2id(mm, mujoco.mjtObj.mjOBJ_GEOM, "floor")
        first_efc = None
        for i in range(q.shape[0]):
            mujoco.mj_resetData(mm, md)
            md.qpos[:] = q[i]; md.qvel[:] = v[i]; md.qacc_warmstart[:] = w[i]; md.ctrl[:] = ctrl[i]
            mujoco.mj_step(mm, md)
            res["qpos1"].append(md.qpos.copy()); res["qvel1"].append(md.qvel.copy()); res["qacc"].append(md.qacc.copy())
            mask = 0; nself = 0
            for c in md.contact[: md.ncon]:
                if c.geom1 == floor:
                    mask |= 1 << int(c.geom2)
                else:
                    nself += 1                               # self-collision contacts: NOT simulated by the restatement
            res["floor_geoms"].append(mask); res["ncon_self"].append(nself)
            if first_efc is None and md.nefc > 0 and nself == 0:
                first_efc = dict(efc_R=md.efc_R[: md.nefc].copy(), efc_aref=md.efc_aref[: md.nefc].copy(), efc_D=md.efc_D[: md.nefc].copy(), state=i)
        for k in ("qpos1", "qvel1", "qacc"):
            res[k] = np.array(res[k])
        res["floor_geoms"] = np.array(res["floor_geoms"], dtype=np.uint64); res["ncon_self"] = np.array(res["ncon_self"])
        res.update(body_mass=mm.body_mass[1:].copy(), body_invweight0=mm.body_invweight0[1:].copy(), dof_invweight0=mm.dof_invweight0.copy(),
                   body_inertia=mm.body_inertia[1:].copy(), body_ipos=mm.body_ipos[1:].copy(), version=np.array(mujoco.__version__))
        if first_efc:
            res.update({"efc_" + k: val for k, val in first_efc.items()})
        out.update({f"{name}.{k}": val for k, val in res.items()})
    np.savez_compressed(os.path.join(HERE, "mj_step_g4.npz"), **out)
    print("mj_step_g4.npz written:", len(out), "arrays; states without self-contact are the ones the 1e-4 criterion applies to")


if __name__ == "__main__":
    main()
