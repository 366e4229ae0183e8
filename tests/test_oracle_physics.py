"""Analytic known-answer tests that pin the oracle's physics restatement (SURVEY.md 8c G1 / G3, App. C KAT list) -- there is no
golden mj_step vector in the reference and no `mujoco` wheel here ("parity unpinned"), so the restatement is held to physics:
model-table checksums, M against an independent numpy sum of J^T I J, free fall in closed form, momentum / energy conservation of
the unforced airborne body with first-order convergence in h, static stance force balance, soft-contact KKT conditions."""
import numpy as np
import pytest

from oracle import oracle as orc
from smplsim_b200.cfg import make_cfg
from smplsim_b200.model import fk_numpy, load_model, mass_matrix_numpy
from util_states import airborne_states, make_models

G = 9.81


@pytest.mark.parametrize("name,nb,nq,nv,nu,mass", [("smpl", 24, 76, 75, 69, 71.81), ("smplx", 52, 160, 159, 153, 73.57)])
def test_model_table_checksums(name, nb, nq, nv, nu, mass):
    """G1: sizes, tree, total mass (survey-computed from the XMLs)."""
    m = load_model(name)
    assert (m.nbody, m.nq, m.nv, m.nu) == (nb, nq, nv, nu)
    assert abs(float(np.sum(m.body_mass)) - mass) < 0.01
    par = np.asarray(m.body_parent)
    assert par[0] == -1 and (par[1:] < np.arange(1, nb)).all() and (par[1:] >= 0).all()
    if name == "smpl":
        assert list(par) == [-1, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 17, 11, 19, 20, 21, 22]   # torch_smpl_humanoid_batch.py:45
        assert m.body_names[0] == "Pelvis" and m.body_names[-1] == "R_Hand"
    assert np.allclose(m.dof_armature[6:], 0.01) and np.allclose(m.dof_armature[:6], 0.0)


@pytest.mark.parametrize("robot", ["smpl_humanoid", "smplx_humanoid"])
def test_mass_matrix_equals_independent_jacobian_sum(robot):
    cfg, om = make_models(robot=robot)
    m = om.model
    q, v = airborne_states(m, 3, seed=2)
    for i in range(3):
        e = orc.OracleEnv(om)
        e.qpos[:] = q[i]; e.qvel[:] = v[i]
        e.forward()
        M_ref, _ = mass_matrix_numpy(m, e.qpos.copy())
        M = np.array(e.M)
        M = np.tril(M) + np.tril(M, -1).T
        assert np.abs(M - M_ref).max() < 1e-9 * max(1.0, np.abs(M_ref).max())


def _momentum_energy(m, e):
    """Linear momentum, angular momentum about the origin, kinetic + potential energy from body-frame velocities (sensor = framelinvel /
    frameangvel of the last forward pass)."""
    e.forward()
    xpos, xquat, _ = fk_numpy(m, e.qpos.copy())
    P = np.zeros(3); L = np.zeros(3); E = 0.0
    for b in range(m.nbody):
        w, x, y, z = xquat[b]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        c = xpos[b] + R @ m.body_ipos[b]
        om_ = e.sensor[1, b]
        vc = e.sensor[0, b] + np.cross(om_, c - xpos[b])
        I = np.asarray(m.body_inertia[b])
        Il = np.array([[I[0], I[3], I[4]], [I[3], I[1], I[5]], [I[4], I[5], I[2]]]) if I.size == 6 else np.diag(I)
        Iw = R @ Il @ R.T
        mb = m.body_mass[b]
        P += mb * vc
        L += mb * np.cross(c, vc) + Iw @ om_
        E += 0.5 * mb * vc @ vc + 0.5 * om_ @ Iw @ om_ + mb * G * c[2]
    return P, L, E


def _unforced_drift(h_scale, n_steps, seed=5):
    cfg = make_cfg(env="speed", overrides={"env.control_mode": "torque", "env.sim_timestep_inv": int(450 * h_scale)})
    om = orc.OracleModel.from_cfg(cfg)
    m = om.model
    q, v = airborne_states(m, 1, seed=seed, vel=1.0)
    e = orc.OracleEnv(om)
    e.qpos[:] = q[0]; e.qvel[:] = v[0]; e.ctrl[:] = 0.0
    P0, L0, E0 = _momentum_energy(m, e)
    # joint armature is a reflected rotor inertia: it carries kinetic energy / momentum the body sums above do not see;
    # keep the joints nearly locked by zero joint velocity so the comparison is clean
    for _ in range(n_steps * h_scale):
        e.mj_step()
    P1, L1, E1 = _momentum_energy(m, e)
    t = n_steps / 450.0
    mass = float(np.sum(m.body_mass))
    return np.abs(P1 - P0 - np.array([0, 0, -mass * G * t])).max(), E1 - E0, (P0, P1)


def test_free_fall_closed_form_and_linear_momentum():
    cfg = make_cfg(env="speed", overrides={"env.control_mode": "torque"})
    om = orc.OracleModel.from_cfg(cfg)
    m = om.model
    e = orc.OracleEnv(om)
    e.qpos[2] = 5.0; e.qpos[3] = 1.0
    n, h = 90, 1.0 / 450.0
    for _ in range(n):
        e.mj_step()
    # semi-implicit Euler: v_n = -g h n ; z_n = z0 - g h^2 n (n + 1) / 2   (rigid: zero joint velocity stays zero)
    assert abs(e.qvel[2] + G * h * n) < 1e-9
    assert abs(e.qpos[2] - (5.0 - G * h * h * n * (n + 1) / 2)) < 1e-9
    assert np.abs(e.qvel[6:]).max() < 1e-9 and np.abs(e.qvel[3:6]).max() < 1e-9
    # tumbling articulated body: total linear momentum changes by -m g t only -- exactly in continuous time, to first order in h for
    # semi-implicit Euler (M(q) moves between the velocity and the position update)
    dP1, _, (P0, _) = _unforced_drift(1, 45)
    dP2, _, _ = _unforced_drift(2, 45)
    dP4, _, _ = _unforced_drift(4, 45)
    assert dP1 < 2e-3 * max(1.0, np.abs(P0).max()) and dP2 < 0.65 * dP1 and dP4 < 0.65 * dP2, (dP1, dP2, dP4, P0)


def test_energy_drift_is_first_order_in_h():
    """Unforced airborne tumbling: the energy error of semi-implicit Euler over a fixed time shrinks ~ linearly with h."""
    _, dE1, _ = _unforced_drift(1, 30)
    _, dE2, _ = _unforced_drift(2, 30)
    _, dE4, _ = _unforced_drift(4, 30)
    assert abs(dE1) < 0.05 * 300                  # sane magnitude (KE + PE ~ kJ scale)
    assert abs(dE2) < 0.75 * abs(dE1) and abs(dE4) < 0.75 * abs(dE2), (dE1, dE2, dE4)


def test_static_stance_force_balance_and_kkt():
    """Zero-action stable PD from the Default pose: after 5 env steps the floor contacts are on the feet only; the pose is not statically
    stable and the body comes to rest on the ground within 2 s -- at rest the constraint force on the root's z translation balances
    m g, every active pyramid row pushes (f >= 0) with f = -D r, inactive rows carry no force (soft-contact KKT)."""
    cfg, om = make_models(env="speed")
    m = om.model
    e = orc.OracleEnv(om)
    e.reset()
    feet = {m.body_names.index(n) for n in ("L_Ankle", "L_Toe", "R_Ankle", "R_Toe")}
    for t in range(60):
        e.step(np.zeros(m.nu))
        if t == 4:
            con = e.contacts()
            assert len(con["geom"]) >= 4 and {int(m.geom_body[g - 1]) for g in con["geom"]} <= feet
    e.forward()
    mass = float(np.sum(m.body_mass))
    assert np.abs(e.qvel).max() < 0.5
    assert abs(e.qfrc_constraint[2] - mass * G) < 0.03 * mass * G, (e.qfrc_constraint[2], mass * G)
    efc = e.efc()
    r = efc["J"] @ np.array(e.qacc) - efc["aref"]
    f = efc["force"]
    assert (f >= -1e-9).all()
    assert np.abs(f[r < 0] + (efc["D"] * r)[r < 0]).max() < 1e-6 * max(1.0, np.abs(f).max())
    assert np.abs(f[r >= 0]).max() < 1e-9 if (r >= 0).any() else True
    # newton's third law in joint space: qfrc_constraint = J^T f
    assert np.abs(efc["J"].T @ f - np.array(e.qfrc_constraint)).max() < 1e-8 * max(1.0, np.abs(f).max())


@pytest.mark.parametrize("name", ["smpl", "smplx"])
def test_oracle_matches_mujoco_g4(name):
    """G4: the restatement against real MuJoCo output (tests/golden/make_golden_mujoco.py).  Skipped until that fixture can be
    generated -- there is no `mujoco` wheel in this image, hence DESIGN.md's "parity unpinned" for the physics."""
    import os
    from conftest import GOLDEN
    path = os.path.join(GOLDEN, "mj_step_g4.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/mj_step_g4.npz absent (needs a real mujoco wheel): physics parity unpinned")
    g = np.load(path)
    cfg, om = make_models(robot=f"{name}_humanoid", control_mode="torque")
    m = om.model
    assert np.abs(np.asarray(m.body_mass) - g[f"{name}.body_mass"]).max() < 1e-6
    n_checked = 0
    for i in range(g[f"{name}.qpos"].shape[0]):
        if g[f"{name}.ncon_self"][i]:
            continue                                         # self-collision is not simulated
        e = orc.OracleEnv(om)
        e.qpos[:] = g[f"{name}.qpos"][i]; e.qvel[:] = g[f"{name}.qvel"][i]; e.qacc_warm[:] = g[f"{name}.qacc_warm"][i]
        e.ctrl[:] = g[f"{name}.ctrl"][i]
        e.mj_step()
        ref_q, ref_v = g[f"{name}.qpos1"][i], g[f"{name}.qvel1"][i]
        assert np.abs(e.qpos - ref_q).max() / max(1.0, np.abs(ref_q).max()) < 1e-6, i
        assert np.abs(e.qvel - ref_v).max() / max(1.0, np.abs(ref_v).max()) < 1e-6, i
        assert e.contact_mask == int(g[f"{name}.floor_geoms"][i]), i
        n_checked += 1
    assert n_checked > 0
