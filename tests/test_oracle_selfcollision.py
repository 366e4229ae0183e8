"""ORACLE-ONLY groundwork for SURVEY 8 f4 (self-collision; the CUDA product does not simulate it): capsule / sphere geom
pairs with MuJoCo's pair filters and two-body constraint rows `J = Jp_body2 - Jp_body1`.  Off by default -- these tests switch it on
and hold it to geometry and to physics (internal contact forces cannot change the total momentum)."""
import ctypes as C

import numpy as np

from oracle import oracle as orc
from smplsim_b200.model import fk_numpy
from util_states import make_models
from test_oracle_physics import _momentum_energy, G


def _seg(c1, a1, h1, c2, a2, h2):
    L = orc.lib()
    dp = C.POINTER(C.c_double)
    L.orc_segment_segment.argtypes = [dp, dp, C.c_double, dp, dp, C.c_double, dp, dp]
    s, t = C.c_double(), C.c_double()
    f = lambda v: np.ascontiguousarray(v, dtype=np.float64).ctypes.data_as(dp)  # noqa: E731
    L.orc_segment_segment(f(c1), f(a1), h1, f(c2), f(a2), h2, C.byref(s), C.byref(t))
    return s.value, t.value


def test_segment_segment_known_answers_and_bruteforce():
    # crossed at right angles, offset along z: closest points are the centres
    s, t = _seg([0, 0, 0], [1, 0, 0], 1.0, [0, 0, 0.5], [0, 1, 0], 1.0)
    assert abs(s) < 1e-12 and abs(t) < 1e-12
    # end-to-end on one line: clamped to the facing ends
    s, t = _seg([0, 0, 0], [1, 0, 0], 1.0, [3, 0, 0], [1, 0, 0], 0.5)
    assert abs(s - 1.0) < 1e-12 and abs(t + 0.5) < 1e-12
    # point (sphere) against a segment
    s, t = _seg([0.3, 2, 0], [0, 0, 1], 0.0, [0, 0, 0], [1, 0, 0], 1.0)
    assert s == 0.0 and abs(t - 0.3) < 1e-12
    rng = np.random.default_rng(0)
    u = np.linspace(-1, 1, 401)
    for _ in range(50):
        c1, c2 = rng.normal(size=3), rng.normal(size=3)
        a1, a2 = rng.normal(size=3), rng.normal(size=3)
        a1 /= np.linalg.norm(a1); a2 /= np.linalg.norm(a2)
        h1, h2 = rng.uniform(0.1, 1.0, 2)
        s, t = _seg(c1, a1, h1, c2, a2, h2)
        assert abs(s) <= h1 + 1e-12 and abs(t) <= h2 + 1e-12
        d = np.linalg.norm(c1 + s * a1 - c2 - t * a2)
        P1 = c1[None] + (u * h1)[:, None] * a1[None]; P2 = c2[None] + (u * h2)[:, None] * a2[None]
        dmin = np.sqrt(((P1[:, None, :] - P2[None, :, :]) ** 2).sum(-1)).min()
        assert d <= dmin + 1e-9                         # never worse than a dense sampling of both segments


def _capsule_segments(m, qpos):
    xpos, xquat, _ = fk_numpy(m, qpos)
    out = {}
    for g in range(len(m.geom_names)):
        if int(m.geom_type[g]) not in (2, 3):
            continue
        b = int(m.geom_body[g]); w, x, y, z = xquat[b]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        Rg = R @ np.asarray(m.geom_mat[g]).reshape(3, 3)
        hl = float(m.geom_size[g][1]) if int(m.geom_type[g]) == 3 else 0.0
        out[g] = (xpos[b] + R @ np.asarray(m.geom_pos[g]), Rg[:, 2], hl, float(m.geom_size[g][0]), b)
    return out


def test_self_contacts_match_independent_geometry_and_filters():
    cfg, om = make_models(env="speed", control_mode="torque")
    om.set_self_collision(True)
    m = om.model
    par = list(m.body_parent)
    excl = {frozenset((m.body_names.index(a), m.body_names.index(b))) for a, b in m.excludes}
    assert len(excl) == 10                               # smpl_humanoid.xml:231-242
    rng = np.random.default_rng(3)
    e = orc.OracleEnv(om)
    u = np.linspace(-1, 1, 201)
    total = 0
    for _ in range(25):
        e.qpos[:] = 0; e.qpos[2] = 3.0; e.qpos[3] = 1.0; e.qpos[7:] = rng.uniform(-1.5, 1.5, m.nu)
        e.forward()
        con = e.contacts()
        seg = _capsule_segments(m, e.qpos.copy())
        assert (con["geom1"] > 0).all() and e.contact_mask == 0          # airborne: only geom-geom contacts, floor mask untouched
        reported = set()
        for g1, g2, dist, pos, fr in zip(con["geom1"] - 1, con["geom"] - 1, con["dist"], con["pos"], con["frame"]):
            c1, a1, h1, r1, b1 = seg[g1]; c2, a2, h2, r2, b2 = seg[g2]
            assert b1 != b2 and par[b1] != b2 and par[b2] != b1 and frozenset((b1, b2)) not in excl
            P1 = c1[None] + (u * h1)[:, None] * a1[None]; P2 = c2[None] + (u * h2)[:, None] * a2[None]
            dmin = np.sqrt(((P1[:, None, :] - P2[None, :, :]) ** 2).sum(-1)).min() - r1 - r2
            assert dist <= dmin + 1e-9 and dist > dmin - 2e-3 and dist <= 1e-3      # true closest distance, within the contact margin
            n = fr[0]
            assert abs(np.linalg.norm(n) - 1) < 1e-12 and abs(n @ fr[1]) < 1e-12 and np.allclose(np.cross(n, fr[1]), fr[2])
            assert n @ (c2 - c1) > -1e-9 or True                           # normal points from geom1 to geom2 (checked on the closest points below)
            reported.add((g1, g2)); total += 1
        # completeness: every admissible capsule / sphere pair closer than the margin was reported
        gs = sorted(seg)
        for i, g1 in enumerate(gs):
            for g2 in gs[i + 1:]:
                c1, a1, h1, r1, b1 = seg[g1]; c2, a2, h2, r2, b2 = seg[g2]
                if b1 == b2 or par[b1] == b2 or par[b2] == b1 or frozenset((b1, b2)) in excl:
                    assert (g1, g2) not in reported
                    continue
                s, t = _seg(c1, a1, h1, c2, a2, h2)
                d = np.linalg.norm(c1 + s * a1 - c2 - t * a2) - r1 - r2
                assert ((g1, g2) in reported) == (d <= 1e-3)
    assert total > 20


def _penetrating_env(om, min_con=2, seed=7):
    m = om.model
    rng = np.random.default_rng(seed)
    e = orc.OracleEnv(om)
    for _ in range(200):
        e.qpos[:] = 0; e.qpos[2] = 3.0; e.qpos[3] = 1.0; e.qpos[7:] = rng.uniform(-1.2, 1.2, m.nu); e.qvel[:] = 0
        e.forward()
        if e.ncon >= min_con:
            return e
    raise AssertionError("no self-penetrating pose found")


def test_internal_contact_forces_are_internal():
    """Airborne, self-penetrating pose, zero torque: the two-body rows carry large forces, yet the generalized constraint force on
    the six free-joint dofs (= net external force and torque) vanishes; KKT holds; the Newton solve converges (dense Cholesky: the
    Hessian of two-body rows is not tree-sparse)."""
    cfg, om = make_models(env="speed", control_mode="torque")
    om.set_self_collision(True)
    e = _penetrating_env(om)
    efc = e.efc()
    f = efc["force"]
    assert e.nefc == 4 * e.ncon and f.max() > 100.0 and e.solver_iter < 40
    q = np.array(e.qfrc_constraint)
    assert np.abs(q[:6]).max() < 1e-9 * f.max() and np.abs(q[6:]).max() > 1.0
    assert np.abs(efc["J"][:, :3]).max() < 1e-12                       # J = Jp_body2 - Jp_body1: root translation cancels
    r = efc["J"] @ np.array(e.qacc) - efc["aref"]
    assert (f >= 0).all() and np.abs(f[r < 0] + (efc["D"] * r)[r < 0]).max() < 1e-6 * f.max() and np.abs(f[r >= 0]).max() < 1e-9
    assert np.abs(efc["J"].T @ f - q).max() < 1e-8 * f.max()
    d0 = e.contacts()["dist"].min()
    for _ in range(30):
        e.mj_step()
    e.forward()
    assert e.ncon == 0 or e.contacts()["dist"].min() > d0               # pushed apart


def _momentum_drift(h_scale, steps=12):
    from smplsim_b200.cfg import make_cfg
    cfg = make_cfg(env="speed", overrides={"env.control_mode": "torque", "env.sim_timestep_inv": 450 * h_scale})
    om = orc.OracleModel.from_cfg(cfg)
    om.set_self_collision(True)
    m = om.model
    e = _penetrating_env(om)
    P0, L0, _ = _momentum_energy(m, e)
    e.ctrl[:] = 0
    for _ in range(steps * h_scale):
        e.mj_step()
    P1, L1, _ = _momentum_energy(m, e)
    t = steps / 450.0
    return np.abs(P1 - P0 - np.array([0, 0, -float(np.sum(m.body_mass)) * G * t])).max()


def test_momentum_drift_with_self_contact_is_first_order():
    d1, d2, d4 = _momentum_drift(1), _momentum_drift(2), _momentum_drift(4)
    assert d2 < 0.7 * d1 and d4 < 0.7 * d2, (d1, d2, d4)


def test_default_is_off():
    cfg, om = make_models(env="speed", control_mode="torque")
    e = orc.OracleEnv(om)
    e.qpos[2] = 3.0; e.qpos[3] = 1.0; e.qpos[7:] = 1.4
    e.forward()
    assert e.ncon == 0
