"""Second and third witnesses for the oracle's mj_forward on contact states (SURVEY 8c G4 is impossible here: no `mujoco`).

The oracle solves MuJoCo's soft-constraint problem in its PRIMAL form (Newton + exact line search on qacc, CRB mass matrix,
analytic Jacobian).  This file re-derives the same accelerations along a path that shares no code and no algorithm with it:

  * mass matrix: smplsim_b200.model.mass_matrix_numpy (sum_b J_b^T I_b J_b from numpy kinematics);
  * constraint Jacobian: NUMERICAL -- central differences of the contact point carried by its body (numpy kinematics) along
    every dof, projected on the pyramid directions n +- mu t1, n +- mu t2 built from the reported contact frame;
  * solver: the DUAL problem  min_f 1/2 f^T (A + R) f + f^T (J a0 - aref),  f >= 0,  A = J M^-1 J^T, R = 1 / D,  solved exactly
    by Lawson-Hanson non-negative least squares (scipy.optimize.nnls) on the Cholesky factor; qacc = a0 + M^-1 J^T f.

Agreement of qacc to 1e-7 relative and of the row forces to 1e-6 on stumbling / fallen rollout states pins the oracle's
solver, Jacobian and mass matrix against an independent derivation; what stays unpinned against real MuJoCo is the
*modelling* (contact frames of near-vertical capsules, diagApprox, solref mixing -- the VERIFY items of SURVEY App. A)."""
import numpy as np
import pytest
from scipy.linalg import cho_factor, cho_solve, cholesky, solve_triangular
from scipy.optimize import nnls

from oracle import oracle as orc
from smplsim_b200.mjcf import quat_to_mat
from smplsim_b200.model import fk_numpy, mass_matrix_numpy
from util_states import make_models, rollout_states


def _integrate(m, q, dq, eps):
    """qpos moved by eps along the velocity direction dq (free-joint translation in the world frame, rotation in the body frame)."""
    q2 = q.copy()
    q2[0:3] += eps * dq[0:3]
    w = dq[3:6] * eps
    ang = np.linalg.norm(w)
    if ang > 0:
        ax = w / ang
        dqt = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax])
        a, b = q[3:7], dqt
        q2[3:7] = [a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                   a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]]
    q2[7:] += eps * dq[6:]
    return q2


def _point_world(m, q, body, local):
    xp, xq, _ = fk_numpy(m, q)
    return xp[body] + quat_to_mat(xq[body]) @ local


def _numeric_contact_jacobian(m, q, con, mu):
    """4 pyramid rows per contact by central differences of the body-fixed contact point."""
    xp, xq, _ = fk_numpy(m, q)
    rows = []
    for c in range(len(con["dist"])):
        g = int(con["geom"][c]) - 1
        b = int(m.geom_body[g])
        n, t1, t2 = con["frame"][c]
        local = quat_to_mat(xq[b]).T @ (con["pos"][c] - xp[b])
        Jp = np.zeros((3, m.nv))
        eps = 1e-6
        for j in range(m.nv):
            e = np.zeros(m.nv); e[j] = 1.0
            Jp[:, j] = (_point_world(m, _integrate(m, q, e, eps), b, local) - _point_world(m, _integrate(m, q, e, -eps), b, local)) / (2 * eps)
        for d in (n + mu * t1, n - mu * t1, n + mu * t2, n - mu * t2):
            rows.append(d @ Jp)
    return np.array(rows)


def test_dual_nnls_with_numeric_jacobian_reproduces_the_oracle():
    cfg, om = make_models(control_mode="torque")
    m = om.model
    q, v, w = rollout_states(make_models(control_mode="uhc_pd")[1], 64, seed=11)
    rng = np.random.default_rng(2)
    ncheck = 0
    worst_a = worst_f = worst_j = worst_m = 0.0
    for i in range(64):
        e = orc.OracleEnv(om)
        e.qpos[:] = q[i]; e.qvel[:] = v[i]; e.qacc_warm[:] = w[i]; e.ctrl[:] = rng.uniform(-80, 80, m.nu)
        e.forward()
        if e.ncon == 0:
            continue
        efc, con = e.efc(), e.contacts()
        J, aref, D, force = efc["J"], efc["aref"], efc["D"], efc["force"]
        nlim = J.shape[0] - 4 * e.ncon
        # ---- witness 1: mass matrix
        M = mass_matrix_numpy(m, e.qpos.copy())
        M = M[0] if isinstance(M, tuple) else M
        worst_m = max(worst_m, np.abs(M - e.M).max() / np.abs(e.M).max())
        # ---- witness 2: numerical Jacobian of the contact rows (limit rows, if any, are +-1 on one dof: kept from the oracle)
        crow = [r for r in range(J.shape[0]) if np.count_nonzero(J[r]) > 1 or nlim == 0]
        if nlim:
            lim_rows = [r for r in range(J.shape[0]) if np.count_nonzero(J[r]) == 1 and abs(abs(J[r]).max() - 1) < 1e-12]
            crow = [r for r in range(J.shape[0]) if r not in lim_rows]
            assert len(lim_rows) == nlim
        Jn = J.copy()
        Jc = _numeric_contact_jacobian(m, e.qpos.copy(), con, float(m.friction[0]))
        assert Jc.shape[0] == len(crow)
        worst_j = max(worst_j, np.abs(Jc - J[crow]).max())
        Jn[crow] = Jc
        # ---- witness 3: dual NNLS
        cf = cho_factor(M)
        a0 = e.qacc_smooth.copy()
        A = Jn @ cho_solve(cf, Jn.T) + np.diag(1.0 / D)
        b = Jn @ a0 - aref
        U = cholesky(A)                       # A = U^T U
        f, _ = nnls(U, -solve_triangular(U, b, trans="T"), maxiter=20 * len(b))
        qacc = a0 + cho_solve(cf, Jn.T @ f)
        worst_a = max(worst_a, np.abs(qacc - e.qacc).max() / max(1.0, np.abs(e.qacc).max()))
        worst_f = max(worst_f, np.abs(f - force).max() / max(1.0, np.abs(force).max()))
        ncheck += 1
    print(f"{ncheck} contact states: M {worst_m:.1e}, numeric J {worst_j:.1e}, qacc {worst_a:.1e}, forces {worst_f:.1e}")
    assert ncheck >= 30
    assert worst_m < 1e-9 and worst_j < 1e-6 and worst_a < 1e-6 and worst_f < 1e-5
