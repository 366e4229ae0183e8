"""Seeded state sets for parity tests, produced with the CPU oracle (test infrastructure)."""
import numpy as np

from oracle import oracle as orc
from smplsim_b200.cfg import make_cfg


def make_models(env="speed", robot="smpl_humanoid", seed=0, **ov):
    overrides = {(k if "." in k else f"env.{k}"): v for k, v in ov.items()}
    cfg = make_cfg(env=env, robot=robot, overrides=overrides)
    om = orc.OracleModel.from_cfg(cfg, seed=seed)
    return cfg, om


def rollout_states(om, n_states, seed=0, every=3, control_sigma=0.3, init_mode=0, max_steps=400):
    """Diverse (qpos, qvel, qacc_warm) snapshots: standing, stumbling, falling, lying -- taken between substeps."""
    rng = np.random.default_rng(seed)
    m = om.model
    out = []
    e = orc.OracleEnv(om, env_id=0)
    while len(out) < n_states:
        e.reset(init_mode)
        sigma = control_sigma * rng.uniform(0.2, 2.0)
        for t in range(max_steps):
            a = np.clip(rng.normal(size=m.nu) * sigma, -1, 1)
            e.step(a)
            if t % every == 0:
                out.append((e.qpos.copy(), e.qvel.copy(), e.qacc_warm.copy()))
                if len(out) >= n_states:
                    break
            if e.qpos[2] < 0.12 and rng.random() < 0.1:
                break
    q = np.array([s[0] for s in out]); v = np.array([s[1] for s in out]); w = np.array([s[2] for s in out])
    return q, v, w


def airborne_states(m, n, seed=0, vel=2.0):
    rng = np.random.default_rng(seed)
    q = np.zeros((n, m.nq)); q[:, 0:2] = rng.uniform(-3, 3, (n, 2)); q[:, 2] = rng.uniform(2.5, 4.0, n)
    quat = rng.normal(size=(n, 4)); q[:, 3:7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    q[:, 7:] = rng.uniform(-1.0, 1.0, (n, m.nu))
    v = rng.normal(size=(n, m.nv)) * vel
    return q, v


def relerr(a, b):
    """Norm-relative error: max|a - b| / max(1, max|b|)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))))


def elem_relerr(a, b, floor):
    """Per-element relative error max_i |a_i - b_i| / max(|b_i|, floor): every component is held to its own magnitude, components
    below `floor` to `floor` (state the floor next to the tolerance)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))
