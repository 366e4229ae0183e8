"""ctypes wrapper around oracle/liboracle.so (CPU fp64 ORACLE -- TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  See mjstep_oracle.c for what is restated and the parity status
("parity unpinned" for mj_step; env-level arithmetic pinned by tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from smplsim_b200.abi import SmplsimEnvCfgC, env_cfg_from, model_from_cfg
from smplsim_b200.model import ModelDesc, mass_matrix_numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "mjstep_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-O3", "-march=x86-64-v2", "-fPIC", "-shared", "-o", so, src, "-lm", "-lpthread"],
                              cwd=_HERE)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        L.orc_model_create.restype = C.c_void_p
        L.orc_model_create.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_model_free.argtypes = [C.c_void_p]
        L.orc_set_meaninertia.argtypes = [C.c_void_p, C.c_double]
        L.orc_obs_dim.argtypes = [C.c_void_p]
        L.orc_data_create.restype = C.c_void_p
        L.orc_data_create.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_data_free.argtypes = [C.c_void_p]
        for f in ("orc_kinematics", "orc_forward", "orc_step"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_void_p]
        L.orc_compute_torque.argtypes = [C.c_void_p, C.c_void_p, dp, dp]
        L.orc_self_obs.argtypes = [C.c_void_p, C.c_int, dp, dp, dp, dp, dp, dp]
        L.orc_compute_observations.argtypes = [C.c_void_p, C.c_void_p, dp]
        L.orc_env_step.argtypes = [C.c_void_p, C.c_void_p, dp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_env_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, dp, dp, dp]
        for f in ("qpos", "qvel", "ctrl", "qacc", "qacc_warm", "qacc_smooth", "qfrc_bias", "qfrc_constraint", "M", "xpos",
                  "xquat", "sensor", "target", "last_torque", "pid_integral", "pid_last_error", "con_dist", "con_pos", "con_frame", "efc_force",
                  "efc_aref", "efc_D", "efc_J", "sub_com"):
            fn = getattr(L, "orc_" + f)
            fn.restype = dp
            fn.argtypes = [C.c_void_p]
        L.orc_con_geom.restype = C.POINTER(C.c_int)
        L.orc_con_geom.argtypes = [C.c_void_p]
        for f in ("orc_ncon", "orc_nefc", "orc_solver_iter"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_contact_mask.restype = C.c_uint64
        L.orc_contact_mask.argtypes = [C.c_void_p]
        L.orc_get_int.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_int.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_philox.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.orc_bench_env_steps.restype = C.c_double
        L.orc_bench_env_steps.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, dp, C.c_int, C.c_int]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleModel:
    def __init__(self, model: ModelDesc, envcfg: SmplsimEnvCfgC):
        self.model = model
        self.envcfg = envcfg
        self._cs = model.c_struct()
        self.ptr = lib().orc_model_create(C.addressof(self._cs), C.addressof(envcfg))
        M0, _ = mass_matrix_numpy(model, model.qpos0)
        lib().orc_set_meaninertia(self.ptr, float(np.mean(np.diag(M0))))
        self.obs_dim = lib().orc_obs_dim(self.ptr)

    @classmethod
    def from_cfg(cls, cfg, seed: int = 0):
        m = model_from_cfg(cfg)
        om = cls(m, env_cfg_from(cfg, m, seed=seed))
        e = cfg.env
        if bool(e.get("self_collision", False) if hasattr(e, "get") else getattr(e, "self_collision", False)):
            om.set_self_collision(True)
        return om

    def set_self_collision(self, enable: bool = True):
        """Geom-geom contacts between capsule / sphere geom pairs with MuJoCo's filters (same body, parent-child,
        contype / conaffinity, <contact><exclude>); what cfg.env.self_collision switches on in the product."""
        m = self.model
        ex = np.array([[m.body_names.index(a), m.body_names.index(b)] for a, b in m.excludes], dtype=np.int32).reshape(-1, 2)
        ct = np.ascontiguousarray(m.geom_contype, dtype=np.int32); ca = np.ascontiguousarray(m.geom_conaffinity, dtype=np.int32)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
        L = lib()
        L.orc_set_self_collision.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]
        L.orc_set_self_collision(self.ptr, int(enable), ip(ct), ip(ca), ip(np.ascontiguousarray(ex)), int(ex.shape[0]))
        return self

    def __del__(self):
        try:
            lib().orc_model_free(self.ptr)
        except Exception:
            pass


class OracleEnv:
    """One fp64 environment (mj_data + task state)."""

    def __init__(self, om: OracleModel, env_id: int = 0):
        self.om = om
        self.m = om.model
        self.ptr = lib().orc_data_create(om.ptr, env_id)
        L = lib()
        nb, nv, nq, nu = self.m.nbody, self.m.nv, self.m.nq, self.m.nu

        def view(name, shape):
            p = getattr(L, "orc_" + name)(self.ptr)
            return np.ctypeslib.as_array(p, shape=shape)

        self.qpos, self.qvel, self.ctrl = view("qpos", (nq,)), view("qvel", (nv,)), view("ctrl", (nu,))
        self.qacc, self.qacc_warm, self.qacc_smooth = view("qacc", (nv,)), view("qacc_warm", (nv,)), view("qacc_smooth", (nv,))
        self.qfrc_bias, self.qfrc_constraint = view("qfrc_bias", (nv,)), view("qfrc_constraint", (nv,))
        self.M = view("M", (nv, nv))
        self.xpos, self.xquat = view("xpos", (nb, 3)), view("xquat", (nb, 4))
        self.sensor = view("sensor", (2, nb, 3))
        self.target = view("target", (4,))
        self.last_torque = view("last_torque", (nu,))
        self.pid_integral, self.pid_last_error = view("pid_integral", (nu,)), view("pid_last_error", (nu,))
        self.sub_com = view("sub_com", (3,))
        self._con_dist = view("con_dist", (256,))
        self._con_pos = view("con_pos", (256, 3))
        self._con_frame = view("con_frame", (256, 3, 3))
        self._con_geom = np.ctypeslib.as_array(L.orc_con_geom(self.ptr), shape=(256,))

    def __del__(self):
        try:
            lib().orc_data_free(self.ptr)
        except Exception:
            pass

    # -- mujoco-like calls
    def kinematics(self):
        lib().orc_kinematics(self.om.ptr, self.ptr)

    def forward(self):
        lib().orc_forward(self.om.ptr, self.ptr)

    def mj_step(self):
        lib().orc_step(self.om.ptr, self.ptr)

    def compute_torque(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        t = np.zeros(self.m.nu)
        lib().orc_compute_torque(self.om.ptr, self.ptr, _dp(a), _dp(t))
        return t

    @property
    def ncon(self):
        return lib().orc_ncon(self.ptr)

    @property
    def nefc(self):
        return lib().orc_nefc(self.ptr)

    @property
    def solver_iter(self):
        return lib().orc_solver_iter(self.ptr)

    @property
    def contact_mask(self):
        return int(lib().orc_contact_mask(self.ptr))

    def contacts(self):
        n = self.ncon
        L = lib()
        L.orc_con_geom1.restype = C.POINTER(C.c_int)
        L.orc_con_geom1.argtypes = [C.c_void_p]
        g1 = np.ctypeslib.as_array(L.orc_con_geom1(self.ptr), shape=(256,))[:n].copy() + 1     # 0 = floor, else MuJoCo geom id
        return dict(geom1=g1, geom=self._con_geom[:n].copy() + 1, dist=self._con_dist[:n].copy(), pos=self._con_pos[:n].copy(),
                    frame=self._con_frame[:n].copy())

    def efc(self):
        n, nv = self.nefc, self.m.nv
        L = lib()
        g = lambda f, shape: np.ctypeslib.as_array(getattr(L, "orc_" + f)(self.ptr), shape=shape).copy()  # noqa: E731
        return dict(J=g("efc_J", (n, nv)), aref=g("efc_aref", (n,)), D=g("efc_D", (n,)), force=g("efc_force", (n,)))

    # -- env-level
    cur_t = property(lambda s: lib().orc_get_int(s.ptr, 0), lambda s, v: lib().orc_set_int(s.ptr, 0, int(v)))
    change_step = property(lambda s: lib().orc_get_int(s.ptr, 1), lambda s, v: lib().orc_set_int(s.ptr, 1, int(v)))
    recovery = property(lambda s: lib().orc_get_int(s.ptr, 2), lambda s, v: lib().orc_set_int(s.ptr, 2, int(v)))
    warn = property(lambda s: lib().orc_get_int(s.ptr, 4), lambda s, v: lib().orc_set_int(s.ptr, 4, int(v)))   # sticky mj_warning bits (1 qpos, 2 qvel, 4 qacc)
    rng_counter = property(lambda s: lib().orc_get_int(s.ptr, 3) & 0xFFFFFFFF, lambda s, v: lib().orc_set_int(s.ptr, 3, int(v)))

    def observations(self):
        o = np.zeros(self.om.obs_dim + 8)
        n = lib().orc_compute_observations(self.om.ptr, self.ptr, _dp(o))
        return o[:n]

    def reset(self, init_mode: int = -1, qpos0=None, qvel0=None):
        o = np.zeros(self.om.obs_dim + 8)
        q0 = np.ascontiguousarray(qpos0 if qpos0 is not None else np.zeros(self.m.nq), dtype=np.float64)
        v0 = np.ascontiguousarray(qvel0 if qvel0 is not None else np.zeros(self.m.nv), dtype=np.float64)
        n = lib().orc_env_reset(self.om.ptr, self.ptr, init_mode, _dp(q0), _dp(v0), _dp(o))
        return o[:n]

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        o = np.zeros(self.om.obs_dim + 8)
        rew = C.c_double()
        term, trunc = C.c_int(), C.c_int()
        n = lib().orc_env_step(self.om.ptr, self.ptr, _dp(a), _dp(o), C.byref(rew), C.byref(term), C.byref(trunc))
        return o[:n], rew.value, bool(term.value), bool(trunc.value)


def self_obs(om: OracleModel, version, qvel, xpos, xquat, linvel=None, angvel=None):
    nb = om.model.nbody
    z = np.zeros((nb, 3))
    o = np.zeros(16 * nb + om.model.nv + 16)
    args = [np.ascontiguousarray(x if x is not None else z, dtype=np.float64) for x in (qvel, xpos, xquat, linvel, angvel)]
    n = lib().orc_self_obs(om.ptr, version, *[_dp(a) for a in args], _dp(o))
    return o[:n]


def philox(counter: int, env_id: int, seed: int):
    out = (C.c_uint32 * 4)()
    lib().orc_philox(counter & 0xFFFFFFFF, env_id & 0xFFFFFFFF, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF, out)
    return list(out)


def bench_env_steps(om: OracleModel, envs, actions: np.ndarray, autoreset: bool = True, nthreads: int = 1) -> float:
    """actions: [nsteps, nenv, nu] float64.  Steps every env nsteps times on nthreads host threads."""
    nsteps, nenv, _ = actions.shape
    arr = (C.c_void_p * nenv)(*[e.ptr for e in envs])
    a = np.ascontiguousarray(actions, dtype=np.float64)
    return lib().orc_bench_env_steps(om.ptr, arr, nenv, nsteps, _dp(a), int(autoreset), nthreads)
