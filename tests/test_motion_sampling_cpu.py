"""f3: the sampling half of the motion library -- PMCP weights (hard / soft / restored termination history),
sample_motions / sample_time(_interval), motion_aa / motion_bodies in the returned state -- against the reference's own methods
(smpl_sim/smpllib/motion_lib_base.py:225-312) run unbound on a namespace (tests/golden/make_golden_motion.py writes
motion_sampling.npz).  The library object is built on a stub env (sampling is plain torch; the gather kernel is covered by the GPU
tests)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from conftest import GOLDEN  # noqa: E402


def _lib(K=6, F=5):
    from smplsim_b200.motion_lib import MotionLibB200, TABLE_KEYS
    env = SimpleNamespace(device=torch.device("cpu"), seed=3, model=SimpleNamespace(nbody=24), _h=None, gpu_launches=0)
    t = {k: np.zeros((K * F, 4), np.float32) for k in TABLE_KEYS}
    t.update(motion_num_frames=np.full(K, F, np.int32), motion_dt=np.full(K, 1 / 30, np.float32),
             motion_lengths=np.linspace(1.0, 2.0, K).astype(np.float32), length_starts=(np.arange(K) * F).astype(np.int32),
             motion_keys=np.array([f"k{i}" for i in range(K)]))
    return MotionLibB200(env, t)


def test_pmcp_weights_match_reference_methods():
    g = np.load(os.path.join(GOLDEN, "motion_sampling.npz"))
    lib = _lib()
    lib.update_hard_sampling_weight(["k1", "k4"])
    assert np.allclose(lib._sampling_prob.numpy(), g["hard"])
    lib.update_hard_sampling_weight([])
    assert np.allclose(lib._sampling_prob.numpy(), g["hard_empty"])
    lib.update_soft_sampling_weight(["k0", "k2"])
    lib.update_soft_sampling_weight(["k2", "k5"])
    assert np.allclose(lib._sampling_prob.numpy(), g["soft2"]) and np.allclose(lib._termination_history.numpy(), g["soft2_hist"])
    lib.set_termination_history(dict(termination_history=g["restore_hist"], failed_keys=["k3"]))
    assert np.allclose(lib._sampling_prob.numpy(), g["restore"]) and lib.curr_failed_keys == ["k3"]
    assert lib.update_sampling_prob(np.ones(3)) is False


def test_sampling_follows_the_weights_and_time_rules():
    lib = _lib()
    lib.update_hard_sampling_weight(["k1", "k4"])
    ids = lib.sample_motions(4000)
    cnt = np.bincount(ids.numpy(), minlength=6)
    assert cnt[[0, 2, 3, 5]].sum() == 0 and abs(cnt[1] - 2000) < 200
    t = lib.sample_time(ids, truncate_time=0.5)
    ml = lib.get_motion_length(ids).numpy() - 0.5
    assert (t.numpy() >= 0).all() and (t.numpy() <= ml + 1e-6).all()
    ti = lib.sample_time_interval(ids)
    assert np.allclose(ti.numpy() * 30, np.round(ti.numpy() * 30), atol=1e-4)
    assert (lib.get_motion_num_steps().numpy() == 5).all()
