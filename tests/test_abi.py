"""CPU checks of the C-ABI library: it loads, and exports every symbol include/smplsim.h declares."""
import ctypes
import os
import re

from smplsim_b200 import _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_builds_and_exports_header_symbols():
    so = _lib.build()
    L = ctypes.CDLL(so)
    hdr = open(os.path.join(ROOT, "include", "smplsim.h")).read()
    declared = set(re.findall(r"\b(smplsim_[a-z_]+)\s*\(", hdr))
    assert declared == set(_lib.HEADER_SYMBOLS), declared ^ set(_lib.HEADER_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    L.smplsim_version.restype = ctypes.c_int
    assert L.smplsim_version() >= 100


def test_ctypes_struct_sizes_match_header():
    """sizeof of the ctypes mirrors == sizeof in C (compiled with gcc from the header)."""
    import subprocess
    import tempfile
    from smplsim_b200.abi import SmplsimAuxC, SmplsimEnvCfgC, SmplsimStateC
    from smplsim_b200.model import SmplsimModelDescC
    src = '#include <stdio.h>\n#include "smplsim.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(SmplsimModelDesc), sizeof(SmplsimEnvCfg), sizeof(SmplsimState), sizeof(SmplsimAux));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        sizes = list(map(int, subprocess.check_output([os.path.join(d, "t")]).split()))
    assert sizes == [ctypes.sizeof(SmplsimModelDescC), ctypes.sizeof(SmplsimEnvCfgC), ctypes.sizeof(SmplsimStateC), ctypes.sizeof(SmplsimAuxC)]


def test_missing_gpu_fails_loudly():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from smplsim_b200.batched import HumanoidBatchB200
    from smplsim_b200.cfg import make_cfg
    with pytest.raises(RuntimeError):
        HumanoidBatchB200(make_cfg(env="speed"), num_envs=2)
