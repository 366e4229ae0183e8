"""TEST INFRASTRUCTURE: the product's kernel sources run on the host SIMT emulator (tests/emu/shim/cuda_runtime.h).

`EmuBatch` mirrors smplsim_b200.batched.HumanoidBatchB200 on numpy arrays and calls the SAME C ABI, exported by
tests/emu/libsmplsim_emu.so = smplsim_b200/csrc/smplsim_capi.cu compiled with g++ against the emulator shim.  It exists so
that `pytest -m "not gpu"` exercises the kernel logic against the oracle where no GPU is present; it is never imported by
smplsim_b200/ (the product has no CPU path)."""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

import numpy as np

from smplsim_b200 import _lib
from smplsim_b200.abi import SmplsimAuxC, SmplsimStateC, env_cfg_from, model_from_cfg

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.abspath(os.path.join(_HERE, "..", ".."))
SO = os.environ.get("SMPLSIM_EMU_SO") or os.path.join(_HERE, "libsmplsim_emu.so")
_L = None


def build(force: bool = False) -> str:
    srcs = glob.glob(os.path.join(_ROOT, "smplsim_b200", "csrc", "*")) + glob.glob(os.path.join(_HERE, "*.cpp")) + glob.glob(os.path.join(_HERE, "shim", "*"))
    newest = max(os.path.getmtime(s) for s in srcs)
    if "SMPLSIM_EMU_SO" in os.environ:
        return SO
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        # nvcc contracts a * b + c into FFMA; do the same where the host has FMA so that the rounding pattern is comparable
        fma = ["-mfma", "-ffp-contract=fast"] if "fma" in open("/proc/cpuinfo").read() else ["-ffp-contract=off"]
        cmd = ["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", *fma, "-Wno-unknown-pragmas", "-I", os.path.join(_HERE, "shim"),
               "-x", "c++", os.path.join(_ROOT, "smplsim_b200", "csrc", "smplsim_capi.cu"), "-x", "c++", os.path.join(_HERE, "emu.cpp"), "-o", SO]
        subprocess.check_call(cmd)
    return SO


def lib():
    global _L
    if _L is None:
        build()
        L = C.CDLL(SO)
        for name, (res, args) in _lib._SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _L = L
    return _L


class EmuBatch:
    """HumanoidBatchB200 with its library binding pointed at the emulator build (torch CPU tensors as 'device' memory)."""

    def __new__(cls, cfg, num_envs=None, seed=None, with_aux=True, **kw):
        import torch
        from smplsim_b200.batched import HumanoidBatchB200

        class _Emu(HumanoidBatchB200):
            def _require_device(self, device):
                pass

            def _L(self):
                return lib()

            def _device_index(self):
                return 0

            def _stream(self):
                return None

        return _Emu(cfg, num_envs=num_envs, device="cpu", seed=seed, with_aux=with_aux, **kw)
