"""TEST INFRASTRUCTURE (not product): builds the CPU SIMT emulator harness around judo_b200/csrc/*.cuh with g++ and loads it.

The shipped library (judo_b200/libb200mpc.so) is built by nvcc and never references this package; tests use it to run the
kernels' device code on the CPU against the oracle in the no-GPU tier."""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "..", "..", "judo_b200", "csrc")
_OUT = os.path.join(_HERE, "_build", "libwarpsim_kernels.so")


def build() -> str:
    srcs = [os.path.join(_HERE, "sim_kernels.cpp"), os.path.join(_HERE, "warpsim.cpp")]
    deps = srcs + [os.path.join(_HERE, "warpsim.h")] + glob.glob(os.path.join(_CSRC, "*.cuh"))
    if not os.path.exists(_OUT) or any(os.path.getmtime(d) > os.path.getmtime(_OUT) for d in deps):
        os.makedirs(os.path.dirname(_OUT), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-DB2_HOST_SIM", "-ffp-contract=off", "-I", _HERE, "-I", _CSRC,
                               "-fPIC", "-shared", "-o", _OUT, *srcs])
    return _OUT


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib
