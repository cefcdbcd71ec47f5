"""Build libb200mpc.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m judo_b200.build          # rebuild if sources changed
    python -m judo_b200.build --force
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libb200mpc.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.cpp")) +
                  [os.path.join(_HERE, "..", "include", "b200mpc.h")])


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    defs = []
    if os.path.exists(os.path.join(CSRC, "leap.cuh")):
        defs.append("-DB200MPC_WITH_LEAP")
    # host-only glue (sampler / spline basis / trace segments): plain g++, strict IEEE rounding (no FMA contraction: the candidates must
    # be bit-identical to NumPy's), SIMD variants selected at load time (the library is built here and runs on another CPU)
    glue_o = os.path.join(_HERE, "host_glue.o")
    gcmd = [os.environ.get("CXX", "g++"), "-O3", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-math-errno", "-c",
            os.path.join(CSRC, "host_glue.cpp"), "-o", glue_o]
    gres = subprocess.run(gcmd, capture_output=True, text=True)
    if gres.returncode != 0:
        raise RuntimeError("g++ failed:\n" + (gres.stdout + gres.stderr)[-6000:])
    cmd = [NVCC, *FLAGS, *defs, "-o", LIB_PATH, os.path.join(CSRC, "b200mpc.cu"), glue_o]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = " ".join(gcmd) + "\n" + gres.stdout + gres.stderr + res.stdout + res.stderr
    with open(os.path.join(_HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-6000:])
    if verbose:
        print(log)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
