"""Build `libhydpy_b200.so` in-tree with nvcc for sm_100a (`python -m hydpy_b200.build`)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhydpy_b200.so")
SOURCES = ("engine.cu",)
DEPENDS = ("engine.cu", "land_kernel.cuh", "land_kernel_v2.cuh", "land_kernel_coupled.cuh", "fastmath.cuh",
           "river_kernel.cuh", os.path.join("..", "..", "include", "hydpy_b200.h"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # IEEE evaluation like the reference's gcc -O2 x86-64 build: no implicit FMA contraction; explicit fma() in the
    # engine's own math routines is unaffected
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libhydpy_b200.so cannot be built")
    return exe


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPENDS)


def build(force: bool = False, verbose: bool = False, variant: str = "", defines: tuple[str, ...] = ()) -> str:
    """`variant` / `defines`: a development build `libhydpy_b200_<variant>.so` with extra -D flags beside the product
    library (select it with HYDPY_B200_LIB=… for A/B measurements on the GPU box)."""
    out = LIB if not variant else os.path.join(HERE, f"libhydpy_b200_{variant}.so")
    if not variant and not force and not is_stale():
        return LIB
    cmd = [nvcc(), *NVCC_FLAGS, *defines, "-o", out, *[os.path.join(CSRC, s) for s in SOURCES], "-ldl"]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
        print(" ".join(cmd))
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(proc.stdout + proc.stderr)
    return out


if __name__ == "__main__":
    name = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    build(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, variant=name,
          defines=tuple(a for a in sys.argv[1:] if a.startswith("-D")))
