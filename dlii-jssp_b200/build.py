"""Build libjssp_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libjssp_b200.so")
SOURCES = ["jssp_b200.cu"]
HEADERS = ["jssp_kernels.cuh", "jssp_common.cuh", "jssp_seed.cuh", "jssp_pack.cuh", "jssp_eval.cuh", "jssp_eval_small.cuh", "jssp_fop.cuh",
           "jssp_aux.cuh", "jssp_device.cuh", os.path.join("..", "..", "include", "jssp_b200.h")]

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "--fmad=false",                                   # float64 parity with the reference's unfused Python arithmetic
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libjssp_b200.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("JSSP_NVCC_EXTRA", "").split()      # e.g. -DJSSP_EVAL_THREADS=128 -DJSSP_EVAL_MIN_BLOCKS=5 (tuning experiments)
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
