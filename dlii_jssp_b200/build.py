"""Build libjssp_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery).

The library carries a BUILD ID = hash of every source it was compiled from plus the compiler flags (``jssp_build_id()``,
also findable in the file as the string ``JSSP_BUILD_ID=<id>``).  ``needs_build()`` compares it with the hash of the sources
on disk, so a stale library is never loaded silently and the check does not depend on file times (which a copy to another
box does not preserve)."""
from __future__ import annotations

import hashlib
import os
import re
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(HERE, "..", "include")
LIB = os.path.join(HERE, "libjssp_b200.so")
SOURCES = ["jssp_b200.cu"]

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "--fmad=false",                                   # float64 parity with the reference's unfused Python arithmetic; the hot
                                                      # loops of the evaluation kernels ask for fused multiply-adds explicitly
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off",
    "-ldl",
]
_MARK = b"JSSP_BUILD_ID="


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libjssp_b200.so cannot be built (there is no CPU fallback)")


def _extra_flags() -> list:
    return os.environ.get("JSSP_NVCC_EXTRA", "").split()      # e.g. -DJSSP_EVAL_THREADS=128 (tuning experiments)


def dependencies() -> list:
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    return deps + sorted(os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE) if f.endswith(".h"))


def source_id() -> str:
    """16 hex digits over the contents of csrc/*, include/*.h and the compiler flags."""
    h = hashlib.sha256()
    for d in dependencies():
        h.update(os.path.basename(d).encode() + b"\0")
        h.update(open(d, "rb").read())
    h.update(" ".join(NVCC_FLAGS + _extra_flags()).encode())
    return h.hexdigest()[:16]


def library_id(path: str = LIB):
    """Build id embedded in the library file, None when the file is missing or carries none."""
    try:
        blob = open(path, "rb").read()
    except OSError:
        return None
    m = re.search(_MARK + rb"([0-9a-f]{16})", blob)
    return m.group(1).decode() if m else None


def needs_build() -> bool:
    return library_id() != source_id()


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    sid = source_id()
    tmp = LIB + ".tmp"
    cmd = [_nvcc(), *NVCC_FLAGS, *_extra_flags(), f'-DJSSP_BUILD_ID_STR="{sid}"', "-o", tmp, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    os.replace(tmp, LIB)                                       # never leaves a half-written library behind
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
