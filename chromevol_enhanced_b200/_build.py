"""In-tree build of the sm_100a shared library (explicit nvcc; no JIT cache, no torch headers).

The built ``csrc/libcevb200.so`` is git-ignored but travels with the working tree, so a GPU box
that has the tree does not need to rebuild.  ``build()`` is idempotent: it rebuilds only when a
source is newer than the library.
"""
from __future__ import annotations

import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(CSRC, "libcevb200.so")
SOURCES = ["abi.cu", "prepare.cu", "expm.cu", "prune.cu", "fused.cu", "mapper.cu"]
HEADERS = ["common.cuh", os.path.join("..", "..", "include", "chromevol_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the CUDA extension cannot be built")
    return exe


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    lib_m = os.path.getmtime(LIB_PATH)
    for rel in SOURCES + HEADERS:
        path = os.path.join(CSRC, rel)
        if os.path.exists(path) and os.path.getmtime(path) > lib_m:
            return True
    return False


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into csrc/libcevb200.so."""
    if not force and not needs_build():
        return LIB_PATH
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + srcs
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB_PATH
