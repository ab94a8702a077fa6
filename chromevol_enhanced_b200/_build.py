"""In-tree build of the sm_100a shared library (explicit nvcc; no JIT cache, no torch headers).

The built ``csrc/libcevb200.so`` is git-ignored but travels with the working tree, so a GPU box
that has the tree does not need to rebuild.  ``build()`` is idempotent: it rebuilds only when the
library was built from other sources than the ones in the tree (a content hash, stored next to
the library).
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(CSRC, "libcevb200.so")
SOURCES = ["abi.cu", "prepare.cu", "expm.cu", "prune.cu", "reconstruct.cu", "fused.cu", "mapper.cu"]
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


HASH_PATH = LIB_PATH + ".srchash"


def source_hash():
    """sha256 over the sources, headers and compiler flags the library is built from."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for rel in SOURCES + HEADERS:
        path = os.path.join(CSRC, rel)
        if os.path.exists(path):
            with open(path, "rb") as fh:
                h.update(rel.encode() + b"\0" + fh.read())
    return h.hexdigest()


def needs_build():
    """True when the library is missing or was built from other sources than the ones in the
    tree.  Contents, not time stamps: the built library travels with a snapshot of the tree (it is
    git-ignored, not gpurun-ignored) and copies do not keep modification times."""
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as fh:
        return fh.read().strip() != source_hash()


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into csrc/libcevb200.so (under a file lock, into a
    temporary name that is renamed into place: several ranks may import the package at once)."""
    if not force and not needs_build():
        return LIB_PATH
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():      # another process built it meanwhile
                return LIB_PATH
            srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
            tmp = f"{LIB_PATH}.{os.getpid()}.tmp"
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + srcs
            proc = subprocess.run(cmd, capture_output=True, text=True)
            if proc.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
            os.replace(tmp, LIB_PATH)
            with open(HASH_PATH, "w") as fh:
                fh.write(source_hash())
            if verbose:
                print(proc.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH
