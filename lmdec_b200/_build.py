"""Builds the in-tree CUDA library ``liblmdec_b200.so`` (sm_100a only) with nvcc.

The library is the product path: nothing in ``lmdec_b200`` works without it and there is no CPU fallback.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblmdec_b200.so")
SOURCES = [os.path.join(CSRC, "lmdec_b200.cu")]
HEADERS = [os.path.join(CSRC, "geno_mma.cuh"), os.path.join(CSRC, "dense_mma.cuh"), os.path.join(CSRC, "umma.cuh"),
           os.path.join(HERE, "..", "include", "lmdec_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: lmdec_b200 needs the CUDA toolkit to build its sm_100a library")


def source_hash() -> str:
    """sha256 (first 16 hex digits) over the CUDA sources and headers: compiled into the library
    (`lmdec_source_hash()`) so that a stale binary is detected by content -- file times do not survive the copy to
    the GPU box."""
    import hashlib
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def built_hash(path: str = LIB):
    """the source hash the library at `path` was built from (None if it carries none).  Read from the file's bytes
    (marker string compiled in next to lmdec_source_hash), NOT through dlopen: loading the stale library here would
    make the loader hand the same stale handle back after the rebuild."""
    try:
        with open(path, "rb") as f:
            blob = f.read()
    except OSError:
        return None
    marker = b"LMDEC_SRC_HASH="
    i = blob.find(marker)
    if i < 0:
        return None
    return blob[i + len(marker):i + len(marker) + 16].decode("ascii", "replace")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    return built_hash(LIB) != source_hash()


def have_nvcc() -> bool:
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    # several ranks may get here at once: build under a lock, into a temporary file, and rename
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not is_stale():
            return LIB
        tmp = LIB + f".tmp{os.getpid()}"
        cmd = [_nvcc()] + NVCC_FLAGS + [f'-DLMDEC_SRC_HASH="{source_hash()}"', "-o", tmp] + SOURCES
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        os.replace(tmp, LIB)
        if verbose:
            print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
