"""Build libwaveglow_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m text2speech_b200.build [--force] [-v]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libwaveglow_b200.so")
SOURCES = ["api.cu", "common.cu", "wn_tc.cu", "wn_tc2.cu", "wn_skip16.cu", "ref_f32.cu", "flow.cu", "stft.cu", "stft_tc2.cu", "fft.cu", "wn_wgrad.cu", "train.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libwaveglow_b200.so")


def source_hash() -> str:
    """SHA-256 over csrc/* and the public header (names + contents): what the library must have been built from."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    deps.append(os.path.join(HERE, "..", "include", "waveglow_b200.h"))
    for d in deps:
        h.update(os.path.basename(d).encode() + b"\0")
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def built_hash(path: str = LIB_PATH):
    """The source hash embedded in an existing library (read from the file, without loading it), or None."""
    import re
    if not os.path.exists(path):
        return None
    with open(path, "rb") as f:
        m = re.search(rb"WGB_SOURCE_HASH=([0-9a-f]{64}|unknown)", f.read())
    return m.group(1).decode() if m else None


def needs_build() -> bool:
    """True when the library is missing or was built from other sources than the tree holds (content hash, so copies
    of the tree whose mtimes were not preserved do not trigger rebuilds)."""
    return built_hash() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:          # several ranks may find the library stale at the same time
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or needs_build():
                _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def _build_locked(verbose: bool) -> None:
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    digest = source_hash()
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if src == "api.cu":
            cmd.insert(1, f'-DWGB_SOURCE_HASH="{digest}"')
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out.strip():
            print(out)
    tmp = LIB_PATH + ".tmp"
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "--cudart", "static", "-o", tmp, *objs]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}")
    os.replace(tmp, LIB_PATH)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
