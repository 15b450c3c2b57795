"""In-tree nvcc build of librgp_psi.so for sm_100a.

The .so is written next to the package (``rgp_b200/_lib/``) so it travels with the
source snapshot to the GPU box; it is git-ignored.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "librgp_psi.so")
STAMP = os.path.join(LIBDIR, "librgp_psi.stamp")
MICRO = os.path.join(LIBDIR, "microbench")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; librgp_psi.so cannot be built (no CPU fallback exists)")


def _source_hash() -> str:
    h = hashlib.sha256()
    paths = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    paths.append(os.path.join(HERE, "..", "include", "rgp_psi.h"))
    for p in paths:
        if os.path.isfile(p) and not p.endswith("microbench.cu"):
            h.update(p.encode())
            with open(p, "rb") as f:
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile rgp_b200/csrc/rgp_psi.cu -> rgp_b200/_lib/librgp_psi.so (sm_100a).

    Safe under concurrent callers (the ranks of a torchrun job importing the package at once): the build
    runs under an exclusive file lock, writes to a temporary name and renames it into place, so no process
    ever dlopens a half-written library."""
    import fcntl
    os.makedirs(LIBDIR, exist_ok=True)
    if not force and is_current():
        return LIB
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():            # another process built it while we waited
                return LIB
            tmp = LIB + ".tmp.%d" % os.getpid()
            cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", tmp, os.path.join(CSRC, "rgp_psi.cu")]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
                print(" ".join(cmd), file=sys.stderr)
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            if verbose:
                print(res.stderr, file=sys.stderr)
            os.replace(tmp, LIB)
            with open(STAMP, "w") as f:
                f.write(_source_hash())
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


DEBUG_LIB = os.path.join(LIBDIR, "librgp_psi_debug.so")


PAD8_LIB = os.path.join(LIBDIR, "librgp_psi_pad8.so")


def build_variant(path: str, defines) -> str:
    """A/B build of the product sources with extra -D defines (loaded through RGP_PSI_LIB)."""
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, *["-D" + d for d in defines], "-shared", "-o", path, os.path.join(CSRC, "rgp_psi.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return path


def build_debug() -> str:
    """Experiment build (-DRGP_DEBUG): the ablation / occupancy knobs of the timing probes exist only
    here (scripts/bwd_ablate.py loads it through RGP_PSI_LIB); the product library does not have them."""
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-DRGP_DEBUG", "-shared", "-o", DEBUG_LIB, os.path.join(CSRC, "rgp_psi.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return DEBUG_LIB


def build_microbench() -> str:
    """Standalone hardware microbenchmarks: microbench (DFMA / DMMA peaks, smem broadcast costs) and mixbench
    (how DMMA and scalar FP64 instructions share the pipe).  Returns the path of the first."""
    os.makedirs(LIBDIR, exist_ok=True)
    for name in ("microbench", "mixbench"):
        src = os.path.join(CSRC, name + ".cu")
        exe = os.path.join(LIBDIR, name)
        if os.path.exists(exe) and os.path.getmtime(exe) >= os.path.getmtime(src):
            continue
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
               "-std=c++17", "-o", exe, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return MICRO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_microbench())
