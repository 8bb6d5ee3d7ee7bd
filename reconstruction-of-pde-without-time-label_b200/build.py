"""Build libblindno_b200.so (hand-written sm_100a CUDA + the C ABI of include/blindno_b200.h).

nvcc cross-compiles without a GPU; the .so is written in-tree (``lib/``) so that it travels
to the GPU box with the repo snapshot.  ``python -m blindno_b200.build`` or
``__graft_entry__.build()``.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBPATH = os.path.join(LIBDIR, "libblindno_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
SOURCES = ["spectral.cu", "pointwise.cu", "fused1d.cu", "tc_gemm.cu", "tc_layer.cu", "bagattn.cu", "api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
] + os.environ.get("BDN_NVCC_EXTRA", "").split()      # (experiments, e.g. -DBDN_PDL_LATE=1; part of the build fingerprint)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libblindno_b200.so cannot be built")


def _fingerprint() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join(INCLUDE, "blindno_b200.h")]
    for f in files:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        with open(path, "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if the sources changed since the last build.  Returns the .so path.

    Safe under concurrent callers (one process per GPU under torchrun all import the package at once): the check and
    the compile run under an exclusive file lock, nvcc writes to a private temporary file, and the library is put in
    place with an atomic rename -- a process never sees a half-written .so."""
    import fcntl
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.stamp")
    fp = _fingerprint()

    def current() -> bool:
        return os.path.exists(LIBPATH) and os.path.exists(stamp) and open(stamp).read().strip() == fp

    if not force and current():
        return LIBPATH
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and current():          # another process built it while this one waited for the lock
                return LIBPATH
            srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
            tmp = f"{LIBPATH}.tmp.{os.getpid()}"
            cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-o", tmp, *srcs]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), file=sys.stderr)
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            if verbose:
                print(res.stderr, file=sys.stderr)
            os.replace(tmp, LIBPATH)
            with open(stamp + ".tmp", "w") as fh:
                fh.write(fp)
            os.replace(stamp + ".tmp", stamp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIBPATH


def ensure_current() -> str:
    """What every load of the library goes through: the in-tree .so must have been built from the sources that
    are in the tree now.  With nvcc present a missing or stale library is rebuilt (the stamp comparison is a hash
    of csrc/, the header and the flags); without nvcc a stale library is an error, never silently loaded."""
    stamp = os.path.join(LIBDIR, "build.stamp")
    fp = _fingerprint()
    if os.path.exists(LIBPATH) and os.path.exists(stamp) and open(stamp).read().strip() == fp:
        return LIBPATH
    try:
        _nvcc()
    except RuntimeError:
        raise RuntimeError(f"{LIBPATH} is missing or was built from other sources than the tree holds "
                           f"(fingerprint {fp[:12]}), and nvcc is not available to rebuild it")
    return build()


def fingerprint() -> str:
    """Fingerprint of the sources the loaded library was built from (bench.py prints it)."""
    return _fingerprint()


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
