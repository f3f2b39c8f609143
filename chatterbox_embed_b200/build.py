"""Build libcbx.so in-tree with nvcc for sm_100a (no torch, no pybind: a plain C-ABI shared library).

    python -m chatterbox_embed_b200.build [--force] [-v]

Safe under torchrun (one process per GPU all importing the package at once): the build runs under an exclusive file lock,
objects and the library are written under temporary names and moved into place with an atomic rename, and staleness is
decided by a content hash of the sources and flags (mtimes do not survive rsync / checkout), so a rank never dlopens a
half-written library and only one rank compiles.

CBX_DEV_TOOLS=1 builds libcbx_dev.so instead: the same library plus the kernel unit-test / timing entry points the scripts
under tools/ use (cbx_test_tgemm, cbx_test_shift_gemm, cbx_lstm_max_clusters, the v1 recurrence kernel, option "probe").
None of those are in the product library.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
DEV = os.environ.get("CBX_DEV_TOOLS") == "1"
LIB = os.path.join(HERE, "libcbx_dev.so" if DEV else "libcbx.so")
SOURCES = ["host_plan.cpp", "weights.cu", "ve.cu", "frontend_tc.cu", "lstm_tc.cu", "fcm_tc.cu", "fcm_block_tc.cu", "local_tc.cu", "xv.cu", "resample.cu", "promptmel_tc.cu", "project.cu", "tc.cu", "api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
] + (["-DCBX_DEV_TOOLS"] if DEV else [])


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libcbx.so cannot be built (there is no CPU fallback)")
    return exe


def _objdir() -> str:
    return os.path.join(HERE, "build_dev" if DEV else "build")


def source_hash() -> str:
    """Content hash of everything the library is made from (sources, headers, the public header, the flags)."""
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.startswith("."))
    files.append(os.path.join(HERE, "..", "include", "cbx.h"))
    for p in files:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + SOURCES).encode())
    return h.hexdigest()


def _hash_file() -> str:
    return LIB + ".hash"


def _stale(want: str) -> bool:
    if not os.path.exists(LIB) or not os.path.exists(_hash_file()):
        return True
    try:
        with open(_hash_file()) as f:
            return f.read().strip() != want
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    want = source_hash()
    if not force and not _stale(want):
        return LIB
    objdir = _objdir()
    os.makedirs(objdir, exist_ok=True)
    with open(os.path.join(objdir, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)              # the other ranks of a torchrun launch wait here ...
        try:
            if not force and not _stale(want):         # ... and find the library the first one built
                return LIB
            nvcc = _nvcc()
            objs, procs = [], []
            for src in SOURCES:
                obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
                objs.append(obj)
                cmd = [nvcc, *NVCC_FLAGS, "-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj + ".tmp"]
                if verbose:
                    cmd.insert(1, "-Xptxas=-v")
                procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            failed = False
            for src, obj, p in procs:
                out, _ = p.communicate()
                if p.returncode != 0 or verbose:
                    sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
                failed |= p.returncode != 0
                if p.returncode == 0:
                    os.replace(obj + ".tmp", obj)
            if failed:
                raise RuntimeError("nvcc failed")
            tmp = LIB + f".tmp{os.getpid()}"
            subprocess.check_call([nvcc, "-shared", "-o", tmp, *objs, "-lcudart"])
            os.replace(tmp, LIB)                       # atomic: a concurrent dlopen sees the old or the new file, never half of one
            with open(_hash_file() + ".tmp", "w") as f:
                f.write(want + "\n")
            os.replace(_hash_file() + ".tmp", _hash_file())
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
