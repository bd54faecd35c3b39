"""Build libgplvm_b200.so (the C-ABI library of include/gplvm_b200.h) in-tree with nvcc for sm_100a.

    python -m gplvmhaskell_b200.build [--force]

The shared object is written next to this file so that it travels to the GPU box with the repo
snapshot; it is git-ignored.  No JIT cache, no torch.utils.cpp_extension: plain nvcc.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libgplvm_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-diag-suppress", "177"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(paths):
    h = hashlib.sha256(" ".join(FLAGS).encode())
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "gplvm_b200.h"))
    for p in sorted(paths + hdrs):
        h.update(os.path.basename(p).encode())   # names only: the stamp must survive a move of the checkout (gpurun box)
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


class _BuildLock:
    """Inter-process lock: torchrun starts one process per GPU and every rank may call build_library()."""

    def __enter__(self):
        import fcntl
        os.makedirs(OBJ, exist_ok=True)
        self.f = open(os.path.join(OBJ, ".lock"), "w")
        fcntl.flock(self.f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *a):
        import fcntl
        fcntl.flock(self.f, fcntl.LOCK_UN)
        self.f.close()


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources()
    stamp = _stamp(srcs)
    stamp_file = os.path.join(OBJ, "stamp")

    def fresh():
        return os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp

    if not force and fresh():
        return LIB
    with _BuildLock():
        if not force and fresh():      # another process built it while we waited
            return LIB
        return _build_locked(srcs, stamp, stamp_file, verbose)


def _build_locked(srcs, stamp, stamp_file, verbose):
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):
            return LIB  # GPU box without a toolkit: use the prebuilt library that travelled with the snapshot
        raise RuntimeError("nvcc not found and no prebuilt libgplvm_b200.so")
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = [NVCC, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, LIB)               # atomic: a concurrent dlopen never sees a half-written file
    with open(stamp_file + ".tmp", "w") as f:
        f.write(stamp)
    os.replace(stamp_file + ".tmp", stamp_file)
    return LIB


CLI_SRC = os.path.join(HERE, "..", "cli", "gplvmhaskell_exe.cpp")
CLI = os.path.join(HERE, "GPLVMHaskell-exe")


def build_cli(force: bool = False) -> str:
    """g++ the GPLVMHaskell-exe twin (cli/gplvmhaskell_exe.cpp) against the in-tree library (rpath $ORIGIN)."""
    lib = build_library()
    if (not force and os.path.exists(CLI) and os.path.getmtime(CLI) >= os.path.getmtime(CLI_SRC)
            and os.path.getmtime(CLI) >= os.path.getmtime(lib)):
        return CLI
    with _BuildLock():
        tmp = CLI + ".tmp.%d" % os.getpid()
        cmd = ["g++", "-O2", "-std=c++17", CLI_SRC, "-o", tmp, "-L" + HERE, "-lgplvm_b200", "-Wl,-rpath,$ORIGIN",
               "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed for the CLI twin:\n%s\n%s" % (r.stdout, r.stderr))
        os.replace(tmp, CLI)
    return CLI


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_cli(force="--force" in sys.argv))
