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
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources()
    stamp = _stamp(srcs)
    stamp_file = os.path.join(OBJ, "stamp")
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
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
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ldl", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


CLI_SRC = os.path.join(HERE, "..", "cli", "gplvmhaskell_exe.cpp")
CLI = os.path.join(HERE, "GPLVMHaskell-exe")


def build_cli(force: bool = False) -> str:
    """g++ the GPLVMHaskell-exe twin (cli/gplvmhaskell_exe.cpp) against the in-tree library (rpath $ORIGIN)."""
    lib = build_library()
    if (not force and os.path.exists(CLI) and os.path.getmtime(CLI) >= os.path.getmtime(CLI_SRC)
            and os.path.getmtime(CLI) >= os.path.getmtime(lib)):
        return CLI
    cmd = ["g++", "-O2", "-std=c++17", CLI_SRC, "-o", CLI, "-L" + HERE, "-lgplvm_b200", "-Wl,-rpath,$ORIGIN",
           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed for the CLI twin:\n%s\n%s" % (r.stdout, r.stderr))
    return CLI


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_cli(force="--force" in sys.argv))
