"""In-tree build of the CUDA extension (sm_100a only) -- no JIT cache, no pip install.

``libns_coder.so`` lands next to this file so that it travels with the repo
snapshot to the GPU box.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB = os.path.join(PKG_DIR, "libns_coder.so")
HOSTLIB = os.path.join(PKG_DIR, "libns_hostmath.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # only the explicit __fma_rn calls fuse: the fp64 contract of DESIGN.md
    "-Xcompiler", "-fPIC", "-shared",
]

SOURCES = ["ns_coder.cu", "ns_codecs.cu"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into one shared library for sm_100a."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc", ".h"))]
    deps.append(os.path.join(os.path.dirname(PKG_DIR), "include", "ns_coder.h"))
    if force or _stale(LIB, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed: %s" % " ".join(cmd))
    return LIB


def build_hostmath(force: bool = False) -> str:
    """Host build of csrc/ns_math.cuh for CPU unit tests (not part of the product)."""
    src = os.path.join(CSRC, "ns_hosttest.cpp")
    deps = [src, os.path.join(CSRC, "ns_math.cuh"), os.path.join(CSRC, "ns_exp_table.inc")]
    if force or _stale(HOSTLIB, deps):
        cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", HOSTLIB, src]
        subprocess.run(cmd, check=True)
    return HOSTLIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_hostmath())
