"""In-tree build of libeod_memory.so (hand-written sm_100a kernels + the C ABI of include/eod_memory.h).

nvcc cross-compiles without a GPU; the resulting .so sits next to this file, is git-ignored, and travels to
the GPU box with the gpurun snapshot.  No JIT cache, no torch.utils.cpp_extension: the library has no torch
types in its ABI, so plain nvcc is enough.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libeod_memory.so")
SOURCES = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))
HEADERS = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cuh"))) + [os.path.join(HERE, "..", "include", "eod_memory.h")]
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo"]


def needs_build() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libeod_memory.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return SO_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
