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
NVCC_FLAGS = ["-Xcompiler", "-fPIC", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"]
OBJ_DIR = os.path.join(HERE, "_obj")


def needs_build() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def _compile(nvcc: str, src: str, obj: str, verbose: bool) -> str:
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed on {os.path.basename(src)}")
    return res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """One object per .cu (compiled in parallel, only when the source or a header is newer), then one link."""
    if not force and not needs_build():
        return SO_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    t_hdr = max(os.path.getmtime(h) for h in HEADERS)
    jobs = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), t_hdr):
            jobs.append((src, obj))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        for log in pool.map(lambda j: _compile(nvcc, j[0], j[1], verbose), jobs):
            if verbose:
                sys.stderr.write(log)
    objs = [os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o") for src in SOURCES]
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO_PATH] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libeod_memory.so")
    return SO_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
