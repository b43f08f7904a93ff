"""Builds compute_b200/lib/libcompute_b200.so (the C-ABI library of include/compute_b200.h) with nvcc for sm_100a.

In-tree build: the .so is git-ignored but travels to the GPU box with the gpurun snapshot.
Usage: python -m compute_b200.build [--force]
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libcompute_b200.so")
SOURCES = ["runtime.cu", "reduce.cu", "scan.cu", "radix_sort.cu", "radix_pass_ws.cu", "radix_exchange_ws.cu", "radix_exchange.cu", "radix_field.cu", "stream_ops.cu", "set_ops.cu"]
HEADERS = ["common.cuh", "ops.cuh", "radix_common.cuh", "tma.cuh", "tile_state.cuh", "scan_ws.cuh", os.path.join("..", "..", "include", "compute_b200.h")]
NVCC = os.environ.get("NVCC", "nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
] + os.environ.get("BCB_EXTRA_NVCC_FLAGS", "").split()


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _compile(src: str) -> str:
    obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS]
    if _mtime(obj) >= max(_mtime(d) for d in deps):
        return obj
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(_compile, SOURCES))
    if force or _mtime(LIB) < max(_mtime(o) for o in objs):
        tmp = LIB + ".tmp"  # link next to the target, then rename: a reader never sees a half-written library
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
