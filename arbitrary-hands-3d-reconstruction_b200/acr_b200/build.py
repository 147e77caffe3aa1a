"""In-tree build of lib/libacr_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m acr_b200.build [--force] [--verbose]

One object per .cu/.cpp under csrc/, rebuilt when the source or any header is newer, linked
into one shared library that exposes the C ABI of /include/acr_b200.h.  The library links the
CUDA runtime statically and resolves the driver API (cuTensorMapEncodeTiled) at run time via
cudaGetDriverEntryPoint, so it loads on machines without libcuda (symbol checks on CPU).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(PKG, "build")
LIB = os.path.join(LIBDIR, "libacr_b200.so")
INCLUDE = os.path.abspath(os.path.join(PKG, "..", "include"))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-I", INCLUDE, "--expt-relaxed-constexpr"]


def _newest_header() -> float:
    ts = [os.path.getmtime(os.path.join(INCLUDE, "acr_b200.h"))]
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".h", ".hpp")):
            ts.append(os.path.getmtime(os.path.join(CSRC, f)))
    return max(ts)


def _compile(src: str, obj: str, verbose: bool) -> str:
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))
    hdr = _newest_header()
    jobs, objs = [], []
    for f in srcs:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJDIR, os.path.splitext(f)[0] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr):
            jobs.append((src, obj))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for (src, _), log in zip(jobs, ex.map(lambda j: _compile(j[0], j[1], verbose), jobs)):
                if verbose:
                    print(f"== {os.path.basename(src)}\n{log}")
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
