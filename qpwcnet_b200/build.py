#!/usr/bin/env python3
"""Build libqpwc.so (the C-ABI CUDA library, include/qpwc.h) in-tree with nvcc for sm_100a.

    python -m qpwcnet_b200.build [--force] [--verbose]

The library is the product's only compute path; there is no JIT and no fallback.  The built
``qpwcnet_b200/lib/libqpwc.so`` is git-ignored but travels with the tree to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "libqpwc.so")
SOURCES = ["qpwc_api.cu", "qpwc_warp.cu", "qpwc_corr_direct.cu", "qpwc_corr_tiled.cu", "qpwc_upsample.cu", "qpwc_occlusion.cu", "qpwc_corr_nchw.cu", "qpwc_corr_bwd_nchw.cu", "qpwc_corr_bwd_tiled.cu", "qpwc_corr_tc.cu"]
HEADERS = ["qpwc_common.cuh", "qpwc_async.cuh", "qpwc_upsample.cuh", os.path.join("..", "..", "include", "qpwc.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=true",              # contraction is controlled explicitly (__f*_rn where TF rounds each op)
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "-Xptxas", "-v",
    "--shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build qpwcnet_b200/lib/libqpwc.so)")


def needs_build() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(os.path.normpath(d)) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    os.makedirs(LIBDIR, exist_ok=True)
    # this image exports CC/CXX=/opt/gcc/bin/*, nvcc's default host compiler (g++ on PATH) is fine
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", SO] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(LIBDIR, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError(f"nvcc failed (exit {res.returncode}); see {log}")
    if verbose:
        print(res.stdout)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
