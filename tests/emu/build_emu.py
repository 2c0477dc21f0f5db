#!/usr/bin/env python3
"""Compile the qpwc kernels for the CPU emulation harness (tests only; see README.md)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "qpwcnet_b200", "csrc")
SO = os.path.join(HERE, "_build", "libqpwc_emu.so")
SOURCES = ["qpwc_api.cu", "qpwc_warp.cu", "qpwc_corr_direct.cu", "qpwc_corr_tiled.cu", "qpwc_upsample.cu", "qpwc_occlusion.cu", "qpwc_corr_nchw.cu", "qpwc_corr_bwd_nchw.cu", "qpwc_corr_tc.cu"]


def build(force=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "qpwc.h"), __file__]
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-ffp-contract=off",
           "-DQPWC_EMU", "-include", os.path.join(HERE, "cuda_emu.h"), "-Wno-unknown-pragmas",
           "-o", SO]
    for s in SOURCES:
        cmd += ["-x", "c++", os.path.join(CSRC, s)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("emu build failed")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
