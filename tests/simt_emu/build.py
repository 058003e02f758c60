#!/usr/bin/env python
"""Build tests/simt_emu/libpgpd_emu.so: the CUDA-core kernels + host orchestration of libpgpd
compiled by g++ against the SIMT emulator (cuda_emu.h).  TEST INFRASTRUCTURE ONLY -- the product
package never loads this library."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "pointnetgpd_b200", "csrc")
OUT = os.path.join(HERE, "libpgpd_emu.so")


def newest_src():
    t = 0.0
    for d in (SRC, HERE, os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            if f.endswith((".cu", ".cuh", ".h")):
                t = max(t, os.path.getmtime(os.path.join(d, f)))
    return t


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest_src():
        return OUT
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-shared", "-fPIC", "-DPGPD_EMU", "-x", "c++",
           os.path.join(SRC, "pgpd_api.cu"), "-I", HERE, "-I", SRC, "-I", os.path.join(ROOT, "include"),
           "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas", "-Wno-sign-compare", "-o", OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
