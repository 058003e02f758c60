"""Build pointnetgpd_b200/libpgpd.so in-tree with nvcc for sm_100a.

    python -m pointnetgpd_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
working-tree snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpgpd.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _newest_src():
    t = 0.0
    for d in (SRC, os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            if f.endswith((".cu", ".cuh", ".h")):
                t = max(t, os.path.getmtime(os.path.join(d, f)))
    return t


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: tuning builds of kernel variants (scripts/, never loaded by the package itself)."""
    if out is None and not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_src():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-I", os.path.join(ROOT, "include"), "-o", out or OUT,
                                                               os.path.join(SRC, "pgpd_api.cu"), "-lcuda"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libpgpd.so")
    if out is None:
        with open(os.path.join(HERE, "libpgpd.ptxas.log"), "w") as f:
            f.write(res.stdout + res.stderr)
    return out or OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
