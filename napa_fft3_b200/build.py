"""Builds napa_fft3_b200/libnapafft_b200.so in-tree with nvcc for sm_100a (cross-compiles without a
GPU).  `python -m napa_fft3_b200.build [--force]`"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "libnapafft_b200.so")
SRC = os.path.join(HERE, "csrc", "napafft.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "napafft_kernels.cuh"), os.path.join(ROOT, "include", "napafft.h")]


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build(force=False, verbose=False):
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in DEPS):
        return SO
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-o", SO, SRC]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    cmd[1:1] = os.environ.get("NAPA_NVCC_EXTRA", "").split()
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
