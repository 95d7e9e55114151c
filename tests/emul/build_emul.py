"""Builds tests/emul/libnapafft_emul.so: the shipped csrc/napafft.cu compiled by g++ against the
test-only CUDA emulation shim (cuda_emul.h).  Used by the `-m "not gpu"` tests to execute the real
planner and kernel source on a CPU.  Never used by the product package."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libnapafft_emul.so")
SRCS = [os.path.join(ROOT, "napa_fft3_b200", "csrc", "napafft.cu"), os.path.join(HERE, "cuda_emul.cpp")]
DEPS = SRCS + [os.path.join(ROOT, "napa_fft3_b200", "csrc", "napafft_kernels.cuh"),
               os.path.join(ROOT, "include", "napafft.h"), os.path.join(HERE, "cuda_emul.h")]


def build(force=False, defines=(), so=SO):
    """defines / so: a differently configured build next to the default one (the mutant of tests/test_emul_orders.py)"""
    if not force and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in DEPS):
        return so
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DNAPA_EMULATE", "-fvisibility=hidden",
           *["-D" + d for d in defines], "-I", HERE, "-x", "c++", SRCS[0], SRCS[1], "-o", so, "-lm", "-lpthread"]
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    print(build(force=True))
