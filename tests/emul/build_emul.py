"""Builds tests/emul/libnapafft_emul.so: the shipped csrc/*.cu compiled by g++ against the
test-only CUDA emulation shim (cuda_emul.h).  Used by the `-m "not gpu"` tests to execute the real
planner and kernel source on a CPU.  Never used by the product package."""
import os
import subprocess
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libnapafft_emul.so")
CSRC = os.path.join(ROOT, "napa_fft3_b200", "csrc")
HOST, TABLES, SHIM = os.path.join(CSRC, "napafft.cu"), os.path.join(CSRC, "napafft_tables.cu"), os.path.join(HERE, "cuda_emul.cpp")
DEPS = [HOST, TABLES, SHIM, os.path.join(CSRC, "napafft_kernels.cuh"), os.path.join(ROOT, "include", "napafft.h"),
        os.path.join(HERE, "cuda_emul.h")]
# the same translation units as napa_fft3_b200/build.py
UNITS = [("host", HOST, []), ("pass0", TABLES, ["-DNAPA_TABLE_PASS=0"]), ("pass1", TABLES, ["-DNAPA_TABLE_PASS=1"]),
         ("pass2", TABLES, ["-DNAPA_TABLE_PASS=2"]), ("fused", TABLES, ["-DNAPA_TABLE_FUSED"]),
         ("fused_tma", TABLES, ["-DNAPA_TABLE_FUSED_TMA"]), ("shim", SHIM, [])]


def build(force=False, defines=(), so=SO):
    """defines / so: a differently configured build next to the default one (the mutant of tests/test_emul_orders.py)"""
    if not force and os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in DEPS):
        return so
    common = ["g++", "-O2", "-std=c++17", "-fPIC", "-DNAPA_EMULATE", "-fvisibility=hidden", *["-D" + d for d in defines], "-I", HERE]
    with tempfile.TemporaryDirectory(prefix="napafft_emul_") as tmp:
        def compile_unit(unit):
            name, src, defs = unit
            obj = os.path.join(tmp, name + ".o")
            subprocess.check_call(common + defs + ["-x", "c++", "-c", src, "-o", obj])
            return obj
        with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as pool:
            objs = list(pool.map(compile_unit, UNITS))
        subprocess.check_call(["g++", "-shared", "-o", so] + objs + ["-lm", "-lpthread"])
    return so


if __name__ == "__main__":
    print(build(force=True))
