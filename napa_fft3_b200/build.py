"""Builds napa_fft3_b200/libnapafft_b200.so in-tree with nvcc for sm_100a (cross-compiles without a
GPU).  `python -m napa_fft3_b200.build [--force] [-v]`

Six translation units compiled in parallel (the host side + small kernels, and five slices of the kernel-template
instantiations of csrc/napafft_tables.cu), then one link: about 1.5 minutes on 8 cores instead of five."""
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "libnapafft_b200.so")
CSRC = os.path.join(HERE, "csrc")
SRC = os.path.join(CSRC, "napafft.cu")
TABLES = os.path.join(CSRC, "napafft_tables.cu")
DEPS = [SRC, TABLES, os.path.join(CSRC, "napafft_kernels.cuh"), os.path.join(ROOT, "include", "napafft.h")]
# (object name, source, extra defines)
UNITS = [("host", SRC, []), ("pass0", TABLES, ["-DNAPA_TABLE_PASS=0"]), ("pass1", TABLES, ["-DNAPA_TABLE_PASS=1"]),
         ("pass2", TABLES, ["-DNAPA_TABLE_PASS=2"]), ("fused", TABLES, ["-DNAPA_TABLE_FUSED"]),
         ("fused_tma", TABLES, ["-DNAPA_TABLE_FUSED_TMA"])]


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build(force=False, verbose=False):
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in DEPS):
        return SO
    common = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC,-fvisibility=hidden"]
    if verbose:
        common += ["-Xptxas", "-v"]
    common += os.environ.get("NAPA_NVCC_EXTRA", "").split()
    with tempfile.TemporaryDirectory(prefix="napafft_build_") as tmp:
        def compile_unit(unit):
            name, src, defs = unit
            obj = os.path.join(tmp, name + ".o")
            out = subprocess.run(common + defs + ["-c", "-o", obj, src], capture_output=True, text=True)
            if verbose or out.returncode:
                sys.stderr.write(out.stdout + out.stderr)
            if out.returncode:
                raise subprocess.CalledProcessError(out.returncode, out.args)
            return obj
        with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as pool:
            objs = list(pool.map(compile_unit, UNITS))
        subprocess.check_call([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO] + objs + ["-ldl"])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
