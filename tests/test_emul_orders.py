"""Barrier-placement check on the CPU: the emulator switches fibers only at barriers and waits, so a barrier that is
MISSING between one thread's shared-memory write and another thread's read gives a wrong result whenever the reader gets
its turn first.  The parity checks are therefore repeated with the fibers scheduled in reverse and in shuffled order
(NAPA_EMUL_ORDER, tests/emul/cuda_emul.cpp), and a mutant build without the exchange barrier proves that this bites.
(compute-sanitizer's racecheck is closed on the GPU pool: profiles/r02_compute_sanitizer_closed.txt.)"""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))

PROBE = r"""
import os, sys
sys.path.insert(0, %(here)r); sys.path.insert(0, os.path.dirname(%(here)r))
import numpy as np
import parity_suite as S
from napa_fft3_b200 import Binding, NapaFFT
b = Binding(%(so)r); b.init(0)
napa = NapaFFT(b)
for n in (1 << 6, 1 << 8, 1 << 10, 1 << 12, 1 << 13, 1 << 14, 1 << 16):
    for io in (True, False):
        S.check_fft(napa, n, io, "sqrt", "real")
        S.check_ifft(napa, n, io, True, "complex")
    S.check_bit_reverse(napa, n)
    S.check_real(napa, n, None)
    S.check_2rfft(napa, n)
S.check_layout_bit_exact(napa, 1 << 12)
print("orders ok")
"""


def run_probe(so, order, extra=None):
    env = dict(os.environ, NAPA_EMUL_ORDER=order, **(extra or {}))
    return subprocess.run([sys.executable, "-c", PROBE % {"here": HERE, "so": so}], env=env, capture_output=True, text=True, timeout=1200)


@pytest.mark.parametrize("order", ["reverse", "shuffle"])
def test_parity_holds_under_other_thread_orders(order):
    from emul.build_emul import build
    out = run_probe(build(), order)
    assert out.returncode == 0 and "orders ok" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]


@pytest.mark.parametrize("order", ["reverse", "shuffle"])
def test_fused_kernels_under_other_thread_orders(order):
    from emul.build_emul import build
    build()
    for extra in ({"NAPAFFT_FUSED": "1"}, {"NAPAFFT_FUSED": "1", "NAPAFFT_FUSED_TMA": "1"}):
        env = dict(os.environ, NAPA_EMUL_ORDER=order, **extra)
        out = subprocess.run([sys.executable, os.path.join(HERE, "fused_probe.py"), "emul"], env=env, capture_output=True, text=True, timeout=1200)
        assert out.returncode == 0 and "fused ok" in out.stdout, str(extra) + out.stdout[-1500:] + out.stderr[-1500:]


def test_a_missing_barrier_is_caught():
    """mutant: the barrier between an exchange's scatter and the next stage's gather compiled out"""
    from emul.build_emul import build, HERE as EMUL
    mutant = build(defines=("NAPA_MUTANT_NO_EXCHANGE_BARRIER",), so=os.path.join(EMUL, "libnapafft_emul_mutant.so"))
    results = [run_probe(mutant, order) for order in ("forward", "reverse", "shuffle")]
    assert any(r.returncode != 0 for r in results), "no thread order exposed the missing barrier"
