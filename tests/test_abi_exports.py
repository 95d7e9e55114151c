"""CPU tier: the product C-ABI library builds for sm_100a, loads without a GPU, exports every
symbol include/napafft.h declares, and FAILS LOUDLY (no CPU fallback) when no device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def product_so():
    from napa_fft3_b200.build import build
    return build()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "napafft.h")).read()
    return sorted(set(re.findall(r"\b(napafft_[a-z0-9_]+)\s*\(", src)) - {"napafft_status"})


def test_header_and_binding_agree():
    from napa_fft3_b200._lib import ABI_SYMBOLS
    assert header_symbols() == sorted(ABI_SYMBOLS)


def test_product_library_exports_abi(product_so):
    lib = C.CDLL(product_so)
    for sym in header_symbols():
        assert hasattr(lib, sym), sym
    assert lib.napafft_abi_version() == 2


def test_emulation_library_exports_same_abi():
    from emul.build_emul import build
    lib = C.CDLL(build())
    for sym in header_symbols():
        assert hasattr(lib, sym), sym


def test_no_device_fails_loudly(product_so):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import napa_fft3_b200 as napa
    with pytest.raises(napa.NapaFFTError) as ei:
        napa.fft(np.arange(8, dtype=np.complex128))
    assert ei.value.status in (5, 8)


def test_product_package_never_touches_oracle_or_emulation():
    pkg = os.path.join(ROOT, "napa_fft3_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".lisp", ".asd")):
                text = open(os.path.join(dirpath, f)).read()
                assert "napa_oracle" not in text and "libnapafft_emul" not in text, f
                assert "import oracle" not in text and "from oracle" not in text, f


def test_argument_errors_via_emulated_abi():
    """status codes / messages of the C ABI (host logic only; runs on the emulation build)."""
    from emul.build_emul import build
    from napa_fft3_b200 import Binding, NapaFFTError
    b = Binding(build())
    b.init(0)
    x = np.zeros(16, np.complex128)
    p = C.c_void_p(x.ctypes.data)
    for args, code in [((p, p, 12, 1, 12, 1, 0, 1, None, 0, 0), 1), ((None, p, 8, 1, 8, 1, 0, 1, None, 0, 0), 3),
                       ((p, p, 8, 2, 4, 1, 0, 1, None, 0, 0), 2), ((p, p, 8, 1, 8, 3, 0, 1, None, 0, 0), 4),
                       ((p, p, 8, 1, 8, 1, 7, 1, None, 0, 0), 4), ((p, p, 8, 1, 8, 1, 0, 1, None, 2, 0), 3)]:
        with pytest.raises(NapaFFTError) as ei:
            b.c2c(*args)
        assert ei.value.status == code
    with pytest.raises(NapaFFTError):
        b.rfft(p, p, 1, 1, 0)
    assert b.scale_factor(8, -1, 2) == (1.0 / 8.0) / (1.0 / np.sqrt(8.0))     # interface.lisp:44-48
    from oracle import napa_oracle as O
    tw = np.zeros(512, np.complex128)
    b.radix2_twiddle(1024, -1, C.c_void_p(tw.ctypes.data))
    assert np.array_equal(tw, O.radix2_twiddle(1024, -1.0))                     # real.lisp:53-64 bit-exact
    b.radix2_twiddle(1024, 1, C.c_void_p(tw.ctypes.data))
    assert np.array_equal(tw, O.radix2_twiddle(1024, 1.0))


def test_example_helpers(tmp_path):
    """example.lisp:1-79 host helpers: raw s32 files round-trip with the reference's scaling rules"""
    import numpy as np
    from napa_fft3_b200 import examples as E
    x = np.array([0.0, 0.5, -0.5, 0.999, -1.0, 1e-10])
    f = E.emit_raw32_file(str(tmp_path / "a.raw"), x)
    raw = np.fromfile(f, dtype=np.int32)
    assert list(raw) == [int(np.floor(v * (2**31 - 1))) for v in x]
    back = E.read_raw32_file(f)
    assert np.array_equal(back, raw * 2.0 ** -31) and np.max(np.abs(back - x)) < 2.0 ** -30
    assert np.array_equal(E.read_raw32_file(f, 3), back[:3])
    assert np.array_equal(E.impulse([1, 3], 4, 2), np.array([0, 2, 0, 2], dtype=complex))
    v = np.array([1 + 0j, -3j, 0.5, 2])
    assert np.array_equal(E.amplitude_clamp_fraction(v, 0.5), np.array([0, -3j, 0, 2]))
    assert np.array_equal(E.amplitude_clamp_k(v, 2), np.array([0, -3j, 0, 2]))
    assert E.energy(v) == 1 + 9 + 0.25 + 4
    assert np.array_equal(E.m_plus([1, 2], [3, 4]), np.array([2, 3], dtype=complex))
