"""CPU tier: the shipped planner + kernel source (napa_fft3_b200/csrc) compiled against the
test-only CUDA emulation, driven through the same C ABI and Python host as on the GPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_suite as S  # noqa: E402


@pytest.fixture(scope="module")
def napa():
    from emul.build_emul import build
    from napa_fft3_b200 import Binding, NapaFFT
    b = Binding(build())
    b.init(0)
    return NapaFFT(b)


SIZES = [1 << k for k in range(0, 15)]


@pytest.mark.parametrize("n", SIZES)
def test_fft_all_options(napa, n):
    for in_order in (True, False):
        for scale, wt in ((None, None), (True, "real"), ("sqrt", "complex")):
            S.check_fft(napa, n, in_order, scale, wt)


@pytest.mark.parametrize("n", SIZES)
def test_ifft_all_options(napa, n):
    for in_order in (True, False):
        for scale, wt in ((True, None), (None, "real"), ("sqrt", "complex")):
            S.check_ifft(napa, n, in_order, scale, wt)


@pytest.mark.parametrize("n", [1 << k for k in range(1, 15)])
def test_layout_bit_exact(napa, n):
    S.check_layout_bit_exact(napa, n)


@pytest.mark.parametrize("n", [1 << k for k in range(0, 16)])
def test_bit_reverse(napa, n):
    S.check_bit_reverse(napa, n)


@pytest.mark.parametrize("n", [1, 2, 8, 32, 1024, 1 << 14])
def test_offsets_exact(napa, n):
    S.check_offsets_exact(napa, n)


@pytest.mark.parametrize("n", [1 << k for k in range(1, 16)])
def test_real(napa, n):
    for scale in (None, True, "sqrt"):
        S.check_real(napa, n, scale)


@pytest.mark.parametrize("n", [1, 2, 4, 64, 1 << 13])
def test_2rfft(napa, n):
    if n >= 2:
        S.check_2rfft(napa, n)


@pytest.mark.parametrize("n", [4, 64, 1024])
def test_ergun(napa, n):
    S.check_ergun(napa, n, True)
    S.check_ergun(napa, n, False)


def test_readme_kats(napa):
    S.check_readme_kats(napa)


def test_dst_rules(napa):
    S.check_dst_rules(napa)


def test_windowed(napa):
    S.check_windowed(napa)


def test_window_functions(napa):
    S.check_window_functions(napa)


def test_batch_and_stride(napa):
    """new export: many transforms per call (rows), incl. a per-batch window"""
    from oracle import napa_oracle as O
    x = np.stack([S.rand_c(100 + i, 256) for i in range(19)])
    w = np.stack([S.rand_r(200 + i, 256, 1, 2) for i in range(19)])
    got = napa.fft_batch(x.copy(), in_order=False, scale="sqrt", window=w)
    for i in range(19):
        assert S.nre(got[i], O.fft(x[i], in_order=False, scale="sqrt", window=w[i])) <= S.tol(256)
    r = np.stack([S.rand_r(300 + i, 512) for i in range(5)])
    spec = napa.rfft_batch(r)
    for i in range(5):
        assert S.nre(spec[i], O.rfft(r[i])) <= S.tol(512)
    back = napa.rifft_batch(spec)
    assert S.nre(back, r) < 1e-12


@pytest.mark.parametrize("n,batch", [(256, 300), (1024, 70), (4096, 17), (8192, 9), (64, 1100)])
def test_many_tiles_single_pass(napa, n, batch):
    """many tiles per launch, incl. a ragged
    last tile, the bit-reversed load path and a fused window"""
    from oracle import napa_oracle as O
    x = np.stack([S.rand_c(500 + i, n) for i in range(batch)])
    w = S.rand_r(77, n, 1, 2)
    got = napa.fft_batch(x.copy(), in_order=False, scale="sqrt", window=w)
    back = napa.fft_batch(got.copy(), direction=-1, in_order=False, scale="sqrt")
    for i in (0, 1, batch // 2, batch - 1):
        want = O.fft(x[i], in_order=False, scale="sqrt", window=w)
        assert S.nre(got[i], want) <= S.tol(n)
        assert S.nre(back[i], O.ifft(want.copy(), in_order=False, scale="sqrt")) <= S.tol(n)
    assert S.nre(back, x * w) < 1e-13


def test_fused_kernel_schedules():
    """the fused single-launch kernels (opt-in, NAPAFFT_FUSED=1; direct-load and TMA-fed warp-specialised variants)
    under several schedules: the emulator keeps every CTA of a cooperative launch alive (tests/emul/cuda_emul.cpp:
    launch_cooperative) and emulates mbarriers, so the tile counters, the waits on them, the producer / compute-thread
    hand-off, ring-slot reuse, ragged last units and odd unit counts all really run; the default plans (one launch per
    digit pass) are covered by every other test of this file"""
    import subprocess
    for extra in ({}, {"NAPAFFT_FUSED_TMA": "1"}, {"NAPAFFT_FUSED_UPG": "1", "NAPAFFT_FUSED_GRID": "2", "NAPAFFT_FUSED_LEAD": "1"},
                  {"NAPAFFT_FUSED_TMA": "1", "NAPAFFT_FUSED_UPG": "3", "NAPAFFT_FUSED_GRID": "5", "NAPAFFT_FUSED_LEAD": "3"}):
        env = dict(os.environ, NAPAFFT_FUSED="1", **extra)
        out = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "fused_probe.py"), "emul"],
                             env=env, capture_output=True, text=True, timeout=1200)
        assert out.returncode == 0 and "fused ok" in out.stdout, str(extra) + out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("digits,n", [("5,5,6", 1 << 16), ("7,8", 1 << 15), ("6,5,6", 1 << 17)])
def test_three_digit_plans(napa, digits, n, monkeypatch):
    """the plans of 2^24 and beyond (three digit passes; natural order without a bit-reversal pass)
    at emulator-sized n, selected with the NAPAFFT_DIGITS tuning override"""
    monkeypatch.setenv("NAPAFFT_DIGITS", digits)
    for in_order in (True, False):
        S.check_fft(napa, n, in_order, "sqrt", "real")
        S.check_ifft(napa, n, in_order, True, "complex")
    S.check_layout_bit_exact(napa, n)
    S.check_offsets_exact(napa, n)


@pytest.mark.parametrize("n,chunk", [(1000, 64), (4096, 256), (777, 1024), (5000, 2), (3 * 8192 + 5, 16384)])
def test_filter_chunks(napa, n, chunk):
    """overlap-add driver (ragged last chunks, signal shorter than a chunk, two-pass chunk size)"""
    S.check_filter_chunks(napa, n, chunk)


@pytest.mark.parametrize("n", [2, 8, 256, 4096, 1 << 14])
def test_typed_inputs(napa, n):
    S.check_typed_inputs(napa, n)


def test_natural_order_in_place_fallback():
    """single transforms too large for a scratch copy (2^31 and up by default) take the in-place
    bit-reversed pipeline + bit-reversal pass; forced here at small sizes with NAPAFFT_MAX_SCRATCH_KB"""
    import subprocess
    env = dict(os.environ, NAPAFFT_MAX_SCRATCH_KB="1")
    out = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "fused_probe.py"), "emul"],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "fused ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_golden_vectors(napa):
    """shipped csrc (emulated) against the committed oracle fixtures"""
    S.check_golden(napa)


@pytest.mark.parametrize("n,batch", [(2, 7), (16, 301), (256, 13), (1024, 5)])
def test_bit_reverse_batches(napa, n, batch):
    """several small vectors per CTA (ragged last CTA), complex and real"""
    x = np.stack([S.rand_c(i, n) for i in range(min(batch, 64))] * ((batch + 63) // 64))[:batch]
    p = S.revperm(n)
    assert np.array_equal(napa.bit_reverse_batch(x.copy()), x[:, p])
    xr = np.ascontiguousarray(x.real)
    assert np.array_equal(napa.bit_reverse_batch(xr.copy()), xr[:, p])


@pytest.mark.parametrize("n", [8, 256, 1 << 14])
def test_get_windowed_fft(napa, n):
    S.check_get_windowed_fft(napa, n)


@pytest.mark.parametrize("lgN", [16, 20])
def test_distributed_pipeline_single_rank(napa, lgN):
    """napafft_dist_c2c_dev with one rank (the exchange is a copy): chunked column FFTs with the fused global twiddle,
    the [chunk][rank][row][Bc] exchange layout, row FFTs reading it in place; library-driven and host-carried exchange"""
    import torch
    from napa_fft3_b200.dist import DistributedFFT
    N = 1 << lgN
    x = S.rand_c(1005, N)
    ref = np.fft.fft(x)
    for native in (True, False):
        f = DistributedFFT(lgN, napa.b, None, torch.device("cpu"), native=native)
        xl = torch.from_numpy(f.scatter_time(x))
        Xl = f.forward(xl)
        assert S.nre(Xl.numpy(), ref[f.freq_index()]) <= S.tol(N)
        assert S.nre(f.inverse(Xl).numpy(), f.scatter_time(x)) <= S.tol(N)
    assert f.nch > 1


def test_distributed_pipeline_two_pass_rows(napa, monkeypatch):
    """rows of the distributed transform long enough for a two-digit plan (2^14 and up on the GPU) read / write the
    exchange layout in their column-like first pass and their transposed last store; forced here at 2^11 rows with the
    digit override"""
    import torch
    from napa_fft3_b200.dist import DistributedFFT
    monkeypatch.setenv("NAPAFFT_DIGITS", "7,4")
    lgN = 22
    x = S.rand_c(1006, 1 << lgN)
    ref = np.fft.fft(x)
    f = DistributedFFT(lgN, napa.b, None, torch.device("cpu"))
    before = napa.b.launch_count()
    Xl = f.forward(torch.from_numpy(f.scatter_time(x)))
    assert napa.b.launch_count() - before >= f.nch + 2 * 4          # column chunks + two passes per row block
    assert S.nre(Xl.numpy(), ref[f.freq_index()]) <= S.tol(1 << lgN)
    assert S.nre(f.inverse(Xl).numpy(), f.scatter_time(x)) <= S.tol(1 << lgN)
