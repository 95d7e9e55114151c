"""The reference's own (REPL-driven) property suites, re-expressed against the CPU oracle.

tests.lisp:66-104 (round-trip pairs), :106-208 (windows, incl. the bit-reversed window for the
in-order inverse), :210-265 (offset invariance, exact), ergun-test.lisp:21-105 (linearity, impulse,
shift theorem; m= absolute tolerance 1e-6, test-support.lisp:35-49).
"""
import numpy as np
import pytest

SIZES = [1 << k for k in range(0, 13)]


def rand_c(rng, n):  # test-support.lisp:19-25
    return rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)


def delta(x, y):  # tests.lisp:3-16
    d = np.abs(x - y)
    m = np.maximum(np.abs(x), np.abs(y))
    return float(np.max(np.where(d > 0, d / np.where(m > 0, m, 1), 0.0))) if len(x) else 0.0


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("split", ["1", "sqrt", "1/n"])
@pytest.mark.parametrize("in_order", [True, False])
def test_pairs(oracle, n, split, in_order):
    rng = np.random.Generator(np.random.Philox(n + 17))
    x = rand_c(rng, n)
    fs, bs = {"1": (None, True), "sqrt": ("sqrt", "sqrt"), "1/n": (True, None)}[split]
    y = oracle.ifft(oracle.fft(x, in_order=in_order, scale=fs), in_order=in_order, scale=bs)
    assert delta(x, y) / max(np.log2(n), 1) < 1e-13


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("wf,wi", [(True, True), (True, False), (False, True)])
@pytest.mark.parametrize("in_order", [True, False])
def test_windows(oracle, n, wf, wi, in_order):
    rng = np.random.Generator(np.random.Philox(n + 29))
    window = rng.uniform(1, 2, n)
    x = rand_c(rng, n)
    v = x.copy()
    if wf:
        v = v / window
    v = oracle.fft(v, in_order=in_order, scale="sqrt", window=window if wf else None)
    if wi:
        v = v / window
        # in-order inverse: the window is applied after the bit reversal, so it must be
        # bit-reversed itself (tests.lisp:156-172)
        w = oracle.bit_reverse(window) if in_order else window
    v = oracle.ifft(v, in_order=in_order, scale="sqrt", window=w if wi else None)
    assert delta(x, v) / max(np.log2(n), 1) < 1e-13


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("direction", [+1, -1])
def test_offsets_exact(oracle, n, direction):
    """tests.lisp:210-265: start=13 / window-start=7 give exactly the same bits as 0 / 0."""
    rng = np.random.Generator(np.random.Philox(n + 41))
    window = rng.uniform(1, 2, n)
    x = rand_c(rng, n)
    ref = oracle.lowlevel(x.copy(), 0, n, direction, 1.0, window, 0)
    owin = np.concatenate([np.zeros(7), window])
    ox = np.concatenate([np.zeros(13, dtype=complex), x])
    for w, ws in ((owin, 7), (window, 0)):
        for d, st in ((ox, 13), (x, 0)):
            got = oracle.lowlevel(d.copy(), st, n, direction, 1.0, w, ws)[st:st + n]
            assert np.array_equal(got, ref)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("in_order", [True, False])
def test_ergun(oracle, n, in_order):
    rng = np.random.Generator(np.random.Philox(n + 53))
    lg = int(np.log2(n))
    rev = np.array([int(oracle.lib().napa_oracle_bit_reverse_integer(i, lg)) for i in range(n)])
    f = lambda v: oracle.fft(v, in_order=in_order)
    for _ in range(4):
        a, b = rand_c(rng, n), rand_c(rng, n)
        assert np.max(np.abs(f(a) + f(b) - f(a + b))) < 1e-6                    # linearity
        r = np.zeros(n, complex); r[0] = 1
        assert np.max(np.abs(f(a) + f(r - a) - np.ones(n))) < 1e-6             # impulse
        a[-1] = 0
        y1, y2 = f(a), f(np.roll(a, 1))                                         # shift theorem
        if not in_order:
            y1, y2 = y1[rev], y2[rev]   # rev is an involution: natural[k] = out[rev(k)]
        assert np.max(np.abs(y1 * np.exp(-2j * np.pi * np.arange(n) / n) - y2)) < 1e-6


def test_window_functions(oracle):
    """windowing.lisp:6-51 quirks: hann is double; rectangular is 1.0f0; blackman-harris and
    blackman use single-float coefficients widened to double; triangle/bartlett are single."""
    n = 64
    i = np.arange(n)
    assert np.array_equal(oracle.window_vector("rectangular", n), np.ones(n))
    hann = 0.5 * (1.0 - np.cos(((2 * np.pi) * i) / (n - 1)))
    assert np.array_equal(oracle.window_vector("hann", n), hann)
    tri = oracle.window_vector("triangle", n)
    assert np.array_equal(tri, tri.astype(np.float32).astype(np.float64))       # single precision
    want = (np.float32(2.0 / n) * (np.float32(n) * np.float32(0.5)
            - np.abs(i.astype(np.float32) - np.float32(0.5) * np.float32(n - 1)))).astype(np.float64)
    assert np.array_equal(tri, want)
    bh = oracle.window_vector("blackman-harris", n)
    a = [np.float64(np.float32(c)) for c in (0.35875, 0.48829, 0.14128, 0.01168)]
    want = ((a[0] - a[1] * np.cos(((2 * np.pi) * i) / (n - 1))) + a[2] * np.cos(((4 * np.pi) * i) / (n - 1))) \
        - a[3] * np.cos(((6 * np.pi) * i) / (n - 1))
    assert np.array_equal(bh, want)
    assert np.array_equal(oracle.window_vector("hann", n, bit_reverse=True), oracle.bit_reverse(hann))


def test_extract_centered_window(oracle):  # windowing.lisp:67-111, zero padding off both ends
    sig = np.arange(1, 11, dtype=np.float64)
    assert np.array_equal(oracle.extract_centered_window(sig, 5, 4), [4, 5, 6, 7])
    assert np.array_equal(oracle.extract_centered_window(sig, 0, 4), [0, 0, 1, 2])
    assert np.array_equal(oracle.extract_centered_window(sig, 10, 4), [9, 10, 0, 0])
    assert np.array_equal(oracle.extract_centered_window(sig, 30, 4), [0, 0, 0, 0])
    assert np.array_equal(oracle.extract_centered_window(sig, 5, 16),
                          [0, 0, 0] + list(range(1, 11)) + [0, 0, 0])


def test_windowed_fft_matches_manual(oracle):
    rng = np.random.Generator(np.random.Philox(7))
    sig = rand_c(rng, 300)
    got = oracle.windowed_fft(sig, 100, 64)
    chunk = sig[68:132] * oracle.window_vector("hann", 64)
    assert np.linalg.norm(got - np.fft.fft(chunk)) < 1e-12 * np.linalg.norm(got)


def test_size_keeps_tail(oracle):
    """easy-interface.lisp:44-67: :size transforms the first n elements; the rest of dst is a copy."""
    rng = np.random.Generator(np.random.Philox(9))
    x = rand_c(rng, 24)
    out = oracle.fft(x, size=16)
    assert out is not x and np.array_equal(out[16:], x[16:])
    assert np.allclose(out[:16], np.fft.fft(x[:16]))
    same = oracle.fft(x, dst=x, size=16)
    assert same is x


def test_golden_vectors_reproduce(oracle):
    """the oracle still reproduces tests/golden/oracle_vectors.npz bit for bit (fixture drift guard:
    compiler flags, libm, edits to oracle/)"""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import parity_suite as S
    S.check_golden(oracle, exact=True)
