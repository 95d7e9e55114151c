"""Pins the CPU oracle against every REPL transcript in the reference README.

The reference (Common Lisp, SBCL) cannot run in this image; these printed known answers are the
only outputs of the real implementation available, so the oracle must reproduce them exactly
(SBCL prints the shortest digits that round-trip, the same as Python's float repr, so parsing
the printed value gives the exact double).  README line numbers are into /root/reference/README.md.
"""
import numpy as np
import pytest


def C(re, im):
    return complex(re, im)


def exact(a, b):
    a = np.asarray(a)
    b = np.asarray(b, dtype=a.dtype)
    assert a.shape == b.shape
    assert np.array_equal(a, b), f"\n got {a!r}\nwant {b!r}"


def test_fft_iota8_in_order(oracle):  # README:106-110
    got = oracle.fft([0, 1, 2, 3, 4, 5, 6, 7])
    exact(got, [C(28.0, 0.0), C(-4.0, 9.65685424949238), C(-4.0, 4.0), C(-4.0, 1.6568542494923806),
                C(-4.0, 0.0), C(-4.0, -1.6568542494923806), C(-4.0, -4.0), C(-4.0, -9.65685424949238)])


def test_fft_iota8_bit_reversed(oracle):  # README:113-116 pins the :in-order nil layout
    got = oracle.fft([0, 1, 2, 3, 4, 5, 6, 7], in_order=False)
    exact(got, [C(28.0, 0.0), C(-4.0, 0.0), C(-4.0, 4.0), C(-4.0, -4.0), C(-4.0, 9.65685424949238),
                C(-4.0, -1.6568542494923806), C(-4.0, 1.6568542494923806), C(-4.0, -9.65685424949238)])


def test_fft_scales(oracle):  # README:119-128
    exact(oracle.fft([0, 1, 2, 3], scale=None), [C(6, 0), C(-2, 2), C(-2, 0), C(-2, -2)])
    exact(oracle.fft([0, 1, 2, 3], scale=True), [C(1.5, 0), C(-0.5, 0.5), C(-0.5, 0), C(-0.5, -0.5)])
    got = oracle.fft([0, 1, 2, 3, 5, 6, 7, 8], size=4, scale="sqrt")
    exact(got, [C(3, 0), C(-1, 1), C(-1, 0), C(-1, -1)])


def test_ifft_round_trips(oracle):  # README:174-184
    exact(oracle.ifft(oracle.fft([0, 1, 2, 3])), [C(0, 0), C(1, 0), C(2, 0), C(3, 0)])
    got = oracle.ifft(oracle.fft([0, 1, 2, 3], in_order=False, scale="sqrt"), in_order=False, scale="sqrt")
    exact(got, [C(0, 0), C(1, 0), C(2, 0), C(3, 0)])


def test_ifft_window_convolution(oracle):  # README:190-193
    w = oracle.fft([0, 0.5, 0, 0], in_order=False)
    got = oracle.ifft(oracle.fft([0, 1, 2, 3]), window=w)
    exact(got, [C(1.5, 0), C(0, 0), C(0.5, 0), C(1.0, 0)])


def test_bit_reverse_real(oracle):  # README:217-223
    v = np.array([0.0, 1.0, 2.0, 3.0])
    r = oracle.bit_reverse(v)
    exact(r, [0.0, 2.0, 1.0, 3.0])
    exact(oracle.bit_reverse(r, r), [0.0, 1.0, 2.0, 3.0])


def test_rfft_last_ulp(oracle):  # README:276-278
    got = oracle.rfft([0, 1, 2, 3])
    exact(got, [C(6.0, 0.0), C(-2.0, 2.0), C(-2.0, 0.0), C(-1.9999999999999998, -2.0)])


def test_rifft_rfft(oracle):  # README:307-308
    exact(oracle.rifft(oracle.rfft([0, 1, 2, 3])), [0.0, 1.0, 2.0, 3.0])


def test_polynomial_product(oracle):  # README:761-776
    a = oracle.fft(np.array([1, 0, 0, 5, 0, 0, 0, 0]), in_order=False)
    exact(a, [C(6.0, 0.0), C(-4.0, 0.0), C(1.0, 5.0), C(1.0, -5.0),
              C(-2.5355339059327378, -3.5355339059327378), C(4.535533905932738, 3.5355339059327378),
              C(4.535533905932738, -3.5355339059327378), C(-2.5355339059327378, 3.5355339059327378)])
    b = oracle.fft(np.array([1, 2, 3, 4, 0, 0, 0, 0]), in_order=False)
    exact(b, [C(10.0, 0.0), C(-2.0, 0.0), C(-2.0, 2.0), C(-2.0, -2.0),
              C(-0.41421356237309515, -7.242640687119286), C(2.414213562373095, 1.2426406871192857),
              C(2.414213562373095, -1.2426406871192857), C(-0.41421356237309515, 7.242640687119286)])
    got = oracle.ifft(b, window=a, in_order=False)
    exact(got, [C(0.9999999999999982, 0.0), C(1.9999999999999991, 0.0), C(3.0, 0.0), C(9.0, 0.0),
                C(10.000000000000002, 0.0), C(15.0, 0.0), C(20.0, 0.0), C(-8.881784197001252e-16, 0.0)])


def test_2rfft_and_rifft(oracle):  # README:794-804
    both = oracle.two_rfft([1, 0, 0, 5, 0, 0, 0, 0], [1, 2, 3, 4, 0, 0, 0, 0])
    exact(both[:4], [C(6.0, 0.0), C(-2.5355339059327378, -3.5355339059327378), C(1.0, 5.0),
                     C(4.535533905932738, -3.535533905932738)])
    x, y = both[:8].copy(), both[8:].copy()
    got = oracle.rifft(x * y)
    exact(got, [0.9999999999999991, 2.0, 2.9999999999999964, 9.0, 10.0, 15.0, 20.000000000000004,
                -8.881784197001252e-16])


def test_get_fft_all_ones(oracle):  # README:859-865 (get-fft 8 = DIF then bit reversal, in place)
    data = np.full(8, 1 + 0j)
    oracle.lowlevel(data, 0, 8, +1)
    oracle.bit_reverse(data, data)
    exact(data, [C(8, 0)] + [C(0, 0)] * 7)


def test_sqrt_inverse_factor_quirk(oracle):  # interface.lisp:44-48: (1/n)/(1/sqrt n), not 1/sqrt n
    assert oracle.scale_factor(8, False, "sqrt") == (1.0 / 8.0) / (1.0 / np.sqrt(8.0))
    assert oracle.scale_factor(8, True, "sqrt") == 1.0 / np.sqrt(8.0)
    assert oracle.scale_factor(8, False, "sqrt") != oracle.scale_factor(8, True, "sqrt")


@pytest.mark.parametrize("lg", range(0, 15))
def test_matches_numpy(oracle, lg):
    """Sanity beyond the KATs: the restated split radix is a DFT (vs numpy pocketfft)."""
    n = 1 << lg
    rng = np.random.Generator(np.random.Philox(1000 + lg))
    x = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    ref = np.fft.fft(x)
    tol = 1e-13 * max(lg, 1)
    assert np.linalg.norm(oracle.fft(x) - ref) <= tol * np.linalg.norm(ref)
    rev = np.array([int(oracle.lib().napa_oracle_bit_reverse_integer(i, lg)) for i in range(n)])
    exact(oracle.fft(x, in_order=False), oracle.fft(x)[rev])          # out[p] = X[rev(p)]
    back = oracle.ifft(oracle.fft(x, in_order=False), in_order=False)
    assert np.linalg.norm(back - x) <= tol * np.linalg.norm(x)
    if n >= 2:
        xr = rng.uniform(-1, 1, n)
        refr = np.fft.fft(xr)
        rtol = max(tol, 4e-16 * n)  # recurrence twiddles: error grows with n
        assert np.linalg.norm(oracle.rfft(xr) - refr) <= rtol * np.linalg.norm(refr)
        assert np.linalg.norm(oracle.rifft(oracle.rfft(xr)) - xr) <= 2 * rtol * np.linalg.norm(xr)
