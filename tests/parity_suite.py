"""Parity checks shared by the two execution tiers:

  * tests/test_emul_parity.py  (-m "not gpu"): the shipped csrc compiled against the test-only CUDA
    emulation (tests/emul), small sizes, proves planner + kernel index logic on a CPU;
  * tests/test_gpu_parity.py   (-m gpu): the product library on a B200 through the C ABI.

Every check compares against the CPU oracle (oracle/), which is pinned to the reference README.
Tolerance (BASELINE.json north_star): normwise relative error <= 1e-13 * log2(n) for FP64 data;
bit-exact for permutations, layouts and offset invariance.
"""
import os

import numpy as np

from oracle import napa_oracle as O


def tol(n):
    return 1e-13 * max(np.log2(n), 1.0)


def nre(got, want):
    d = np.linalg.norm(np.asarray(got) - np.asarray(want))
    w = np.linalg.norm(np.asarray(want))
    return d / w if w > 0 else d


def rand_c(seed, n):
    rng = np.random.Generator(np.random.Philox(seed))
    return rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)


def rand_r(seed, n, lo=-1.0, hi=1.0):
    rng = np.random.Generator(np.random.Philox(seed))
    return rng.uniform(lo, hi, n)


def revperm(n):
    lg = int(np.log2(n))
    idx = np.arange(n, dtype=np.uint64)
    out = np.zeros(n, dtype=np.uint64)
    for b in range(lg):
        out |= ((idx >> np.uint64(b)) & np.uint64(1)) << np.uint64(lg - 1 - b)
    return out.astype(np.int64)


def check_fft(napa, n, in_order, scale, wtype, seed=1):
    x = rand_c(seed + n, n)
    w = None
    if wtype == "real":
        w = rand_r(seed + 7, n, 1.0, 2.0)
    elif wtype == "complex":
        w = rand_c(seed + 9, n)
    got = napa.fft(x, in_order=in_order, scale=scale, window=w)
    want = O.fft(x, in_order=in_order, scale=scale, window=w)
    assert got is not x
    e = nre(got, want)
    assert e <= tol(n), (n, in_order, scale, wtype, e)
    return e


def check_ifft(napa, n, in_order, scale, wtype, seed=2):
    x = rand_c(seed + n, n)
    w = None
    if wtype == "real":
        w = rand_r(seed + 7, n, 1.0, 2.0)
    elif wtype == "complex":
        w = rand_c(seed + 9, n)
    got = napa.ifft(x, in_order=in_order, scale=scale, window=w)
    want = O.ifft(x, in_order=in_order, scale=scale, window=w)
    e = nre(got, want)
    assert e <= tol(n), (n, in_order, scale, wtype, e)
    return e


def check_layout_bit_exact(napa, n):
    """:in-order nil position<->bin mapping: out[p] = X[rev(p)] exactly (README:113-116)."""
    x = rand_c(11 + n, n)
    nat = napa.fft(x, in_order=True)
    rev = napa.fft(x, in_order=False)
    # same arithmetic, different store index -> identical bits.  2^17 and 2^19 .. 2^22 use a different digit
    # plan per output order (napafft.cu digits_of, tuned per layout), so there the two agree to rounding.
    per_order_plans = n in (1 << 17, 1 << 19, 1 << 20, 1 << 21, 1 << 22) and not os.environ.get("NAPAFFT_DIGITS")
    if not per_order_plans:
        assert np.array_equal(rev, nat[revperm(n)]), n
    else:
        assert nre(rev, nat[revperm(n)]) <= tol(n), n
    # impulse train pins the mapping independently of rounding: delta at j -> bins e^{-2 pi i jk/n}
    for j in (0, 1, n // 2, n - 1):
        d = np.zeros(n, complex); d[j % n] = 1
        got = napa.fft(d, in_order=False)
        k = revperm(n)
        want = np.exp(-2j * np.pi * ((j % n) * k % n) / n)
        assert np.max(np.abs(got - want)) < 1e-12, (n, j)
    # inverse consumes the bit-reversed layout
    back = napa.ifft(rev, in_order=False)
    assert nre(back, x) <= tol(n)


def check_bit_reverse(napa, n):
    p = revperm(n)
    xc = rand_c(21 + n, n)
    assert np.array_equal(napa.bit_reverse(xc), xc[p])
    assert np.array_equal(napa.bit_reverse(xc), O.bit_reverse(xc))
    xr = rand_r(22 + n, n)
    assert np.array_equal(napa.bit_reverse(xr), xr[p])
    y = xc.copy()
    assert napa.bit_reverse(y, y) is y and np.array_equal(y, xc[p])      # in place
    assert np.array_equal(napa.bit_reverse(y, y), xc)                     # involution


def check_offsets_exact(napa, n):
    """tests.lisp:210-265: start=13, window-start=7 give exactly the bits of start 0."""
    for direction in (1, -1):
        f = napa.ensure_fft(direction, None, "real-sample", n)
        window = rand_r(31 + n, n, 1.0, 2.0)
        x = rand_c(32 + n, n)
        ref = f(x.copy(), 0, window, 0)
        owin = np.concatenate([np.zeros(7), window])
        ox = np.concatenate([np.zeros(13, complex), x])
        for w, ws in ((owin, 7), (window, 0)):
            for d, st in ((ox, 13), (x, 0)):
                got = f(d.copy(), st, w, ws)[st:st + n]
                assert np.array_equal(got, ref), (n, direction, ws, st)
        want = O.lowlevel(x.copy(), 0, n, direction, 1.0, window, 0)
        assert nre(ref, want) <= tol(n)


def check_real(napa, n, scale, seed=3):
    x = rand_r(seed + n, n)
    got = napa.rfft(x, scale=scale)
    want = O.rfft(x, scale=scale)
    e1 = nre(got, want)
    assert e1 <= tol(n), ("rfft", n, scale, e1)
    inv_scale = {None: True, True: None, "sqrt": "sqrt"}[scale]
    spec = want.copy()
    back = napa.rifft(spec, scale=inv_scale)
    want_back = O.rifft(want.copy(), scale=inv_scale)
    e2 = nre(back, want_back)
    assert e2 <= tol(n), ("rifft", n, scale, e2)
    assert nre(back, x) <= 4 * max(tol(n), 4e-16 * n)
    return e1, e2


def check_2rfft(napa, n, seed=4):
    a, b = rand_r(seed + n, n), rand_r(seed + n + 1, n)
    got = napa.two_rfft(a, b)
    want = O.two_rfft(a, b)
    assert nre(got, want) <= tol(n), n
    assert nre(got[:n], np.fft.fft(a)) <= tol(n) and nre(got[n:], np.fft.fft(b)) <= tol(n)


def check_ergun(napa, n, in_order):
    """ergun-test.lisp:21-105: linearity, impulse, shift theorem; abs tolerance 1e-6."""
    rng = np.random.Generator(np.random.Philox(n + 53))
    f = napa.get_fft(n, in_order=in_order)
    p = revperm(n)
    rc = lambda: rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    for _ in range(3):
        a, b = rc(), rc()
        s = a + b
        fa, fb, fs = f(a.copy()), f(b.copy()), f(s.copy())
        assert np.max(np.abs(fa + fb - fs)) < 1e-6
        r = np.zeros(n, complex); r[0] = 1
        assert np.max(np.abs(fa + f(r - a) - 1)) < 1e-6
        a[-1] = 0
        y1, y2 = f(a.copy()), f(np.roll(a, 1))
        if not in_order:
            y1, y2 = y1[p], y2[p]
        assert np.max(np.abs(y1 * np.exp(-2j * np.pi * np.arange(n) / n) - y2)) < 1e-6


def check_readme_kats(napa):
    """The README transcripts through the product path (tolerance, not last-bit: the GPU uses a
    different factorisation than the reference's split radix)."""
    C = complex
    close = lambda a, b: np.allclose(a, b, rtol=0, atol=1e-13)
    assert close(napa.fft([0, 1, 2, 3, 4, 5, 6, 7]),
                 [28, C(-4, 9.65685424949238), C(-4, 4), C(-4, 1.6568542494923806), -4,
                  C(-4, -1.6568542494923806), C(-4, -4), C(-4, -9.65685424949238)])
    assert close(napa.fft([0, 1, 2, 3, 4, 5, 6, 7], in_order=False),
                 [28, -4, C(-4, 4), C(-4, -4), C(-4, 9.65685424949238), C(-4, -1.6568542494923806),
                  C(-4, 1.6568542494923806), C(-4, -9.65685424949238)])
    assert close(napa.fft([0, 1, 2, 3], scale=True), [1.5, C(-.5, .5), -.5, C(-.5, -.5)])
    assert close(napa.fft([0, 1, 2, 3, 5, 6, 7, 8], size=4, scale="sqrt"), [3, C(-1, 1), -1, C(-1, -1)])
    assert close(napa.ifft(napa.fft([0, 1, 2, 3])), [0, 1, 2, 3])
    assert close(napa.ifft(napa.fft([0, 1, 2, 3], in_order=False, scale="sqrt"), in_order=False, scale="sqrt"),
                 [0, 1, 2, 3])
    assert close(napa.ifft(napa.fft([0, 1, 2, 3]), window=napa.fft([0, 0.5, 0, 0], in_order=False)),
                 [1.5, 0, 0.5, 1.0])
    v = np.array([0.0, 1.0, 2.0, 3.0])
    r = napa.bit_reverse(v)
    assert np.array_equal(r, [0, 2, 1, 3]) and np.array_equal(napa.bit_reverse(r, r), v)
    assert close(napa.rfft([0, 1, 2, 3]), [6, C(-2, 2), -2, C(-2, -2)])
    assert close(napa.rifft(napa.rfft([0, 1, 2, 3])), [0, 1, 2, 3])
    a = napa.fft(np.array([1, 0, 0, 5, 0, 0, 0, 0]), in_order=False)
    b = napa.fft(np.array([1, 2, 3, 4, 0, 0, 0, 0]), in_order=False)
    assert close(napa.ifft(b, window=a, in_order=False), [1, 2, 3, 9, 10, 15, 20, 0])
    both = napa.two_rfft([1, 0, 0, 5, 0, 0, 0, 0], [1, 2, 3, 4, 0, 0, 0, 0])
    assert close(napa.rifft(both[:8] * both[8:]), [1, 2, 3, 9, 10, 15, 20, 0])
    data = np.full(8, 1 + 0j)
    assert close(napa.get_fft(8)(data), [8, 0, 0, 0, 0, 0, 0, 0])


def check_dst_rules(napa):
    """easy-interface.lisp:3-8, 44-67: aliasing / :size / dst rules and the error behaviour."""
    from napa_fft3_b200 import NapaFFTError
    import pytest
    x = rand_c(5, 24)
    out = napa.fft(x, size=16)
    assert out is not x and np.array_equal(out[16:], x[16:]) and nre(out[:16], np.fft.fft(x[:16])) < 1e-13
    y = x.copy()
    assert napa.fft(y, dst=y, size=16) is y and np.array_equal(y[16:], x[16:])
    d = np.zeros(32, complex)
    assert napa.fft(x, dst=d, size=16) is d and np.array_equal(d[16:24], x[16:24])
    lst = napa.fft([1, 2, 3, 4])
    assert lst.dtype == np.complex128 and len(lst) == 4
    with pytest.raises(NapaFFTError):
        napa.fft(np.zeros(12, complex))                     # not a power of two
    with pytest.raises(NapaFFTError):
        napa.fft(np.zeros(8, complex), size=16)             # vec shorter than size
    with pytest.raises(NapaFFTError):
        napa.fft(np.zeros(8, complex), scale="bogus")       # ecase on scaling
    with pytest.raises(TypeError):
        napa.fft(np.zeros(8, complex), window=np.zeros(8, np.float32))
    with pytest.raises(TypeError):
        napa.bit_reverse([1, 2, 3, 4])                      # etypecase: lists are not accepted


def check_windowed(napa, n=64):
    sig = rand_c(77, 5 * n)
    for center in (0, n, 5 * n - 3):
        got = napa.windowed_fft(sig, center, n)
        want = O.windowed_fft(sig, center, n)
        assert nre(got, want) <= tol(n)
    got = napa.windowed_ifft(sig[:n], window_fn=__import__("napa_fft3_b200").hann)
    want = O.windowed_ifft(sig[:n].copy(), window_fn="hann")
    assert nre(got, want) <= tol(n)
    r = rand_r(78, 5 * n)
    assert nre(napa.windowed_rfft(r, 2 * n, n), O.windowed_rfft(r, 2 * n, n)) <= tol(n)
    spec = O.rfft(rand_r(79, n))
    assert nre(napa.windowed_rifft(spec.copy()), O.windowed_rifft(spec.copy())) <= tol(n)


def check_get_windowed_fft(napa, n):
    import pytest
    """interface.lisp:191-217 get-windowed-fft: closures (vec window) for every direction / order / window type,
    default scales (nil forward, :inv inverse), against the oracle's fft / ifft with the same window"""
    wr, wc = rand_r(31, n, 0.5, 1.5), rand_c(32, n)
    for forward in (True, False):
        for in_order in (True, False):
            for wtype, w in (("real-sample", wr), ("complex-sample", wc)):
                fun = napa.get_windowed_fft(n, wtype, forward=forward, in_order=in_order)
                x = rand_c(33 + n, n)
                got = fun(x.copy(), w)
                if forward:
                    want = O.fft(x.copy(), in_order=in_order, window=w)
                else:
                    want = O.ifft(x.copy(), in_order=in_order, scale="inv", window=w)
                assert nre(got, want) <= tol(n), (n, forward, in_order, wtype)
    # explicit scale, and the result is the argument itself (in place)
    fun = napa.get_windowed_fft(n, "float", forward=True, scale="sqrt")
    x = rand_c(34, n)
    y = x.copy()
    assert fun(y, wr) is y and nre(y, O.fft(x.copy(), scale="sqrt", window=wr)) <= tol(n)
    with pytest.raises(Exception):
        napa.get_windowed_fft(n, None)


def check_window_functions(napa):
    import napa_fft3_b200 as P
    n = 64
    pairs = [(P.rectangular, "rectangular", 0), (P.hann, "hann", 0), (P.blackman, "blackman", 0),
             (P.triangle, "triangle", 0), (P.bartlett, "bartlett", 0), (P.blackman_harris, "blackman-harris", 0)]
    for fn, name, prm in pairs:
        assert np.array_equal(napa.window_vector(fn, n), O.window_vector(name, n, param=prm)), name
    g = P.gaussian(np.float32(0.5))
    assert np.array_equal(napa.window_vector(g, n), O.window_vector("gauss*", n, param=0.5))
    # blackman* with another alpha, the general cosine-series and gaussian*bartlett^x (windowing.lisp:12-20, 36-45)
    f32 = np.float32
    bs = lambda i, m: P.blackman_star(f32(0.08), i, m)
    assert np.array_equal(napa.window_vector(bs, n), O.window_vector("blackman*", n, param=0.08))
    for coeffs in ((0.3635819, 0.4891775, 0.1365995, 0.0106411),                  # Nuttall, double coefficients
                   (f32(0.21557895), f32(0.41663158), f32(0.277263158), f32(0.083578947))):   # flat top (4 terms), single
        cs = lambda i, m: P.cosine_series(i, m, *coeffs)
        want = np.array([O.cosine_series(i, n, *coeffs) for i in range(n)])
        assert np.array_equal(napa.window_vector(cs, n), want), coeffs
    for sigma, power in ((0.4, 2), (0.25, 3), (0.5, 1)):
        gb = P.gaussian_bartlett_x(f32(sigma), power)
        assert gb is P.gaussian_bartlett_x(f32(sigma), power)                       # same closure for the same parameters
        want = np.array([O.gaussian_bartlett_x(i, n, f32(sigma), power) for i in range(n)])
        assert np.array_equal(napa.window_vector(gb, n), want), (sigma, power)
    assert P.gaussian(f32(0.5)) is g
    assert napa.window_vector(P.hann, n) is napa.window_vector(P.hann, n)          # cached
    assert np.array_equal(napa.window_vector(P.hann, n, bit_reverse=True), O.window_vector("hann", n, bit_reverse=True))


def check_filter_chunks(napa, n, chunk):
    """example.lisp:81-114: overlap-add streaming filter against the oracle's statement-for-statement version
    (energies are summed in a different order on the device: agreement to rounding, not bits)."""
    x = rand_r(900 + n + chunk, n)
    # a smooth low-pass response over the bins, like the README's filters
    k = np.minimum(np.arange(chunk), chunk - np.arange(chunk)) / chunk
    filt = 1.0 / (1.0 + (8 * k) ** 4)
    got = napa.filter_chunks(x, filt)
    want = O.filter_chunks(x, filt)
    assert got.shape == want.shape and got.dtype == np.complex128
    assert nre(got, want) <= 2 * tol(chunk), (n, chunk, nre(got, want))
    return nre(got, want)


def check_typed_inputs(napa, n):
    """complex-samplify (support.lisp:64-93) on the device -- real-double, real-single and
    complex-single inputs give exactly the transform of their widened copy (widening is exact)."""
    xr = rand_r(700 + n, n)
    xc = rand_c(701 + n, n)
    w = rand_r(702 + n, n, 1.0, 2.0)
    for v in (xr, xr.astype(np.float32), xc.astype(np.complex64)):
        wide = v.astype(np.complex128)
        for kw in ({}, {"in_order": False, "scale": "sqrt"}, {"window": w}):
            assert np.array_equal(napa.fft(v, **kw), napa.fft(wide.copy(), **kw)), (n, v.dtype, kw)
            assert nre(napa.fft(v, **kw), O.fft(wide.copy(), **kw)) <= tol(n)
        assert np.array_equal(napa.ifft(v, in_order=False), napa.ifft(wide.copy(), in_order=False))
        d = np.empty(n + 3, dtype=np.complex128)
        assert napa.fft(v, dst=d) is d and np.array_equal(d[:n], napa.fft(wide.copy()))


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_vectors.npz")


def check_golden(napa, exact=False):
    """the committed fixtures of tests/golden/make_golden.py (outputs of the pinned oracle).  exact=True is
    for the oracle itself (must reproduce its own fixtures bit for bit); the CUDA path is held to the FP64
    tolerance, bit-exact for the permutations."""
    g = np.load(GOLDEN)
    sizes = sorted({int(k.split("_")[0][1:]) for k in g.files if k.startswith("n")})
    cmp = (lambda a, b, n: np.array_equal(a, b)) if exact else (lambda a, b, n: nre(a, b) <= tol(n))
    for n in sizes:
        k = "n%d_" % n
        xc, xr, wr = g[k + "xc"], g[k + "xr"], g[k + "wr"]
        assert cmp(napa.fft(xc.copy()), g[k + "fft"], n), (n, "fft")
        assert cmp(napa.fft(xc.copy(), in_order=False, scale="sqrt", window=wr), g[k + "fft_rev_sqrt_wr"], n), (n, "fft rev")
        if k + "wc" not in g.files:
            continue
        wc = g[k + "wc"]
        assert cmp(napa.ifft(xc.copy(), in_order=False, scale=True, window=wc), g[k + "ifft_rev_wc"], n), (n, "ifft rev")
        assert cmp(napa.ifft(xc.copy(), in_order=True, scale="sqrt", window=wr), g[k + "ifft_nat_wr"], n), (n, "ifft nat")
        assert np.array_equal(napa.bit_reverse(xc.copy()), g[k + "bitrev_c"]), (n, "bit-reverse complex")
        assert np.array_equal(napa.bit_reverse(xr.copy()), g[k + "bitrev_r"]), (n, "bit-reverse real")
        assert cmp(napa.rfft(xr.copy()), g[k + "rfft"], n), (n, "rfft")
        assert cmp(napa.rifft(napa.rfft(xr.copy()), scale=True), g[k + "rifft"], n), (n, "rifft")
        assert cmp(napa.two_rfft(xr.copy(), wr.copy()), g[k + "2rfft"], n), (n, "%2rfft")
    got = napa.filter_chunks(g["oa_signal"], g["oa_filter"])
    assert (np.array_equal(got, g["oa_out"]) if exact else nre(got, g["oa_out"]) <= 2 * tol(64)), "filter-chunks"
