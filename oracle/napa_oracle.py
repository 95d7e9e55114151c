"""ctypes front-end of the CPU ORACLE (oracle/napa_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package napa_fft3_b200.

The functions mirror the reference's easy interface (easy-interface.lisp:32-95, real.lisp,
windowing.lisp) with numpy arrays standing in for Lisp vectors, so that parity tests read like
the reference's REPL examples.  Parity status: pinned bit-for-bit against the README known
answers (see tests/test_oracle_kat.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnapa_oracle.so")

SCALE = {None: 0, 1: 0, False: 0, "nil": 0, True: 1, "inv": 1, ":inv": 1, "t": 1, "sqrt": 2, ":sqrt": 2}
WINDOW_IDS = {"rectangular": 0, "hann": 1, "blackman": 2, "triangle": 3, "bartlett": 4,
              "blackman-harris": 5, "blackman*": 6, "gauss*": 7}


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "napa_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libnapa_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, sz, i, d, vp = C.POINTER(C.c_double), C.c_size_t, C.c_int, C.c_double, C.c_void_p
        L.napa_oracle_bit_reverse_integer.restype = C.c_uint64
        L.napa_oracle_bit_reverse_integer.argtypes = [C.c_uint64, i]
        L.napa_oracle_make_twiddle.argtypes = [sz, d, vp]
        L.napa_oracle_scale_factor.restype = d
        L.napa_oracle_scale_factor.argtypes = [sz, i, i]
        L.napa_oracle_lowlevel.argtypes = [vp, sz, sz, i, d, vp, i, sz]
        L.napa_oracle_bit_reverse_inplace.argtypes = [vp, sz, sz, i]
        L.napa_oracle_fft.argtypes = [vp, sz, i, i, vp, i]
        L.napa_oracle_ifft.argtypes = [vp, sz, i, i, vp, i]
        L.napa_oracle_radix2_twiddle.argtypes = [sz, d, vp]
        L.napa_oracle_swizzle.argtypes = [vp, sz, i]
        L.napa_oracle_rfft.argtypes = [vp, vp, sz, i]
        L.napa_oracle_rifft.argtypes = [vp, vp, sz, i]
        L.napa_oracle_2rfft.argtypes = [vp, vp, vp, sz, i]
        L.napa_oracle_window_fn.restype = d
        L.napa_oracle_window_fn.argtypes = [i, C.c_long, C.c_long, d]
        L.napa_oracle_window_vector.argtypes = [i, sz, d, i, vp]
        L.napa_oracle_extract_centered_window.argtypes = [vp, sz, C.c_long, sz, vp, i]
        L.napa_oracle_c2c_batch.argtypes = [vp, sz, sz, i, i, i, vp, i]
        L.napa_oracle_warm.argtypes = [sz]
        L.napa_oracle_warm.restype = None
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _is_pow2(n):
    return n >= 1 and (n & (n - 1)) == 0


def complex_samplify(vec, size=None):
    """support.lisp:64-93: identity on complex128 arrays (aliasing!), else a fresh array of `size`
    elements of which only min(size, len(vec)) are filled."""
    if isinstance(vec, np.ndarray) and vec.dtype == np.complex128 and vec.ndim == 1 and vec.flags.c_contiguous:
        return vec
    src = np.asarray(vec)
    size = len(src) if size is None else size
    out = np.zeros(size, dtype=np.complex128)
    m = min(size, len(src))
    out[:m] = src[:m].astype(np.complex128)
    return out


def real_samplify(vec, size=None):
    """support.lisp:95-112"""
    if isinstance(vec, np.ndarray) and vec.dtype == np.float64 and vec.ndim == 1 and vec.flags.c_contiguous:
        return vec
    src = np.asarray(vec)
    size = len(src) if size is None else size
    out = np.zeros(size, dtype=np.float64)
    m = min(size, len(src))
    out[:m] = src[:m].astype(np.float64)
    return out


def _window_args(window):
    if window is None:
        return None, 0
    w = np.ascontiguousarray(window)
    if w.dtype == np.complex128:
        return w, 2
    if w.dtype == np.float64:
        return w, 1
    raise TypeError("window must be float64 or complex128 (easy-interface.lisp:39-42)")


def _resolve_dst(old, vec, dst):
    """easy-interface.lisp:3-8, 51-53: copy-or-replace."""
    if old is vec or dst is not None:
        if dst is vec:
            return dst
        if dst is not None:
            m = min(len(dst), len(vec))
            dst[:m] = vec[:m]
            return dst
        return vec.copy()
    return vec


def fft(vec, dst=None, size=None, in_order=True, scale=None, window=None):
    """easy-interface.lisp:44-67"""
    n = size if size is not None else len(vec)
    old = vec
    vec = complex_samplify(vec, n)
    dst = _resolve_dst(old, vec, dst)
    assert _is_pow2(n) and len(vec) >= n and len(dst) >= n
    w, wt = _window_args(window)
    rc = lib().napa_oracle_fft(_ptr(dst), n, int(bool(in_order)), SCALE[scale], _ptr(w), wt)
    assert rc == 0
    return dst


def ifft(vec, dst=None, size=None, in_order=True, scale=True, window=None):
    """easy-interface.lisp:69-95"""
    n = size if size is not None else len(vec)
    old = vec
    vec = complex_samplify(vec, n)
    dst = _resolve_dst(old, vec, dst)
    assert _is_pow2(n) and len(vec) >= n and len(dst) >= n
    w, wt = _window_args(window)
    rc = lib().napa_oracle_ifft(_ptr(dst), n, int(bool(in_order)), SCALE[scale], _ptr(w), wt)
    assert rc == 0
    return dst


def bit_reverse(vec, dst=None, size=None):
    """easy-interface.lisp:10-37 (complex or real arrays)"""
    assert isinstance(vec, np.ndarray) and vec.dtype in (np.complex128, np.float64)
    size = len(vec) if size is None else size
    assert len(vec) >= size
    if dst is None:
        dst = vec.copy()
    elif dst is not vec:
        m = min(len(dst), len(vec))
        dst[:m] = vec[:m]
    assert len(dst) >= size
    rc = lib().napa_oracle_bit_reverse_inplace(_ptr(dst), 0, size, 0 if vec.dtype == np.complex128 else 1)
    assert rc == 0
    return dst


def lowlevel(vec, start, n, direction, scale_factor=1.0, window=None, window_start=0):
    """interface.lisp:55-74: the generated closure (vec start [window window-start] twiddle)."""
    w, wt = _window_args(window)
    rc = lib().napa_oracle_lowlevel(_ptr(vec), start, n, direction, float(scale_factor), _ptr(w), wt, window_start)
    assert rc == 0
    return vec


def scale_factor(n, forward, scale):
    return lib().napa_oracle_scale_factor(n, int(bool(forward)), SCALE[scale])


def make_twiddle(n, direction=1.0):
    out = np.zeros(n, dtype=np.complex128)
    assert lib().napa_oracle_make_twiddle(n, float(direction), _ptr(out)) == 0
    return out


def radix2_twiddle(n, direction):
    out = np.zeros(n // 2, dtype=np.complex128)
    lib().napa_oracle_radix2_twiddle(n, float(direction), _ptr(out))
    return out


def rfft(vec, dst=None, size=None, scale=None):
    """real.lisp:101-135"""
    size = len(vec) if size is None else size
    v = real_samplify(vec, size)
    assert len(v) >= size and _is_pow2(size)
    if dst is None:
        dst = np.zeros(size, dtype=np.complex128)
    assert len(dst) >= size
    assert lib().napa_oracle_rfft(_ptr(v), _ptr(dst), size, SCALE[scale]) == 0
    return dst


def rifft(vec, dst=None, size=None, scale=True):
    """real.lisp:153-185 (vec is destroyed when it is already a complex128 array)"""
    size = len(vec) if size is None else size
    v = complex_samplify(vec, size)
    assert len(v) >= size and _is_pow2(size)
    if dst is None:
        dst = np.zeros(size, dtype=np.float64)
    assert len(dst) >= size
    assert lib().napa_oracle_rifft(_ptr(v), _ptr(dst), size, SCALE[scale]) == 0
    return dst


def two_rfft(v1, v2, dst=None, size=None, scale=None):
    """real.lisp:33-48 (%2rfft)"""
    n = size if size is not None else min(len(v1), len(v2))
    a, b = real_samplify(v1, n), real_samplify(v2, n)
    if dst is None:
        dst = np.zeros(2 * n, dtype=np.complex128)
    assert len(dst) == 2 * n, "oracle models the documented case (length dst) = 2n"
    assert lib().napa_oracle_2rfft(_ptr(np.ascontiguousarray(a[:n])), _ptr(np.ascontiguousarray(b[:n])),
                                   _ptr(dst), n, SCALE[scale]) == 0
    return dst


def window_fn(name, i, n, param=0.0):
    return lib().napa_oracle_window_fn(WINDOW_IDS[name], i, n, float(param))


def cosine_series(i, n, a0, a1, a2, a3):
    """windowing.lisp:43-45"""
    f = lib().napa_oracle_cosine_series
    f.restype = C.c_double
    f.argtypes = [C.c_long, C.c_long, C.c_double, C.c_double, C.c_double, C.c_double]
    return f(i, n, float(a0), float(a1), float(a2), float(a3))


def gaussian_bartlett_x(i, n, sigma, power):
    """windowing.lisp:36-41, single-float sigma, integer exponent"""
    f = lib().napa_oracle_gaussian_bartlett_x
    f.restype = C.c_double
    f.argtypes = [C.c_long, C.c_long, C.c_double, C.c_long]
    return f(i, n, float(sigma), int(power))


def window_vector(name, n, bit_reverse=False, param=0.0):
    """windowing.lisp:53-65"""
    out = np.zeros(n, dtype=np.float64)
    assert lib().napa_oracle_window_vector(WINDOW_IDS[name], n, float(param), int(bool(bit_reverse)), _ptr(out)) == 0
    return out


def extract_centered_window(signal, center, size):
    """windowing.lisp:85-111"""
    signal = np.ascontiguousarray(signal)
    elt = 0 if signal.dtype == np.complex128 else 1
    out = np.zeros(size, dtype=signal.dtype)
    lib().napa_oracle_extract_centered_window(_ptr(signal), len(signal), int(center), size, _ptr(out), elt)
    return out


def windowed_fft(signal, center, length, window_fn="hann", dst=None, in_order=True, scale=None):
    """windowing.lisp:113-130"""
    assert _is_pow2(length)
    sig = complex_samplify(signal)
    chunk = extract_centered_window(sig, center, length)
    w = window_vector(window_fn, length)
    return fft(chunk, dst=dst, in_order=in_order, scale=scale, window=w)


def windowed_ifft(signal, window_fn="rectangular", size=None, dst=None, in_order=True, scale=True):
    """windowing.lisp:132-148 (window is NOT bit-reversed)"""
    v = complex_samplify(signal)
    n = size if size is not None else len(v)
    w = window_vector(window_fn, n)
    return ifft(v, dst=dst, size=n, in_order=in_order, scale=scale, window=w)


def windowed_rfft(signal, center, length, window_fn="hann", dst=None, scale=None):
    """real.lisp:137-151"""
    sig = real_samplify(signal)
    chunk = extract_centered_window(sig, center, length)
    chunk = chunk * window_vector(window_fn, length)
    return rfft(chunk, dst=dst, scale=scale)


def windowed_rifft(vec, window_fn="rectangular", dst=None, size=None, scale=True):
    """real.lisp:187-198: multiplies the spectrum in place and hard-codes :scale t"""
    size = len(vec) if size is None else size
    v = complex_samplify(vec, size)
    w = window_vector(window_fn, size)
    v[:size] *= w
    return rifft(v, dst=dst, size=size, scale=True)


def c2c_batch(data, n, direction, in_order=True, scale=None, window=None):
    """In-place batch driver used for the timed CPU baseline (rows of `data` are transforms)."""
    assert data.dtype == np.complex128 and data.flags.c_contiguous
    batch = data.size // n
    w, wt = _window_args(window)
    rc = lib().napa_oracle_c2c_batch(_ptr(data), n, batch, direction, int(bool(in_order)), SCALE[scale], _ptr(w), wt)
    assert rc == 0
    return data


def filter_chunks(vector, filter):
    """example.lisp:69-114 (energy, filter-chunks), statement for statement on top of the oracle's
    bit_reverse / windowed_fft / ifft: the checker for napafft_filter_chunks."""
    vector = np.asarray(vector, dtype=np.float64)
    filt = np.asarray(filter, dtype=np.float64)
    n = len(vector)
    destination = np.zeros(n, dtype=np.complex128)
    chunk_size = len(filt)
    half_size = chunk_size // 2
    filt_rev = bit_reverse(filt.copy())

    def energy(x):                       # example.lisp:69-79: sequential sum of re^2 + im^2
        acc = 0.0
        for v in x:
            acc += v.real * v.real + v.imag * v.imag
        return acc
    for i in range(0, n, half_size):
        end = min(n, i + chunk_size)
        chunk = np.zeros(chunk_size, dtype=np.complex128)
        chunk[:end - i] = vector[i:end]
        old_energy = energy(chunk)
        chunk = windowed_fft(chunk, half_size, chunk_size, window_fn="triangle", dst=chunk, in_order=False)
        chunk = ifft(chunk, dst=chunk, in_order=False, window=filt_rev)
        new_energy = energy(chunk)
        scale = 0.5 * np.sqrt(old_energy / new_energy)
        destination[i:end] += scale * chunk[:end - i]
    return destination
