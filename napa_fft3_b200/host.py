"""Host-side mirror of the reference's public interface, on top of the C ABI.

SBCL is not available in this image, so the keyword front-ends that would be Common Lisp
(lisp/ holds the sb-alien version) are restated here in Python with numpy arrays standing in for
`(simple-array (complex double-float) (*))` / `(simple-array double-float (*))`.  Names, argument
meaning, defaults, aliasing rules and error behaviour follow the reference:

  easy-interface.lisp:32-95   bit-reverse, fft, ifft
  interface.lisp:12-217       %ensure-fft, %ensure-reverse, %ensure-twiddles, get-fft,
                              get-windowed-fft, get-reverse
  windowing.lisp:6-148        window functions, window-vector, extract-*, windowed-fft/-ifft
  real.lisp:33-198            %2rfft, rfft, rifft, windowed-rfft, windowed-rifft
  support.lisp:64-112         complex-samplify, real-samplify

All arithmetic of the transforms runs in the CUDA library; nothing here computes a DFT.
"""
from __future__ import annotations

import ctypes as C
import math
import threading

import numpy as np

from . import _lib
from ._lib import Binding, NapaFFTError

_SCALE = {None: 0, 1: 0, "nil": 0, True: 1, "t": 1, "inv": 1, ":inv": 1, "sqrt": 2, ":sqrt": 2}


def _scale_code(scale):
    """interface.lisp:6-7 (deftype scaling) / ecase of interface.lisp:41-48"""
    try:
        if scale is False:
            raise KeyError
        return _SCALE[scale]
    except (KeyError, TypeError):
        raise NapaFFTError(4, f"{scale!r} is not one of NIL, 1, T, :INV, SQRT, :SQRT") from None


def _ptr(a, offset_elems=0):
    return C.c_void_p(a.ctypes.data + offset_elems * a.itemsize)


def power_of_two_p(x):  # support.lisp:61-62
    return isinstance(x, (int, np.integer)) and x >= 1 and (x & (x - 1)) == 0


def lb(n):  # support.lisp:58-59
    return int(n - 1).bit_length()


def _is_complex_sample_array(v):
    return isinstance(v, np.ndarray) and v.dtype == np.complex128 and v.ndim == 1 and v.flags.c_contiguous


def _is_real_sample_array(v):
    return isinstance(v, np.ndarray) and v.dtype == np.float64 and v.ndim == 1 and v.flags.c_contiguous


def complex_samplify(vec, size=None):
    """support.lisp:64-93: identity (NO copy) on complex-sample arrays; otherwise a fresh array of
    `size` elements of which min(size, length) are filled (map-into)."""
    if _is_complex_sample_array(vec):
        return vec
    src = np.asarray(vec)
    if src.ndim != 1:
        raise TypeError("complex-samplify: not a sequence")
    size = len(src) if size is None else size
    out = np.zeros(size, dtype=np.complex128)
    m = min(size, len(src))
    out[:m] = src[:m]
    return out


def real_samplify(vec, size=None):
    """support.lisp:95-112"""
    if _is_real_sample_array(vec):
        return vec
    src = np.asarray(vec)
    if src.ndim != 1 or np.iscomplexobj(src):
        raise TypeError("real-samplify: not a sequence of reals")
    size = len(src) if size is None else size
    out = np.zeros(size, dtype=np.float64)
    m = min(size, len(src))
    out[:m] = src[:m]
    return out


# ---- window functions (windowing.lisp:6-51) with the reference's float contagion ------------
_f32 = np.float32
_PI = math.pi


def rectangular(i, n):
    return _f32(1.0)


def hann(i, n):
    return 0.5 * (1.0 - math.cos((2 * _PI * i) / (n - 1)))


def blackman_star(alpha, i, n):
    a0 = (1 - alpha) / 2
    a1 = _f32(0.5)
    a2 = alpha / 2
    return (a0 + (-(a1 * np.float64(math.cos((2 * _PI * i) / (n - 1)))))) + a2 * np.float64(math.cos((4 * _PI * i) / (n - 1)))


def blackman(i, n):
    return blackman_star(_f32(0.16), i, n)


def triangle(i, n):
    return (_f32(2) / _f32(n)) * (_f32(n) * _f32(0.5) - abs(_f32(i) - _f32(0.5) * _f32(n - 1)))


def bartlett(i, n):
    return (_f32(2) / _f32(n - 1)) * (_f32(n - 1) * _f32(0.5) - abs(_f32(i) - _f32(0.5) * _f32(n - 1)))


def gauss_star(sigma, i, n):
    m = _f32(0.5) * _f32(n - 1)
    r = (_f32(i) - m) / (sigma * m)
    e = _f32(-0.5) * (r * r)
    if isinstance(e, np.float32):
        return _f32(math.exp(float(e)))       # SBCL: (coerce (%exp (coerce x 'double-float)) 'single-float)
    return np.float64(math.exp(float(e)))


_gaussian_cache, _gb_cache = {}, {}


def gaussian(sigma):
    if sigma not in _gaussian_cache:
        _gaussian_cache[sigma] = lambda i, n: gauss_star(sigma, i, n)
    return _gaussian_cache[sigma]


def _expt(base, power):
    """CL EXPT as SBCL evaluates it for a float base: an integer power goes through INTEXP (repeated squaring in the
    base's own precision, one rounding per multiplication -- not a correctly rounded pow); a float power is computed
    in double precision and rounded to single when both arguments are single."""
    if isinstance(power, (int, np.integer)) and not isinstance(power, bool):
        p = abs(int(power))
        total = base if p & 1 else type(base)(1)
        p >>= 1
        while p:
            base = base * base
            if p & 1:
                total = base * total
            p >>= 1
        return total if power >= 0 else type(base)(1) / total
    r = math.pow(float(base), float(power)) if base >= 0 else complex(float(base)) ** float(power)
    if isinstance(base, np.float32) and isinstance(power, np.float32) and not isinstance(r, complex):
        return _f32(r)
    return r


def gaussian_bartlett_x(sigma, triangle_exponent):
    key = (sigma, triangle_exponent)
    if key not in _gb_cache:
        _gb_cache[key] = lambda i, n: np.real(_expt(bartlett(i, n), triangle_exponent)) * gauss_star(sigma, i, n)
    return _gb_cache[key]


def cosine_series(i, n, a0, a1, a2, a3):
    def f(scale, x):
        return scale * np.float64(math.cos((x * _PI * i) / (n - 1)))
    return ((a0 + (-f(a1, 2))) + f(a2, 4)) + (-f(a3, 6))


def blackman_harris(i, n):
    return cosine_series(i, n, _f32(0.35875), _f32(0.48829), _f32(0.14128), _f32(0.01168))


def clip_in_window(x, start, end):
    return max(start, min(x, end))


def extract_window_into(vector, start, length, destination):
    """windowing.lisp:69-83: VECTOR is zero outside its legal indices."""
    assert length <= len(destination)
    s = clip_in_window(start, 0, len(vector))
    e = clip_in_window(start + length, 0, len(vector))
    if length != e - s:
        destination[:] = 0
    off = s - start
    if -1 < off < len(destination):
        destination[off:off + (e - s)] = vector[s:e]
    return destination


def extract_window(vector, start, length, element_type=None):
    vector = np.asarray(vector)
    return extract_window_into(vector, start, length, np.zeros(length, dtype=element_type or vector.dtype))


def extract_centered_window_into(vector, center, size, destination):
    return extract_window_into(vector, center - size // 2, size, destination)


def extract_centered_window(vector, center, size, element_type=None):
    vector = np.asarray(vector)
    return extract_centered_window_into(vector, center, size, np.zeros(size, dtype=element_type or vector.dtype))


class TwiddleHandle:
    """What %ensure-twiddles returns here: the twiddle tables live on the device inside the
    library, so the value handed back to (and accepted from) callers is an opaque token."""

    def __init__(self, n, forwardp):
        self.n, self.forwardp = n, bool(forwardp)

    def __len__(self):
        return self.n


class NapaFFT:
    """The reference's public API bound to one library instance."""

    def __init__(self, binding: Binding | None = None):
        self._b = binding
        self._window_cache = {}
        self._lock = threading.Lock()

    @property
    def b(self) -> Binding:
        if self._b is None:
            self._b = _lib.default_binding()
        return self._b

    # ---- easy-interface.lisp -----------------------------------------------------------
    @staticmethod
    def _window_args(window, n):
        if window is None:
            return None, 0, None
        if _is_complex_sample_array(window):
            wt = _lib.WINDOW_COMPLEX
        elif _is_real_sample_array(window):
            wt = _lib.WINDOW_REAL
        else:  # get-window-type etypecase, easy-interface.lisp:39-42
            raise TypeError("window must be a complex-sample-array or a real-sample-array")
        if len(window) < n:
            raise NapaFFTError(2, f"window of length {len(window)} shorter than the transform ({n})")
        return _ptr(window), wt, window

    @staticmethod
    def _typed_input(vec, n):
        """(array, NAPA_SAMPLE_*) when `vec` is a dense numpy vector of exactly n real-double, complex-single or
        real-single samples -- the inputs complex-samplify (support.lisp:64-93) would widen on the host."""
        if not isinstance(vec, np.ndarray) or vec.ndim != 1 or len(vec) != n or not vec.flags.c_contiguous:
            return None
        code = {np.dtype(np.float64): _lib.SAMPLE_F64, np.dtype(np.complex64): _lib.SAMPLE_C32,
                np.dtype(np.float32): _lib.SAMPLE_F32}.get(vec.dtype)
        return None if code is None else (vec, code)

    def _c2c(self, vec, dst, size, in_order, scale, window, direction):
        n = size if size is not None else len(vec)
        if dst is not None and not _is_complex_sample_array(dst):
            raise TypeError("dst must be a complex-sample-array or NIL")
        old = vec
        fast = self._typed_input(vec, n)
        if fast is not None and power_of_two_p(n):
            # complex-samplify on the device: the raw samples cross PCIe (8 or 4 bytes instead of 16)
            raw, code = fast
            if dst is None:
                dst = np.empty(n, dtype=np.complex128)     # the coerced copy of the reference, transformed in place
            elif len(dst) < n:
                raise NapaFFTError(2, f"destination of length {len(dst)} shorter than size {n}")
            wptr, wt, _keep = self._window_args(window, n)
            self.b.c2c_typed(_ptr(raw), code, _ptr(dst), n, 1, direction, _scale_code(scale), int(bool(in_order)), wptr, wt, 0)
            return dst
        vec = complex_samplify(vec, n)
        if not power_of_two_p(n):
            raise NapaFFTError(1, f"transform size {n} is not a power of two")
        if len(vec) < n:
            raise NapaFFTError(2, f"vector of length {len(vec)} shorter than size {n}")
        # copy-or-replace (easy-interface.lisp:3-8, 51-53): same object -> in place; dst given ->
        # (replace dst src); complex input without dst -> fresh copy; coerced input -> itself.
        if old is vec or dst is not None:
            if dst is None:
                dst = np.empty_like(vec)
                tail_from = n
            elif dst is vec:
                tail_from = None
            else:
                tail_from = n
        else:
            dst, tail_from = vec, None
        if len(dst) < n:
            raise NapaFFTError(2, f"destination of length {len(dst)} shorter than size {n}")
        wptr, wt, _keep = self._window_args(window, n)
        sc = _scale_code(scale)
        self.b.c2c(_ptr(vec), _ptr(dst), n, 1, n, direction, sc, int(bool(in_order)), wptr, wt, 0)
        if tail_from is not None:
            m = min(len(dst), len(vec))
            if m > tail_from:
                dst[tail_from:m] = vec[tail_from:m]
        return dst

    def fft(self, vec, dst=None, size=None, in_order=True, scale=None, window=None):
        """easy-interface.lisp:44-67"""
        return self._c2c(vec, dst, size, in_order, scale, window, _lib.FWD)

    def ifft(self, vec, dst=None, size=None, in_order=True, scale=True, window=None):
        """easy-interface.lisp:69-95"""
        return self._c2c(vec, dst, size, in_order, scale, window, _lib.INV)

    def bit_reverse(self, vec, dst=None, size=None):
        """easy-interface.lisp:10-37"""
        if _is_complex_sample_array(vec):
            elt, ok = _lib.ELT_COMPLEX, _is_complex_sample_array
        elif _is_real_sample_array(vec):
            elt, ok = _lib.ELT_REAL, _is_real_sample_array
        else:
            raise TypeError("bit-reverse: vec must be a complex-sample-array or a real-sample-array")
        size = len(vec) if size is None else size
        if dst is not None and not ok(dst):
            raise TypeError("bit-reverse: dst must have the element type of vec")
        if len(vec) < size:
            raise NapaFFTError(2, f"vector of length {len(vec)} shorter than size {size}")
        if not power_of_two_p(size):
            raise NapaFFTError(1, f"size {size} is not a power of two")
        if dst is None:
            dst = np.empty_like(vec)
        if len(dst) < size:
            raise NapaFFTError(2, f"destination of length {len(dst)} shorter than size {size}")
        self.b.bit_reverse(_ptr(vec), _ptr(dst), size, 1, size, elt)
        if dst is not vec:
            m = min(len(dst), len(vec))
            if m > size:
                dst[size:m] = vec[size:m]
        return dst

    # ---- interface.lisp (low-level generators) -----------------------------------------
    @staticmethod
    def _direction(direction):
        if direction in (1, "fwd", ":fwd"):
            return _lib.FWD
        if direction in (-1, "inv", ":inv", "bwd", ":bwd"):
            return _lib.INV
        raise NapaFFTError(4, f"{direction!r} is not one of 1, -1, :FWD, :INV, :BWD")

    @staticmethod
    def _windowing(windowing):
        if windowing is None:
            return _lib.WINDOW_NONE
        if windowing in ("float", "real-sample", float, np.float64):
            return _lib.WINDOW_REAL
        if windowing in ("complex", "complex-sample", complex, np.complex128):
            return _lib.WINDOW_COMPLEX
        raise NapaFFTError(4, f"{windowing!r} is not one of NIL, FLOAT, REAL-SAMPLE, COMPLEX, COMPLEX-SAMPLE")

    def ensure_twiddles(self, n, forwardp):
        """interface.lisp:139-163"""
        if not power_of_two_p(n):
            raise NapaFFTError(1, f"{n} is not a power of two")
        return TwiddleHandle(n, forwardp)

    def ensure_fft(self, direction, scaling, windowing, n):
        """interface.lisp:81-97: returns (lambda (vec start [window window-start] twiddle)), which
        transforms vec[start, start+n) in place: DIF forward leaves it bit-reversed, DIT inverse
        expects it bit-reversed (forward.lisp:54-117, inverse.lisp:47-112)."""
        d, sc, wt = self._direction(direction), _scale_code(scaling), self._windowing(windowing)
        if not power_of_two_p(n):
            raise NapaFFTError(1, f"{n} is not a power of two")
        if lb(n) > 32:
            raise NapaFFTError(7, "n exceeds 2^32 (interface.lisp:86)")
        b = self.b
        if wt == _lib.WINDOW_NONE:
            def fun(vec, start, twiddle=None):
                b.c2c(_ptr(vec, start), _ptr(vec, start), n, 1, n, d, sc, 0, None, 0, 0)
                return vec
        else:
            def fun(vec, start, window, window_start, twiddle=None):
                b.c2c(_ptr(vec, start), _ptr(vec, start), n, 1, n, d, sc, 0, _ptr(window, window_start), wt, 0)
                return vec
        return fun

    def ensure_reverse(self, n, eltype="complex-sample"):
        """interface.lisp:103-128: (lambda (vec start))"""
        if eltype not in ("complex-sample", "real-sample"):
            raise NapaFFTError(4, f"{eltype!r} is not COMPLEX-SAMPLE or REAL-SAMPLE")
        if not power_of_two_p(n):
            raise NapaFFTError(1, f"{n} is not a power of two")
        elt = _lib.ELT_COMPLEX if eltype == "complex-sample" else _lib.ELT_REAL
        b = self.b

        def fun(vec, start):
            b.bit_reverse(_ptr(vec, start), _ptr(vec, start), n, 1, n, elt)
            return vec
        return fun

    def get_reverse(self, n, eltype="complex-sample"):
        """interface.lisp:130-133"""
        fun = self.ensure_reverse(n, eltype)
        return lambda vec: fun(vec, 0)

    _UNSET = object()

    def get_fft(self, size, forward=True, scale=_UNSET, in_order=True):
        """interface.lisp:165-189: in-place one-argument closure; the fused device path does the
        transform and the reordering in one call (forward: DIF then reversal; inverse: reversal
        then DIT)."""
        if scale is self._UNSET:
            scale = None if forward else "inv"
        d, sc = (_lib.FWD if forward else _lib.INV), _scale_code(scale)
        if not power_of_two_p(size):
            raise NapaFFTError(1, f"{size} is not a power of two")
        b, io = self.b, int(bool(in_order))

        def fun(vec):
            b.c2c(_ptr(vec), _ptr(vec), size, 1, size, d, sc, io, None, 0, 0)
            return vec
        return fun

    def get_windowed_fft(self, size, window_type, forward=True, scale=_UNSET, in_order=True):
        """interface.lisp:191-217"""
        if window_type is None:
            raise NapaFFTError(4, "window-type must not be NIL (interface.lisp:196)")
        if scale is self._UNSET:
            scale = None if forward else "inv"
        d, sc, wt = (_lib.FWD if forward else _lib.INV), _scale_code(scale), self._windowing(window_type)
        if not power_of_two_p(size):
            raise NapaFFTError(1, f"{size} is not a power of two")
        b, io = self.b, int(bool(in_order))

        def fun(vec, window):
            # for the in-order inverse the reference reverses vec first and then applies window[p]
            # at memory position p; the library indexes the window by rev(k) to the same effect
            b.c2c(_ptr(vec), _ptr(vec), size, 1, size, d, sc, io, _ptr(window), wt, 0)
            return vec
        return fun

    # ---- windowing.lisp -----------------------------------------------------------------
    def window_vector(self, function, n, bit_reverse=False):
        """windowing.lisp:53-65 (cached, and the cached vector itself is returned)"""
        if bit_reverse and not power_of_two_p(n):
            raise NapaFFTError(1, f"{n} is not a power of two")
        key = (function, n, bool(bit_reverse))
        with self._lock:
            v = self._window_cache.get(key)
        if v is None:
            v = np.empty(n, dtype=np.float64)
            for i in range(n):
                v[i] = float(function(i, n))
            if bit_reverse:
                self.bit_reverse(v, v)
            with self._lock:
                self._window_cache[key] = v
        return v

    def windowed_fft(self, signal_vector, center, length, window_fn=hann, dst=None, in_order=True, scale=None):
        """windowing.lisp:113-130"""
        if not power_of_two_p(length):
            raise NapaFFTError(1, f"FFT size {length} is not a power of two")
        signal_vector = complex_samplify(signal_vector)
        input_window = extract_centered_window(signal_vector, center, length)
        window = self.window_vector(window_fn, length)
        return self.fft(input_window, dst=dst, in_order=in_order, scale=scale, window=window)

    def windowed_ifft(self, signal_vector, window_fn=rectangular, size=None, dst=None, in_order=True, scale=True):
        """windowing.lisp:132-148 (the window is NOT bit-reversed, as in the reference)"""
        vector = complex_samplify(signal_vector)
        length = size if size is not None else len(vector)
        if not power_of_two_p(length):
            raise NapaFFTError(1, f"{length} is not a power of two")
        window = self.window_vector(window_fn, length)
        return self.ifft(vector, dst=dst, size=length, in_order=in_order, scale=scale, window=window)

    # ---- real.lisp ----------------------------------------------------------------------
    def rfft(self, vec, dst=None, size=None, scale=None):
        """real.lisp:101-135: n reals -> n complex bins (full spectrum, natural order)"""
        size = len(vec) if size is None else size
        sc = _scale_code(scale)
        vec = real_samplify(vec, size)
        if len(vec) < size:
            raise NapaFFTError(2, f"vector of length {len(vec)} shorter than size {size}")
        if not power_of_two_p(size):
            raise NapaFFTError(1, f"size {size} is not a power of two")
        if dst is None:
            dst = np.zeros(size, dtype=np.complex128)
        elif not _is_complex_sample_array(dst):
            raise TypeError("dst must be a complex-sample-array")
        if len(dst) < size:
            raise NapaFFTError(2, f"destination of length {len(dst)} shorter than size {size}")
        self.b.rfft(_ptr(vec), _ptr(dst), size, 1, sc)
        return dst

    def rifft(self, vec, dst=None, size=None, scale=True):
        """real.lisp:153-185.  The reference destroys vec; this implementation leaves it intact."""
        size = len(vec) if size is None else size
        sc = _scale_code(scale)
        vec = complex_samplify(vec, size)
        if len(vec) < size:
            raise NapaFFTError(2, f"vector of length {len(vec)} shorter than size {size}")
        if not power_of_two_p(size):
            raise NapaFFTError(1, f"size {size} is not a power of two")
        if dst is None:
            dst = np.zeros(size, dtype=np.float64)
        elif not _is_real_sample_array(dst):
            raise TypeError("dst must be a real-sample-array")
        if len(dst) < size:
            raise NapaFFTError(2, f"destination of length {len(dst)} shorter than size {size}")
        self.b.rifft(_ptr(vec), _ptr(dst), size, 1, sc)
        return dst

    def two_rfft(self, v1, v2, dst=None, size=None, scale=None):
        """real.lisp:33-48 (%2rfft): dst[0,n) = FFT(v1), dst[n,2n) = FFT(v2)"""
        n = size if size is not None else min(len(v1), len(v2))
        sc = _scale_code(scale)
        v1, v2 = real_samplify(v1, n), real_samplify(v2, n)
        if not power_of_two_p(n):
            raise NapaFFTError(1, f"size {n} is not a power of two")
        if dst is None:
            dst = np.zeros(2 * n, dtype=np.complex128)
        elif not _is_complex_sample_array(dst):
            raise TypeError("dst must be a complex-sample-array")
        if len(dst) < 2 * n:
            raise NapaFFTError(2, f"destination of length {len(dst)} shorter than 2n = {2 * n}")
        self.b.__getattr__("2rfft")(_ptr(v1), _ptr(v2), _ptr(dst), n, 1, sc)
        return dst

    # ---- example.lisp ---------------------------------------------------------------------
    def filter_chunks(self, vector, filter):
        """example.lisp:81-114 (README:708-741): overlap-add streaming filter.  Half-overlapping chunks of
        len(filter) samples, each `windowed-fft ... :window-fn 'triangle :in-order nil`, then
        `ifft :in-order nil :window (bit-reverse filter)`, renormalised to its input energy and
        accumulated into a complex destination of len(vector).  One device pipeline, all chunks batched."""
        vector = real_samplify(vector)
        filter = real_samplify(filter)
        chunk = len(filter)
        if not power_of_two_p(chunk) or chunk < 2:
            raise NapaFFTError(1, f"filter length {chunk} is not a power of two >= 2")
        filter_rev = self.bit_reverse(filter)
        window = self.window_vector(triangle, chunk)
        dst = np.zeros(len(vector), dtype=np.complex128)
        self.b.filter_chunks(_ptr(vector), len(vector), _ptr(filter_rev), _ptr(window), chunk, _ptr(dst))
        return dst

    def windowed_rfft(self, signal_vector, center, length, window_fn=hann, dst=None, scale=None):
        """real.lisp:137-151"""
        if not power_of_two_p(length):
            raise NapaFFTError(1, f"{length} is not a power of two")
        signal_vector = real_samplify(signal_vector)
        input_window = extract_centered_window(signal_vector, center, length)
        window = self.window_vector(window_fn, length)
        input_window *= window
        return self.rfft(input_window, dst=dst, scale=scale)

    def windowed_rifft(self, vec, window_fn=rectangular, dst=None, size=None, scale=True):
        """real.lisp:187-198: multiplies the spectrum by the window IN PLACE and, like the
        reference, ignores its :scale argument (hard-coded :scale t)."""
        size = len(vec) if size is None else size
        vec = complex_samplify(vec, size)
        if not power_of_two_p(size):
            raise NapaFFTError(1, f"{size} is not a power of two")
        window = self.window_vector(window_fn, size)
        vec[:size] *= window
        return self.rifft(vec, dst=dst, size=size, scale=True)

    # ---- new exports: batched host arrays (H5: the reference has no batch parameter) ------
    @staticmethod
    def _check_batch(x, dtype, what):
        if not (isinstance(x, np.ndarray) and x.ndim == 2 and x.dtype == dtype and x.flags.c_contiguous):
            raise TypeError(f"{what} must be a C-contiguous (batch, n) array of {np.dtype(dtype).name}")
        if not power_of_two_p(x.shape[1]):
            raise NapaFFTError(1, f"size {x.shape[1]} is not a power of two")

    @staticmethod
    def _check_out(out, shape, dtype):
        if not (isinstance(out, np.ndarray) and out.shape == tuple(shape) and out.dtype == dtype and out.flags.c_contiguous
                and out.flags.writeable):
            raise TypeError(f"out must be a writable C-contiguous {tuple(shape)} array of {np.dtype(dtype).name}")

    def fft_batch(self, x, direction=1, in_order=True, scale=None, window=None, out=None):
        """x: (batch, n) complex128, C-contiguous; one call, one staging pipeline.  `window`: n weights shared by the
        batch, or (batch, n) weights, float64 or complex128 (the C ABI takes raw pointers: everything is checked here)."""
        self._check_batch(x, np.complex128, "x")
        batch, n = x.shape
        if out is None:
            out = x
        else:
            self._check_out(out, x.shape, np.complex128)
        wpb = 0
        wptr, wt, w = None, 0, None
        if window is not None:
            w = np.ascontiguousarray(window)
            if w.dtype not in (np.float64, np.complex128):
                raise TypeError("window must be float64 (real-sample-array) or complex128 (complex-sample-array)")
            if w.ndim not in (1, 2) or w.shape[-1] < n or (w.ndim == 2 and w.shape != (batch, n)):
                raise NapaFFTError(2, f"window of shape {w.shape} does not cover {batch} transform(s) of {n} points")
            wt = _lib.WINDOW_COMPLEX if w.dtype == np.complex128 else _lib.WINDOW_REAL
            wpb = int(w.ndim == 2)
            wptr = _ptr(w)
        self.b.c2c(_ptr(x), _ptr(out), n, batch, n, self._direction(direction), _scale_code(scale),
                   int(bool(in_order)), wptr, wt, wpb)
        return out

    def rfft_batch(self, x, scale=None, out=None):
        self._check_batch(x, np.float64, "x")
        batch, n = x.shape
        if out is None:
            out = np.empty((batch, n), dtype=np.complex128)
        else:
            self._check_out(out, (batch, n), np.complex128)
        self.b.rfft(_ptr(x), _ptr(out), n, batch, _scale_code(scale))
        return out

    def rifft_batch(self, x, scale=True, out=None):
        self._check_batch(x, np.complex128, "x")
        batch, n = x.shape
        if out is None:
            out = np.empty((batch, n), dtype=np.float64)
        else:
            self._check_out(out, (batch, n), np.float64)
        self.b.rifft(_ptr(x), _ptr(out), n, batch, _scale_code(scale))
        return out

    def bit_reverse_batch(self, x, out=None):
        if not (isinstance(x, np.ndarray) and x.dtype in (np.complex128, np.float64)):
            raise TypeError("x must be a (batch, n) array of complex128 or float64")
        self._check_batch(x, x.dtype, "x")
        batch, n = x.shape
        if out is None:
            out = x
        else:
            self._check_out(out, x.shape, x.dtype)
        self.b.bit_reverse(_ptr(x), _ptr(out), n, batch, n,
                           _lib.ELT_COMPLEX if x.dtype == np.complex128 else _lib.ELT_REAL)
        return out
