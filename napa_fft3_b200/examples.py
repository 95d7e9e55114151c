"""Host-side helpers of the reference's example.lisp (audio demos, README:560-741): raw signed-32-bit sample
files, test signals and the spectrum clamps.  Plain numpy bookkeeping around the transforms -- the transforms
themselves (and `filter_chunks`, example.lisp:81-114) run on the device through `napa_fft3_b200.host`."""
from __future__ import annotations

import numpy as np


def emit_raw32_file(file, data, max=1.0):
    """example.lisp:1-11: real parts scaled by 1/max, times 2^31 - 1, floored, written as native s32."""
    x = np.real(np.asarray(data)).astype(np.float64) / max
    np.floor(x * float((1 << 31) - 1)).astype(np.int32).tofile(file)
    return file


def read_raw32_file(file, n=None, max=1.0):
    """example.lisp:13-24: s32 samples times (2 max)^-31 as a real-sample-array."""
    seq = np.fromfile(file, dtype=np.int32, count=-1 if n is None else n)
    return seq.astype(np.float64) * (2.0 * max) ** -31


def impulse(i, n, value=1.0):
    """example.lisp:26-30"""
    vec = np.zeros(n, dtype=np.complex128)
    for k in (i if isinstance(i, (list, tuple)) else [i]):
        vec[k] = complex(float(value))
    return vec


def noise(n, range=0.5, rng=None):
    """example.lisp:32-36 (uniform in [-range, range); pass a numpy Generator for reproducibility)"""
    rng = np.random.default_rng() if rng is None else rng
    return (rng.uniform(0.0, 2 * range, n) - range).astype(np.complex128)


def m_plus(x, y, scale=0.5):
    """example.lisp:38-42 (m+)"""
    return (scale * (np.asarray(x) + np.asarray(y))).astype(np.complex128)


def amplitude_clamp_fraction(vector, fraction):
    """example.lisp:44-52: zero every bin below fraction * max |x|"""
    v = np.asarray(vector, dtype=np.complex128)
    limit = np.abs(v).max() * fraction
    return np.where(np.abs(v) < limit, 0.0 + 0.0j, v)


def amplitude_clamp_k(vector, k):
    """example.lisp:54-67: keep the k bins of largest magnitude (stable order among ties)"""
    v = np.asarray(vector, dtype=np.complex128)
    order = np.argsort(-np.abs(v), kind="stable")[:k]
    out = np.zeros_like(v)
    out[order] = v[order]
    return out


def energy(x):
    """example.lisp:69-79"""
    x = np.asarray(x, dtype=np.complex128)
    return float(np.sum(x.real * x.real + x.imag * x.imag))
