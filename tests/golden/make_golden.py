"""Generates tests/golden/oracle_vectors.npz: inputs and outputs of the PINNED CPU oracle
(oracle/napa_oracle.c, bit-exact on every README transcript of the reference, tests/test_oracle_kat.py)
on seeded inputs.  The reference itself (Common Lisp) cannot run in this image, so these are the
closest thing to reference outputs that can travel to the GPU box; the CUDA path is compared with
them in tests/test_gpu_parity.py::test_golden_vectors and the oracle is checked to still reproduce
them bit for bit in tests/test_oracle_properties.py::test_golden_vectors_reproduce.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import napa_oracle as O  # noqa: E402

SIZES = (2, 8, 64, 1024, 4096)


def inputs(n):
    rng = np.random.Generator(np.random.Philox(20240 + n))
    xc = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    xr = rng.uniform(-1, 1, n)
    wr = rng.uniform(1, 2, n)
    wc = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
    return xc, xr, wr, wc


def vectors():
    out = {}
    for n in SIZES:
        xc, xr, wr, wc = inputs(n)
        k = "n%d_" % n
        out[k + "xc"], out[k + "xr"], out[k + "wr"], out[k + "wc"] = xc, xr, wr, wc
        out[k + "fft"] = O.fft(xc.copy())
        out[k + "fft_rev_sqrt_wr"] = O.fft(xc.copy(), in_order=False, scale="sqrt", window=wr)
        if n > 1024:
            del out[k + "wc"]
            continue                      # the largest size only pins the plain transforms (keeps the file small)
        out[k + "ifft_rev_wc"] = O.ifft(xc.copy(), in_order=False, scale=True, window=wc)
        out[k + "ifft_nat_wr"] = O.ifft(xc.copy(), in_order=True, scale="sqrt", window=wr)
        out[k + "bitrev_c"] = O.bit_reverse(xc.copy())
        out[k + "bitrev_r"] = O.bit_reverse(xr.copy())
        out[k + "rfft"] = O.rfft(xr.copy())
        out[k + "rifft"] = O.rifft(O.rfft(xr.copy()), scale=True)
        out[k + "2rfft"] = O.two_rfft(xr.copy(), wr.copy()) if n >= 2 else np.zeros(0)
    rng = np.random.Generator(np.random.Philox(777))
    sig = rng.uniform(-1, 1, 1000)
    kk = np.minimum(np.arange(64), 64 - np.arange(64)) / 64
    filt = 1.0 / (1.0 + (8 * kk) ** 4)
    out["oa_signal"], out["oa_filter"], out["oa_out"] = sig, filt, O.filter_chunks(sig, filt)
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.npz")
    np.savez_compressed(path, **vectors())
    print(path, os.path.getsize(path), "bytes")
