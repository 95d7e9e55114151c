"""ctypes binding of the C ABI declared in include/napafft.h.

The product library is napa_fft3_b200/libnapafft_b200.so (nvcc, sm_100a).  There is no CPU
fallback: if the library is missing or no B200 is visible, every compute call raises NapaFFTError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_SO = os.path.join(_HERE, "libnapafft_b200.so")

STATUS_NAMES = {0: "NAPA_OK", 1: "NAPA_E_NOT_POW2", 2: "NAPA_E_SHORT_BUFFER", 3: "NAPA_E_NULL", 4: "NAPA_E_BAD_ENUM",
                5: "NAPA_E_CUDA", 6: "NAPA_E_NCCL", 7: "NAPA_E_UNSUPPORTED", 8: "NAPA_E_NO_DEVICE"}

FWD, INV = 1, -1
SCALE_NONE, SCALE_INV, SCALE_SQRT = 0, 1, 2
WINDOW_NONE, WINDOW_REAL, WINDOW_COMPLEX = 0, 1, 2
ELT_COMPLEX, ELT_REAL = 0, 1
SAMPLE_C64, SAMPLE_F64, SAMPLE_C32, SAMPLE_F32 = 0, 1, 2, 3

# every symbol include/napafft.h declares (tests check that the .so exports all of them)
ABI_SYMBOLS = [
    "napafft_init", "napafft_shutdown", "napafft_abi_version", "napafft_last_error", "napafft_launch_count",
    "napafft_c2c", "napafft_bit_reverse", "napafft_rfft", "napafft_rifft", "napafft_2rfft",
    "napafft_filter_chunks", "napafft_filter_chunks_dev", "napafft_c2c_typed", "napafft_widen_dev", "napafft_host_register", "napafft_host_unregister",
    "napafft_c2c_dev", "napafft_bit_reverse_dev", "napafft_rfft_dev", "napafft_rifft_dev", "napafft_2rfft_dev",
    "napafft_cmul_dev", "napafft_c2c_cols_dev", "napafft_relayout_dev", "napafft_dev_alloc", "napafft_dev_free", "napafft_dev_upload", "napafft_dev_download",
    "napafft_dev_sync", "napafft_scale_factor", "napafft_radix2_twiddle",
    "napafft_dist_unique_id", "napafft_dist_init", "napafft_dist_shutdown", "napafft_dist_c2c_dev", "napafft_dist_c2c",
    "napafft_dist_geometry", "napafft_dist_cols_dev", "napafft_dist_rows_dev", "napafft_dist_exchange_mode",
]


class NapaFFTError(RuntimeError):
    """Raised for every non-zero napafft_status; mirrors the Lisp conditions of the reference
    (assert / check-type / ecase failures, easy-interface.lisp:55-57, interface.lisp:12-31)."""

    def __init__(self, status, message):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class Binding:
    """Typed view of one shared library exporting the napafft C ABI."""

    def __init__(self, path: str = PRODUCT_SO):
        if not os.path.exists(path):
            raise NapaFFTError(5, f"{path} is missing: build it with `python -m napa_fft3_b200.build` "
                                  "(nvcc, sm_100a); napa-fft3-b200 has no CPU fallback")
        self.path = path
        L = self.L = C.CDLL(path)
        vp, sz, i, d = C.c_void_p, C.c_size_t, C.c_int, C.c_double
        L.napafft_init.argtypes = [i]
        L.napafft_last_error.restype = C.c_char_p
        L.napafft_launch_count.restype = C.c_uint64
        L.napafft_c2c.argtypes = [vp, vp, sz, sz, sz, i, i, i, vp, i, i]
        L.napafft_bit_reverse.argtypes = [vp, vp, sz, sz, sz, i]
        L.napafft_rfft.argtypes = [vp, vp, sz, sz, i]
        L.napafft_rifft.argtypes = [vp, vp, sz, sz, i]
        L.napafft_2rfft.argtypes = [vp, vp, vp, sz, sz, i]
        L.napafft_c2c_typed.argtypes = [vp, i, vp, sz, sz, i, i, i, vp, i, i]
        L.napafft_widen_dev.argtypes = [vp, i, vp, sz, vp]
        L.napafft_filter_chunks.argtypes = [vp, sz, vp, vp, sz, vp]
        L.napafft_filter_chunks_dev.argtypes = [vp, sz, vp, vp, sz, vp, vp]
        L.napafft_host_register.argtypes = [vp, sz]
        L.napafft_host_unregister.argtypes = [vp]
        L.napafft_c2c_dev.argtypes = [vp, vp, sz, sz, sz, i, i, i, vp, i, i, vp]
        L.napafft_bit_reverse_dev.argtypes = [vp, vp, sz, sz, sz, i, vp]
        L.napafft_rfft_dev.argtypes = [vp, vp, sz, sz, i, vp]
        L.napafft_rifft_dev.argtypes = [vp, vp, sz, sz, i, vp]
        L.napafft_2rfft_dev.argtypes = [vp, vp, vp, sz, sz, i, vp]
        L.napafft_cmul_dev.argtypes = [vp, vp, sz, sz, sz, vp]
        L.napafft_c2c_cols_dev.argtypes = [vp, vp, sz, sz, i, d, i, C.c_longlong, vp]
        L.napafft_relayout_dev.argtypes = [vp, vp, sz, sz, sz, i, vp]
        L.napafft_dev_alloc.argtypes = [C.POINTER(vp), sz]
        L.napafft_dev_free.argtypes = [vp]
        L.napafft_dev_upload.argtypes = [vp, vp, sz]
        L.napafft_dev_download.argtypes = [vp, vp, sz]
        L.napafft_dist_unique_id.argtypes = [vp]
        L.napafft_dist_init.argtypes = [vp, i, i]
        L.napafft_dist_c2c_dev.argtypes = [vp, vp, i, i, vp]
        L.napafft_dist_c2c.argtypes = [vp, vp, i, i]
        L.napafft_dist_geometry.argtypes = [i, i, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz), C.POINTER(sz), C.POINTER(i)]
        L.napafft_dist_cols_dev.argtypes = [vp, vp, i, i, i, i, i, vp]
        L.napafft_dist_rows_dev.argtypes = [vp, vp, i, i, i, vp]
        L.napafft_scale_factor.restype = d
        L.napafft_scale_factor.argtypes = [sz, i, i]
        L.napafft_radix2_twiddle.argtypes = [sz, i, vp]

    def check(self, status: int):
        if status != 0:
            raise NapaFFTError(status, (self.L.napafft_last_error() or b"").decode())

    def __getattr__(self, name):
        # b.c2c(...) -> napafft_c2c(...) with status checking
        fn = getattr(self.L, "napafft_" + name)
        if fn.restype is C.c_int:
            def call(*a):
                self.check(fn(*a))
            return call
        return fn


_default = None


def default_binding() -> Binding:
    """The product library, bound to CUDA device LOCAL_RANK (one process per GPU) on first use."""
    global _default
    if _default is None:
        b = Binding(PRODUCT_SO)
        b.init(int(os.environ.get("LOCAL_RANK", "0")) if "NAPAFFT_DEVICE" not in os.environ
               else int(os.environ["NAPAFFT_DEVICE"]))
        _default = b
    return _default
