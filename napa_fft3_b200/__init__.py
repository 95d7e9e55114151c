"""napa-fft3-b200: the hot path of pkhuong/Napa-FFT3 (complex-double power-of-two DFTs, bit
reversal, fused scale/window, real wrappers) as hand-written CUDA for B200 (sm_100a) behind a C
ABI (include/napafft.h), with this package mirroring the reference's NAPA-FFT export list
(package.lisp:28-49) for a host that is not Common Lisp.

    import napa_fft3_b200 as napa
    napa.fft([0, 1, 2, 3, 4, 5, 6, 7])            # (napa-fft:fft '(0 1 2 3 4 5 6 7))
    napa.ifft(x, in_order=False, window=h)         # (napa-fft:ifft x :in-order nil :window h)

There is no CPU fallback: calls raise NapaFFTError when libnapafft_b200.so or a B200 is missing.
"""
from ._lib import Binding, NapaFFTError, default_binding, PRODUCT_SO  # noqa: F401
from .host import (NapaFFT, TwiddleHandle, complex_samplify, real_samplify, power_of_two_p, lb,  # noqa: F401
                   rectangular, hann, blackman_star, blackman, triangle, bartlett, gauss_star, gaussian,
                   gaussian_bartlett_x, cosine_series, blackman_harris, extract_window, extract_window_into,
                   extract_centered_window, extract_centered_window_into)

_default = NapaFFT()          # binds to the product library lazily, on the first transform

fft = _default.fft
ifft = _default.ifft
bit_reverse = _default.bit_reverse
windowed_fft = _default.windowed_fft
windowed_ifft = _default.windowed_ifft
window_vector = _default.window_vector
rfft = _default.rfft
rifft = _default.rifft
two_rfft = _default.two_rfft                # %2rfft
windowed_rfft = _default.windowed_rfft
windowed_rifft = _default.windowed_rifft
filter_chunks = _default.filter_chunks      # example.lisp:81-114
get_fft = _default.get_fft
get_windowed_fft = _default.get_windowed_fft
get_reverse = _default.get_reverse
ensure_fft = _default.ensure_fft            # %ensure-fft
ensure_twiddles = _default.ensure_twiddles  # %ensure-twiddles
ensure_reverse = _default.ensure_reverse    # %ensure-reverse
fft_batch = _default.fft_batch
rfft_batch = _default.rfft_batch
rifft_batch = _default.rifft_batch
bit_reverse_batch = _default.bit_reverse_batch
