"""Runs two-digit transforms (2^14 .. 2^20: through the fused single-launch kernels when NAPAFFT_FUSED=1; the
tuning variables NAPAFFT_FUSED* are read at library init, hence a separate process).
argv[1]: "emul" (CPU emulation build) or "gpu"."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
import parity_suite as S
from oracle import napa_oracle as O

if sys.argv[1] == "emul":
    from emul.build_emul import build
    from napa_fft3_b200 import Binding, NapaFFT
    b = Binding(build()); b.init(0)
    napa = NapaFFT(b)
    sizes, batch = (14, 15), 5
else:
    import napa_fft3_b200 as P
    P.default_binding()
    napa = P._default
    sizes, batch = (14, 15, 16, 18, 19, 20, 22), 9
for lg in sizes:
    n = 1 << lg
    for io in (True, False):
        S.check_fft(napa, n, io, "sqrt", "real")
        S.check_ifft(napa, n, io, True, "complex")
    x = np.stack([S.rand_c(i, n) for i in range(batch)])
    for io in (True, False):
        got = napa.fft_batch(x.copy(), in_order=io)
        assert max(S.nre(got[i], O.fft(x[i], in_order=io)) for i in range(batch)) <= S.tol(n)
        back = napa.fft_batch(got, direction=-1, in_order=io, scale=True)
        assert S.nre(back, x) <= S.tol(n)
    # the real wrappers ride on the same plans (zip / rifft pre-twiddle fused into the first pass)
    S.check_2rfft(napa, n)
    if "NAPAFFT_MAX_SCRATCH_KB" not in os.environ:     # (rifft has no in-place fallback: it reports NAPA_E_UNSUPPORTED)
        S.check_real(napa, 2 * n, "sqrt")
if sys.argv[1] != "emul":
    # many units per launch: ring-slot reuse, ragged last unit, every CTA crossing many phases
    for lg, batch in ((14, 1500), (16, 333), (18, 70)):
        n = 1 << lg
        rng = np.random.Generator(np.random.Philox(lg))
        x = rng.uniform(-1, 1, (batch, n)) + 1j * rng.uniform(-1, 1, (batch, n))
        for io in (True, False):
            got = napa.fft_batch(x.copy(), in_order=io)
            for i in (0, 1, batch // 2, batch - 1):
                assert S.nre(got[i], O.fft(x[i], in_order=io)) <= S.tol(n), (lg, io, i)
            back = napa.fft_batch(got, direction=-1, in_order=io, scale=True)
            assert S.nre(back, x) <= S.tol(n), (lg, io)
print("fused ok")
