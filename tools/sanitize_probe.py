"""Small workload for compute-sanitizer (racecheck / memcheck) on the GPU box: every kernel family once, sizes small
enough for the tool's slowdown.  python tools/sanitize_probe.py [sizes]   (NAPAFFT_FUSED / NAPAFFT_FUSED_TMA select the
fused kernels).  Exits non-zero on a parity failure; the sanitizer reports hazards itself."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import napa_fft3_b200 as P
import parity_suite as S
from oracle import napa_oracle as O

napa = P.NapaFFT(P.default_binding())
sizes = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [6, 10, 13, 14, 16]
for lg in sizes:
    n = 1 << lg
    batch = 3 if lg >= 14 else 9
    x = np.stack([S.rand_c(i, n) for i in range(batch)])
    w = S.rand_r(5, n, 0.5, 1.5)
    for io in (True, False):
        got = napa.fft_batch(x.copy(), in_order=io, scale="sqrt", window=w)
        assert max(S.nre(got[i], O.fft(x[i], in_order=io, scale="sqrt", window=w)) for i in range(batch)) <= S.tol(n), (lg, io)
        back = napa.fft_batch(got, direction=-1, in_order=io, scale="sqrt")
        assert S.nre(back, x * w) <= S.tol(n), (lg, io)
    assert np.array_equal(napa.bit_reverse_batch(x.copy()), x[:, S.revperm(n)])
    r = np.ascontiguousarray(x.real)
    spec = napa.rfft_batch(r)
    assert S.nre(spec[0], O.rfft(r[0])) <= S.tol(n)
    assert S.nre(napa.rifft_batch(spec), r) <= S.tol(n)
print("sanitize probe ok", sizes, {k: v for k, v in os.environ.items() if k.startswith("NAPAFFT")})
