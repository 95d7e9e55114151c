"""A/B of two builds of the library on the same box, interleaved (run on the GPU box):
python tools/ab_plans.py libA.so libB.so [sizes]   -- alternates A, B, A, B ... per size, best of 3 each."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from napa_fft3_b200._lib import Binding

libs = [Binding(os.path.abspath(p)) for p in sys.argv[1:3]]
for b in libs:
    b.init(0)
sizes = [int(a) for a in sys.argv[3].split(",")] if len(sys.argv) > 3 else [5, 6, 9, 17, 19]
st = torch.cuda.current_stream().cuda_stream
x = torch.view_as_complex(torch.rand(((4096 << 20) // 16, 2), dtype=torch.float64, device="cuda:0") * 2 - 1).contiguous()
y = torch.empty_like(x)


def run(b, n, batch, d, io, reps=20):
    f = lambda: b.c2c_dev(x.data_ptr(), y.data_ptr(), n, batch, n, d, 0, io, None, 0, 0, st)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# warm the chip into its steady (power-capped) state first
for _ in range(200):
    libs[0].c2c_dev(x.data_ptr(), y.data_ptr(), 1 << 12, x.numel() >> 12, 1 << 12, 1, 0, 1, None, 0, 0, st)
torch.cuda.synchronize()
for lg in sizes:
    n, batch = 1 << lg, x.numel() >> lg
    for name, d, io in (("fwd nat", 1, 1), ("fwd rev", 1, 0)):
        best = [1e9, 1e9]
        for rnd in range(3):
            for k, b in enumerate(libs):
                best[k] = min(best[k], run(b, n, batch, d, io))
        pct = [100 * 32.0 * n * batch / ms / 1e6 / 6545.3 for ms in best]
        print("2^%-2d %s   A %.3f ms %5.1f%%   B %.3f ms %5.1f%%" % (lg, name, best[0], pct[0], best[1], pct[1]), flush=True)
