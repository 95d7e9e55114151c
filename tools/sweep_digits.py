"""Digit-plan tuning sweep (run on the GPU box): times every admissible digit split of 2^lg through
the NAPAFFT_DIGITS override, natural and bit-reversed output.
python tools/sweep_digits.py [total_MiB] [lg,lg,...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import napa_fft3_b200 as napa

total = (int(sys.argv[1]) if len(sys.argv) > 1 else 2048) << 20
lgs = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(range(14, 25))
b = napa.default_binding()
st = torch.cuda.current_stream().cuda_stream
x = torch.view_as_complex(torch.rand((total // 16, 2), dtype=torch.float64, device="cuda:0") * 2 - 1).contiguous()
y = torch.empty_like(x)


def log2c(l):
    return (11 if l <= 8 else (12 if l <= 12 else 13)) - l


def plans(lg):
    out = []
    for a in range(6, 11):
        if 4 <= lg - a <= 13 and lg - a >= log2c(a):
            out.append((a, lg - a))
    if lg >= 20:
        for a in range(6, 11):
            for c in range(6, 11):
                r = lg - a - c
                if 7 <= r <= 12 and r >= log2c(c) and abs(a - c) <= 1:
                    out.append((a, c, r))
    return out


def timeit(f, reps=8):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            f()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best


for lg in lgs:
    n = 1 << lg
    batch = x.numel() // n
    for d in plans(lg):
        os.environ["NAPAFFT_DIGITS"] = ",".join(map(str, d))
        res = []
        for name, dr, io in (("nat", 1, 1), ("rev", 1, 0), ("inv-rev", -1, 0)):
            ms = timeit(lambda: b.c2c_dev(x.data_ptr(), y.data_ptr(), n, batch, n, dr, 0, io, None, 0, 0, st))
            res.append("%s %6.3f ms %5.1f%%" % (name, ms, 100 * 32.0 * n * batch / ms / 1e6 / 6545.3))
        print("2^%-2d %-10s %s" % (lg, os.environ["NAPAFFT_DIGITS"], "   ".join(res)), flush=True)
    os.environ.pop("NAPAFFT_DIGITS", None)
