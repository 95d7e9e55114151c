"""Device-resident throughput sweep over transform sizes (run on the GPU box):
python tools/sweep_sizes.py [total_MiB]  -> GB/s of algorithmic traffic (32 N B per transform)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import napa_fft3_b200 as napa

total = (int(sys.argv[1]) if len(sys.argv) > 1 else 2048) << 20
b = napa.default_binding()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
x = torch.view_as_complex(torch.rand((total // 16, 2), dtype=torch.float64, device=dev) * 2 - 1).contiguous()
y = torch.empty_like(x)
sizes = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4, 6, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 24]
for lg in sizes:
    n = 1 << lg
    batch = x.numel() // n
    for name, d, io in (("fwd nat", 1, 1), ("fwd rev", 1, 0), ("inv rev", -1, 0)):
        f = lambda: b.c2c_dev(x.data_ptr(), y.data_ptr(), n, batch, n, d, 0, io, None, 0, 0, st)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 10
        for _ in range(reps):
            f()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = 32.0 * n * batch / ms / 1e6
        print("2^%-2d x %-8d %s  %8.3f ms  %7.1f GB/s  %5.1f%% of 6545  %7.1f GFLOP/s" % (lg, batch, name, ms, gbs, 100 * gbs / 6545.3, 5 * n * lg * batch / ms / 1e6), flush=True)
# plain copy for reference
for _ in range(3):
    y.copy_(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y.copy_(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("torch copy %.3f ms %.1f GB/s" % (ms, 2 * x.numel() * 16 / ms / 1e6))
