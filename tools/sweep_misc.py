"""Device-resident throughput of the non-c2c entry points (run on the GPU box):
bit reversal (complex / real), rfft, rifft, %2rfft, filter_chunks.  % of the measured 6545 GB/s copy bandwidth
against each function's algorithmic bytes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import napa_fft3_b200 as napa

b = napa.default_binding()
st = torch.cuda.current_stream().cuda_stream
total = (int(sys.argv[1]) if len(sys.argv) > 1 else 2048) << 20      # bytes of the largest operand
buf_a = torch.empty(total // 8, dtype=torch.float64, device="cuda:0").uniform_(-1, 1)
buf_b = torch.empty(total // 8, dtype=torch.float64, device="cuda:0")
buf_c = torch.empty(total // 8, dtype=torch.float64, device="cuda:0").uniform_(-1, 1)


def timeit(f, reps=10):
    for _ in range(3):
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


def report(name, lg, batch, ms, nbytes):
    print("%-22s 2^%-2d x %-8d %8.3f ms  %7.1f GB/s  %5.1f%% of 6545" % (name, lg, batch, ms, nbytes / ms / 1e6, 100 * nbytes / ms / 1e6 / 6545.3), flush=True)


a, o, c = buf_a.data_ptr(), buf_b.data_ptr(), buf_c.data_ptr()
for lg in (8, 12, 16, 20, 24):
    n = 1 << lg
    bc = total // 16 // n
    report("bit-reverse complex", lg, bc, timeit(lambda: b.bit_reverse_dev(a, o, n, bc, n, 0, st)), 32.0 * n * bc)
    br = total // 8 // n
    report("bit-reverse real", lg, br, timeit(lambda: b.bit_reverse_dev(a, o, n, br, n, 1, st)), 16.0 * n * br)
for lg in (8, 10, 12, 14, 16, 18, 20):
    n = 1 << lg
    bt = total // 16 // n                     # the complex side (n complex per item) fills `total`
    report("rfft", lg, bt, timeit(lambda: b.rfft_dev(c, o, n, bt, 0, st)), 24.0 * n * bt)
    report("rifft", lg, bt, timeit(lambda: b.rifft_dev(a, o, n, bt, 1, st)), 24.0 * n * bt)
    b2 = total // 32 // n
    report("%2rfft", lg, b2, timeit(lambda: b.__getattr__("2rfft_dev")(c, c + 8 * n * b2, o, n, b2, 0, st)), 48.0 * n * b2)
for chunk in (256, 4096, 65536):
    ln = total // 64                          # the chunk matrix is 32 bytes per input sample
    w = torch.ones(chunk, dtype=torch.float64, device="cuda:0")
    ms = timeit(lambda: b.filter_chunks_dev(c, ln, w.data_ptr(), w.data_ptr(), chunk, o, st), reps=5)
    print("filter_chunks chunk %-6d %d samples %8.3f ms  %7.1f Msample/s  (%.1f GB/s of signal in + complex out)" % (
        chunk, ln, ms, ln / ms / 1e3, 24.0 * ln / ms / 1e6), flush=True)
