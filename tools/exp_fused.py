"""A/B of library tuning variables inside ONE process (the variables are read by napafft_init, so each variant is
shutdown -> set env -> init).  python tools/exp_fused.py <total_MiB> <lg,lg,...> "<VAR=val VAR=val>;<...>;..." [reps]
Prints algorithmic GB/s (32 N B per transform) for natural / bit-reversed forward / bit-reversed inverse."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import napa_fft3_b200 as napa

total = int(sys.argv[1]) << 20
sizes = [int(a) for a in sys.argv[2].split(",")]
variants = [v.strip() for v in sys.argv[3].split(";")]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
b = napa.Binding(os.environ["NAPA_SO"]) if "NAPA_SO" in os.environ else napa.default_binding()
if "NAPA_SO" in os.environ:
    b.init(0)
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
x = torch.view_as_complex(torch.rand((total // 16, 2), dtype=torch.float64, device=dev) * 2 - 1).contiguous()
y = torch.empty_like(x)
touched = set()
for rnd in range(2):                      # two interleaved rounds: the chip's power state drifts
    for var in variants:
        for k in touched:
            os.environ.pop(k, None)
        for kv in var.split():
            k, v = kv.split("=")
            os.environ[k] = v
            touched.add(k)
        b.shutdown()
        b.init(0)
        for lg in sizes:
            n = 1 << lg
            batch = x.numel() // n
            out = []
            for name, d, io in (("nat", 1, 1), ("rev", 1, 0), ("inv", -1, 0)):
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
                ms = e0.elapsed_time(e1) / reps
                out.append("%s %7.3f ms %5.1f%%" % (name, ms, 100 * 32.0 * n * batch / ms / 1e6 / 6545.3))
            print("[%d] %-44s 2^%-2d  %s" % (rnd, var or "(default)", lg, "   ".join(out)), flush=True)
