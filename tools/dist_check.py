"""Multi-GPU parity check of the distributed four-step on real GPUs (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 24 [native|host]
lgN <= 26: every rank regenerates the same input on the host, transforms it with the CPU ORACLE (the C restatement of
Napa-FFT3) and compares its slice of the distributed result, normwise, tolerance 1e-13*log2 N; then the inverse round trip.
lgN > 26: the input is a handful of impulses, whose spectrum is known in closed form; every local bin is compared.
Rank 0 prints "dist ok" when every rank passed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

lgN = int(sys.argv[1]) if len(sys.argv) > 1 else 24
mode = sys.argv[2] if len(sys.argv) > 2 else "native"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
os.environ["NAPAFFT_DEVICE"] = str(local)
import napa_fft3_b200 as napa
from napa_fft3_b200.dist import DistributedFFT

b = napa.default_binding()
dev = torch.device("cuda", local)
N = 1 << lgN
f = DistributedFFT(lgN, b, dist if world > 1 else None, dev, native=(mode == "native"))
tol = 1e-13 * lgN
if lgN <= 26:
    from oracle import napa_oracle as O
    rng = np.random.Generator(np.random.Philox(1005))
    x = rng.uniform(-1, 1, N) + 1j * rng.uniform(-1, 1, N)
    ref = O.fft(x.copy())
    xl_host = f.scatter_time(x)
    want = ref[f.freq_index()]
    checker = "oracle"
else:
    rng = np.random.Generator(np.random.Philox(lgN))
    pos = np.unique(np.concatenate([[0, 1, N - 1, N // 2 + 1], rng.integers(0, N, 12)]))
    amp = rng.uniform(-1, 1, len(pos)) + 1j * rng.uniform(-1, 1, len(pos))
    xl_host = np.zeros((f.N1, f.B), dtype=np.complex128)
    for p_, a_ in zip(pos, amp):
        j1, j2 = divmod(int(p_), f.N2)
        if f.rank * f.B <= j2 < (f.rank + 1) * f.B:
            xl_host[j1, j2 - f.rank * f.B] = a_
    k = f.freq_index().astype(np.int64)
    want = np.zeros(k.shape, dtype=np.complex128)
    for p_, a_ in zip(pos, amp):
        ph = (k * int(p_)) % N if int(p_) < (1 << 31) and lgN <= 31 else np.mod(k.astype(object) * int(p_), N).astype(np.int64)
        want += a_ * np.exp(-2j * np.pi * ph.astype(np.float64) / N)
    checker = "closed form (impulses)"
xl = torch.from_numpy(xl_host).to(dev)
Xl = f.forward(xl)
torch.cuda.synchronize()
ef = np.linalg.norm(Xl.cpu().numpy() - want) / np.linalg.norm(want)
back = f.inverse(Xl)
torch.cuda.synchronize()
eb = np.linalg.norm(back.cpu().numpy() - xl_host) / np.linalg.norm(xl_host) if np.linalg.norm(xl_host) > 0 else float(np.linalg.norm(back.cpu().numpy()))
ok = bool(ef <= tol and eb <= tol)
print("rank %d/%d lgN=%d mode=%s chunks=%d forward nre %.3e (vs %s) round-trip nre %.3e tol %.1e %s"
      % (rank, world, lgN, mode, f.nch, ef, checker, eb, tol, "OK" if ok else "FAIL"), flush=True)
if world > 1:
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    ok = t.item() == 0
    dist.barrier()
    dist.destroy_process_group()
if rank == 0 and ok:
    print("dist ok", flush=True)
sys.exit(0 if ok else 1)
