"""Multi-GPU check of the distributed four-step on real GPUs (run under torchrun, one rank per GPU):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 24
Every rank regenerates the same input on the host, compares its slice of the result with numpy's
DFT (normwise, tolerance 1e-13*log2 N) and checks the inverse round trip."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

lgN = int(sys.argv[1]) if len(sys.argv) > 1 else 24
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
rng = np.random.Generator(np.random.Philox(1005))
x = rng.uniform(-1, 1, N) + 1j * rng.uniform(-1, 1, N)
ref = np.fft.fft(x)
f = DistributedFFT(lgN, b, dist if world > 1 else None, dev)
xl = torch.from_numpy(f.scatter_time(x)).to(dev)
Xl = f.forward(xl)
torch.cuda.synchronize()
want = ref[f.freq_index()]
ef = np.linalg.norm(Xl.cpu().numpy() - want) / np.linalg.norm(want)
back = f.inverse(Xl)
torch.cuda.synchronize()
eb = np.linalg.norm(back.cpu().numpy() - f.scatter_time(x)) / np.linalg.norm(f.scatter_time(x))
ok = ef <= 1e-13 * lgN and eb <= 1e-13 * lgN
print("rank %d/%d lgN=%d forward nre %.3e round-trip nre %.3e %s" % (rank, world, lgN, ef, eb, "OK" if ok else "FAIL"), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
