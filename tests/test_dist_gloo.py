"""N > 1 host logic on CPU: world_size 2, gloo backend.  The per-rank transform runs
on the emulation build of the shipped csrc (no GPU here); on the GPU box the same sharding code
drives the product library from bench.py --gpus N."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_shard_bounds_cover_everything():
    from napa_fft3_b200.dist import shard_bounds
    for batch in (0, 1, 5, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from emul.build_emul import build
    from napa_fft3_b200 import Binding, NapaFFT
    from napa_fft3_b200.dist import sharded_apply, max_over_ranks
    from oracle import napa_oracle as O
    import parity_suite as S
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = Binding(build()); b.init(0)
    napa = NapaFFT(b)
    n, batch = 1 << 10, 13                                   # ragged: 7 + 6 rows
    x = np.stack([S.rand_c(900 + i, n) for i in range(batch)])
    w = O.window_vector("hann", n)

    def gather(local):
        out = [None] * world
        dist.all_gather_object(out, local)
        return out
    spec = sharded_apply(lambda rows: napa.fft_batch(rows.copy(), window=w), x, rank, world, gather)
    back = sharded_apply(lambda rows: napa.fft_batch(rows.copy(), direction=-1, scale="inv"), spec, rank, world, gather)
    err = max(S.nre(spec[i], O.fft(x[i], window=w)) for i in range(batch))
    rt = S.nre(back, x * w)
    t = max_over_ranks(float(rank + 1), dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, err, rt, t, spec.shape))


def test_batch_sharding_world2_gloo():
    import torch.multiprocessing as mp
    from emul.build_emul import build
    build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, rt, t, shape in res:
        assert err <= 1e-12 and rt <= 1e-12
        assert t == 2.0                                      # max over ranks
        assert shape == (13, 1024)


def _dist_fft_worker(rank, world, port, q, lgN):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from emul.build_emul import build
    from napa_fft3_b200 import Binding
    from napa_fft3_b200.dist import DistributedFFT
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = Binding(build()); b.init(0)
    N = 1 << lgN
    rng = np.random.Generator(np.random.Philox(1005))
    x = rng.uniform(-1, 1, N) + 1j * rng.uniform(-1, 1, N)           # same vector on every rank
    ref = np.fft.fft(x)
    f = DistributedFFT(lgN, b, dist, torch.device("cpu"))
    xl = torch.from_numpy(f.scatter_time(x))
    Xl = f.forward(xl)
    err_f = np.linalg.norm(Xl.numpy() - ref[f.freq_index()]) / np.linalg.norm(ref[f.freq_index()])
    back = f.inverse(Xl)
    err_b = np.linalg.norm(back.numpy() - f.scatter_time(x)) / np.linalg.norm(f.scatter_time(x))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, float(err_f), float(err_b), f.alltoall_bytes_per_rank()))


@pytest.mark.parametrize("lgN", [16, 21])
def test_distributed_four_step_world2_gloo(lgN):
    """config C5 in miniature: one transform over 2 ranks, one all-to-all, vs numpy's DFT"""
    import torch.multiprocessing as mp
    from emul.build_emul import build
    build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + lgN
    procs = [ctx.Process(target=_dist_fft_worker, args=(r, 2, port, q, lgN)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ef, eb, nbytes in res:
        assert ef <= 1e-13 * lgN and eb <= 1e-13 * lgN, (rank, ef, eb)
        assert nbytes == (1 << lgN) * 16 // 4                      # (P-1)/P * 16 N / P with P = 2
