"""GPU tier (-m gpu): the product library (sm_100a kernels) through the C ABI vs the CPU oracle."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_suite as S  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def napa():
    import napa_fft3_b200 as P
    from napa_fft3_b200.build import build
    build()
    P.default_binding()          # raises if the CUDA library or the GPU is missing: no fallback
    return P._default


@pytest.mark.parametrize("n", [1 << k for k in range(0, 21)])
def test_fft_all_options(napa, n):
    for in_order in (True, False):
        for scale, wt in ((None, None), (True, "real"), ("sqrt", "complex")):
            S.check_fft(napa, n, in_order, scale, wt)


@pytest.mark.parametrize("n", [1 << k for k in range(0, 21)])
def test_ifft_all_options(napa, n):
    for in_order in (True, False):
        for scale, wt in ((True, None), (None, "real"), ("sqrt", "complex")):
            S.check_ifft(napa, n, in_order, scale, wt)


@pytest.mark.parametrize("n", [1 << 22, 1 << 24, 1 << 25])
def test_large_single(napa, n):
    S.check_fft(napa, n, True, None, None)
    S.check_fft(napa, n, False, "sqrt", "real")
    S.check_ifft(napa, n, False, True, "complex")
    S.check_ifft(napa, n, True, True, None)


@pytest.mark.parametrize("n", [1 << k for k in range(1, 21)])
def test_layout_bit_exact(napa, n):
    S.check_layout_bit_exact(napa, n)


@pytest.mark.parametrize("n", [1 << k for k in range(0, 23)])
def test_bit_reverse(napa, n):
    S.check_bit_reverse(napa, n)


@pytest.mark.parametrize("n", [1, 2, 8, 32, 1024, 1 << 14, 1 << 18])
def test_offsets_exact(napa, n):
    S.check_offsets_exact(napa, n)


@pytest.mark.parametrize("n", [1 << k for k in range(1, 21)])
def test_real(napa, n):
    for scale in (None, True, "sqrt"):
        S.check_real(napa, n, scale)


@pytest.mark.parametrize("n", [2, 4, 64, 1 << 13, 1 << 17])
def test_2rfft(napa, n):
    S.check_2rfft(napa, n)


@pytest.mark.parametrize("n", [4, 64, 1024, 1 << 16])
def test_ergun(napa, n):
    S.check_ergun(napa, n, True)
    S.check_ergun(napa, n, False)


def test_readme_kats(napa):
    S.check_readme_kats(napa)


def test_dst_rules(napa):
    S.check_dst_rules(napa)


def test_windowed(napa):
    S.check_windowed(napa)
    S.check_windowed(napa, 1 << 12)


def test_window_functions(napa):
    S.check_window_functions(napa)


def test_batched_configs_vs_oracle(napa):
    """BASELINE configs at reduced batch, every item compared with the oracle.
    C2: windowed (Hann) fft -> ifft :scale :inv, 2^16; C3: fft :in-order nil -> ifft with complex
    window (convolution building block), 2^20; C4: rfft -> rifft, 2^14."""
    import napa_fft3_b200 as P
    from oracle import napa_oracle as O
    n, B = 1 << 16, 24
    x = np.stack([S.rand_c(1002 + i, n) for i in range(B)])
    hann = napa.window_vector(P.hann, n)
    spec = napa.fft_batch(x.copy(), window=hann)
    back = napa.fft_batch(spec.copy(), direction=-1, scale="inv")
    for i in range(B):
        want = O.fft(x[i], window=hann)
        assert S.nre(spec[i], want) <= S.tol(n)
        assert S.nre(back[i], O.ifft(want.copy(), scale=True)) <= S.tol(n)
    n, B = 1 << 20, 6
    x = np.stack([S.rand_c(1003 + i, n) for i in range(B)])
    h = napa.fft(S.rand_c(999, n), in_order=False)
    spec = napa.fft_batch(x.copy(), in_order=False)
    conv = napa.fft_batch(spec.copy(), direction=-1, in_order=False, scale="inv", window=h)
    for i in range(B):
        want = O.fft(x[i], in_order=False)
        assert S.nre(spec[i], want) <= S.tol(n)
        assert S.nre(conv[i], O.ifft(want.copy(), in_order=False, scale=True, window=h)) <= S.tol(n)
    n, B = 1 << 14, 64
    r = np.stack([S.rand_r(1004 + i, n) for i in range(B)])
    spec = napa.rfft_batch(r)
    back = napa.rifft_batch(spec.copy())
    for i in range(B):
        want = O.rfft(r[i])
        assert S.nre(spec[i], want) <= S.tol(n)
        assert S.nre(back[i], O.rifft(want.copy())) <= S.tol(n)


def test_full_size_properties_device_resident(napa):
    """BASELINE full sizes through the device-resident entry points, checked by size-independent
    properties: round trip, linearity, Parseval, and a sample of items against the oracle."""
    import torch
    import napa_fft3_b200 as P
    from oracle import napa_oracle as O
    b = P.default_binding()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(1005)
    for n, B, in_order in ((1 << 16, 512, True), (1 << 20, 64, False), (1 << 10, 4096, True)):
        x = torch.view_as_complex(torch.rand((B, n, 2), dtype=torch.float64, device=dev, generator=g) * 2 - 1).contiguous()
        y = torch.empty_like(x)
        st = torch.cuda.current_stream().cuda_stream
        b.c2c_dev(x.data_ptr(), y.data_ptr(), n, B, n, 1, 0, int(in_order), None, 0, 0, st)
        torch.cuda.synchronize()
        # Parseval: sum |X|^2 = n sum |x|^2
        ex, ey = (x.abs() ** 2).sum(dim=1), (y.abs() ** 2).sum(dim=1)
        assert torch.max(torch.abs(ey / (n * ex) - 1)).item() < 1e-12
        # oracle on a few items
        for i in (0, B // 2, B - 1):
            want = O.fft(x[i].cpu().numpy(), in_order=in_order)
            assert S.nre(y[i].cpu().numpy(), want) <= S.tol(n)
        # linearity: F(a + 2b) = F(a) + 2 F(b)
        z = torch.empty_like(x)
        comb = (x + 2 * x.roll(1, 0)).contiguous()
        b.c2c_dev(comb.data_ptr(), z.data_ptr(), n, B, n, 1, 0, int(in_order), None, 0, 0, st)
        torch.cuda.synchronize()
        ref = y + 2 * y.roll(1, 0)
        assert (torch.linalg.norm(z - ref) / torch.linalg.norm(ref)).item() < 1e-14 * np.log2(n) * 10
        # round trip in place
        b.c2c_dev(y.data_ptr(), y.data_ptr(), n, B, n, -1, 1, int(in_order), None, 0, 0, st)
        torch.cuda.synchronize()
        assert (torch.linalg.norm(y - x) / torch.linalg.norm(x)).item() <= S.tol(n)


def test_launch_counter_moves(napa):
    import napa_fft3_b200 as P
    b = P.default_binding()
    before = b.launch_count()
    napa.fft(np.arange(1024, dtype=np.complex128))
    assert b.launch_count() > before


@pytest.mark.parametrize("lgN", [16, 22, 26, 28])
def test_distributed_building_blocks_single_rank(napa, lgN):
    """the device pipeline of the distributed four-step (chunked column FFTs with the fused global twiddle, exchange
    layout, row FFTs reading it in place -- single-pass rows up to 2^26, two-pass rows at 2^28) on one GPU through
    napafft_dist_c2c_dev: P = 1 makes the exchange a copy on the communication stream"""
    import torch
    import napa_fft3_b200 as P
    from napa_fft3_b200.dist import DistributedFFT
    dev = torch.device("cuda:0")
    for native in ((True, False) if lgN == 22 else (True,)):
        f = DistributedFFT(lgN, P.default_binding(), None, dev, native=native)
        N = 1 << lgN
        if lgN <= 26:
            x = S.rand_c(1005, N)
            xl = torch.from_numpy(f.scatter_time(x)).to(dev)
            want, sel = np.fft.fft(x)[f.freq_index()], None
        else:
            # too large for a host DFT: impulses, whose spectrum is known in closed form, on sampled bins
            rng = np.random.Generator(np.random.Philox(lgN))
            pos = np.unique(np.concatenate([[0, 1, N - 1, N // 2 + 1], rng.integers(0, N, 12)]))
            amp = rng.uniform(-1, 1, len(pos)) + 1j * rng.uniform(-1, 1, len(pos))
            xl = torch.zeros((f.N1, f.B), dtype=torch.complex128, device=dev)
            xl.view(-1)[torch.from_numpy(pos).to(dev)] = torch.from_numpy(amp).to(dev)
            sel = np.unique(np.concatenate([[0, 1, N - 1, N // 2], rng.integers(0, N, 500)]))      # element index of [R][N2] (P = 1)
            k = (sel // f.N2) + f.N1 * (sel % f.N2)                                                 # X[k1 + N1 k2] at row k1, column k2
            ph = (pos[None, :].astype(object) * k[:, None].astype(object)) % N
            want = (amp[None, :] * np.exp(-2j * np.pi * ph.astype(np.float64) / N)).sum(axis=1)
        Xl = f.forward(xl)
        torch.cuda.synchronize()
        got = Xl.cpu().numpy() if sel is None else Xl.view(-1)[torch.from_numpy(sel).to(dev)].cpu().numpy()
        assert S.nre(got, want) <= S.tol(N), (lgN, native)
        back = f.inverse(Xl)
        torch.cuda.synchronize()
        assert (torch.linalg.norm(back - xl) / torch.linalg.norm(xl)).item() <= S.tol(N), (lgN, native)
        del Xl, back, xl
        torch.cuda.empty_cache()


def test_fused_kernel_schedules_gpu():
    """fused single-launch kernels (opt-in, NAPAFFT_FUSED=1): automatic unit / grid shape for the direct-load and the
    TMA-fed variant, one transform per unit on a small grid (every CTA crosses many phases and really waits on the
    counters), wide units with a long lead"""
    import subprocess
    for extra in ({}, {"NAPAFFT_FUSED_TMA": "1"}, {"NAPAFFT_FUSED_UPG": "1", "NAPAFFT_FUSED_GRID": "37", "NAPAFFT_FUSED_LEAD": "1"},
                  {"NAPAFFT_FUSED_TMA": "1", "NAPAFFT_FUSED_UPG": "5", "NAPAFFT_FUSED_LEAD": "3"}):
        env = dict(os.environ, NAPAFFT_FUSED="1", **extra)
        out = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "fused_probe.py"), "gpu"],
                             env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "fused ok" in out.stdout, str(extra) + out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("n,chunk", [(1000, 64), (100000, 1024), (777, 1024), (5000, 2), (1 << 20, 1 << 16), (300001, 4096)])
def test_filter_chunks(napa, n, chunk):
    """overlap-add driver of example.lisp:81-114 as one device pipeline"""
    S.check_filter_chunks(napa, n, chunk)


@pytest.mark.parametrize("n", [1, 2, 8, 256, 4096, 1 << 14, 1 << 18, 1 << 21])
def test_typed_inputs(napa, n):
    S.check_typed_inputs(napa, n)


@pytest.mark.parametrize("n", [256, 1 << 11, 1 << 14])
def test_rfft_fused_epilogue_matches_unfused(napa, n, monkeypatch):
    """the single-launch rfft (epilogue fused into the half-size transform's store) against the two-launch path"""
    from oracle import napa_oracle as O
    x = np.stack([S.rand_r(40 + i, n) for i in range(9)])
    fused = napa.rfft_batch(x.copy())
    monkeypatch.setenv("NAPAFFT_RFFT_UNFUSED", "1")
    unfused = napa.rfft_batch(x.copy())
    assert S.nre(fused, unfused) <= 1e-15
    for i in (0, 8):
        assert S.nre(fused[i], O.rfft(x[i])) <= S.tol(n)


def test_golden_vectors(napa):
    """CUDA path through the C ABI against the committed oracle fixtures (tests/golden/)"""
    S.check_golden(napa)


@pytest.mark.parametrize("lg", [14, 16])
def test_host_batches_across_staging_chunks(napa, lg, monkeypatch):
    """natural-order multi-pass transforms staged over both copy streams: each stream has its own
    scratch, chunks in flight on the two streams must not disturb each other"""
    from oracle import napa_oracle as O
    monkeypatch.setenv("NAPAFFT_STAGE_MB", "8")
    n, batch = 1 << lg, (200 if lg == 14 else 60)
    rng = np.random.Generator(np.random.Philox(lg))
    x = rng.uniform(-1, 1, (batch, n)) + 1j * rng.uniform(-1, 1, (batch, n))
    for rep in range(3):
        got = napa.fft_batch(x.copy())
        for i in range(0, batch, 7):
            assert S.nre(got[i], O.fft(x[i])) <= S.tol(n), (lg, rep, i)
        back = napa.fft_batch(got, direction=-1, scale=True)
        assert S.nre(back, x) <= S.tol(n), (lg, rep)


@pytest.mark.parametrize("n,batch", [(2, 7), (16, 3001), (256, 1300), (1024, 517), (4096, 33)])
def test_bit_reverse_batches(napa, n, batch):
    """several small vectors per CTA (ragged last CTA), complex and real"""
    x = np.stack([S.rand_c(i, n) for i in range(min(batch, 64))] * ((batch + 63) // 64))[:batch]
    p = S.revperm(n)
    assert np.array_equal(napa.bit_reverse_batch(x.copy()), x[:, p])
    xr = np.ascontiguousarray(x.real)
    assert np.array_equal(napa.bit_reverse_batch(xr.copy()), xr[:, p])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [8, 256, 1 << 14])
def test_get_windowed_fft(napa, n):
    S.check_get_windowed_fft(napa, n)


def _sparse_dft_check(napa_binding, lg, in_order, in_place):
    """Device-resident transform of a handful of impulses: X[k] = sum_j a_j W^{p_j k} is known in closed form, so sampled
    bins of a transform far too large for the CPU oracle can be checked to the stated tolerance (1e-13 log2 n)."""
    import torch
    b = napa_binding
    n = 1 << lg
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.Generator(np.random.Philox(lg))
    pos = np.unique(np.concatenate([[0, 1, n - 1, n // 2, n // 2 + 1], rng.integers(0, n, 27)]))
    amp = rng.uniform(-1, 1, len(pos)) + 1j * rng.uniform(-1, 1, len(pos))
    x = torch.zeros(n, dtype=torch.complex128, device=dev)
    x[torch.from_numpy(pos).to(dev)] = torch.from_numpy(amp).to(dev)
    y = x if in_place else torch.empty_like(x)
    b.c2c_dev(x.data_ptr(), y.data_ptr(), n, 1, n, 1, 0, int(in_order), None, 0, 0, st)
    torch.cuda.synchronize()
    ks = np.unique(np.concatenate([[0, 1, 2, n - 1, n // 2, n // 4 + 3], rng.integers(0, n, 250)]))
    mem = ks if in_order else S_rev(ks, lg)
    got = y[torch.from_numpy(mem).to(dev)].cpu().numpy()
    # exact phases: reduce p*k mod n in integers before going to floating point
    ph = (pos[None, :].astype(object) * ks[:, None].astype(object)) % n
    want = (amp[None, :] * np.exp(-2j * np.pi * ph.astype(np.float64) / n)).sum(axis=1)
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert err <= S.tol(n), (lg, in_order, err)
    del x, y
    torch.cuda.empty_cache()


def S_rev(k, lg):
    k = np.asarray(k, dtype=np.uint64)
    r = np.zeros_like(k)
    for i in range(lg):
        r |= ((k >> np.uint64(i)) & np.uint64(1)) << np.uint64(lg - 1 - i)
    return r.astype(np.int64)


@pytest.mark.parametrize("lg,in_order", [(26, True), (28, True), (28, False), (30, True)])
def test_very_large_single_transform(napa, lg, in_order):
    """2^26 .. 2^30 on one GPU (three-digit plans, natural order through the scratch copy), sampled bins"""
    import napa_fft3_b200 as P
    _sparse_dft_check(P.default_binding(), lg, in_order, in_place=False)


def test_two_to_the_31_in_place_fallback(napa):
    """2^31 points (32 GiB) in place and in order: too large for the scratch copy (interface.lisp:86 allows sizes up
    to 2^32), so the natural order comes from the in-place bit-reversed pipeline plus the bit-reversal pass"""
    import torch
    import napa_fft3_b200 as P
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs 40 GiB of free device memory")
    _sparse_dft_check(P.default_binding(), 31, True, in_place=True)


def test_concurrent_host_threads(napa):
    """distinct vectors from several host threads at once (the reference allows it, README:811-814): every thread owns its
    staging streams and buffers, only the launch step is serialised"""
    import threading
    from oracle import napa_oracle as O
    n, batch, nthreads = 1 << 12, 300, 4
    xs = [np.stack([S.rand_c(1000 * t + i, n) for i in range(8)] * ((batch + 7) // 8))[:batch].copy() for t in range(nthreads)]
    outs, errs = [None] * nthreads, []

    def work(t):
        try:
            for _ in range(3):
                got = napa.fft_batch(xs[t].copy(), in_order=(t % 2 == 0), scale="sqrt")
                outs[t] = napa.fft_batch(got, direction=-1, in_order=(t % 2 == 0), scale="sqrt")
                r = napa.rfft(np.ascontiguousarray(xs[t][0].real))
                assert S.nre(r, O.rfft(np.ascontiguousarray(xs[t][0].real))) <= S.tol(n)
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    ts = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for t in range(nthreads):
        assert S.nre(outs[t], xs[t]) <= S.tol(n), t


def test_distributed_fft_multi_gpu():
    """the distributed four-step over every GPU of the box (skipped on a single GPU): tools/dist_check.py under
    torchrun checks sampled rows of the result against the CPU oracle and the round trip"""
    import subprocess
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs")
    world = 1 << (ngpu.bit_length() - 1)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "dist_check.py"), "24"],
                         capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0 and "dist ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
