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


@pytest.mark.parametrize("lgN", [16, 22, 26])
def test_distributed_building_blocks_single_rank(napa, lgN):
    """the device kernels of the distributed four-step (column-axis transform with fused global
    twiddle, re-layout, batched rows) on one GPU: P = 1 makes the exchange a copy"""
    import torch
    import napa_fft3_b200 as P
    from napa_fft3_b200.dist import DistributedFFT
    dev = torch.device("cuda:0")
    f = DistributedFFT(lgN, P.default_binding(), None, dev)
    x = S.rand_c(1005, 1 << lgN)
    ref = np.fft.fft(x)
    Xl = f.forward(torch.from_numpy(f.scatter_time(x)).to(dev))
    torch.cuda.synchronize()
    want = ref[f.freq_index()]
    assert S.nre(Xl.cpu().numpy(), want) <= S.tol(1 << lgN)
    back = f.inverse(Xl)
    torch.cuda.synchronize()
    assert S.nre(back.cpu().numpy().reshape(-1), x.reshape(f.N1, f.N2).reshape(-1)) <= S.tol(1 << lgN)


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
