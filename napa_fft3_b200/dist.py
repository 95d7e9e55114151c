"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed only for control.

Batched transforms (BASELINE configs C2-C4) are independent, so the batch is cut into contiguous
per-rank slices and every rank runs the single-GPU path on its slice: no data-path collective
(SURVEY.md 8e).  `gather_timing` is the max-over-ranks reduction bench.py reports."""
from __future__ import annotations

import numpy as np


def shard_bounds(batch: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of `batch` items owned by `rank`; sizes differ by at most one and
    every item is owned exactly once (ragged batches included)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(x: np.ndarray, rank: int, world: int) -> np.ndarray:
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def sharded_apply(fn, x: np.ndarray, rank: int, world: int, gather=None):
    """Run `fn` (e.g. napa.fft_batch) on this rank's rows.  `gather(local) -> list of per-rank
    arrays` is optional (tests use torch.distributed.all_gather_object; production keeps results
    sharded)."""
    local = fn(np.ascontiguousarray(shard_rows(x, rank, world)))
    if gather is None:
        return local
    return np.concatenate(gather(local), axis=0)


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """bench.py timing rule: a multi-GPU number is the max over ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


_native_ctx = None          # (rank, nranks) of the library-owned communicator of this process


class DistributedFFT:
    """One complex-double transform of N = 2^lgN points spread over the P ranks of a
    torch.distributed group (BASELINE config C5): the four-step N = N1*N2 with ONE
    all-to-all, cut into column chunks so that the exchange of a chunk overlaps the column FFTs of the next.

    Layouts (both are "blocks of a row-major matrix"; no rank ever holds the whole vector):
      time  side: x[j1*N2 + j2] viewed as [N1][N2];  rank g holds columns j2 in [g*B, (g+1)*B),
                  B = N2/P, as a contiguous local [N1][B] array;
      freq  side: X[k1 + N1*k2] viewed as [N1][N2] (k1 = row); rank h holds rows k1 in
                  [h*R, (h+1)*R), R = N1/P, as a contiguous local [R][N2] array (k2 natural).
    forward():  chunk by chunk, local column FFTs over j1 with the twiddle W_N^{j2 k1} fused into the last
                column pass -> all-to-all of contiguous [R][Bc] slabs; then local row FFTs over j2 that read
                the received pieces in place (no re-layout pass).
    inverse():  the same steps backwards, scaled by 1/N.

    Two ways to carry the exchange:
      native (default on CUDA): the C library owns an NCCL communicator (napafft_dist_init; the unique id is
        broadcast once through torch.distributed) and napafft_dist_c2c_dev runs the whole pipeline, the
        exchange on its own stream;
      host-carried (gloo / CPU tests, or native=False): torch.distributed.all_to_all_single per chunk between the
        library's two local stages napafft_dist_cols_dev / napafft_dist_rows_dev.
    """

    def __init__(self, lgN, binding, dist, device, stream=None, native=None):
        import ctypes as C
        import torch
        global _native_ctx
        self.torch, self.b, self.dist, self.device = torch, binding, dist, device
        self.P = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.P > 1 else 0
        self.lgN = lgN
        n1, n2, B, R, nch = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_int()
        binding.dist_geometry(lgN, self.P, C.byref(n1), C.byref(n2), C.byref(B), C.byref(R), C.byref(nch))
        self.N1, self.N2, self.B, self.R, self.nch = n1.value, n2.value, B.value, R.value, nch.value
        self.Bc = self.B // self.nch
        self._stream = stream
        self.native = (device.type == "cuda" or self.P == 1) if native is None else bool(native)
        if self.native:
            if _native_ctx is None:
                uid = (C.c_ubyte * 128)()
                if self.P > 1:
                    if self.rank == 0:
                        binding.dist_unique_id(uid)
                    box = [bytes(uid)]
                    dist.broadcast_object_list(box, src=0)
                    uid = (C.c_ubyte * 128).from_buffer_copy(box[0])
                binding.dist_init(uid, self.rank, self.P)
                _native_ctx = (self.rank, self.P)
            elif _native_ctx != (self.rank, self.P):
                raise ValueError("this process already owns a communicator of another shape")
        else:
            mk = lambda *shape: torch.empty(shape, dtype=torch.complex128, device=device)
            self.xs = mk(self.nch, self.N1, self.Bc)
            self.xr = mk(self.nch, self.N1, self.Bc)

    def _st(self):
        if self._stream is not None:
            return self._stream
        if self.device.type == "cuda":
            return self.torch.cuda.current_stream().cuda_stream
        return None

    def _exchange(self, out, inp):
        """all-to-all of P contiguous equal slabs (complex as pairs of doubles: NCCL has no complex type)"""
        if self.P == 1:
            out.copy_(inp)
            return
        self.dist.all_to_all_single(self.torch.view_as_real(out).view(-1), self.torch.view_as_real(inp).view(-1))

    def forward(self, x_local, out=None):
        """x_local: [N1][B] complex128 (time layout) -> [R][N2] (freq layout), unscaled."""
        t = self.torch
        assert x_local.shape == (self.N1, self.B) and x_local.dtype == t.complex128 and x_local.is_contiguous()
        out = t.empty((self.R, self.N2), dtype=t.complex128, device=self.device) if out is None else out
        st = self._st()
        if self.native:
            self.b.dist_c2c_dev(x_local.data_ptr(), out.data_ptr(), self.lgN, 1, st)
            return out
        for c in range(self.nch):
            self.b.dist_cols_dev(x_local.data_ptr(), self.xs.data_ptr(), self.lgN, self.P, self.rank, c, 1, st)
            self._exchange(self.xr[c], self.xs[c])
        self.b.dist_rows_dev(self.xr.data_ptr(), out.data_ptr(), self.lgN, self.P, 1, st)
        return out

    def inverse(self, X_local, out=None):
        """X_local: [R][N2] (freq layout) -> [N1][B] (time layout), scaled by 1/N."""
        t = self.torch
        assert X_local.shape == (self.R, self.N2) and X_local.dtype == t.complex128 and X_local.is_contiguous()
        out = t.empty((self.N1, self.B), dtype=t.complex128, device=self.device) if out is None else out
        st = self._st()
        if self.native:
            self.b.dist_c2c_dev(X_local.data_ptr(), out.data_ptr(), self.lgN, -1, st)
            return out
        self.b.dist_rows_dev(self.xs.data_ptr(), X_local.data_ptr(), self.lgN, self.P, -1, st)
        for c in range(self.nch):
            self._exchange(self.xr[c], self.xs[c])
            self.b.dist_cols_dev(out.data_ptr(), self.xr.data_ptr(), self.lgN, self.P, self.rank, c, -1, st)
        return out

    def exchange_only(self, steps=1):
        """the all-to-all alone on buffers of the transform's size (bench: NVLink roofline of the exchange)"""
        t = self.torch
        if not hasattr(self, "_xa"):
            self._xa = t.zeros((self.N1, self.B), dtype=t.complex128, device=self.device)
            self._xb = t.empty_like(self._xa)
        for _ in range(steps):
            self._exchange(self._xb, self._xa)

    # helpers for tests / host callers: cut a full host vector into this rank's local arrays
    def scatter_time(self, x_full):
        m = np.asarray(x_full).reshape(self.N1, self.N2)
        return np.ascontiguousarray(m[:, self.rank * self.B:(self.rank + 1) * self.B])

    def freq_index(self):
        """global bin index k of every element of the local freq-side array"""
        k1 = (self.rank * self.R + np.arange(self.R))[:, None]
        k2 = np.arange(self.N2)[None, :]
        return k1 + self.N1 * k2

    def alltoall_bytes_per_rank(self):
        """bytes each rank sends (= receives) over the interconnect per transform"""
        return (self.P - 1) * self.R * self.B * 16
