// napafft.cu -- host side of libnapafft_b200.so: C ABI (include/napafft.h), plan/twiddle caches,
// pass planner, page-locked staging.  Replaces the L3 layer of the reference
// (interface.lisp:12-217: generate-fft / %ensure-fft / %ensure-twiddles / %ensure-reverse) and the
// numeric part of L4 (easy-interface.lisp, real.lisp); keyword parsing and window generation stay
// in the host language (napa_fft3_b200/*.py here, lisp/ for SBCL).
#include "napafft_kernels.cuh"
#include "../../include/napafft.h"

#ifndef NAPA_EMULATE
#include <cuda.h>          // CUtensorMap types only; the encoder is fetched with cudaGetDriverEntryPoint (no libcuda link)
#endif
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

using napa::cplx;
using napa::PassParams;

#define NAPA_API extern "C" __attribute__((visibility("default")))

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? NAPA_E_NO_DEVICE : NAPA_E_CUDA, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define OK(call)                   \
    do {                           \
        int rc_ = (call);          \
        if (rc_ != NAPA_OK) return rc_; \
    } while (0)

static std::atomic<uint64_t> g_launches{0};

static inline bool pow2(size_t n) { return n && !(n & (n - 1)); }
static inline int ilog2(size_t n) { int l = 0; while (((size_t)1 << l) < n) l++; return l; }

// ------------------------------------------------------------------------------------------
// per-process device state (one process per GPU)
// ------------------------------------------------------------------------------------------
struct BigTw { cplx *lo = nullptr, *hi = nullptr; int shift = 0; };
struct PairTable { unsigned *dev = nullptr; int count = 0; };

// Staging resources of ONE calling host thread (host-pointer entry points): two streams, two device buffers and
// two page-locked buffers, so that concurrent callers on distinct vectors (README:811-814 of the reference allows
// them) each run their own H2D -> kernels -> D2H pipeline and only meet at the short, mutex-protected launch step.
struct Staging {
    cudaStream_t stream[2] = {};
    cudaEvent_t done[2] = {};
    void *pinned[2] = {}; size_t pinned_bytes = 0;
    void *dbuf[2] = {}; size_t dbuf_bytes = 0;
    void *dwin = nullptr; size_t dwin_bytes = 0;
};

struct State {
    std::mutex mu;                                   // caches, tables, per-stream scratch, kernel attributes, launches
    std::atomic<bool> inited{false};
    std::vector<Staging *> stagings;                 // every thread's staging block (released by napafft_shutdown)
    int device = -1;
    cplx *stage_tw[14] = {};
    std::map<int, BigTw> big;                        // by log2 M
    std::map<std::pair<size_t, int>, cplx *> r2tw;   // (n, dir) -> n/2 entries (reference recurrence)
    std::map<int, PairTable> pairs;                  // bit-reversal tile pairs by width
    std::map<cudaStream_t, std::pair<cplx *, size_t>> scratch;   // natural-order intermediate, one per stream
    size_t chunk_bytes = (size_t)2 << 30; // bytes of transforms per launch of the multi-pass path = natural-order scratch size (2 GiB: +0.7 % over 1 GiB on C2, 4 GiB +0.8 %)
    std::map<cudaStream_t, std::pair<cplx *, size_t>> oa_buf;    // chunk matrix + energies of the overlap-add filter, one per stream
    size_t zero_copy_bytes = 128 << 10;   // host-pointer jobs up to this size run on mapped page-locked memory, no copies
    size_t max_scratch_bytes = (size_t)16 << 30;   // natural-order scratch copy up to this size, else in-place fallback
    size_t fused_unit_bytes = 20u << 20;  // fused kernel: largest unit of intermediate (two units are live in L2)
    int fused_upg = 0, fused_grid = 0;    // tuning overrides: transforms per unit, CTAs (0 = auto)
    int fused_prefetch = 1;               // L2 prefetch of each CTA's next pass-A tile
    int fused_tma = 0;                    // NAPAFFT_FUSED_TMA=1: TMA-fed warp-specialised variant of the fused kernel (2^14 .. 2^16)
    int fused_lead = 0;                   // units pass A runs ahead of pass B (0 = default 2)
    int dist_chunks = 0;                  // distributed four-step: column chunks whose exchange overlaps the next chunk's column FFTs
    int dist_peer_copy = 1;               // exchange by copy-engine transfers into IPC-mapped peer buffers (0: ncclSend / ncclRecv)
    int dist_row_blocks = 0;              // row blocks of the chunks next to the row stage (0 = 4)
    int dist_tail_chunks = 0;             // how many chunks are exchanged row block by row block (0 = 1)
    bool use_fused = false;               // fused single-launch plans for 2^14 .. 2^20 (NAPAFFT_FUSED=1); measured: no faster than one launch per digit pass
    bool attrs_set = false;
    int num_sms = 148;
    std::map<cudaStream_t, std::pair<unsigned char *, size_t>> counters;   // per-stream group-barrier counters
    int *h_error = nullptr, *d_error = nullptr;                             // mapped error flag
};
static State G;
static thread_local Staging *tl_staging = nullptr;
static Staging &staging() {
    if (!tl_staging) {
        tl_staging = new Staging();
        std::lock_guard<std::mutex> lk(G.mu);
        G.stagings.push_back(tl_staging);
    }
    return *tl_staging;
}
// every exported function: library bound to a device, and THIS thread's current device is that device
static int napafft_enter();
#define LOCKED(expr) [&]() -> int { std::lock_guard<std::mutex> lk_(G.mu); return (expr); }()

static int ensure_init() {
    if (G.inited.load()) return NAPA_OK;
    return napafft_init(-1);
}
static int napafft_enter() {
    OK(ensure_init());
    // the binding is per process, CUDA's current device is per thread: a caller's worker thread starts on device 0
    CU(cudaSetDevice(G.device));
    return NAPA_OK;
}

// W_M^e = exp(-2 pi i e / M), evaluated in long double and rounded once
static inline cplx unit_root(unsigned long long e, unsigned long long M) {
    // reduce to the first octant for accuracy
    const long double PI2 = 6.283185307179586476925286766559005768L;
    e %= M;
    unsigned long long q8 = (e * 8) / M;             // octant (M >= 8 in all uses, or e == 0)
    long double c, s;
    if (M < 8) {
        long double a = PI2 * (long double)e / (long double)M;
        c = cosl(a); s = sinl(a);
    } else {
        unsigned long long oct = M / 8;
        unsigned long long r = e - q8 * oct;          // 0 <= r < M/8
        bool odd = q8 & 1;
        long double a = PI2 * (long double)(odd ? oct - r : r) / (long double)M;  // in [0, pi/4]
        long double ca = cosl(a), sa = sinl(a);
        if (odd) std::swap(ca, sa);                   // angle = (q8+1)*pi/4 - a'
        // now (ca, sa) = cos/sin of the angle within quadrant (q8/2), offset (q8&1 ? pi/2 - a' : a)
        switch (q8 >> 1) {
        case 0: c = ca; s = sa; break;
        case 1: c = -sa; s = ca; break;
        case 2: c = -ca; s = -sa; break;
        default: c = sa; s = -ca; break;
        }
    }
    return make_double2((double)c, (double)(-s));
}

// Table uploads happen once per table.  cudaMemcpy from pageable memory may return before the DMA
// has landed and does not order against cudaStreamNonBlocking streams, so fence explicitly.
static int upload_table(void *dev, const void *host, size_t bytes) {
    CU(cudaMemcpy(dev, host, bytes, cudaMemcpyHostToDevice));
    CU(cudaDeviceSynchronize());
    return NAPA_OK;
}

static int ensure_stage_tw(int l) {
    if (G.stage_tw[l]) return NAPA_OK;
    size_t L = (size_t)1 << l;
    std::vector<cplx> h(L);
    for (size_t m = 0; m < L; m++) h[m] = unit_root(m, L);
    CU(cudaMalloc(&G.stage_tw[l], L * sizeof(cplx)));
    OK(upload_table(G.stage_tw[l], h.data(), L * sizeof(cplx)));
    return NAPA_OK;
}

static int ensure_big_tw(int lgM, BigTw *out) {
    auto it = G.big.find(lgM);
    if (it != G.big.end()) { *out = it->second; return NAPA_OK; }
    BigTw t;
    t.shift = (lgM + 1) / 2;
    unsigned long long M = 1ull << lgM;
    size_t nlo = (size_t)1 << t.shift, nhi = (size_t)1 << (lgM - t.shift);
    std::vector<cplx> lo(nlo), hi(nhi);
    for (size_t i = 0; i < nlo; i++) lo[i] = unit_root(i, M);
    for (size_t i = 0; i < nhi; i++) hi[i] = unit_root((unsigned long long)i << t.shift, M);
    CU(cudaMalloc(&t.lo, nlo * sizeof(cplx)));
    CU(cudaMalloc(&t.hi, nhi * sizeof(cplx)));
    OK(upload_table(t.lo, lo.data(), nlo * sizeof(cplx)));
    OK(upload_table(t.hi, hi.data(), nhi * sizeof(cplx)));
    G.big[lgM] = t;
    *out = t;
    return NAPA_OK;
}

// real.lisp:53-64 %get-radix-2-twiddle, op for op (separately rounded complex product, libm
// cos/sin of the same argument) so that rfft/rifft track the reference's own rounding drift.
static void radix2_twiddle_host(size_t n, int dir, cplx *tw) {
    const double pi = 3.141592653589793;
    volatile double a = ((-2.0 * pi) * (double)dir) * (1.0 / (double)n);
    const double mr = cos(a), mi = sin(a);
    double rr = dir > 0 ? 1.0 : 0.0, ri = dir > 0 ? 0.0 : 1.0;
    for (size_t k = 0; k < n / 2; k++) {
        tw[k] = make_double2(rr, ri);
        volatile double p0 = rr * mr, p1 = ri * mi, p2 = rr * mi, p3 = ri * mr;   // no FMA contraction
        rr = p0 - p1; ri = p2 + p3;
    }
}

static int ensure_r2tw(size_t n, int dir, cplx **out) {
    auto key = std::make_pair(n, dir);
    auto it = G.r2tw.find(key);
    if (it != G.r2tw.end()) { *out = it->second; return NAPA_OK; }
    std::vector<cplx> h(std::max<size_t>(n / 2, 1));
    radix2_twiddle_host(n, dir, h.data());
    cplx *d;
    CU(cudaMalloc(&d, h.size() * sizeof(cplx)));
    OK(upload_table(d, h.data(), h.size() * sizeof(cplx)));
    G.r2tw[key] = d;
    *out = d;
    return NAPA_OK;
}

// One scratch area PER STREAM: launches queued on different streams (the two staging streams of the
// host-pointer entry points, or a caller's own streams) run concurrently and must not share it.
static int ensure_scratch(cudaStream_t st, size_t bytes, cplx **out) {
    auto &slot = G.scratch[st];
    if (slot.second < bytes) {
        if (slot.first) { CU(cudaDeviceSynchronize()); CU(cudaFree(slot.first)); slot.first = nullptr; slot.second = 0; }
        CU(cudaMalloc((void **)&slot.first, bytes));
        slot.second = bytes;
    }
    *out = slot.first;
    return NAPA_OK;
}

// ------------------------------------------------------------------------------------------
// kernel launch table
// ------------------------------------------------------------------------------------------
// tile kinds (lane maps): 0 column-like, 1 row-like, 2 row-like load + transposed store
enum { KIND_COL = 0, KIND_ROW = 1, KIND_ROWT = 2 };
typedef void (*pass_fn)(const PassParams);
// The kernel templates are instantiated in napafft_tables.cu, which the build compiles once per table slice
// (NAPA_TABLE = pass0 / pass1 / pass2 / fused / fused_tma) in parallel: one translation unit took five minutes.
namespace napa {
pass_fn pass_kernel_table_v0(int l, int kind);
pass_fn pass_kernel_table_v1(int l, int kind);
pass_fn pass_kernel_table_v2(int l, int kind);
typedef void (*fused_fn)(const FusedParams);
typedef void (*fused_tma_fn)(const FusedTmaParams);
fused_fn fused_kernel_table(int la, int lb, int pk, bool win);
fused_tma_fn fused_tma_kernel_table(int la, int lb, int pk, bool win);
}
// variant: 0 every fusion path, 1 lean, 2 lean + window (napafft_kernels.cuh: variant_flags)
static int pass_variant(unsigned flags) {
    if ((flags & ~napa::LEAN_FLAGS) == 0) return 1;
    if ((flags & ~napa::WINLEAN_FLAGS) == 0) return 2;
    return 0;
}
static pass_fn pass_kernel(int l, int kind, int variant) {
    if (l < 7 && kind != KIND_ROW) kind = KIND_COL;   // small L: one lane map (see kind_lm); KIND_ROW = dense-rows variant
    return variant == 1 ? napa::pass_kernel_table_v1(l, kind) : (variant == 2 ? napa::pass_kernel_table_v2(l, kind) : napa::pass_kernel_table_v0(l, kind));
}

static int set_kernel_attrs() {
    if (G.attrs_set) return NAPA_OK;
    for (int l = 4; l <= 13; l++)
        for (int kind = 0; kind < 3; kind++) {
            size_t sm = napa::kind_smem_bytes(l, kind);
            if (sm > 48 * 1024) {
                for (int v = 0; v < 3; v++)
                    CU(cudaFuncSetAttribute((const void *)pass_kernel(l, kind, v), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            }
        }
    G.attrs_set = true;
    return NAPA_OK;
}

static int launch_pass(int l, int kind, PassParams &p, long long ntiles, cudaStream_t st) {
    OK(ensure_stage_tw(l));
    p.stage_tw = G.stage_tw[l];
    p.lg_tpi = ilog2((size_t)p.tiles_per_item); p.lg_G = ilog2((size_t)p.G);
    if (ntiles <= 0) return NAPA_OK;
    if (ntiles > 0x7fffffffLL) return fail(NAPA_E_UNSUPPORTED, "grid too large (%lld tiles)", ntiles);
    if (l < 7) {
        // dense rows: back-to-back single-pass transforms, read and written as whole contiguous tiles
        const bool dense = kind == KIND_ROW && (p.flags & napa::PF_Q_IS_BATCH) && p.src_batch_stride == (1LL << l) &&
                           p.dst_batch_stride == (1LL << l) && !(p.flags & napa::PF_RIFFT_PRE);
        kind = dense ? KIND_ROW : KIND_COL;
    }
    pass_fn fn = pass_kernel(l, kind, pass_variant(p.flags));
    NAPA_LAUNCH(fn, (unsigned)ntiles, napa::plan_threads(l), napa::kind_smem_bytes(l, kind), st, p);
    g_launches++;
    return NAPA_OK;
}

// fused two-pass kernels: (first-executed digit, second digit, plan kind, window fusions compiled in)
using napa::fused_fn;
using napa::fused_tma_fn;
static fused_fn fused_kernel(int la, int lb, int pk, bool win) { return napa::fused_kernel_table(la, lb, pk, win); }
// TMA-fed variant: 2048-point tiles only (a column tile is one tensor-map box of <= 256 rows)
static fused_tma_fn fused_tma_kernel(int la, int lb, int pk, bool win) { return napa::fused_tma_kernel_table(la, lb, pk, win); }
#ifndef NAPA_EMULATE
typedef CUresult (*tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// tensor map of a column pass's operand: doubles [item][row i][2 * col], one box = one tile = L rows x C columns
static int encode_col_tmap(unsigned long long *out, const PassParams &cp, int l, long long items) {
    static tmap_encode_fn enc = nullptr;
    if (!enc) {
        void *fp = nullptr; cudaDriverEntryPointQueryResult qr;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
        if (!fp || qr != cudaDriverEntryPointSuccess) return fail(NAPA_E_CUDA, "cuTensorMapEncodeTiled is not available");
        enc = (tmap_encode_fn)fp;
    }
    const cuuint64_t B = (cuuint64_t)cp.s_i, L = 1ull << l, C = 1ull << napa::plan_log2c(l);
    const cuuint64_t gdim[3] = { 2 * B, L, (cuuint64_t)items };
    const cuuint64_t gstr[2] = { B * sizeof(cplx), (cuuint64_t)cp.src_batch_stride * sizeof(cplx) };
    const cuuint32_t box[3] = { (cuuint32_t)(2 * C), (cuuint32_t)L, 1 };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    CUresult cr = enc((CUtensorMap *)out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)cp.src, gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(NAPA_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
    return NAPA_OK;
}
#endif

// column digit of the fused plan for 2^lg (0 = no fused kernel: the digits would need two tile shapes,
// or two units of intermediate would not fit in L2)
static int fused_col_digit(int lg) {
    static const int l1[] = { 7, 7, 8, 0, 9, 9, 10 };      // lg = 14 .. 20
    return lg >= 14 && lg <= 20 ? l1[lg - 14] : 0;
}

static int ensure_counters(cudaStream_t st, size_t bytes, unsigned char **out) {
    auto &slot = G.counters[st];
    if (slot.second < bytes) {
        if (slot.first) { CU(cudaDeviceSynchronize()); CU(cudaFree(slot.first)); }
        size_t cap = std::max<size_t>(bytes, 64 << 10);
        CU(cudaMalloc(&slot.first, cap));
        slot.second = cap;
    }
    if (!G.h_error) {
        CU(cudaHostAlloc((void **)&G.h_error, sizeof(int), cudaHostAllocMapped));
        *G.h_error = 0;
#ifdef NAPA_EMULATE
        G.d_error = G.h_error;
#else
        CU(cudaHostGetDevicePointer((void **)&G.d_error, G.h_error, 0));
#endif
    }
    *out = slot.first;
    return NAPA_OK;
}

// A wait inside the fused kernel that timed out leaves its flag set (never expected: it means a CTA of a
// cooperative launch was lost).  Checked at the start of every later call and by napafft_sync paths.
static int check_device_error() {
    if (G.h_error && *G.h_error) {
        *G.h_error = 0;
        return fail(NAPA_E_CUDA, "a fused-kernel dependency wait timed out; results of the previous launch are invalid");
    }
    return NAPA_OK;
}

// resident CTAs of a fused kernel (cached)
static int fused_resident(const void *fn, int threads, size_t smem, long long *out) {
    static std::map<const void *, int> occ;
    auto it = occ.find(fn);
    if (it == occ.end()) {
        if (smem > 48 * 1024) CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, threads, smem));
        if (nb < 1) return fail(NAPA_E_CUDA, "fused kernel does not fit on an SM");
        it = occ.emplace(fn, nb).first;
    }
    *out = (long long)G.num_sms * it->second;
    return NAPA_OK;
}

// One cooperative launch over the whole batch.  The grid is every co-resident CTA; a unit is about one tile
// per CTA and pass, capped so that two units of intermediate stay well inside L2 (G.fused_unit_bytes).
static int launch_fused(int la, int lb, int pk, bool win, size_t n, size_t batch, napa::FusedParams &fp, bool ring,
                        cudaStream_t st) {
    fused_fn fn = fused_kernel(la, lb, pk, win);
    if (!fn) return fail(NAPA_E_UNSUPPORTED, "no fused kernel for digits (%d, %d)", la, lb);
    // TMA-fed warp-specialised variant where it exists; real-input fusions (zip, rifft pre-twiddle) read two operands
    // per point and stay on the direct-load kernel
    fused_tma_fn tfn = G.fused_tma && !(fp.a.flags & (napa::PF_ZIP2 | napa::PF_RIFFT_PRE)) ? fused_tma_kernel(la, lb, pk, win) : nullptr;
    OK(check_device_error());
    OK(ensure_stage_tw(la));
    OK(ensure_stage_tw(lb));
    fp.a.stage_tw = G.stage_tw[la];
    fp.b.stage_tw = G.stage_tw[lb];
    fp.a.lg_tpi = ilog2((size_t)fp.a.tiles_per_item); fp.a.lg_G = ilog2((size_t)fp.a.G);
    fp.b.lg_tpi = ilog2((size_t)fp.b.tiles_per_item); fp.b.lg_G = ilog2((size_t)fp.b.G);
    const size_t smem = tfn ? napa::fused_tma_smem_bytes(la, lb, pk) : napa::fused_smem_bytes(la, lb, pk);
    const int threads = napa::plan_threads(la) + (tfn ? 32 : 0);
    long long W;
    OK(fused_resident(tfn ? (const void *)tfn : (const void *)fn, threads, smem, &W));
    if (G.fused_grid > 0) W = std::min<long long>(W, G.fused_grid);
    const long long tpi = (long long)(n >> napa::plan_log2pts(la));
    long long upg = std::max<long long>(1, (W + tpi - 1) / tpi);
    upg = std::min<long long>(upg, std::max<long long>(1, (long long)(G.fused_unit_bytes / (n * sizeof(cplx)))));
    if (G.fused_upg > 0) upg = G.fused_upg;
    upg = std::min<long long>(upg, (long long)batch);
#ifdef NAPA_EMULATE
    // the emulator keeps every fiber of a cooperative grid alive: a small grid, several tiles per CTA and phase
    W = std::min<long long>(W, 3);
    if (G.fused_upg <= 0) upg = std::min<long long>(2, (long long)batch);
#endif
    const long long nunits = ((long long)batch + upg - 1) / upg;
    if (nunits > 0x3fffffffLL) return fail(NAPA_E_UNSUPPORTED, "batch too large for one fused launch");
    fp.items = (long long)batch; fp.upg = (int)upg; fp.nunits = (int)nunits;
    fp.lead = std::max(1, G.fused_lead > 0 ? G.fused_lead : (tfn ? 2 : 2));
    if (ring) {
        // natural order: ring of lead + 1 unit-sized slots for the intermediate
        cplx *buf;
        OK(ensure_scratch(st, (size_t)((fp.lead + 1) * upg) * n * sizeof(cplx), &buf));
        fp.a.dst = buf; fp.b.src = buf;
    }
    const size_t cbytes = (size_t)nunits * 2 * sizeof(unsigned);
    unsigned char *ctr;
    OK(ensure_counters(st, cbytes, &ctr));
    CU(cudaMemsetAsync(ctr, 0, cbytes, st));
    fp.done_a = (unsigned *)ctr; fp.done_b = fp.done_a + nunits;
    fp.error = G.d_error;
    fp.prefetch = G.fused_prefetch;
    const unsigned grid = (unsigned)std::min<long long>(W, std::max<long long>(1, upg * tpi));
    if (tfn) {
        napa::FusedTmaParams tp{};
        tp.f = fp;
#ifndef NAPA_EMULATE
        if (pk == 2) OK(encode_col_tmap(tp.tmap_b, fp.b, lb, (long long)batch));
        else OK(encode_col_tmap(tp.tmap_a, fp.a, la, (long long)batch));
        void *args[] = { (void *)&tp };
        CU(cudaLaunchCooperativeKernel((const void *)tfn, dim3(grid), dim3(threads), args, smem, st));
#else
        NAPA_LAUNCH_COOP(tfn, grid, threads, smem, st, tp);
#endif
        g_launches++;
        return NAPA_OK;
    }
#ifdef NAPA_EMULATE
    NAPA_LAUNCH_COOP(fn, grid, threads, smem, st, fp);
#else
    // cooperative launch: every CTA of the grid is resident before any of them waits on another
    void *args[] = { (void *)&fp };
    CU(cudaLaunchCooperativeKernel((const void *)fn, dim3(grid), dim3(threads), args, smem, st));
#endif
    g_launches++;
    return NAPA_OK;
}

// ------------------------------------------------------------------------------------------
// planner
// ------------------------------------------------------------------------------------------
static std::vector<int> digits_of(int lg, bool natural = false) {
    if (lg <= 13 && !(lg >= 8 && getenv("NAPAFFT_DIGITS"))) return { lg };
    if (const char *env = getenv("NAPAFFT_DIGITS")) {       // tuning / test override: "8,12" or "7,7,8" (column digits first)
        std::vector<int> v; int sum = 0; bool ok = true;
        for (const char *c = env; *c;) { int x = atoi(c); v.push_back(x); sum += x; while (*c && *c != ',') c++; if (*c) c++; }
        for (size_t i = 0; i < v.size(); i++) {
            ok = ok && v[i] >= 4 && v[i] <= (i + 1 < v.size() ? 10 : 13);
            int rest = 0;
            for (size_t j = i + 1; j < v.size(); j++) rest += v[j];
            if (ok && i + 1 < v.size()) ok = rest >= napa::plan_log2c(v[i]);     // a column tile spans C contiguous columns
        }
        if (ok && sum == lg && v.size() >= 2 && v.size() <= 3) return v;
    }
    if (lg <= 13) return { lg };
    // Measured on B200 with tools/sweep_digits.py (profiles/r01_sweep_digits.txt, r01_sweep_digits2.txt).  Column digits first.
    // Two passes while both sides keep >= 64..128-byte contiguous segments; the natural-order row pass
    // stores C*16-byte transposed segments (C = rows per tile), so it leaves the two-digit plans
    // sooner than the bit-reversed layouts, whose last pass is fully contiguous.  Three-digit plans
    // run every pass at copy speed (~1/3 of the roofline); larger digits go last.
    // The plans of both output orders coincide up to 2^16, at 2^18 and from 2^23.
    switch (lg) {
    case 14: return { 6, 8 };
    case 15: return { 7, 8 };
    case 16: return { 8, 8 };
    case 17: return natural ? std::vector<int>{ 8, 9 } : std::vector<int>{ 9, 8 };
    case 18: return { 8, 10 };
    case 19: return natural ? std::vector<int>{ 9, 10 } : std::vector<int>{ 7, 12 };
    case 20: return natural ? std::vector<int>{ 9, 11 } : std::vector<int>{ 8, 12 };
    case 21: return natural ? std::vector<int>{ 6, 7, 8 } : std::vector<int>{ 9, 12 };
    case 22: return natural ? std::vector<int>{ 7, 7, 8 } : std::vector<int>{ 10, 12 };
    }
    const int base = lg / 3, rem = lg % 3;
    int l1 = base, l2 = base + (rem >= 2), l3 = base + (rem >= 1);
    if (l2 > 10) { l3 += l2 - 10; l2 = 10; }          // 2^31, 2^32: column digits stay <= 2^10
    return { l1, l2, l3 };
}

// interface.lisp:41-48
static double scale_factor(size_t n, int dir, int mode) {
    double fn = (double)n;
    if (mode == NAPA_SCALE_NONE) return 1.0;
    if (mode == NAPA_SCALE_INV) return 1.0 / fn;
    return dir > 0 ? 1.0 / sqrt(fn) : (1.0 / fn) / (1.0 / sqrt(fn));
}

struct C2COpts {
    int dir, in_order;
    double scale;              // numeric factor applied on the first load
    const void *window; int wtype; int wpb;
    bool rifft_pre = false; const cplx *r2tw = nullptr;
    bool rfft_post = false;    // single-pass only: fuse the rfft epilogue into the store (dst stride = 2n)
    const double *zip_im = nullptr;   // single-pass only (%2rfft): src is a REAL vector, zip_im the second one; split epilogue in the store
    // distributed four-step: the rows are read from (split_in) / written to (split_out) the exchange layout
    // [chunk][rank][row][Bc] instead of a dense [row][n] matrix (napafft_kernels.cuh: split_offset); natural order only
    bool split_in = false, split_out = false;
    int sp_lgbc = 0, sp_lgnch = 0; long long sp_chunk_stride = 0, sp_rank_stride = 0;
};
static void apply_split(PassParams &p, const C2COpts &o, bool in) {
    p.flags |= in ? napa::PF_SPLIT_IN : napa::PF_SPLIT_OUT;
    p.sp_lgbc = o.sp_lgbc; p.sp_lgnch = o.sp_lgnch; p.sp_chunk_stride = o.sp_chunk_stride; p.sp_rank_stride = o.sp_rank_stride;
}

// geometry of digit pass s (0-based) of an item with digits d[0..p)
static void pass_geometry(const std::vector<int> &d, int s, PassParams &p, long long *tiles_per_item) {
    int lgA = 0, lgB = 0;
    for (int i = 0; i < s; i++) lgA += d[i];
    for (int i = s + 1; i < (int)d.size(); i++) lgB += d[i];
    const int l = d[s];
    const long long A = 1LL << lgA, L = 1LL << l, B = 1LL << lgB, C = 1LL << napa::plan_log2c(l);
    if (lgB > 0) {          // column-like
        p.G = (int)(B / C);
        p.s_SA = L * B; p.s_SG = C; p.s_i = B; p.s_q = 1;
        *tiles_per_item = A * (B / C);
    } else {                // row-like
        p.G = 1;
        p.s_SA = C * L; p.s_SG = 0; p.s_i = 1; p.s_q = L;
        *tiles_per_item = A / C;
    }
    p.d_SA = p.s_SA; p.d_SG = p.s_SG; p.d_i = p.s_i; p.d_q = p.s_q;
    p.tiles_per_item = (int)*tiles_per_item;
    p.tw_mul_t = 1; p.tw_mul_a = 0; p.tw_col_offset = 0; p.tw_col_shift = 0;
}

static void first_pass_fusions(PassParams &p, const C2COpts &o, size_t n, bool win_rev) {
    if (o.scale != 1.0) { p.flags |= napa::PF_SCALE; p.scale = o.scale; }
    if (o.wtype == NAPA_WINDOW_REAL) p.flags |= napa::PF_WIN_REAL;
    if (o.wtype == NAPA_WINDOW_COMPLEX) p.flags |= napa::PF_WIN_CPLX;
    if (o.wtype) {
        p.window = o.window;
        p.window_batch_stride = o.wpb ? (long long)n : 0;
        if (win_rev) p.flags |= napa::PF_WIN_REV;
    }
    p.lgN = ilog2(n);
    if (o.rifft_pre) { p.flags |= napa::PF_RIFFT_PRE; p.r2tw = o.r2tw; p.aux_off = (long long)n; }
}

static int exec_bitrev(const void *src, void *dst, size_t n, size_t batch, size_t sstride, size_t dstride, int elt,
                       cudaStream_t st);

// The whole c2c family on device pointers.  src/dst strides in complex elements.
static int exec_c2c(const cplx *src, cplx *dst, size_t n, size_t batch, size_t sstride, size_t dstride,
                    const C2COpts &o, cudaStream_t st) {
    if (batch == 0) return NAPA_OK;
    OK(set_kernel_attrs());
    const int lg = ilog2(n);
    const bool inv = o.dir < 0;
    // window index is the bit-reversed position for an in-order inverse (easy-interface.lisp:83-95)
    const bool win_rev = inv && o.in_order;

    if (lg <= 3 && (o.split_in || o.split_out)) return fail(NAPA_E_UNSUPPORTED, "the exchange layout needs rows of at least 16 points");
    if (lg <= 3) {
        napa::TinyParams t{};
        t.src = src; t.dst = dst; t.src_batch_stride = (long long)sstride; t.dst_batch_stride = (long long)dstride;
        t.batch = (long long)batch; t.lgn = lg; t.flags = 0;
        if (o.scale != 1.0) { t.flags |= napa::PF_SCALE; t.scale = o.scale; }
        if (o.wtype == NAPA_WINDOW_REAL) t.flags |= napa::PF_WIN_REAL;
        if (o.wtype == NAPA_WINDOW_COMPLEX) t.flags |= napa::PF_WIN_CPLX;
        t.window = o.window; t.window_batch_stride = o.wpb ? (long long)n : 0;
        if (win_rev) t.flags |= napa::PF_WIN_REV;
        if (inv) t.flags |= napa::PF_CONJ_IN | napa::PF_CONJ_OUT;
        if (inv && !o.in_order) t.flags |= napa::PF_IN_REV;
        if (!inv && !o.in_order) t.flags |= napa::PF_OUT_REV;
        if (o.rifft_pre) { t.flags |= napa::PF_RIFFT_PRE; t.r2tw = o.r2tw; }
        NAPA_LAUNCH(napa::fft_tiny_kernel, (unsigned)((batch + 255) / 256), 256, 0, st, t);
        g_launches++;
        CU(cudaGetLastError());
        return NAPA_OK;
    }

    std::vector<int> d = digits_of(lg, o.in_order != 0);
    const int p = (int)d.size();

    if (p == 1) {
        PassParams pp{};
        pp.src = src; pp.dst = dst; pp.src_batch_stride = (long long)sstride; pp.dst_batch_stride = (long long)dstride;
        pp.batch = (long long)batch; pp.tiles_per_item = 1; pp.G = 1;
        pp.s_i = pp.d_i = 1;
        pp.flags = napa::PF_Q_IS_BATCH;
        if (inv) pp.flags |= napa::PF_CONJ_IN | napa::PF_CONJ_OUT;
        if (inv && !o.in_order) pp.flags |= napa::PF_IN_REV;
        if (!inv && !o.in_order) pp.flags |= napa::PF_OUT_REV;
        first_pass_fusions(pp, o, n, win_rev);
        if (o.rfft_post) { pp.flags |= napa::PF_RFFT_POST; pp.r2tw = o.r2tw; }
        if (o.zip_im) { pp.flags |= napa::PF_ZIP2 | napa::PF_SPLIT2_POST; pp.src_im = o.zip_im; }
        if (o.split_in) apply_split(pp, o, true);
        if (o.split_out) apply_split(pp, o, false);
        const long long C = 1LL << napa::plan_log2c(lg);
        OK(launch_pass(lg, KIND_ROW, pp, ((long long)batch + C - 1) / C, st));
        CU(cudaGetLastError());
        return NAPA_OK;
    }

    // items per chunk: keep a chunk's working set L2-resident across its passes
    size_t chunk = std::max<size_t>(1, G.chunk_bytes / (n * sizeof(cplx)));
    chunk = std::min(chunk, batch);

    const bool natural = o.in_order != 0;
    if ((o.split_in || o.split_out) && (p > 2 || !natural))
        return fail(NAPA_E_UNSUPPORTED, "the exchange layout is supported for natural-order rows of 2^4 .. 2^22 points");
    const int fcol = G.use_fused && !o.split_in && !o.split_out ? fused_col_digit(lg) : 0;
    if (fcol && !getenv("NAPAFFT_DIGITS")) {
        // one persistent launch for the whole batch, the intermediate stays in L2 (napafft_kernels.cuh: fft_fused_kernel)
        const std::vector<int> fd = { fcol, lg - fcol };
        napa::FusedParams fp{};
        const int pk = natural ? 0 : (inv ? 2 : 1);
        const bool win = o.wtype != 0 || o.zip_im || o.rifft_pre;
        BigTw bt; OK(ensure_big_tw(lg, &bt));
        long long tpi;
        if (natural) {
            // A: column pass src -> ring slot (+ inter-pass twiddle); B: row pass slot -> dst, transposed store
            PassParams &p1 = fp.a, &p2 = fp.b;
            pass_geometry(fd, 0, p1, &tpi);
            p1.src = src;
            p1.src_batch_stride = (long long)sstride; p1.dst_batch_stride = (long long)n; p1.batch = (long long)batch;
            p1.flags = napa::PF_TW_STORE | (inv ? napa::PF_CONJ_IN : 0);
            p1.tw_lo = bt.lo; p1.tw_hi = bt.hi; p1.tw_shift = bt.shift; p1.tw_mask = (1ull << lg) - 1;
            if (o.zip_im) { p1.flags |= napa::PF_ZIP2; p1.src_im = o.zip_im; }
            first_pass_fusions(p1, o, n, win_rev);
            pass_geometry(fd, 1, p2, &tpi);
            p2.dst = dst;
            p2.src_batch_stride = (long long)n; p2.dst_batch_stride = (long long)dstride; p2.batch = (long long)batch;
            p2.flags = inv ? napa::PF_CONJ_OUT : 0;
            p2.d_SA = 1LL << napa::plan_log2c(fd[1]); p2.d_SG = 0; p2.d_i = 1LL << fd[0]; p2.d_q = 1;
            OK(launch_fused(fd[0], fd[1], 0, win, n, batch, fp, true, st));
        } else {
            // bit-reversed layouts: in place on dst after the first pass, no scratch
            const int sa = inv ? 1 : 0, sb = inv ? 0 : 1;      // DIT runs the digits last to first
            PassParams *ps[2] = { &fp.a, &fp.b };
            const int order[2] = { sa, sb };
            for (int step = 0; step < 2; step++) {
                PassParams &q = *ps[step];
                const int sdig = order[step];
                pass_geometry(fd, sdig, q, &tpi);
                q.src = step == 0 ? src : dst; q.dst = dst;
                q.src_batch_stride = (long long)(step == 0 ? sstride : dstride); q.dst_batch_stride = (long long)dstride;
                q.batch = (long long)batch;
                if (sdig == 0) {   // the column-like digit carries the twiddle
                    q.tw_lo = bt.lo; q.tw_hi = bt.hi; q.tw_shift = bt.shift; q.tw_mask = (1ull << lg) - 1;
                    q.flags |= inv ? napa::PF_TW_LOAD : napa::PF_TW_STORE;
                }
                q.flags |= inv ? napa::PF_IN_REV : napa::PF_OUT_REV;
                if (inv && step == 0) q.flags |= napa::PF_CONJ_IN;
                if (inv && step == 1) q.flags |= napa::PF_CONJ_OUT;
                if (step == 0) first_pass_fusions(q, o, n, false);
            }
            OK(launch_fused(fd[order[0]], fd[order[1]], inv ? 2 : 1, win, n, batch, fp, false, st));
        }
        CU(cudaGetLastError());
        return NAPA_OK;
    }
    // Natural order normally goes through a scratch copy of one chunk; a single transform too large for
    // that (> max_scratch_bytes, 2^31 and up by default) falls back to the in-place bit-reversed
    // pipeline plus the in-place bit-reversal pass, which needs no extra memory.
    const bool nat_scratch = natural && chunk * n * sizeof(cplx) <= G.max_scratch_bytes;
    if ((o.split_in || o.split_out) && !nat_scratch) return fail(NAPA_E_UNSUPPORTED, "the exchange layout needs the scratch copy");
    if (nat_scratch) {
        // natural in -> natural out, one HBM pass per digit and no separate bit-reversal pass:
        //   p == 2: column pass src -> scratch [k1][n2], row pass scratch -> dst stored transposed (k1 + L1 k2)
        //   p == 3: column pass src -> dst [k1][n2][n3]; column pass dst -> scratch [k2][k1][n3]; row pass
        //           scratch -> dst stored transposed ((k1 + L1 k2) + L1 L2 k3)
        // The inverse runs the same forward structure between conjugations (bitwise the DIT result of
        // the bit-reversed input up to the order of additions; window index = rev(k), set by win_rev).
        cplx *scratch; OK(ensure_scratch(st, chunk * n * sizeof(cplx), &scratch));
        for (size_t b0 = 0; b0 < batch; b0 += chunk) {
            const size_t nb = std::min(chunk, batch - b0);
            long long tpi;
            // %2rfft: src is a vector of doubles (strides in doubles), zipped with o.zip_im on the first load
            const cplx *cur = o.zip_im ? (const cplx *)((const double *)src + b0 * sstride) : src + b0 * sstride;
            long long cur_stride = (long long)sstride;
            for (int s = 0; s + 1 < p; s++) {
                int lgB = 0;
                for (int i = s + 1; i < p; i++) lgB += d[i];
                PassParams p1{};
                pass_geometry(d, s, p1, &tpi);
                const bool to_scratch = s == p - 2;
                p1.src = cur; p1.dst = to_scratch ? scratch : dst + b0 * dstride;
                p1.src_batch_stride = cur_stride; p1.dst_batch_stride = (long long)(to_scratch ? n : dstride); p1.batch = (long long)nb;
                BigTw bt; OK(ensure_big_tw(d[s] + lgB, &bt));
                p1.flags = napa::PF_TW_STORE | ((inv && s == 0) ? napa::PF_CONJ_IN : 0);
                p1.tw_lo = bt.lo; p1.tw_hi = bt.hi; p1.tw_shift = bt.shift; p1.tw_mask = (1ull << (d[s] + lgB)) - 1;
                if (s == 0) {
                    if (o.zip_im) { p1.flags |= napa::PF_ZIP2; p1.src_im = o.zip_im + b0 * sstride; }
                    if (o.split_in) apply_split(p1, o, true);
                    first_pass_fusions(p1, o, n, win_rev);
                    if (o.wpb && o.wtype) p1.window = (const char *)o.window + b0 * n * (o.wtype == NAPA_WINDOW_REAL ? 8 : 16);
                }
                if (s == 1) {
                    // a = k1 (stride L3 in scratch), output k2 at stride L1*L3: swaps the two leading digits
                    p1.d_SA = 1LL << d[2]; p1.d_i = (1LL << d[0]) << d[2];
                }
                OK(launch_pass(d[s], KIND_COL, p1, tpi * (long long)nb, st));
                cur = p1.dst; cur_stride = p1.dst_batch_stride;
            }
            PassParams p2{};
            pass_geometry(d, p - 1, p2, &tpi);
            p2.src = scratch; p2.dst = dst + b0 * dstride;
            p2.src_batch_stride = (long long)n; p2.dst_batch_stride = (long long)dstride; p2.batch = (long long)nb;
            p2.flags = inv ? napa::PF_CONJ_OUT : 0;
            // row r = ra*C + q holds the leading output digits (k1 + L1 k2); output k_last -> address r + (n / L_last) * k_last
            p2.d_SA = 1LL << napa::plan_log2c(d[p - 1]); p2.d_SG = 0; p2.d_i = (long long)(n >> d[p - 1]); p2.d_q = 1;
            if (o.split_out) apply_split(p2, o, false);
            OK(launch_pass(d[p - 1], KIND_ROWT, p2, tpi * (long long)nb, st));
        }
        CU(cudaGetLastError());
        return NAPA_OK;
    }

    // bit-reversed pipelines (in place on dst after the first pass), any number of digits
    if (natural && inv && o.rifft_pre)
        return fail(NAPA_E_UNSUPPORTED, "rifft of 2^%d points needs a scratch copy of %zu bytes (NAPAFFT_MAX_SCRATCH_KB)", lg + 1, chunk * n * sizeof(cplx));
    if (natural && inv) {
        // reference order: bit-reverse first, then DIT (easy-interface.lisp:83-95)
        OK(exec_bitrev(src, dst, n, batch, sstride, dstride, NAPA_ELT_COMPLEX, st));
        src = dst; sstride = dstride;
    }
    for (size_t b0 = 0; b0 < batch; b0 += chunk) {
        const size_t nb = std::min(chunk, batch - b0);
        for (int step = 0; step < p; step++) {
            const int s = inv ? p - 1 - step : step;     // DIT runs the digits last to first
            PassParams pp{}; long long tpi;
            pass_geometry(d, s, pp, &tpi);
            const bool first = step == 0, last = step == p - 1;
            pp.src = first ? src + b0 * sstride : dst + b0 * dstride;
            pp.dst = dst + b0 * dstride;
            pp.src_batch_stride = (long long)(first ? sstride : dstride); pp.dst_batch_stride = (long long)dstride;
            pp.batch = (long long)nb;
            if (first && o.zip_im) {   // %2rfft: src is a vector of doubles (strides in doubles), zipped with o.zip_im on the first load
                pp.src = (const cplx *)((const double *)src + b0 * sstride);
                pp.flags |= napa::PF_ZIP2; pp.src_im = o.zip_im + b0 * sstride;
            }
            int lgB = 0;
            for (int i = s + 1; i < p; i++) lgB += d[i];
            if (lgB > 0) {
                BigTw bt; OK(ensure_big_tw(d[s] + lgB, &bt));
                pp.tw_lo = bt.lo; pp.tw_hi = bt.hi; pp.tw_shift = bt.shift; pp.tw_mask = (1ull << (d[s] + lgB)) - 1;
                pp.flags |= inv ? napa::PF_TW_LOAD : napa::PF_TW_STORE;
            }
            pp.flags |= inv ? napa::PF_IN_REV : napa::PF_OUT_REV;
            if (inv && first) pp.flags |= napa::PF_CONJ_IN;
            if (inv && last) pp.flags |= napa::PF_CONJ_OUT;
            if (first) {
                C2COpts oo = o;
                first_pass_fusions(pp, oo, n, false);
                if (o.wpb && o.wtype) pp.window = (const char *)o.window + b0 * n * (o.wtype == NAPA_WINDOW_REAL ? 8 : 16);
            }
            OK(launch_pass(d[s], lgB > 0 ? KIND_COL : KIND_ROW, pp, tpi * (long long)nb, st));
        }
    }
    if (natural && !inv) OK(exec_bitrev(dst, dst, n, batch, dstride, dstride, NAPA_ELT_COMPLEX, st));
    CU(cudaGetLastError());
    return NAPA_OK;
}

// ------------------------------------------------------------------------------------------
// Transform along axis 0 of a row-major [n1][cols] matrix (column passes only), used by the
// distributed four-step: dst[k1][c] = sum_j1 src[j1][c] W_n1^{+-j1 k1}, optionally times the
// global twiddle W_M^{+-(col_offset + c) * k1} (dir > 0: applied to the OUTPUT of the forward
// transform; dir < 0: applied to the INPUT of the inverse transform).  Rows of dst are in natural
// k1 order.  n1 = 2^4 .. 2^20, cols a multiple of the tile width.  src != dst.
// ------------------------------------------------------------------------------------------
static int exec_cols(const cplx *src, size_t src_ld, cplx *dst, size_t dst_ld, size_t n1, size_t cols, int dir, double scale, int lgM,
                     long long col_offset, cudaStream_t st) {
    // src / dst are [n1][cols] blocks of row-major matrices with row pitches src_ld / dst_ld >= cols (a column chunk of a
    // wider local matrix on one side, a dense exchange buffer on the other)
    OK(set_kernel_attrs());
    const int lg1 = ilog2(n1);
    const bool inv = dir < 0;
    std::vector<int> d;
    if (lg1 <= 10) d = { lg1 };
    else { int la = lg1 / 2; d = { la, lg1 - la }; }
    for (int l : d) if (l < 4 || l > 10) return fail(NAPA_E_UNSUPPORTED, "column transform length 2^%d unsupported", lg1);
    for (int l : d) if (cols % ((size_t)1 << napa::plan_log2c(l))) return fail(NAPA_E_UNSUPPORTED, "cols must be a multiple of %d", 1 << napa::plan_log2c(l));
    BigTw gt{}; if (lgM > 0) OK(ensure_big_tw(lgM, &gt));
    const long long B = (long long)cols, SL = (long long)src_ld, DL = (long long)dst_ld;
    auto global_tw = [&](PassParams &p, long long mul_t, long long mul_a) {
        p.tw_lo = gt.lo; p.tw_hi = gt.hi; p.tw_shift = gt.shift; p.tw_mask = (1ull << lgM) - 1;
        p.tw_col_offset = col_offset; p.tw_col_shift = 0; p.tw_mul_t = mul_t; p.tw_mul_a = mul_a;
        p.flags |= inv ? napa::PF_TW_LOAD : napa::PF_TW_STORE;
    };
    if (d.size() == 1) {
        const int l = d[0];
        const long long C = 1LL << napa::plan_log2c(l);
        PassParams p{};
        p.src = src; p.dst = dst; p.batch = 1; p.G = (int)(B / C); p.tiles_per_item = p.G;
        p.s_SA = 0; p.s_SG = C; p.s_i = SL; p.s_q = 1;
        p.d_SA = 0; p.d_SG = C; p.d_i = DL; p.d_q = 1;
        p.tw_mul_t = 1;
        if (inv) p.flags |= napa::PF_CONJ_IN | napa::PF_CONJ_OUT;
        if (scale != 1.0) { p.flags |= napa::PF_SCALE; p.scale = scale; }
        if (lgM > 0) global_tw(p, 1, 0);
        OK(launch_pass(l, KIND_COL, p, p.tiles_per_item, st));
        CU(cudaGetLastError());
        return NAPA_OK;
    }
    // two digits n1 = La * Lb.  Forward (DIF): digit a then digit b; row k1 = ka + La*kb.
    // Inverse (DIT, mirror): digit b first (input rows k1 = ka + La*kb), then digit a.
    const int la = d[0], lb = d[1];
    const long long La = 1LL << la, Lb = 1LL << lb;
    cplx *scratch; OK(ensure_scratch(st, n1 * cols * sizeof(cplx), &scratch));      // dense [n1][cols]
    BigTw it; OK(ensure_big_tw(lg1, &it));
    PassParams pa{}, pb{};
    // digit a: rows ia*Lb + jb, tile = (jb, C columns): outer index ra = jb, intra twiddle W_n1^{ka * jb}.  ld = pitch of the side
    auto digit_a = [&](bool src_side, long long ld, long long &SA, long long &SG, long long &si, long long &sq) {
        (void)src_side; SA = ld; SG = 1LL << napa::plan_log2c(la); si = Lb * ld; sq = 1;
    };
    {
        const long long C = 1LL << napa::plan_log2c(la);
        pa.batch = 1; pa.G = (int)(B / C); pa.tiles_per_item = (int)(Lb * (B / C));
        pa.tw_lo = it.lo; pa.tw_hi = it.hi; pa.tw_shift = it.shift; pa.tw_mask = (1ull << lg1) - 1;
        pa.tw_mul_t = 1; pa.tw_mul_a = 0; pa.tw_col_offset = 0; pa.tw_col_shift = 0; pa.tw_use_ra = 1;
    }
    // digit b over the [La][Lb][.] view (outer index ra = ia / ka), natural row order on the k1 side:
    // position (ka, kb) <-> row ka + La*kb
    {
        const long long C = 1LL << napa::plan_log2c(lb);
        pb.batch = 1; pb.G = (int)(B / C); pb.tiles_per_item = (int)(La * (B / C));
        pb.tw_mul_t = 1;
    }
    if (!inv) {
        pa.src = src; pa.dst = scratch; pa.flags |= napa::PF_TW_STORE;
        digit_a(true, SL, pa.s_SA, pa.s_SG, pa.s_i, pa.s_q);
        digit_a(false, B, pa.d_SA, pa.d_SG, pa.d_i, pa.d_q);
        if (scale != 1.0) { pa.flags |= napa::PF_SCALE; pa.scale = scale; }
        pb.src = scratch; pb.dst = dst;
        pb.s_SA = Lb * B; pb.s_SG = 1LL << napa::plan_log2c(lb); pb.s_i = B; pb.s_q = 1;
        pb.d_SA = DL; pb.d_SG = pb.s_SG; pb.d_i = La * DL; pb.d_q = 1;          // store row ka + La*kb
        if (lgM > 0) global_tw(pb, La, 1);                   // k1 = ra + La*(t + m*T)
        OK(launch_pass(la, KIND_COL, pa, pa.tiles_per_item, st));
        OK(launch_pass(lb, KIND_COL, pb, pb.tiles_per_item, st));
    } else {
        // DIT mirror: first the b digit, reading rows k1 = ka + La*kb, global twiddle on its load
        pb.src = src; pb.dst = scratch;
        pb.s_SA = SL; pb.s_SG = 1LL << napa::plan_log2c(lb); pb.s_i = La * SL; pb.s_q = 1;
        pb.d_SA = Lb * B; pb.d_SG = pb.s_SG; pb.d_i = B; pb.d_q = 1;
        pb.flags |= napa::PF_CONJ_IN;
        if (scale != 1.0) { pb.flags |= napa::PF_SCALE; pb.scale = scale; }
        if (lgM > 0) global_tw(pb, La, 1);
        pa.src = scratch; pa.dst = dst; pa.flags |= napa::PF_TW_LOAD | napa::PF_CONJ_OUT;
        digit_a(true, B, pa.s_SA, pa.s_SG, pa.s_i, pa.s_q);
        digit_a(false, DL, pa.d_SA, pa.d_SG, pa.d_i, pa.d_q);
        OK(launch_pass(lb, KIND_COL, pb, pb.tiles_per_item, st));
        OK(launch_pass(la, KIND_COL, pa, pa.tiles_per_item, st));
    }
    CU(cudaGetLastError());
    return NAPA_OK;
}

// ------------------------------------------------------------------------------------------
// bit reversal
// ------------------------------------------------------------------------------------------
static int ensure_pairs(int w, PairTable *out) {
    auto it = G.pairs.find(w);
    if (it != G.pairs.end()) { *out = it->second; return NAPA_OK; }
    const int A = 5, mb = w - 2 * A;
    std::vector<unsigned> h;
    for (unsigned m = 0; m < (1u << mb); m++) {
        unsigned r = 0;
        for (int i = 0; i < mb; i++) r |= ((m >> i) & 1u) << (mb - 1 - i);
        if (m <= r) { h.push_back(m); h.push_back(r); }
    }
    PairTable t;
    t.count = (int)(h.size() / 2);
    CU(cudaMalloc(&t.dev, h.size() * sizeof(unsigned)));
    OK(upload_table(t.dev, h.data(), h.size() * sizeof(unsigned)));
    G.pairs[w] = t;
    *out = t;
    return NAPA_OK;
}

template <typename E>
static int exec_bitrev_t(const E *src, E *dst, size_t n, size_t batch, size_t sstride, size_t dstride, cudaStream_t st) {
    const int w = ilog2(n);
    if (batch == 0) return NAPA_OK;
    if (w <= 10) {
        const int per_cta = (int)std::max<size_t>(1, 2048 / n);          // 2048 elements (<= 32 KiB) per CTA
        const size_t ctas = (batch + per_cta - 1) / per_cta;
        if (ctas > 0x7fffffffULL) return fail(NAPA_E_UNSUPPORTED, "batch too large");
        auto k = napa::bitrev_small_kernel<E>;
        NAPA_LAUNCH(k, (unsigned)ctas, 256, (size_t)per_cta * (n >= 32 ? n / 32 * 33 : n) * sizeof(E), st, src, dst, w, (long long)sstride, (long long)dstride,
                    (long long)batch, per_cta);
    } else {
        PairTable pt; OK(ensure_pairs(w, &pt));
        long long blocks = (long long)pt.count * (long long)batch;
        if (blocks > 0x7fffffffLL) return fail(NAPA_E_UNSUPPORTED, "grid too large");
        auto k = napa::bitrev_tiled_kernel<E, 5>;
        NAPA_LAUNCH(k, (unsigned)blocks, 256, 0, st, src, dst, w, (long long)sstride, (long long)dstride, pt.count,
                    (const unsigned *)pt.dev);
    }
    g_launches++;
    CU(cudaGetLastError());
    return NAPA_OK;
}

static int exec_bitrev(const void *src, void *dst, size_t n, size_t batch, size_t sstride, size_t dstride, int elt,
                       cudaStream_t st) {
    if (elt == NAPA_ELT_COMPLEX) return exec_bitrev_t<cplx>((const cplx *)src, (cplx *)dst, n, batch, sstride, dstride, st);
    return exec_bitrev_t<double>((const double *)src, (double *)dst, n, batch, sstride, dstride, st);
}

// ------------------------------------------------------------------------------------------
// real wrappers on device pointers
// ------------------------------------------------------------------------------------------
static double real_scale(int mode) { return mode == NAPA_SCALE_NONE ? 1.0 : mode == NAPA_SCALE_INV ? 0.5 : 1.0 / sqrt(2.0); }

static int exec_rfft(const double *src, cplx *dst, size_t n, size_t batch, int scale, cudaStream_t st) {
    const size_t h = n / 2;
    C2COpts o{};
    o.dir = NAPA_FWD; o.in_order = 1; o.scale = real_scale(scale) * scale_factor(h, NAPA_FWD, scale);
    cplx *tw; OK(ensure_r2tw(n, +1, &tw));
    const int lgh = ilog2(h);
    if (lgh >= 7 && lgh <= 13 && !getenv("NAPAFFT_RFFT_UNFUSED")) {
        // one launch: the half-size transform finishes the real spectrum in its store (24 n bytes of traffic instead of 40 n)
        o.rfft_post = true; o.r2tw = tw;
        return exec_c2c((const cplx *)src, dst, h, batch, h, n, o, st);
    }
    OK(exec_c2c((const cplx *)src, dst, h, batch, h, n, o, st));
    const long long per = (long long)(h / 2) + 1, total = per * (long long)batch;
    NAPA_LAUNCH(napa::rfft_post_kernel, (unsigned)((total + 255) / 256), 256, 0, st, dst, (long long)n, (long long)batch,
                (long long)n, (const cplx *)tw);
    g_launches++;
    CU(cudaGetLastError());
    return NAPA_OK;
}

static int exec_rifft(const cplx *src, double *dst, size_t n, size_t batch, int scale, cudaStream_t st) {
    const size_t h = n / 2;
    cplx *tw; OK(ensure_r2tw(n, -1, &tw));
    C2COpts o{};
    o.dir = NAPA_INV; o.in_order = 1; o.scale = real_scale(scale) * scale_factor(h, NAPA_INV, scale);
    o.rifft_pre = true; o.r2tw = tw;
    return exec_c2c(src, (cplx *)dst, h, batch, n, h, o, st);
}

static int exec_2rfft(const double *v1, const double *v2, cplx *dst, size_t n, size_t batch, int scale, cudaStream_t st) {
    const int lgn = ilog2(n);
    if (lgn >= 7 && lgn <= 13 && !getenv("NAPAFFT_RFFT_UNFUSED")) {
        // one launch: the two real vectors are zipped on load and the two spectra separated in the store
        C2COpts z{};
        z.dir = NAPA_FWD; z.in_order = 1; z.scale = scale_factor(n, NAPA_FWD, scale); z.zip_im = v2;
        return exec_c2c((const cplx *)v1, dst, n, batch, n, 2 * n, z, st);
    }
    if (lgn >= 14 && !getenv("NAPAFFT_RFFT_UNFUSED")) {
        // multi-pass sizes: the zip is fused into the first pass's load, the split stays a separate launch
        C2COpts z{};
        z.dir = NAPA_FWD; z.in_order = 1; z.scale = scale_factor(n, NAPA_FWD, scale); z.zip_im = v2;
        OK(exec_c2c((const cplx *)v1, dst, n, batch, n, 2 * n, z, st));
        const long long per2 = (long long)(n / 2) + 1, tot3 = per2 * (long long)batch;
        NAPA_LAUNCH(napa::split2_post_kernel, (unsigned)((tot3 + 255) / 256), 256, 0, st, dst, (long long)(2 * n), (long long)batch,
                    (long long)n);
        g_launches++;
        CU(cudaGetLastError());
        return NAPA_OK;
    }
    const long long total = (long long)n * (long long)batch;
    NAPA_LAUNCH(napa::zip2_kernel, (unsigned)((total + 255) / 256), 256, 0, st, v1, v2, dst, (long long)n, (long long)batch,
                (long long)n, (long long)(2 * n));
    g_launches++;
    C2COpts o{};
    o.dir = NAPA_FWD; o.in_order = 1; o.scale = scale_factor(n, NAPA_FWD, scale);
    OK(exec_c2c(dst, dst, n, batch, 2 * n, 2 * n, o, st));
    const long long per = (long long)(n / 2) + 1, tot2 = per * (long long)batch;
    NAPA_LAUNCH(napa::split2_post_kernel, (unsigned)((tot2 + 255) / 256), 256, 0, st, dst, (long long)(2 * n), (long long)batch,
                (long long)n);
    g_launches++;
    CU(cudaGetLastError());
    return NAPA_OK;
}

// ------------------------------------------------------------------------------------------
// example.lisp:81-114 (filter-chunks) on the device.  vector: len reals; window: chunk reals (the
// triangle window of the reference's windowed-fft call); filter_rev: chunk reals, the filter in
// bit-reversed order; dst: len complex.
// ------------------------------------------------------------------------------------------
static int exec_filter_chunks(const double *vec, size_t len, const double *filter_rev, const double *window, size_t chunk,
                              cplx *dst, cudaStream_t st) {
    if (len == 0) return NAPA_OK;
    const int lgc = ilog2(chunk);
    const size_t half = chunk / 2, nchunks = (len + half - 1) / half;
    const size_t need = nchunks * chunk * sizeof(cplx) + 2 * nchunks * sizeof(double) + 256;
    auto &oa = G.oa_buf[st];
    if (oa.second < need) {
        if (oa.first) { CU(cudaDeviceSynchronize()); CU(cudaFree(oa.first)); oa.first = nullptr; oa.second = 0; }
        CU(cudaMalloc((void **)&oa.first, need));
        oa.second = need;
    }
    cplx *X = oa.first;
    double *e_old = (double *)(X + nchunks * chunk), *e_new = e_old + nchunks;
    NAPA_LAUNCH(napa::oa_gather_kernel, (unsigned)nchunks, 256, 0, st, vec, (long long)len, X, e_old, lgc);
    g_launches++;
    C2COpts f{};
    f.dir = NAPA_FWD; f.in_order = 0; f.scale = 1.0; f.window = window; f.wtype = NAPA_WINDOW_REAL;
    OK(exec_c2c(X, X, chunk, nchunks, chunk, chunk, f, st));
    C2COpts b{};
    b.dir = NAPA_INV; b.in_order = 0; b.scale = scale_factor(chunk, NAPA_INV, NAPA_SCALE_INV); b.window = filter_rev; b.wtype = NAPA_WINDOW_REAL;
    OK(exec_c2c(X, X, chunk, nchunks, chunk, chunk, b, st));
    NAPA_LAUNCH(napa::oa_energy_kernel, (unsigned)nchunks, 256, 0, st, (const cplx *)X, e_new, lgc);
    NAPA_LAUNCH(napa::oa_add_kernel, (unsigned)((len + 255) / 256), 256, 0, st, (const cplx *)X, (const double *)e_old,
                (const double *)e_new, dst, (long long)len, lgc);
    g_launches += 2;
    CU(cudaGetLastError());
    return NAPA_OK;
}

// ------------------------------------------------------------------------------------------
// argument checks shared by host and device entry points
// ------------------------------------------------------------------------------------------
static int check_c2c_args(const void *src, const void *dst, size_t n, size_t batch, size_t stride, int dir, int scale,
                          const void *window, int wtype) {
    if (!src || !dst) return fail(NAPA_E_NULL, "src/dst must not be NULL");
    if (!pow2(n)) return fail(NAPA_E_NOT_POW2, "transform size %zu is not a power of two", n);
    if (n > ((size_t)1 << 32)) return fail(NAPA_E_UNSUPPORTED, "transform size %zu exceeds 2^32 (interface.lisp:86)", n);
    if (batch > 1 && stride < n) return fail(NAPA_E_SHORT_BUFFER, "stride %zu shorter than the transform (%zu)", stride, n);
    if (dir != NAPA_FWD && dir != NAPA_INV) return fail(NAPA_E_BAD_ENUM, "direction %d is not 1 / -1", dir);
    if (scale < 0 || scale > 2) return fail(NAPA_E_BAD_ENUM, "scaling %d is not 0 / 1 / 2", scale);
    if (wtype < 0 || wtype > 2) return fail(NAPA_E_BAD_ENUM, "windowing %d is not 0 / 1 / 2", wtype);
    if (wtype && !window) return fail(NAPA_E_NULL, "window type %d given with a NULL window", wtype);
    return NAPA_OK;
}

// ------------------------------------------------------------------------------------------
// host staging
// ------------------------------------------------------------------------------------------
static int ensure_staging(size_t chunk_bytes_needed) {
    Staging &S = staging();
    if (!S.stream[0]) {
        for (int i = 0; i < 2; i++) {
            CU(cudaStreamCreateWithFlags(&S.stream[i], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&S.done[i], cudaEventDisableTiming));
        }
    }
    if (S.dbuf_bytes < chunk_bytes_needed) {
        CU(cudaDeviceSynchronize());
        for (int i = 0; i < 2; i++) {
            if (S.dbuf[i]) CU(cudaFree(S.dbuf[i]));
            CU(cudaMalloc(&S.dbuf[i], chunk_bytes_needed));
        }
        S.dbuf_bytes = chunk_bytes_needed;
    }
    return NAPA_OK;
}
static int ensure_pinned(size_t bytes) {
    Staging &S = staging();
    if (S.pinned_bytes >= bytes) return NAPA_OK;
    CU(cudaDeviceSynchronize());
    for (int i = 0; i < 2; i++) {
        if (S.pinned[i]) CU(cudaFreeHost(S.pinned[i]));
        CU(cudaHostAlloc(&S.pinned[i], bytes, cudaHostAllocDefault));
    }
    S.pinned_bytes = bytes;
    return NAPA_OK;
}
static bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
static int upload_window(const void *window, int wtype, size_t count, const void **dev) {
    Staging &S = staging();
    *dev = nullptr;
    if (!wtype) return NAPA_OK;
    size_t bytes = count * (wtype == NAPA_WINDOW_REAL ? 8 : 16);
    if (S.dwin_bytes < bytes) {
        CU(cudaDeviceSynchronize());
        if (S.dwin) CU(cudaFree(S.dwin));
        CU(cudaMalloc(&S.dwin, bytes));
        S.dwin_bytes = bytes;
    }
    CU(cudaMemcpyAsync(S.dwin, window, bytes, cudaMemcpyHostToDevice, S.stream[0]));
    CU(cudaStreamSynchronize(S.stream[0]));
    *dev = S.dwin;
    return NAPA_OK;
}

// Generic chunked host pipeline: for every chunk of items, H2D the inputs (item b at
// src + b*in_pitch) into a dense device region, run `body`, D2H the dense output region.
// Two chunks are in flight (two streams, two device buffers) so copies overlap compute.
struct HostJob {
    const char *src = nullptr, *src2 = nullptr;            // src2: optional second input, same geometry
    size_t in_item_bytes = 0, in_pitch = 0;
    char *dst = nullptr; size_t out_item_bytes = 0, out_pitch = 0;
    bool inplace = false;                                  // output region aliases the input region
    size_t batch = 0;
};
template <typename Body>
static int run_host_job(const HostJob &j, Body body) {
    Staging &S = staging();
    const size_t in_total = j.in_item_bytes * (j.src2 ? 2 : 1);
    const size_t per_item = j.inplace ? std::max(in_total, j.out_item_bytes) : in_total + j.out_item_bytes;
    size_t target = (size_t)256 << 20;                                    // device bytes per chunk
    if (const char *e = getenv("NAPAFFT_STAGE_MB")) target = (size_t)atoll(e) << 20;
    const size_t small_in = in_total * j.batch, small_out = j.out_item_bytes * j.batch;
    if (small_in + small_out <= G.zero_copy_bytes) {
        // Small job (a single 2^10 transform is 32 KiB): no copy launches at all.  The operands are packed into
        // the page-locked ring, the kernels read and write that host memory directly over PCIe (unified
        // addressing), one stream synchronisation, unpack.  Three driver calls instead of five.
        OK(ensure_staging(0));
        const size_t in_span = (small_in + 255) & ~(size_t)255;
        OK(ensure_pinned(j.inplace ? std::max(in_span, small_out) : in_span + small_out));
        char *hin = (char *)S.pinned[0], *hout = j.inplace ? hin : hin + in_span;
        for (int which = 0; which < (j.src2 ? 2 : 1); which++) {
            const char *hs = which ? j.src2 : j.src;
            char *ph = hin + (which ? j.batch * j.in_item_bytes : 0);
            for (size_t i = 0; i < j.batch; i++) memcpy(ph + i * j.in_item_bytes, hs + i * j.in_pitch, j.in_item_bytes);
        }
        cudaStream_t st = S.stream[0];
        OK(body(hin, hout, j.batch, 0, st));
        CU(cudaStreamSynchronize(st));
        for (size_t i = 0; i < j.batch; i++) memcpy(j.dst + i * j.out_pitch, hout + i * j.out_item_bytes, j.out_item_bytes);
        return NAPA_OK;
    }
    const size_t items = std::max<size_t>(1, std::min(j.batch, target / std::max<size_t>(per_item, 1)));
    OK(ensure_staging(items * per_item + 256));
    const bool src_pinned = is_pinned(j.src) && (!j.src2 || is_pinned(j.src2));
    const bool dst_pinned = is_pinned(j.dst);
    if (!src_pinned || !dst_pinned) OK(ensure_pinned(items * std::max(in_total, j.out_item_bytes)));
    size_t pending_out[2] = { 0, 0 }, pending_b0[2] = { 0, 0 };
    auto drain = [&](int slot) -> int {
        if (!pending_out[slot]) return NAPA_OK;
        CU(cudaEventSynchronize(S.done[slot]));
        if (!dst_pinned) {
            const char *ph = (const char *)S.pinned[slot];
            for (size_t i = 0; i < pending_out[slot]; i++)
                memcpy(j.dst + (pending_b0[slot] + i) * j.out_pitch, ph + i * j.out_item_bytes, j.out_item_bytes);
        }
        pending_out[slot] = 0;
        return NAPA_OK;
    };
    int slot = 0;
    for (size_t b0 = 0; b0 < j.batch; b0 += items, slot ^= 1) {
        const size_t nb = std::min(items, j.batch - b0);
        OK(drain(slot));
        cudaStream_t st = S.stream[slot];
        char *dev_in = (char *)S.dbuf[slot];
        char *dev_out = j.inplace ? dev_in : dev_in + ((nb * in_total + 255) & ~(size_t)255);    // keep the output 256-byte aligned
        for (int which = 0; which < (j.src2 ? 2 : 1); which++) {
            const char *hs = which ? j.src2 : j.src;
            char *dd = dev_in + (which ? nb * j.in_item_bytes : 0);
            if (src_pinned) {
                CU(cudaMemcpy2DAsync(dd, j.in_item_bytes, hs + b0 * j.in_pitch, j.in_pitch, j.in_item_bytes, nb,
                                     cudaMemcpyHostToDevice, st));
            } else {
                char *ph = (char *)S.pinned[slot] + (which ? nb * j.in_item_bytes : 0);
                for (size_t i = 0; i < nb; i++) memcpy(ph + i * j.in_item_bytes, hs + (b0 + i) * j.in_pitch, j.in_item_bytes);
                CU(cudaMemcpyAsync(dd, ph, nb * j.in_item_bytes, cudaMemcpyHostToDevice, st));
            }
        }
        OK(body(dev_in, dev_out, nb, b0, st));
        if (dst_pinned) {
            CU(cudaMemcpy2DAsync(j.dst + b0 * j.out_pitch, j.out_pitch, dev_out, j.out_item_bytes, j.out_item_bytes, nb,
                                 cudaMemcpyDeviceToHost, st));
        } else {
            CU(cudaMemcpyAsync(S.pinned[slot], dev_out, nb * j.out_item_bytes, cudaMemcpyDeviceToHost, st));
        }
        CU(cudaEventRecord(S.done[slot], st));
        pending_out[slot] = nb; pending_b0[slot] = b0;
    }
    OK(drain(0));
    OK(drain(1));
    return NAPA_OK;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
NAPA_API int napafft_abi_version(void) { return NAPAFFT_ABI_VERSION; }
NAPA_API const char *napafft_last_error(void) { return g_err; }
NAPA_API uint64_t napafft_launch_count(void) { return g_launches.load(); }

NAPA_API int napafft_init(int device) {
    std::lock_guard<std::mutex> lk(G.mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(NAPA_E_NO_DEVICE, "no CUDA device available (%s); napa-fft3-b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0) { if (G.inited) return NAPA_OK; CU(cudaGetDevice(&device)); }
    if (device >= count) return fail(NAPA_E_NO_DEVICE, "device %d requested but only %d present", device, count);
    if (G.inited && G.device != device) return fail(NAPA_E_UNSUPPORTED, "already bound to device %d (one process per GPU)", G.device);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(NAPA_E_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    // tunables: defaults first (a re-init after napafft_shutdown must not inherit the previous settings)
    { State d; G.chunk_bytes = d.chunk_bytes; G.zero_copy_bytes = d.zero_copy_bytes; G.max_scratch_bytes = d.max_scratch_bytes;
      G.fused_unit_bytes = d.fused_unit_bytes; G.fused_upg = d.fused_upg; G.fused_grid = d.fused_grid; G.use_fused = d.use_fused;
      G.fused_prefetch = d.fused_prefetch; G.fused_tma = d.fused_tma; G.fused_lead = d.fused_lead; G.dist_chunks = d.dist_chunks; G.dist_peer_copy = d.dist_peer_copy; G.dist_row_blocks = d.dist_row_blocks; G.dist_tail_chunks = d.dist_tail_chunks; }
    if (const char *env = getenv("NAPAFFT_CHUNK_MB")) G.chunk_bytes = (size_t)atoll(env) << 20;
    if (const char *env = getenv("NAPAFFT_ZERO_COPY_KB")) G.zero_copy_bytes = (size_t)atoll(env) << 10;
    if (const char *env = getenv("NAPAFFT_MAX_SCRATCH_KB")) G.max_scratch_bytes = (size_t)atoll(env) << 10;
    if (const char *env = getenv("NAPAFFT_FUSED_UNIT_MB")) G.fused_unit_bytes = (size_t)atoll(env) << 20;
    if (const char *env = getenv("NAPAFFT_FUSED_UPG")) G.fused_upg = atoi(env);
    if (const char *env = getenv("NAPAFFT_FUSED_GRID")) G.fused_grid = atoi(env);
    if (const char *env = getenv("NAPAFFT_FUSED")) G.use_fused = atoi(env) != 0;
    if (const char *env = getenv("NAPAFFT_FUSED_PREFETCH")) G.fused_prefetch = atoi(env);
    if (const char *env = getenv("NAPAFFT_FUSED_TMA")) G.fused_tma = atoi(env);
    if (const char *env = getenv("NAPAFFT_FUSED_LEAD")) G.fused_lead = atoi(env);
    if (const char *env = getenv("NAPAFFT_DIST_CHUNKS")) G.dist_chunks = atoi(env);
    if (const char *env = getenv("NAPAFFT_DIST_PEER_COPY")) G.dist_peer_copy = atoi(env);
    if (const char *env = getenv("NAPAFFT_DIST_ROW_BLOCKS")) G.dist_row_blocks = atoi(env);
    if (const char *env = getenv("NAPAFFT_DIST_TAIL_CHUNKS")) G.dist_tail_chunks = atoi(env);
    G.num_sms = prop.multiProcessorCount;
    G.device = device;
    G.inited = true;
    return NAPA_OK;
}

static void dist_release();
NAPA_API int napafft_shutdown(void) {
    std::lock_guard<std::mutex> lk(G.mu);
    if (!G.inited) return NAPA_OK;
    cudaDeviceSynchronize();
    dist_release();
    for (int l = 0; l < 14; l++) if (G.stage_tw[l]) { cudaFree(G.stage_tw[l]); G.stage_tw[l] = nullptr; }
    for (auto &kv : G.big) { cudaFree(kv.second.lo); cudaFree(kv.second.hi); }
    G.big.clear();
    for (auto &kv : G.r2tw) cudaFree(kv.second);
    G.r2tw.clear();
    for (auto &kv : G.pairs) cudaFree(kv.second.dev);
    G.pairs.clear();
    for (auto &kv : G.scratch) cudaFree(kv.second.first);
    G.scratch.clear();
    for (auto &kv : G.oa_buf) cudaFree(kv.second.first);
    G.oa_buf.clear();
    for (auto &kv : G.counters) cudaFree(kv.second.first);
    G.counters.clear();
    if (G.h_error) { cudaFreeHost(G.h_error); G.h_error = nullptr; G.d_error = nullptr; }
    for (Staging *S : G.stagings) {
        for (int i = 0; i < 2; i++) {
            if (S->dbuf[i]) cudaFree(S->dbuf[i]);
            if (S->pinned[i]) cudaFreeHost(S->pinned[i]);
            if (S->stream[i]) cudaStreamDestroy(S->stream[i]);
            if (S->done[i]) cudaEventDestroy(S->done[i]);
        }
        if (S->dwin) cudaFree(S->dwin);
        *S = Staging();                              // the owning thread re-creates its resources on its next call
    }
    G.inited = false;
    return NAPA_OK;
}

NAPA_API double napafft_scale_factor(size_t n, int dir, int scale) { return scale_factor(n, dir, scale); }

NAPA_API int napafft_radix2_twiddle(size_t n, int dir, double *out) {
    if (!out) return fail(NAPA_E_NULL, "out must not be NULL");
    if (!pow2(n)) return fail(NAPA_E_NOT_POW2, "size %zu is not a power of two", n);
    radix2_twiddle_host(n, dir, (cplx *)out);
    return NAPA_OK;
}

// ---- device-resident entry points ----
NAPA_API int napafft_c2c_dev(const void *src, void *dst, size_t n, size_t batch, size_t stride, int dir, int scale,
                             int in_order, const void *window, int window_type, int window_per_batch, void *stream) {
    OK(check_c2c_args(src, dst, n, batch, stride, dir, scale, window, window_type));
    if (((uintptr_t)src | (uintptr_t)dst) & 15) return fail(NAPA_E_UNSUPPORTED, "device pointers must be 16-byte aligned");
    if (window_type == NAPA_WINDOW_COMPLEX && ((uintptr_t)window & 15)) return fail(NAPA_E_UNSUPPORTED, "a complex window must be 16-byte aligned");
    if (window_type == NAPA_WINDOW_REAL && ((uintptr_t)window & 7)) return fail(NAPA_E_UNSUPPORTED, "a real window must be 8-byte aligned");
    OK(napafft_enter());
    std::lock_guard<std::mutex> lk(G.mu);
    C2COpts o{};
    o.dir = dir; o.in_order = in_order; o.scale = scale_factor(n, dir, scale);
    o.window = window; o.wtype = window_type; o.wpb = window_per_batch;
    return exec_c2c((const cplx *)src, (cplx *)dst, n, batch, stride, stride, o, (cudaStream_t)stream);
}

NAPA_API int napafft_bit_reverse_dev(const void *src, void *dst, size_t n, size_t batch, size_t stride, int elt, void *stream) {
    if (!src || !dst) return fail(NAPA_E_NULL, "src/dst must not be NULL");
    if (!pow2(n)) return fail(NAPA_E_NOT_POW2, "size %zu is not a power of two", n);
    if (elt != NAPA_ELT_COMPLEX && elt != NAPA_ELT_REAL) return fail(NAPA_E_BAD_ENUM, "element type %d", elt);
    if (batch > 1 && stride < n) return fail(NAPA_E_SHORT_BUFFER, "stride %zu shorter than n %zu", stride, n);
    OK(napafft_enter());
    std::lock_guard<std::mutex> lk(G.mu);
    return exec_bitrev(src, dst, n, batch, stride, stride, elt, (cudaStream_t)stream);
}

static int check_real_args(const void *a, const void *b, size_t n, int scale) {
    if (!a || !b) return fail(NAPA_E_NULL, "src/dst must not be NULL");
    if (!pow2(n)) return fail(NAPA_E_NOT_POW2, "size %zu is not a power of two", n);
    if (n < 2) return fail(NAPA_E_UNSUPPORTED, "real transforms need n >= 2 (real.lisp:10-15 asserts on a size-0 half transform)");
    if (scale < 0 || scale > 2) return fail(NAPA_E_BAD_ENUM, "scaling %d is not 0 / 1 / 2", scale);
    return NAPA_OK;
}

NAPA_API int napafft_rfft_dev(const void *src, void *dst, size_t n, size_t batch, int scale, void *stream) {
    OK(check_real_args(src, dst, n, scale));
    if (((uintptr_t)src | (uintptr_t)dst) & 15) return fail(NAPA_E_UNSUPPORTED, "device pointers must be 16-byte aligned");
    OK(napafft_enter());
    std::lock_guard<std::mutex> lk(G.mu);
    return exec_rfft((const double *)src, (cplx *)dst, n, batch, scale, (cudaStream_t)stream);
}
NAPA_API int napafft_rifft_dev(const void *src, void *dst, size_t n, size_t batch, int scale, void *stream) {
    OK(check_real_args(src, dst, n, scale));
    if (((uintptr_t)src | (uintptr_t)dst) & 15) return fail(NAPA_E_UNSUPPORTED, "device pointers must be 16-byte aligned");
    // item b writes dst[b*n/2 ..] (complex view) while other CTAs still read that region of src: no in-place batches
    if ((const void *)src == (const void *)dst && batch > 1) return fail(NAPA_E_UNSUPPORTED, "napafft_rifft_dev cannot run a batch in place (src == dst)");
    OK(napafft_enter());
    std::lock_guard<std::mutex> lk(G.mu);
    return exec_rifft((const cplx *)src, (double *)dst, n, batch, scale, (cudaStream_t)stream);
}
NAPA_API int napafft_2rfft_dev(const void *v1, const void *v2, void *dst, size_t n, size_t batch, int scale, void *stream) {
    OK(check_real_args(v1, dst, n, scale));
    if (!v2) return fail(NAPA_E_NULL, "v2 must not be NULL");
    if ((((uintptr_t)v1 | (uintptr_t)v2) & 7) || ((uintptr_t)dst & 15)) return fail(NAPA_E_UNSUPPORTED, "v1 / v2 must be 8-byte and dst 16-byte aligned");
    OK(napafft_enter());
    std::lock_guard<std::mutex> lk(G.mu);
    return exec_2rfft((const double *)v1, (const double *)v2, (cplx *)dst, n, batch, scale, (cudaStream_t)stream);
}
NAPA_API int napafft_cmul_dev(void *a, const void *h, size_t n, size_t batch, size_t h_stride, void *stream) {
    if (!a || !h) return fail(NAPA_E_NULL, "a/h must not be NULL");
    OK(napafft_enter());
    const long long total = (long long)n * (long long)batch;
    if (total == 0) return NAPA_OK;
    std::lock_guard<std::mutex> lk(G.mu);
    NAPA_LAUNCH(napa::cmul_pointwise_kernel, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream, (cplx *)a,
                (const cplx *)h, (long long)n, (long long)batch, (long long)h_stride);
    g_launches++;
    CU(cudaGetLastError());
    return NAPA_OK;
}

static int check_filter_args(const void *vec, const void *filter_rev, const void *window, size_t chunk, const void *dst) {
    if (!vec || !filter_rev || !window || !dst) return fail(NAPA_E_NULL, "vector / filter / window / dst must not be NULL");
    if (!pow2(chunk) || chunk < 2) return fail(NAPA_E_NOT_POW2, "chunk size %zu is not a power of two >= 2", chunk);
    if (chunk > ((size_t)1 << 32)) return fail(NAPA_E_UNSUPPORTED, "chunk size %zu exceeds 2^32", chunk);
    return NAPA_OK;
}
NAPA_API int napafft_filter_chunks_dev(const void *vector, size_t len, const void *filter_rev, const void *window, size_t chunk,
                                       void *dst, void *stream) {
    OK(check_filter_args(vector, filter_rev, window, chunk, dst));
    OK(napafft_enter());
    std::lock_guard<std::mutex> lk(G.mu);
    return exec_filter_chunks((const double *)vector, len, (const double *)filter_rev, (const double *)window, chunk, (cplx *)dst,
                              (cudaStream_t)stream);
}
NAPA_API int napafft_filter_chunks(const double *vector, size_t len, const double *filter_rev, const double *window, size_t chunk,
                                   double *dst) {
    OK(check_filter_args(vector, filter_rev, window, chunk, dst));
    if (len == 0) return NAPA_OK;
    OK(napafft_enter());
    OK(ensure_staging(0));
    cudaStream_t st = staging().stream[0];
    double *d_vec = nullptr, *d_f = nullptr, *d_w = nullptr; cplx *d_dst = nullptr;
    int rc = NAPA_OK;
    auto cu = [&](cudaError_t e) { if (e != cudaSuccess && rc == NAPA_OK) rc = fail(NAPA_E_CUDA, "%s", cudaGetErrorString(e)); return e == cudaSuccess; };
    if (cu(cudaMalloc(&d_vec, len * 8)) && cu(cudaMalloc(&d_f, chunk * 8)) && cu(cudaMalloc(&d_w, chunk * 8)) &&
        cu(cudaMalloc(&d_dst, len * sizeof(cplx)))) {
        cu(cudaMemcpyAsync(d_vec, vector, len * 8, cudaMemcpyHostToDevice, st));
        cu(cudaMemcpyAsync(d_f, filter_rev, chunk * 8, cudaMemcpyHostToDevice, st));
        cu(cudaMemcpyAsync(d_w, window, chunk * 8, cudaMemcpyHostToDevice, st));
        if (rc == NAPA_OK) rc = LOCKED(exec_filter_chunks(d_vec, len, d_f, d_w, chunk, d_dst, st));
        if (rc == NAPA_OK) cu(cudaMemcpyAsync(dst, d_dst, len * sizeof(cplx), cudaMemcpyDeviceToHost, st));
        cu(cudaStreamSynchronize(st));
    }
    cudaFree(d_vec); cudaFree(d_f); cudaFree(d_w); cudaFree(d_dst);
    return rc;
}

NAPA_API int napafft_c2c_cols_dev(const void *src, void *dst, size_t n1, size_t cols, int dir, double scale,
                                  int lg_twiddle_modulus, long long col_offset, void *stream) {
    if (!src || !dst) return fail(NAPA_E_NULL, "src/dst must not be NULL");
    if (src == dst) return fail(NAPA_E_UNSUPPORTED, "column transform is out of place");
    if (!pow2(n1) || !pow2(cols)) return fail(NAPA_E_NOT_POW2, "n1 = %zu / cols = %zu must be powers of two", n1, cols);
    if (dir != NAPA_FWD && dir != NAPA_INV) return fail(NAPA_E_BAD_ENUM, "direction %d is not 1 / -1", dir);
    OK(napafft_enter());
    std::lock_guard<std::mutex> lk(G.mu);
    return exec_cols((const cplx *)src, cols, (cplx *)dst, cols, n1, cols, dir, scale, lg_twiddle_modulus, col_offset, (cudaStream_t)stream);
}

// dst[r][g][c] = src[g][r][c]   (g < groups, r < rows, c < cols): re-layout around the all-to-all
NAPA_API int napafft_relayout_dev(const void *src, void *dst, size_t groups, size_t rows, size_t cols, int inverse, void *stream) {
    if (!src || !dst) return fail(NAPA_E_NULL, "src/dst must not be NULL");
    OK(napafft_enter());
    for (size_t g = 0; g < groups; g++) {
        if (!inverse)
            CU(cudaMemcpy2DAsync((char *)dst + g * cols * 16, groups * cols * 16, (const char *)src + g * rows * cols * 16, cols * 16,
                                 cols * 16, rows, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        else
            CU(cudaMemcpy2DAsync((char *)dst + g * rows * cols * 16, cols * 16, (const char *)src + g * cols * 16, groups * cols * 16,
                                 cols * 16, rows, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
    return NAPA_OK;
}

NAPA_API int napafft_dev_alloc(void **out, size_t bytes) {
    if (!out) return fail(NAPA_E_NULL, "out must not be NULL");
    OK(napafft_enter());
    CU(cudaMalloc(out, bytes ? bytes : 16));
    return NAPA_OK;
}
NAPA_API int napafft_dev_free(void *p) { if (p) { OK(napafft_enter()); CU(cudaFree(p)); } return NAPA_OK; }
NAPA_API int napafft_dev_upload(void *d, const void *h, size_t bytes) {
    if (!d || !h) return fail(NAPA_E_NULL, "NULL pointer");
    OK(napafft_enter());
    CU(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice));
    return NAPA_OK;
}
NAPA_API int napafft_dev_download(void *h, const void *d, size_t bytes) {
    if (!d || !h) return fail(NAPA_E_NULL, "NULL pointer");
    OK(napafft_enter());
    CU(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost));
    return NAPA_OK;
}
NAPA_API int napafft_dev_sync(void) {
    OK(napafft_enter());
    CU(cudaDeviceSynchronize());
    return LOCKED(check_device_error());
}

NAPA_API int napafft_host_register(void *ptr, size_t bytes) {
    if (!ptr) return fail(NAPA_E_NULL, "NULL pointer");
    OK(napafft_enter());
    CU(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return NAPA_OK;
}
NAPA_API int napafft_host_unregister(void *ptr) {
    if (!ptr) return fail(NAPA_E_NULL, "NULL pointer");
    CU(cudaHostUnregister(ptr));
    return NAPA_OK;
}

// ------------------------------------------------------------------------------------------
// Distributed four-step: ONE transform of N = 2^lgN points over P processes, one GPU each (BASELINE config 5,
// SURVEY.md 8e).  N = N1 * N2 (N1 = 2^(lgN/2)); x[j1*N2 + j2] is an [N1][N2] matrix.
//   time layout (rank g):  columns j2 in [g*B, (g+1)*B), B = N2/P, a local [N1][B] array;
//   freq layout (rank h):  rows k1 in [h*R, (h+1)*R), R = N1/P, of X[k1 + N1*k2], a local [R][N2] array.
// forward:  the B local columns are cut into `nch` chunks of Bc columns.  Chunk c: column FFTs over j1 with
//           W_N^{j2 k1} fused into their last pass write the dense block send_c = [N1][Bc], whose P slabs of R rows
//           are the messages -- no packing; the all-to-all of chunk c (ncclSend/ncclRecv on the library's
//           communication stream) then overlaps the column FFTs of chunk c + 1.  The rows of the received
//           [chunk][rank][R][Bc] pieces are transformed in place of any re-layout pass: the first row pass reads
//           them through split_offset().
// inverse:  the mirror image (row pass stores into the exchange layout, chunked exchange, twiddled column passes
//           writing the caller's [N1][B] array), scaled by 1/N.
// The exchange itself is a callback so that hosts that own their communicator (Python + torch.distributed in the
// tests) can drive the same pipeline through napafft_dist_cols_dev / napafft_dist_rows_dev.
// ------------------------------------------------------------------------------------------
struct DistGeom { int lgN, lg1, lg2, P, nch; size_t N1, N2, B, R, Bc; };
static int dist_geom(int lgN, int P, int nch, DistGeom *g) {
    if (lgN < 8 || lgN > 40) return fail(NAPA_E_UNSUPPORTED, "distributed transform of 2^%d points unsupported", lgN);
    if (P < 1 || !pow2((size_t)P)) return fail(NAPA_E_UNSUPPORTED, "the number of ranks (%d) must be a power of two", P);
    g->lgN = lgN; g->lg1 = lgN / 2; g->lg2 = lgN - g->lg1; g->P = P;
    g->N1 = (size_t)1 << g->lg1; g->N2 = (size_t)1 << g->lg2;
    if (g->N1 % P || g->N2 % P) return fail(NAPA_E_UNSUPPORTED, "%d ranks do not divide 2^%d x 2^%d", P, g->lg1, g->lg2);
    g->B = g->N2 / P; g->R = g->N1 / P;
    // a chunk holds whole tiles of every pass that touches it (a tile spans at most 128 adjacent columns / points)
    const size_t min_bc = 128;
    if (nch <= 0) nch = P <= 2 ? 8 : 4;
    while (nch > 1 && (g->B % nch || g->B / nch < min_bc)) nch >>= 1;
    if (g->B < min_bc) return fail(NAPA_E_UNSUPPORTED, "fewer than %zu columns per rank (2^%d points over %d ranks)", min_bc, lgN, P);
    if (!pow2((size_t)nch)) return fail(NAPA_E_UNSUPPORTED, "the number of chunks must be a power of two");
    g->nch = nch; g->Bc = g->B / nch;
    return NAPA_OK;
}

// column stage of chunk c.  forward: x_local [N1][B] -> xbuf chunk c; inverse: xbuf chunk c -> x_local
static int dist_cols(const DistGeom &g, int rank, int c, int dir, cplx *x_local, cplx *xbuf, cudaStream_t st) {
    cplx *chunk = xbuf + (size_t)c * g.N1 * g.Bc;
    const long long col0 = (long long)rank * (long long)g.B + (long long)c * (long long)g.Bc;
    if (dir > 0) return exec_cols(x_local + (size_t)c * g.Bc, g.B, chunk, g.Bc, g.N1, g.Bc, +1, 1.0, g.lgN, col0, st);
    return exec_cols(chunk, g.Bc, x_local + (size_t)c * g.Bc, g.B, g.N1, g.Bc, -1, 1.0 / (double)((size_t)1 << g.lgN), g.lgN, col0, st);
}
// row stage.  forward: exchange layout -> X_local [R][N2]; inverse: X_local -> exchange layout (unscaled)
// (rows [rb * R / nrb, (rb + 1) * R / nrb) only: a row is an item, Bc elements apart in the exchange layout)
static int dist_rows(const DistGeom &g, int dir, cplx *xbuf, cplx *X_local, cudaStream_t st, int rb = 0, int nrb = 1) {
    C2COpts o{};
    o.dir = dir; o.in_order = 1; o.scale = 1.0;
    o.sp_lgbc = ilog2(g.Bc); o.sp_lgnch = ilog2((size_t)g.nch);
    o.sp_chunk_stride = (long long)(g.N1 * g.Bc); o.sp_rank_stride = (long long)(g.R * g.Bc);
    const size_t rows = g.R / (size_t)nrb, r0 = (size_t)rb * rows;
    xbuf += r0 * g.Bc; X_local += r0 * g.N2;
    if (dir > 0) { o.split_in = true; return exec_c2c(xbuf, X_local, g.N2, rows, g.Bc, g.N2, o, st); }
    o.split_out = true;
    return exec_c2c(X_local, xbuf, g.N2, rows, g.N2, g.Bc, o, st);
}

#ifndef NAPA_EMULATE
// NCCL is bound at run time (dlopen by SONAME: a process that already carries libnccl.so.2, e.g. through PyTorch,
// shares that copy), so the library itself has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi NC;
static int nccl_load() {
    if (NC.lib) return NAPA_OK;
    const char *names[] = { getenv("NAPAFFT_NCCL_LIBRARY"), "libnccl.so.2", "libnccl.so" };
    for (const char *nm : names) { if (nm && (NC.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break; }
    if (!NC.lib) return fail(NAPA_E_NCCL, "libnccl.so.2 not found (%s); set NAPAFFT_NCCL_LIBRARY", dlerror());
#define NAPA_NCCL_SYM(field, sym) do { *(void **)(&NC.field) = dlsym(NC.lib, sym); if (!NC.field) { NC.lib = nullptr; return fail(NAPA_E_NCCL, "%s missing from the NCCL library", sym); } } while (0)
    NAPA_NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); NAPA_NCCL_SYM(CommInitRank, "ncclCommInitRank"); NAPA_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    NAPA_NCCL_SYM(GroupStart, "ncclGroupStart"); NAPA_NCCL_SYM(GroupEnd, "ncclGroupEnd"); NAPA_NCCL_SYM(Send, "ncclSend"); NAPA_NCCL_SYM(Recv, "ncclRecv");
    NAPA_NCCL_SYM(GetErrorString, "ncclGetErrorString"); NAPA_NCCL_SYM(AllGather, "ncclAllGather"); NAPA_NCCL_SYM(AllReduce, "ncclAllReduce");
#undef NAPA_NCCL_SYM
    return NAPA_OK;
}
#define NCCL(call)                                                                                  \
    do {                                                                                            \
        ncclResult_t r_ = (call);                                                                   \
        if (r_ != ncclSuccess) return fail(NAPA_E_NCCL, "%s failed: %s", #call, NC.GetErrorString(r_)); \
    } while (0)
#endif

struct DistState {
    int rank = 0, nranks = 0;                    // 0 ranks = not initialised
#ifndef NAPA_EMULATE
    ncclComm_t comm = nullptr;
#endif
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_ready[16] = {}, ev_done[16] = {};
    cplx *send = nullptr, *recv = nullptr; size_t buf_elems = 0;
    // peer-copy exchange: every rank's recv buffer mapped into this process (CUDA IPC); the all-to-all is then P - 1
    // copy-engine transfers over NVLink that use no SM, so it overlaps the column FFTs without competing with them
    cplx *peer_recv[64] = {};
    bool peer_ok = false;
    int *sync_word = nullptr;                    // operand of the tiny all-reduces that fence the copies across ranks
};
static DistState D;

// Map every peer's receive buffer (collective: all ranks allocate the same size in the same call).  Any failure on any
// rank leaves peer_ok false everywhere and the exchange falls back to ncclSend / ncclRecv.
static int dist_map_peers() {
    D.peer_ok = false;
#ifndef NAPA_EMULATE
    if (D.nranks <= 1 || D.nranks > 64 || !G.dist_peer_copy) return NAPA_OK;
    if (!D.sync_word) CU(cudaMalloc((void **)&D.sync_word, 256));
    cudaIpcMemHandle_t mine;
    int ok = cudaIpcGetMemHandle(&mine, D.recv) == cudaSuccess;
    cudaGetLastError();
    unsigned char *hbuf = nullptr;
    CU(cudaMalloc((void **)&hbuf, (size_t)D.nranks * sizeof mine));
    CU(cudaMemcpy(hbuf + (size_t)D.rank * sizeof mine, &mine, sizeof mine, cudaMemcpyHostToDevice));
    NCCL(NC.AllGather(hbuf + (size_t)D.rank * sizeof mine, hbuf, sizeof mine, ncclChar, D.comm, D.comm_stream));
    CU(cudaStreamSynchronize(D.comm_stream));
    std::vector<cudaIpcMemHandle_t> all(D.nranks);
    CU(cudaMemcpy(all.data(), hbuf, (size_t)D.nranks * sizeof mine, cudaMemcpyDeviceToHost));
    CU(cudaFree(hbuf));
    for (int r = 0; r < D.nranks && ok; r++) {
        if (r == D.rank) { D.peer_recv[r] = D.recv; continue; }
        void *pp = nullptr;
        if (cudaIpcOpenMemHandle(&pp, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        D.peer_recv[r] = (cplx *)pp;
    }
    // everyone or no one
    CU(cudaMemcpy(D.sync_word, &ok, sizeof ok, cudaMemcpyHostToDevice));
    NCCL(NC.AllReduce(D.sync_word, D.sync_word, 1, ncclInt, ncclMin, D.comm, D.comm_stream));
    CU(cudaStreamSynchronize(D.comm_stream));
    CU(cudaMemcpy(&ok, D.sync_word, sizeof ok, cudaMemcpyDeviceToHost));
    D.peer_ok = ok != 0;
#endif
    return NAPA_OK;
}
static void dist_unmap_peers() {
#ifndef NAPA_EMULATE
    for (int r = 0; r < 64; r++) {
        if (D.peer_recv[r] && r != D.rank) cudaIpcCloseMemHandle(D.peer_recv[r]);
        D.peer_recv[r] = nullptr;
    }
    cudaGetLastError();
#endif
    D.peer_ok = false;
}
// cross-rank fence on the communication stream: returns (in stream order) once every rank has reached it
static int dist_fence() {
#ifndef NAPA_EMULATE
    if (D.nranks > 1) NCCL(NC.AllReduce(D.sync_word, D.sync_word, 1, ncclInt, ncclSum, D.comm, D.comm_stream));
#endif
    return NAPA_OK;
}

static int dist_buffers(size_t elems) {
    if (D.buf_elems >= elems) return NAPA_OK;
    CU(cudaDeviceSynchronize());
    dist_unmap_peers();
    if (D.send) CU(cudaFree(D.send));
    if (D.recv) CU(cudaFree(D.recv));
    D.send = D.recv = nullptr; D.buf_elems = 0;
    CU(cudaMalloc((void **)&D.send, elems * sizeof(cplx)));
    CU(cudaMalloc((void **)&D.recv, elems * sizeof(cplx)));
    D.buf_elems = elems;
    OK(dist_map_peers());
    return NAPA_OK;
}
static int dist_events() {
    if (D.comm_stream) return NAPA_OK;
    // highest priority: the exchange kernels need only a few SMs but must get them while column passes fill the machine
    int prio_lo = 0, prio_hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CU(cudaStreamCreateWithPriority(&D.comm_stream, cudaStreamNonBlocking, prio_hi));
    for (int i = 0; i < 16; i++) {
        CU(cudaEventCreateWithFlags(&D.ev_ready[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&D.ev_done[i], cudaEventDisableTiming));
    }
    return NAPA_OK;
}
// all-to-all of chunk c on the communication stream: slab h of `from` goes to rank h, slab g of `to` comes from rank g;
// optionally only the rows [rb * R / nrb, (rb + 1) * R / nrb) of every slab (a contiguous piece of it)
static int dist_exchange_chunk(const DistGeom &g, int c, const cplx *from, cplx *to, int rb = 0, int nrb = 1) {
    const size_t slab = g.R * g.Bc, piece = slab / (size_t)nrb;
    const size_t off = (size_t)c * g.N1 * g.Bc + (size_t)rb * piece;
    if (g.P == 1) {
        CU(cudaMemcpyAsync(to + off, from + off, piece * sizeof(cplx), cudaMemcpyDeviceToDevice, D.comm_stream));
        return NAPA_OK;
    }
#ifdef NAPA_EMULATE
    return fail(NAPA_E_NCCL, "no NCCL in the emulation build");
#else
    if (D.peer_ok && to == D.recv) {
        // piece h of my chunk goes straight into rank h's receive buffer, at my rank's place; copy engines, no SM
        for (int i = 0; i < g.P; i++) {
            const int h = (D.rank + i) % g.P;
            CU(cudaMemcpyAsync(D.peer_recv[h] + off + (size_t)D.rank * slab, from + off + (size_t)h * slab, piece * sizeof(cplx),
                               cudaMemcpyDeviceToDevice, D.comm_stream));
        }
        return NAPA_OK;
    }
    NCCL(NC.GroupStart());
    for (int h = 0; h < g.P; h++) {
        NCCL(NC.Send(from + off + (size_t)h * slab, 2 * piece, ncclDouble, h, D.comm, D.comm_stream));
        NCCL(NC.Recv(to + off + (size_t)h * slab, 2 * piece, ncclDouble, h, D.comm, D.comm_stream));
    }
    NCCL(NC.GroupEnd());
    return NAPA_OK;
#endif
}

static int exec_dist(cplx *src_local, cplx *dst_local, int lgN, int dir, int nch, cudaStream_t st) {
    DistGeom g;
    OK(dist_geom(lgN, std::max(D.nranks, 1), nch, &g));
    if (g.nch > 8) return fail(NAPA_E_UNSUPPORTED, "at most 8 chunks");
    OK(dist_events());
    OK(dist_buffers(g.N1 * g.B));
    // Peer copies write into OTHER ranks' receive buffers: fence A (every rank is done reading its buffer: the
    // communication stream is already behind the previous call's compute) before the first copy, fence B (every rank's
    // copies have landed) before the data is used.  ncclSend / ncclRecv pairs synchronise by themselves.
    const bool fenced = D.peer_ok && g.P > 1;
    // The chunk next to the row stage (the last one forward, the first one backward) is exchanged in `nrb` blocks of rows,
    // so that the row FFTs of a block run while the later blocks are still on the wire: the exchange then overlaps the
    // column FFTs at one end and the row FFTs at the other.
    int nrb = G.dist_row_blocks > 0 ? G.dist_row_blocks : 4;
    while (nrb > 1 && (g.R % (size_t)nrb || g.R / (size_t)nrb < 8 || g.nch + nrb > 15)) nrb >>= 1;
    // `tail` chunks are exchanged row block by row block ACROSS those chunks, so that block rb is complete -- and its row
    // FFTs start -- after 1/nrb of that traffic.  Measured (2^30, 8 GPUs, 4 chunks x 4 row blocks): tail 1 4.54 ms,
    // tail 2 4.64 ms, tail 4 6.11 ms (holding chunks back until the last column pass idles the links): default 1.
    const int tail = nrb > 1 ? std::max(1, std::min(g.nch, G.dist_tail_chunks > 0 ? G.dist_tail_chunks : 1)) : 1;
    if (dir > 0) {
        if (fenced) OK(dist_fence());
        for (int c = 0; c < g.nch; c++) {
            OK(dist_cols(g, D.rank, c, +1, src_local, D.send, st));
            CU(cudaEventRecord(D.ev_ready[c], st));
            if (c < g.nch - tail) {
                CU(cudaStreamWaitEvent(D.comm_stream, D.ev_ready[c], 0));
                OK(dist_exchange_chunk(g, c, D.send, D.recv));
            }
        }
        CU(cudaStreamWaitEvent(D.comm_stream, D.ev_ready[g.nch - 1], 0));
        for (int rb = 0; rb < nrb; rb++) {
            for (int c = g.nch - tail; c < g.nch; c++) OK(dist_exchange_chunk(g, c, D.send, D.recv, rb, nrb));
            if (fenced) OK(dist_fence());
            CU(cudaEventRecord(D.ev_done[rb], D.comm_stream));
        }
        for (int rb = 0; rb < nrb; rb++) {
            CU(cudaStreamWaitEvent(st, D.ev_done[rb], 0));
            OK(dist_rows(g, +1, D.recv, dst_local, st, rb, nrb));
        }
    } else {
        if (fenced) OK(dist_fence());
        for (int rb = 0; rb < nrb; rb++) {
            OK(dist_rows(g, -1, D.send, src_local, st, rb, nrb));
            CU(cudaEventRecord(D.ev_ready[rb], st));
            CU(cudaStreamWaitEvent(D.comm_stream, D.ev_ready[rb], 0));
            for (int c = 0; c < tail; c++) OK(dist_exchange_chunk(g, c, D.send, D.recv, rb, nrb));
        }
        for (int c = 0; c < g.nch; c++) {
            if (c >= tail) OK(dist_exchange_chunk(g, c, D.send, D.recv));
            if (c >= tail - 1) {
                // (the first `tail` chunks complete together, with the last row block)
                if (fenced) OK(dist_fence());
                CU(cudaEventRecord(D.ev_done[c], D.comm_stream));
            }
        }
        for (int c = 0; c < g.nch; c++) {
            CU(cudaStreamWaitEvent(st, D.ev_done[c < tail ? tail - 1 : c], 0));
            OK(dist_cols(g, D.rank, c, -1, dst_local, D.recv, st));
        }
    }
    // the buffers are reused by the next call: keep the communication stream behind this call's compute
    CU(cudaEventRecord(D.ev_ready[15], st));
    CU(cudaStreamWaitEvent(D.comm_stream, D.ev_ready[15], 0));
    return NAPA_OK;
}

NAPA_API int napafft_dist_unique_id(void *out128) {
    if (!out128) return fail(NAPA_E_NULL, "out must not be NULL");
#ifdef NAPA_EMULATE
    memset(out128, 0, 128);
    return NAPA_OK;
#else
    OK(nccl_load());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    NCCL(NC.GetUniqueId((ncclUniqueId *)out128));
    return NAPA_OK;
#endif
}
NAPA_API int napafft_dist_init(const void *unique_id128, int rank, int nranks) {
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(NAPA_E_BAD_ENUM, "rank %d of %d", rank, nranks);
    if (!pow2((size_t)nranks)) return fail(NAPA_E_UNSUPPORTED, "the number of ranks (%d) must be a power of two", nranks);
    OK(napafft_enter());
    std::lock_guard<std::mutex> lk(G.mu);
    if (D.nranks) return fail(NAPA_E_UNSUPPORTED, "the distributed context is already initialised (napafft_dist_shutdown first)");
    if (nranks > 1) {
#ifdef NAPA_EMULATE
        return fail(NAPA_E_NCCL, "no NCCL in the emulation build");
#else
        if (!unique_id128) return fail(NAPA_E_NULL, "unique id must not be NULL");
        OK(nccl_load());
        ncclUniqueId id;
        memcpy(&id, unique_id128, sizeof id);
        NCCL(NC.CommInitRank(&D.comm, nranks, id, rank));
#endif
    }
    D.rank = rank; D.nranks = nranks;
    return NAPA_OK;
}
static void dist_release() {
    dist_unmap_peers();
    if (D.sync_word) cudaFree(D.sync_word);
#ifndef NAPA_EMULATE
    if (D.comm) { NC.CommDestroy(D.comm); D.comm = nullptr; }
#endif
    if (D.send) cudaFree(D.send);
    if (D.recv) cudaFree(D.recv);
    if (D.comm_stream) {
        cudaStreamDestroy(D.comm_stream);
        for (int i = 0; i < 16; i++) { cudaEventDestroy(D.ev_ready[i]); cudaEventDestroy(D.ev_done[i]); }
    }
    D = DistState();
}
NAPA_API int napafft_dist_shutdown(void) {
    if (!G.inited.load()) return NAPA_OK;
    OK(napafft_enter());
    std::lock_guard<std::mutex> lk(G.mu);
    cudaDeviceSynchronize();
    dist_release();
    return NAPA_OK;
}
static int check_dist_args(const void *a, const void *b, int dir) {
    if (!a || !b) return fail(NAPA_E_NULL, "src/dst must not be NULL");
    if (a == b) return fail(NAPA_E_UNSUPPORTED, "the distributed transform is out of place");
    if (dir != NAPA_FWD && dir != NAPA_INV) return fail(NAPA_E_BAD_ENUM, "direction %d is not 1 / -1", dir);
    if (((uintptr_t)a | (uintptr_t)b) & 15) return fail(NAPA_E_UNSUPPORTED, "device pointers must be 16-byte aligned");
    return NAPA_OK;
}
NAPA_API int napafft_dist_c2c_dev(const void *src_local, void *dst_local, int lgN, int dir, void *stream) {
    OK(check_dist_args(src_local, dst_local, dir));
    OK(napafft_enter());
    std::lock_guard<std::mutex> lk(G.mu);
    if (!D.nranks) return fail(NAPA_E_NCCL, "napafft_dist_init has not been called");
    return exec_dist((cplx *)src_local, (cplx *)dst_local, lgN, dir, G.dist_chunks, (cudaStream_t)stream);
}
NAPA_API int napafft_dist_c2c(const double *src_local, double *dst_local, int lgN, int dir) {
    if (!src_local || !dst_local) return fail(NAPA_E_NULL, "src/dst must not be NULL");
    if (dir != NAPA_FWD && dir != NAPA_INV) return fail(NAPA_E_BAD_ENUM, "direction %d is not 1 / -1", dir);
    OK(napafft_enter());
    if (!D.nranks) return fail(NAPA_E_NCCL, "napafft_dist_init has not been called");
    if (lgN < 8 || lgN > 40) return fail(NAPA_E_UNSUPPORTED, "distributed transform of 2^%d points unsupported", lgN);
    const size_t bytes = (((size_t)1 << lgN) / (size_t)D.nranks) * sizeof(cplx);
    OK(ensure_staging(0));
    cudaStream_t st = staging().stream[0];
    void *d_in = nullptr, *d_out = nullptr;
    int rc = NAPA_OK;
    auto cu = [&](cudaError_t e) { if (e != cudaSuccess && rc == NAPA_OK) rc = fail(NAPA_E_CUDA, "%s", cudaGetErrorString(e)); return e == cudaSuccess; };
    if (cu(cudaMalloc(&d_in, bytes)) && cu(cudaMalloc(&d_out, bytes))) {
        cu(cudaMemcpyAsync(d_in, src_local, bytes, cudaMemcpyHostToDevice, st));
        if (rc == NAPA_OK) rc = LOCKED(exec_dist((cplx *)d_in, (cplx *)d_out, lgN, dir, G.dist_chunks, st));
        if (rc == NAPA_OK) cu(cudaMemcpyAsync(dst_local, d_out, bytes, cudaMemcpyDeviceToHost, st));
        cu(cudaStreamSynchronize(st));
    }
    cudaFree(d_in); cudaFree(d_out);
    return rc;
}
// how the exchange is carried: 0 = not initialised, 1 = single rank (device copy), 2 = ncclSend / ncclRecv,
// 3 = copy-engine transfers into IPC-mapped peer buffers (decided when the buffers are first allocated)
NAPA_API int napafft_dist_exchange_mode(void) {
    std::lock_guard<std::mutex> lk(G.mu);
    if (!D.nranks) return 0;
    if (D.nranks == 1) return 1;
    return D.peer_ok || (!D.buf_elems && G.dist_peer_copy) ? 3 : 2;
}
NAPA_API int napafft_dist_geometry(int lgN, int nranks, size_t *n1, size_t *n2, size_t *cols_per_rank, size_t *rows_per_rank, int *nchunks) {
    OK(napafft_enter());
    DistGeom g;
    OK(dist_geom(lgN, nranks, LOCKED(G.dist_chunks), &g));
    if (n1) *n1 = g.N1; if (n2) *n2 = g.N2; if (cols_per_rank) *cols_per_rank = g.B; if (rows_per_rank) *rows_per_rank = g.R;
    if (nchunks) *nchunks = g.nch;
    return NAPA_OK;
}
// the two local stages for hosts that carry the exchange themselves: xbuf is the [nch][N1][Bc] exchange buffer
NAPA_API int napafft_dist_cols_dev(void *x_local, void *xbuf, int lgN, int nranks, int rank, int chunk, int dir, void *stream) {
    OK(check_dist_args(x_local, xbuf, dir));
    OK(napafft_enter());
    std::lock_guard<std::mutex> lk(G.mu);
    DistGeom g;
    OK(dist_geom(lgN, nranks, G.dist_chunks, &g));
    if (rank < 0 || rank >= nranks || chunk < 0 || chunk >= g.nch) return fail(NAPA_E_BAD_ENUM, "rank %d / chunk %d out of range", rank, chunk);
    return dist_cols(g, rank, chunk, dir, (cplx *)x_local, (cplx *)xbuf, (cudaStream_t)stream);
}
NAPA_API int napafft_dist_rows_dev(void *xbuf, void *X_local, int lgN, int nranks, int dir, void *stream) {
    OK(check_dist_args(xbuf, X_local, dir));
    OK(napafft_enter());
    std::lock_guard<std::mutex> lk(G.mu);
    DistGeom g;
    OK(dist_geom(lgN, nranks, G.dist_chunks, &g));
    return dist_rows(g, dir, (cplx *)xbuf, (cplx *)X_local, (cudaStream_t)stream);
}

// ---- host-pointer entry points ----
NAPA_API int napafft_c2c(const double *src, double *dst, size_t n, size_t batch, size_t stride, int dir, int scale,
                         int in_order, const void *window, int window_type, int window_per_batch) {
    OK(check_c2c_args(src, dst, n, batch, stride, dir, scale, window, window_type));
    if (batch == 0) return NAPA_OK;
    OK(napafft_enter());
    if (batch == 1) stride = n;
    OK(ensure_staging(16));
    const void *dwin;
    OK(upload_window(window, window_type, window_per_batch ? n * batch : n, &dwin));
    HostJob j;
    j.src = (const char *)src; j.in_item_bytes = n * 16; j.in_pitch = stride * 16;
    j.dst = (char *)dst; j.out_item_bytes = n * 16; j.out_pitch = stride * 16;
    j.inplace = true; j.batch = batch;
    const double sf = scale_factor(n, dir, scale);
    return run_host_job(j, [&](char *din, char *dout, size_t nb, size_t b0, cudaStream_t st) -> int {
        C2COpts o{};
        o.dir = dir; o.in_order = in_order; o.scale = sf;
        o.wtype = window_type; o.wpb = window_per_batch;
        o.window = window_per_batch && dwin ? (const char *)dwin + b0 * n * (window_type == NAPA_WINDOW_REAL ? 8 : 16) : dwin;
        return LOCKED(exec_c2c((const cplx *)din, (cplx *)dout, n, nb, n, n, o, st));
    });
}

static size_t sample_bytes(int src_type) { return src_type == NAPA_SAMPLE_C64 ? 16 : (src_type == NAPA_SAMPLE_F32 ? 4 : 8); }
static int exec_widen(const void *src, int src_type, cplx *dst, size_t total, cudaStream_t st) {
    if (total == 0) return NAPA_OK;
    if (src_type == NAPA_SAMPLE_C64) { CU(cudaMemcpyAsync(dst, src, total * sizeof(cplx), cudaMemcpyDeviceToDevice, st)); return NAPA_OK; }
    NAPA_LAUNCH(napa::widen_kernel, (unsigned)((total + 255) / 256), 256, 0, st, src, src_type, dst, (long long)total);
    g_launches++;
    CU(cudaGetLastError());
    return NAPA_OK;
}
NAPA_API int napafft_widen_dev(const void *src, int src_type, void *dst, size_t count, void *stream) {
    if (!src || !dst) return fail(NAPA_E_NULL, "src/dst must not be NULL");
    if (src_type < 0 || src_type > 3) return fail(NAPA_E_BAD_ENUM, "sample type %d is not 0 .. 3", src_type);
    OK(napafft_enter());
    return LOCKED(exec_widen(src, src_type, (cplx *)dst, count, (cudaStream_t)stream));
}
// napafft_c2c on an input that still needs complex-samplify: the raw samples cross PCIe, the
// widening happens on the device.  Items are dense (n samples apart) on both sides.
NAPA_API int napafft_c2c_typed(const void *src, int src_type, double *dst, size_t n, size_t batch, int dir, int scale,
                               int in_order, const void *window, int window_type, int window_per_batch) {
    OK(check_c2c_args(src, dst, n, batch, n, dir, scale, window, window_type));
    if (src_type < 0 || src_type > 3) return fail(NAPA_E_BAD_ENUM, "sample type %d is not 0 .. 3", src_type);
    if (src_type == NAPA_SAMPLE_C64)
        return napafft_c2c((const double *)src, dst, n, batch, n, dir, scale, in_order, window, window_type, window_per_batch);
    if ((const void *)src == (const void *)dst) return fail(NAPA_E_UNSUPPORTED, "a typed input cannot alias the complex destination");
    if (batch == 0) return NAPA_OK;
    OK(napafft_enter());
    OK(ensure_staging(16));
    const void *dwin;
    OK(upload_window(window, window_type, window_per_batch ? n * batch : n, &dwin));
    const size_t es = sample_bytes(src_type);
    HostJob j;
    j.src = (const char *)src; j.in_item_bytes = n * es; j.in_pitch = n * es;
    j.dst = (char *)dst; j.out_item_bytes = n * 16; j.out_pitch = n * 16;
    j.inplace = false; j.batch = batch;
    const double sf = scale_factor(n, dir, scale);
    return run_host_job(j, [&](char *din, char *dout, size_t nb, size_t b0, cudaStream_t st) -> int {
        OK(LOCKED(exec_widen(din, src_type, (cplx *)dout, nb * n, st)));
        C2COpts o{};
        o.dir = dir; o.in_order = in_order; o.scale = sf;
        o.wtype = window_type; o.wpb = window_per_batch;
        o.window = window_per_batch && dwin ? (const char *)dwin + b0 * n * (window_type == NAPA_WINDOW_REAL ? 8 : 16) : dwin;
        return LOCKED(exec_c2c((const cplx *)dout, (cplx *)dout, n, nb, n, n, o, st));
    });
}

NAPA_API int napafft_bit_reverse(const void *src, void *dst, size_t n, size_t batch, size_t stride, int elt) {
    if (!src || !dst) return fail(NAPA_E_NULL, "src/dst must not be NULL");
    if (!pow2(n)) return fail(NAPA_E_NOT_POW2, "size %zu is not a power of two", n);
    if (elt != NAPA_ELT_COMPLEX && elt != NAPA_ELT_REAL) return fail(NAPA_E_BAD_ENUM, "element type %d", elt);
    if (batch > 1 && stride < n) return fail(NAPA_E_SHORT_BUFFER, "stride %zu shorter than n %zu", stride, n);
    if (batch == 0) return NAPA_OK;
    OK(napafft_enter());
    if (batch == 1) stride = n;
    const size_t es = elt == NAPA_ELT_COMPLEX ? 16 : 8;
    HostJob j;
    j.src = (const char *)src; j.in_item_bytes = n * es; j.in_pitch = stride * es;
    j.dst = (char *)dst; j.out_item_bytes = n * es; j.out_pitch = stride * es;
    j.inplace = true; j.batch = batch;
    return run_host_job(j, [&](char *din, char *dout, size_t nb, size_t, cudaStream_t st) -> int {
        return LOCKED(exec_bitrev(din, dout, n, nb, n, n, elt, st));
    });
}

NAPA_API int napafft_rfft(const double *src, double *dst, size_t n, size_t batch, int scale) {
    OK(check_real_args(src, dst, n, scale));
    if (batch == 0) return NAPA_OK;
    OK(napafft_enter());
    HostJob j;
    j.src = (const char *)src; j.in_item_bytes = n * 8; j.in_pitch = n * 8;
    j.dst = (char *)dst; j.out_item_bytes = n * 16; j.out_pitch = n * 16;
    j.batch = batch;
    return run_host_job(j, [&](char *din, char *dout, size_t nb, size_t, cudaStream_t st) -> int {
        return LOCKED(exec_rfft((const double *)din, (cplx *)dout, n, nb, scale, st));
    });
}

NAPA_API int napafft_rifft(const double *src, double *dst, size_t n, size_t batch, int scale) {
    OK(check_real_args(src, dst, n, scale));
    if (batch == 0) return NAPA_OK;
    OK(napafft_enter());
    HostJob j;
    j.src = (const char *)src; j.in_item_bytes = n * 16; j.in_pitch = n * 16;
    j.dst = (char *)dst; j.out_item_bytes = n * 8; j.out_pitch = n * 8;
    j.batch = batch;
    return run_host_job(j, [&](char *din, char *dout, size_t nb, size_t, cudaStream_t st) -> int {
        return LOCKED(exec_rifft((const cplx *)din, (double *)dout, n, nb, scale, st));
    });
}

NAPA_API int napafft_2rfft(const double *v1, const double *v2, double *dst, size_t n, size_t batch, int scale) {
    OK(check_real_args(v1, dst, n, scale));
    if (!v2) return fail(NAPA_E_NULL, "v2 must not be NULL");
    if (batch == 0) return NAPA_OK;
    OK(napafft_enter());
    HostJob j;
    j.src = (const char *)v1; j.src2 = (const char *)v2; j.in_item_bytes = n * 8; j.in_pitch = n * 8;
    j.dst = (char *)dst; j.out_item_bytes = 2 * n * 16; j.out_pitch = 2 * n * 16;
    j.batch = batch;
    return run_host_job(j, [&](char *din, char *dout, size_t nb, size_t, cudaStream_t st) -> int {
        return LOCKED(exec_2rfft((const double *)din, (const double *)(din + nb * n * 8), (cplx *)dout, n, nb, scale, st));
    });
}
