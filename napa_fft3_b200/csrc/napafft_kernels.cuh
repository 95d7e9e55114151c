// napafft_kernels.cuh -- sm_100a device code for the Napa-FFT3 hot path (complex-double
// power-of-two DFTs, bit reversal, fused scale/window, real wrappers).
//
// Design: a transform of size N = L_1 * ... * L_p is run as p "digit passes".
// One CTA owns a TILE of C interleaved length-L sub-transforms (C*L = 4096 or 8192 points):
//   * column-like tile: C adjacent columns of an [A][L][B] view (element (a,i,b) at a*L*B+i*B+b),
//     every global access is a C*16-byte contiguous segment;
//   * row-like tile: C contiguous rows of L points.
// Each thread keeps 16 points in registers; a pass is a Stockham autosort FFT whose first radix
// stage is fed straight from global memory and whose last radix stage stores straight to global
// memory, with shared memory used only for the exchanges in between.  Scale, window, the
// inter-pass twiddle, conjugation (inverse transforms) and the bit-reversal of the reference's
// :in-order nil layout are fused into those global loads and stores.
//
// Reference semantics being reproduced (file:line into pkhuong/Napa-FFT3):
//   forward.lisp:54-117 (DIF: natural in, bit-reversed out, scale THEN window on the first load),
//   inverse.lisp:47-112 (DIT: bit-reversed in, natural out, window indexed by memory position),
//   gen-support.lisp:44-60 (%scale, %window), bit-reversal.lisp:146-252, real.lisp:10-31,101-185.
#pragma once
#ifdef NAPA_EMULATE
#include "cuda_emul.h"   // tests/emul: CPU functional emulation for the no-GPU test tier only
#else
#include <cuda_runtime.h>
#define NAPA_DYN_SMEM(T, name)                                      \
    extern __shared__ __align__(128) unsigned char name##_raw[];    \
    T *name = reinterpret_cast<T *>(name##_raw)
#define NAPA_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif
#include <math.h>
#include <stdint.h>

namespace napa {

typedef double2 cplx;

// ------------------------------------------------------------------------------------------
// complex helpers (nvcc contracts a*b+c into DFMA: 2 DMUL + 2 DFMA per complex product)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cplx mul_mi(cplx a) { return make_double2(a.y, -a.x); }   // * (-i)
__device__ __forceinline__ cplx mul_pi(cplx a) { return make_double2(-a.y, a.x); }   // * (+i)

#define NAPA_SQRT1_2 0.70710678118654752440
#define NAPA_COS_PI_8 0.92387953251128675613
#define NAPA_SIN_PI_8 0.38268343236508977173

// forward (e^{-2 pi i jk/R}) butterflies, natural in -> natural out, in registers
__device__ __forceinline__ void fft2(cplx &a0, cplx &a1) {
    cplx t = a0; a0 = cadd(t, a1); a1 = csub(t, a1);
}
__device__ __forceinline__ void fft4(cplx &a0, cplx &a1, cplx &a2, cplx &a3) {
    cplx s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = mul_mi(csub(a1, a3));
    a0 = cadd(s02, s13); a2 = csub(s02, s13); a1 = cadd(d02, d13); a3 = csub(d02, d13);
}
// W_8^1 * z and W_8^3 * z
__device__ __forceinline__ cplx mul_w8_1(cplx z) { return make_double2((z.x + z.y) * NAPA_SQRT1_2, (z.y - z.x) * NAPA_SQRT1_2); }
__device__ __forceinline__ cplx mul_w8_3(cplx z) { return make_double2((z.y - z.x) * NAPA_SQRT1_2, -(z.x + z.y) * NAPA_SQRT1_2); }

template <int R> struct Radix;
template <> struct Radix<2> {
    static __device__ __forceinline__ void run(cplx *w) { fft2(w[0], w[1]); }
};
template <> struct Radix<4> {
    static __device__ __forceinline__ void run(cplx *w) { fft4(w[0], w[1], w[2], w[3]); }
};
template <> struct Radix<8> {
    // j = a + 2d, k = b + 4c:  W_8^{jk} = W_8^{ab} W_2^{ac} W_4^{db}
    static __device__ __forceinline__ void run(cplx *w) {
        fft4(w[0], w[2], w[4], w[6]);          // a = 0: u0[b] in w[0],w[2],w[4],w[6]
        fft4(w[1], w[3], w[5], w[7]);          // a = 1: u1[b] in w[1],w[3],w[5],w[7]
        cplx u1 = mul_w8_1(w[3]), u2 = mul_mi(w[5]), u3 = mul_w8_3(w[7]);
        cplx o0 = cadd(w[0], w[1]), o4 = csub(w[0], w[1]);
        cplx o1 = cadd(w[2], u1), o5 = csub(w[2], u1);
        cplx o2 = cadd(w[4], u2), o6 = csub(w[4], u2);
        cplx o3 = cadd(w[6], u3), o7 = csub(w[6], u3);
        w[0] = o0; w[1] = o1; w[2] = o2; w[3] = o3; w[4] = o4; w[5] = o5; w[6] = o6; w[7] = o7;
    }
};
template <> struct Radix<16> {
    // j = a + 4d, k = b + 4c:  W_16^{jk} = W_16^{ab} W_4^{ac} W_4^{db}
    static __device__ __forceinline__ void run(cplx *w) {
#pragma unroll
        for (int a = 0; a < 4; a++) fft4(w[a], w[a + 4], w[a + 8], w[a + 12]);   // w[a+4b] = u[a][b]
        const cplx W1 = make_double2(NAPA_COS_PI_8, -NAPA_SIN_PI_8);
        const cplx W3 = make_double2(NAPA_SIN_PI_8, -NAPA_COS_PI_8);
        w[1 + 4] = cmul(w[1 + 4], W1);            // a=1,b=1: W^1
        w[1 + 8] = mul_w8_1(w[1 + 8]);            // a=1,b=2: W^2
        w[1 + 12] = cmul(w[1 + 12], W3);          // a=1,b=3: W^3
        w[2 + 4] = mul_w8_1(w[2 + 4]);            // a=2,b=1: W^2
        w[2 + 8] = mul_mi(w[2 + 8]);              // a=2,b=2: W^4 = -i
        w[2 + 12] = mul_w8_3(w[2 + 12]);          // a=2,b=3: W^6
        w[3 + 4] = cmul(w[3 + 4], W3);            // a=3,b=1: W^3
        w[3 + 8] = mul_w8_3(w[3 + 8]);            // a=3,b=2: W^6
        {                                          // a=3,b=3: W^9 = -W^1
            cplx t = cmul(w[3 + 12], W1); w[3 + 12] = make_double2(-t.x, -t.y);
        }
        cplx o[16];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            cplx x0 = w[4 * b], x1 = w[4 * b + 1], x2 = w[4 * b + 2], x3 = w[4 * b + 3];
            fft4(x0, x1, x2, x3);                  // output c -> X[b + 4c]
            o[b] = x0; o[b + 4] = x1; o[b + 8] = x2; o[b + 12] = x3;
        }
#pragma unroll
        for (int i = 0; i < 16; i++) w[i] = o[i];
    }
};

// ------------------------------------------------------------------------------------------
// radix plans: every thread owns 16 points, so a radix-R stage is 16/R butterflies per thread
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int plan_nstages(int l) { return l <= 4 ? 1 : l <= 8 ? 2 : l <= 12 ? 3 : 4; }
__host__ __device__ constexpr int plan_log2radix(int l, int s) {
    // l: 4 [16] 5 [8,4] 6 [8,8] 7 [16,8] 8 [16,16] 9 [16,16,2] 10 [16,16,4] 11 [16,16,8]
    //    12 [16,16,16] 13 [16,16,8,4]   (interleaved A/B on one box, tools/ab_plans.py: 16*16*2 beats 8*8*8 by 3 % at
    //    L = 512; 16*2 / 16*4 lose 1-3 % to 8*4 / 8*8 at L = 32 / 64; 16*16*16*2 equals 16*16*8*4 at L = 8192)
    switch (l) {
    case 4: return 4;
    case 5: return s == 0 ? 3 : 2;
    case 6: return 3;
#ifdef NAPA_PLAN_OLD   /* A/B build of the earlier 8*8*8 plan (tools/ab_plans.py) */
    case 9: return 3;
#else
    case 9: return s < 2 ? 4 : 1;
#endif
    case 7: return s == 0 ? 4 : 3;
    case 8: return 4;
    case 10: return s < 2 ? 4 : 2;
    case 11: return s < 2 ? 4 : 3;
    case 12: return 4;
    case 13: return s < 2 ? 4 : (s == 2 ? 3 : 2);
    }
    return 0;
}
// tile shape: C*L = 2048 / 4096 / 8192 points per CTA; 16 points per thread
#ifndef NAPA_SMALL_TILE_MAXL
#define NAPA_SMALL_TILE_MAXL 8     // L <= 2^8: 2048-point tiles on 128 threads (4 CTAs/SM, more barrier domains)
#endif
__host__ __device__ constexpr int plan_log2pts(int l) { return l <= NAPA_SMALL_TILE_MAXL ? 11 : (l <= 12 ? 12 : 13); }
__host__ __device__ constexpr int plan_log2c(int l) { return plan_log2pts(l) - l; }
__host__ __device__ constexpr int plan_min_ctas(int l) { return plan_log2pts(l) == 11 ? 4 : (plan_log2pts(l) == 12 ? 2 : 1); }
__host__ __device__ constexpr int plan_threads(int l) { return (1 << (l - 4)) << plan_log2c(l); }
// shared-memory index padding (only needed when C < 8; see tools/bank_check.py)
__host__ __device__ constexpr int plan_padshift(int l) { return plan_log2c(l) >= 3 ? 31 : 4; }
__host__ __device__ constexpr size_t plan_smem_bytes(int l) {
    return plan_nstages(l) == 1 ? 0
           : (size_t)(((1u << l) + ((1u << l) >> plan_padshift(l))) << plan_log2c(l)) * sizeof(cplx);
}

enum : unsigned {
    PF_Q_IS_BATCH = 1u << 0,   // q indexes batch items (single-pass row tiles) instead of rows/cols
    PF_IN_REV     = 1u << 1,   // logical input i lives at position rev_L(i) along the tile's i axis
    PF_OUT_REV    = 1u << 2,   // logical output k is stored at position rev_L(k)
    PF_CONJ_IN    = 1u << 3,
    PF_CONJ_OUT   = 1u << 4,
    PF_TW_LOAD    = 1u << 5,   // multiply loads by W_M^{i*col}   (DIT-order pipelines)
    PF_TW_STORE   = 1u << 6,   // multiply stores by W_M^{k*col}  (DIF-order pipelines)
    PF_WIN_REAL   = 1u << 7,
    PF_WIN_CPLX   = 1u << 8,
    PF_WIN_REV    = 1u << 9,   // window index = rev_{lgN}(offset in transform) (ifft :in-order t)
    PF_SCALE      = 1u << 10,
    PF_RIFFT_PRE  = 1u << 11,  // x = (a+b) + (a-b)*r2tw[off], b = src[off + N]   (real.lisp:170-177)
    PF_RFFT_POST  = 1u << 12,  // single-pass row tiles: rfft epilogue (real.lisp:16-30, 128-134) fused into the store
    PF_ZIP2       = 1u << 13,  // %2rfft: the input is two REAL vectors, z = src[.] + i src_im[.] (strides in doubles)
    PF_SPLIT2_POST = 1u << 14, // %2rfft: single-pass row tiles store the two spectra E (at i) and O (at L + i), real.lisp:16-30
    PF_SPLIT_IN   = 1u << 15,  // distributed four-step: the item is read from the exchange layout [chunk][rank][row][Bc] (split_offset)
    PF_SPLIT_OUT  = 1u << 16,  // ... or written to it
};

struct PassParams {
    const cplx *src;
    cplx *dst;
    long long src_batch_stride, dst_batch_stride;   // complex elements between batch items
    long long batch;                                // number of batch items in this launch
    int src_hint, dst_hint;                         // HINT_* L2 eviction hints of the data path
    int tiles_per_item;                             // tiles inside one item (1 if PF_Q_IS_BATCH); a power of two
    int G;                                          // tile r -> (r / G, r % G); a power of two
    int lg_tpi, lg_G;                               // their logarithms (set by the launcher)
    long long s_SA, s_SG, s_i, s_q;                 // src offset = (r/G)*SA + (r%G)*SG + pos(i)*s_i + q*s_q
    long long d_SA, d_SG, d_i, d_q;
    unsigned flags;
    double scale;
    const void *window;                             // double[] or cplx[]
    long long window_batch_stride;                  // 0 = shared window
    int lgN;                                        // log2 of the full transform (PF_WIN_REV)
    const cplx *stage_tw;                           // W_L^m, m < L
    const cplx *tw_lo, *tw_hi;                      // W_M^e = hi[e >> tw_shift] * lo[e & mask]
    int tw_shift;
    unsigned long long tw_mask;                     // M - 1
    // exponent of register m = colx * (tw_mul_a * ra + tw_mul_t * (t + m*T)),  colx = tw_col_offset + (col >> tw_col_shift)
    // (plain passes: tw_mul_t = 1, tw_mul_a = 0, offset 0, shift 0)
    long long tw_col_offset, tw_mul_t, tw_mul_a;
    int tw_col_shift;
    const cplx *r2tw;                               // reference recurrence table (rifft pre-pass)
    long long aux_off;                              // N for PF_RIFFT_PRE
    const double *src_im;                           // PF_ZIP2: the second real input
    // PF_SPLIT_IN / PF_SPLIT_OUT: in-row index j -> j % Bc + ((j / Bc) % nch) * sp_chunk_stride + (j / (Bc * nch)) * sp_rank_stride
    int sp_lgbc, sp_lgnch;
    long long sp_chunk_stride, sp_rank_stride;
    int tw_use_ra;                                  // inter-pass twiddle: the column factor is the tile's outer index ra, not its column
};
// Row j of a distributed transform lives in P * nch pieces of Bc points (one per source rank and column chunk,
// exactly as the all-to-all delivered them): logical in-row index -> element offset inside the item.
__device__ __forceinline__ long long split_offset(const PassParams &p, const long long j) {
    const long long lo = j & ((1LL << p.sp_lgbc) - 1);
    const long long ch = (j >> p.sp_lgbc) & ((1LL << p.sp_lgnch) - 1);
    const long long rk = j >> (p.sp_lgbc + p.sp_lgnch);
    return lo + ch * p.sp_chunk_stride + rk * p.sp_rank_stride;
}

__device__ __forceinline__ cplx ldg_c(const cplx *p) { return __ldg(p); }
// Data-path loads/stores carry an L2 eviction hint (HINT_*): streaming operands (user src / dst,
// touched once) are evict-first, the L2-resident intermediate of the two-pass pipeline is
// evict-last, so that the stream does not push the intermediate out to HBM.
enum { HINT_NORMAL = 0, HINT_STREAM = 1, HINT_KEEP = 2 };
#ifdef NAPA_EMULATE
struct L2Policies { int dummy; };
__device__ __forceinline__ L2Policies make_policies() { return L2Policies{0}; }
template <bool COHERENT> __device__ __forceinline__ cplx ld_data(const cplx *p, int, const L2Policies &) { return *p; }
__device__ __forceinline__ void st_data(cplx *p, cplx v, int, const L2Policies &) { *p = v; }
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) { return *(const volatile unsigned *)p; }
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) { return *(const volatile unsigned *)p; }
__device__ __forceinline__ void fence_acq_rel_gpu() {}
__device__ __forceinline__ void napa_nanosleep(unsigned) { napa_emul::yield_spin(); }   // let the other CTAs' fibers run
// mbarriers and bulk copies, emulated: a copy is done synchronously at issue time and the barrier's
// transaction count is ignored, so the issuing thread's expect_tx arrival is deferred to mbar_emul_complete()
// (called right after the copies; a no-op on the device).  Word layout: [63] phase, [62:32] count, [31:0] pending.
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) { *bar = ((unsigned long long)count << 32) | count; }
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    unsigned long long w = *bar;
    unsigned pending = (unsigned)w - 1;
    const unsigned long long count = (w >> 32) & 0x7fffffffull, phase = w >> 63;
    if (pending == 0) w = ((phase ^ 1ull) << 63) | (count << 32) | count;
    else w = (phase << 63) | (count << 32) | pending;
    *bar = w;
    napa_emul::note_progress();
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *, unsigned) {}
__device__ __forceinline__ void mbar_emul_complete(unsigned long long *bar) { mbar_arrive(bar); }
__device__ __forceinline__ bool mbar_test(unsigned long long *bar, unsigned parity) { return (unsigned)(*(volatile unsigned long long *)bar >> 63) != parity; }
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) { while (!mbar_test(bar, parity)) napa_emul::yield_any(); }
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void napa_syncwarp() {}
__device__ __forceinline__ void bulk_g2s(cplx *s, const cplx *g, unsigned bytes, unsigned long long *, int, const L2Policies &) {
    for (unsigned i = 0; i < bytes / sizeof(cplx); i++) s[i] = g[i];
}
template <int N> struct SyncCompute { static __device__ __forceinline__ void sync() { napa_emul::bar_sync(1, N); } };
#else
struct L2Policies { unsigned long long normal, first, last; };
__device__ __forceinline__ L2Policies make_policies() {
    L2Policies pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol.normal));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol.first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol.last));
    return pol;
}
__device__ __forceinline__ unsigned long long pick_policy(const L2Policies &pol, int hint) {
    return hint == HINT_STREAM ? pol.first : (hint == HINT_KEEP ? pol.last : pol.normal);
}
// COHERENT: L2-only (.cg) because another SM may have written the line earlier in this launch
template <bool COHERENT> __device__ __forceinline__ cplx ld_data(const cplx *p, int hint, const L2Policies &pol) {
    cplx v;
    if (!COHERENT && hint < 0) return *p;              // un-hinted launch: plain load, no policy operand
    const unsigned long long policy = pick_policy(pol, hint);
    if (COHERENT)
        asm volatile("ld.global.cg.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(policy) : "memory");
    else
        asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(policy) : "memory");
    return v;
}
__device__ __forceinline__ void st_data(cplx *p, cplx v, int hint, const L2Policies &pol) {
    if (hint < 0) { *p = v; return; }                   // un-hinted launch: plain store
    const unsigned long long policy = pick_policy(pol, hint);
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" :: "l"(p), "d"(v.x), "d"(v.y), "l"(policy) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void napa_nanosleep(unsigned ns) { __nanosleep(ns); }
// mbarrier + bulk asynchronous copy (TMA engine, no tensor map): global -> shared, completion
// counted in bytes on an mbarrier.  The copy goes through L2 only, hence sees other SMs' writes.
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_emul_complete(unsigned long long *) {}
__device__ __forceinline__ bool mbar_test(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    while (!mbar_test(bar, parity)) {}
}
// orders earlier generic-proxy accesses (made visible to this thread) before its later async-proxy operations
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void napa_syncwarp() { __syncwarp(); }
// one whole column tile (L rows of C points, a 3-d box of the tensor map) in a single instruction
__device__ __forceinline__ void tma_tile_g2s(cplx *s, const void *tmap, int x, int y, int z, unsigned long long *bar, int hint, const L2Policies &pol) {
    const unsigned long long policy = pick_policy(pol, hint);
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
                 :: "r"((unsigned)__cvta_generic_to_shared(s)), "l"(tmap), "r"((unsigned)__cvta_generic_to_shared(bar)),
                    "r"(x), "r"(y), "r"(z), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_g2s(cplx *s, const cplx *g, unsigned bytes, unsigned long long *bar, int hint, const L2Policies &pol) {
    const unsigned long long policy = pick_policy(pol, hint);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"((unsigned)__cvta_generic_to_shared(s)), "l"(g), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)), "l"(policy) : "memory");
}
#endif
#ifndef NAPA_EMULATE
// barrier over the N compute threads of a warp-specialised CTA (named barrier 1; the producer warp stays out)
template <int N> struct SyncCompute { static __device__ __forceinline__ void sync() { asm volatile("bar.sync 1, %0;" :: "n"(N) : "memory"); } };
#endif
struct SyncAll { static __device__ __forceinline__ void sync() { __syncthreads(); } };

// rev_w(x): reverse the low w bits of x (w = 0 -> 0)
__device__ __forceinline__ int rev_bits(int x, int w) { return w ? (int)(__brev((unsigned)x) >> (32 - w)) : 0; }
__device__ __forceinline__ long long rev_bits64(long long x, int w) {
    return w ? (long long)(__brevll((unsigned long long)x) >> (64 - w)) : 0;
}

// ------------------------------------------------------------------------------------------
// Lane maps and shared-memory layouts (validated by tools/bank_check.py)
//   MAP_Q: transform index q fastest across lanes  -> a quarter warp covers 8 adjacent COLUMNS
//          (one 128-byte segment of a column-like tile)
//   MAP_T: position t fastest across lanes         -> a quarter warp covers 8 adjacent elements
//          of one ROW (row-like tiles, single-pass transforms)
// A 128-bit access costs one LSU data-pipe wavefront per distinct 128-byte line per quarter warp,
// so the map must follow the contiguous direction of each side of the tile in global memory; the
// map can change from the load side to the store side at the first shared-memory exchange.
// Layouts: interleaved [pos][q] when both sides are MAP_Q, else planar [q][pos] with a pitch that
// keeps both maps conflict-free.
// ------------------------------------------------------------------------------------------
enum { MAP_Q = 0, MAP_T = 1 };

template <int LOG2L, int MAP>
__device__ __forceinline__ void lane_ids(const int tid, int &t, int &q) {
    if constexpr (MAP == MAP_Q) {
        constexpr int LOG2C = plan_log2c(LOG2L);
        q = tid & ((1 << LOG2C) - 1);
        t = tid >> LOG2C;
    } else {
        constexpr int LOG2T = LOG2L - 4;
        t = tid & ((1 << LOG2T) - 1);
        q = tid >> LOG2T;
    }
}

__host__ __device__ constexpr int planar_padshift(int l) { return plan_log2radix(l, 0); }
__host__ __device__ constexpr int planar_pitch(int l) {
    const int lc = plan_log2c(l);
    const int skew = lc >= 3 ? 1 : (lc == 2 ? 2 : (lc == 1 ? 4 : 0));
    return (1 << l) + ((1 << l) >> planar_padshift(l)) + skew;
}
__host__ __device__ constexpr size_t plan_smem_bytes_planar(int l) {
    return plan_nstages(l) == 1 ? 0 : (size_t)(planar_pitch(l) << plan_log2c(l)) * sizeof(cplx);
}

template <int LOG2L, bool PLANAR>
__device__ __forceinline__ int smem_phys(int pos, int q) {
    if constexpr (PLANAR) return q * planar_pitch(LOG2L) + pos + (pos >> planar_padshift(LOG2L));
    else return ((pos + (pos >> plan_padshift(LOG2L))) << plan_log2c(LOG2L)) + q;
}
// smem_phys(t + m*T, q) - smem_phys(t, q): a compile-time constant whenever the padding period
// divides T (then (t + m*T) >> PADSH == (t >> PADSH) + ((m*T) >> PADSH))
template <int LOG2L, bool PLANAR>
__device__ __forceinline__ int smem_delta(int t, int m) {
    constexpr int T = (1 << LOG2L) / 16;
    constexpr int PADSH = PLANAR ? planar_padshift(LOG2L) : plan_padshift(LOG2L);
    constexpr int LOG2C = PLANAR ? 0 : plan_log2c(LOG2L);
    if constexpr (PADSH >= 31) return (m * T) << LOG2C;
    else if constexpr ((T & ((1 << PADSH) - 1)) == 0) return (m * T + ((m * T) >> PADSH)) << LOG2C;
    else return ((m * T + (((t + m * T) >> PADSH) - (t >> PADSH))) << LOG2C);
}
// reorder buffer for bit-reversed row-side accesses: planar, low 3 position bits XOR reversed top 3
template <int LOG2L>
__device__ __forceinline__ int reorder_phys(int pos, int q) {
    const int top = pos >> (LOG2L - 3);
    const int r3 = ((top & 1) << 2) | (top & 2) | (top >> 2);
    return (q << LOG2L) + (pos ^ r3);
}

// v[u + s*NB] *= w^s for s = 1 .. R-1
template <int R, int NB>
__device__ __forceinline__ void twiddle_powers(cplx (&v)[16], const int u, const cplx w1) {
    if constexpr (R >= 2) v[u + 1 * NB] = cmul(v[u + 1 * NB], w1);
    if constexpr (R >= 4) {
        const cplx w2 = cmul(w1, w1), w3 = cmul(w2, w1);
        v[u + 2 * NB] = cmul(v[u + 2 * NB], w2);
        v[u + 3 * NB] = cmul(v[u + 3 * NB], w3);
        if constexpr (R >= 8) {
            const cplx w4 = cmul(w2, w2);
            const cplx w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
            v[u + 4 * NB] = cmul(v[u + 4 * NB], w4);
            v[u + 5 * NB] = cmul(v[u + 5 * NB], w5);
            v[u + 6 * NB] = cmul(v[u + 6 * NB], w6);
            v[u + 7 * NB] = cmul(v[u + 7 * NB], w7);
            if constexpr (R >= 16) {
                const cplx w8 = cmul(w4, w4);
                v[u + 8 * NB] = cmul(v[u + 8 * NB], w8);
                v[u + 9 * NB] = cmul(v[u + 9 * NB], cmul(w8, w1));
                v[u + 10 * NB] = cmul(v[u + 10 * NB], cmul(w8, w2));
                v[u + 11 * NB] = cmul(v[u + 11 * NB], cmul(w8, w3));
                v[u + 12 * NB] = cmul(v[u + 12 * NB], cmul(w8, w4));
                v[u + 13 * NB] = cmul(v[u + 13 * NB], cmul(w8, w5));
                v[u + 14 * NB] = cmul(v[u + 14 * NB], cmul(w8, w6));
                v[u + 15 * NB] = cmul(v[u + 15 * NB], cmul(w8, w7));
            }
        }
    }
}

// Stockham autosort stages.  Stage 0 runs under the load-side ids (t0, q0) and scatters; every
// later stage gathers and runs under the store-side ids (t1, q1).  PRESYNC0: sm was used before
// stage 0 (reorder of a bit-reversed load) and everyone must be done with it before the scatter.
// Hook protocol of run_tile: loads_issued() runs right after the tile's global loads have been issued (their
// latency hides whatever it does), operator() after the last exchange has been gathered.
struct NoHook {
    __device__ __forceinline__ void operator()() const {}
    __device__ __forceinline__ void loads_issued() const {}
};

template <int LOG2L, int STAGE, int LOG2NS, bool PLANAR, bool PRESYNC0, class Sync, class Hook>
struct Stages {
    static constexpr int L = 1 << LOG2L;
    static constexpr int T = L / 16;
    static constexpr int NSTAGES = plan_nstages(LOG2L);
    static constexpr int LOG2R = plan_log2radix(LOG2L, STAGE);
    static constexpr int R = 1 << LOG2R;
    static constexpr int NB = 16 / R;
    static constexpr int NS = 1 << LOG2NS;

    // seed of butterfly u of this stage (thread position t on the store side): W_{NS*R}^{(t + u*T) % NS}
    static __device__ __forceinline__ cplx load_seed(const cplx *__restrict__ tw, const int t, const int u) {
        const int jm = (t + u * T) & (NS - 1);
        return ldg_c(tw + jm * (L / (NS * R)));
    }
    // PRE: seeds of this stage, fetched by the previous stage before its exchange (nullptr: fetch here)
    static __device__ __forceinline__ void run(cplx (&v)[16], cplx *sm, const cplx *__restrict__ tw,
                                               const int t0, const int q0, const int t1, const int q1, Hook &hook,
                                               const cplx *pre = nullptr) {
        const int t = STAGE == 0 ? t0 : t1, q = STAGE == 0 ? q0 : q1;
        if constexpr (STAGE > 0) {
            // exchange: previous stage wrote its scattered outputs; gather t + m*T
            // (one base address + compile-time offsets when the padding step divides T)
            const cplx *g = sm + smem_phys<LOG2L, PLANAR>(t, q);
#pragma unroll
            for (int m = 0; m < 16; m++) v[m] = g[smem_delta<LOG2L, PLANAR>(t, m)];
            if constexpr (STAGE + 1 == NSTAGES) hook();
            // twiddle W_{NS*R}^{(jb % NS) * s} on input s of butterfly jb = t + u*T.  Only the seed
            // w = W^{jb % NS} is loaded; its powers come from a product tree (depth <= 4): the
            // LSU<->register datapath, not the FP64 pipe, is the scarce resource of this kernel.
#pragma unroll
            for (int u = 0; u < NB; u++) {
                const cplx w1 = pre ? pre[u] : load_seed(tw, t, u);
                twiddle_powers<R, NB>(v, u, w1);
            }
        }
#pragma unroll
        for (int u = 0; u < NB; u++) {
            cplx w[R];
#pragma unroll
            for (int s = 0; s < R; s++) w[s] = v[u + s * NB];
            Radix<R>::run(w);
#pragma unroll
            for (int s = 0; s < R; s++) v[u + s * NB] = w[s];
        }
        if constexpr (STAGE + 1 < NSTAGES) {
            using Next = Stages<LOG2L, STAGE + 1, LOG2NS + LOG2R, PLANAR, PRESYNC0, Sync, Hook>;
            // the next stage's seeds depend only on the thread: fetch them now, so that their latency
            // (an L2 round trip whenever L1 was invalidated by a fence) hides behind the exchange
            constexpr bool PREFETCH = Next::NB <= 2;
            cplx wn[PREFETCH ? Next::NB : 1];
            if constexpr (PREFETCH) {
#pragma unroll
                for (int u = 0; u < Next::NB; u++) wn[u] = Next::load_seed(tw, t1, u);
            }
            if constexpr (STAGE > 0 || PRESYNC0) Sync::sync();   // everyone has gathered before we overwrite
#pragma unroll
            for (int u = 0; u < NB; u++) {
                int jb = t + u * T;
                int base = ((jb >> LOG2NS) << (LOG2NS + LOG2R)) + (jb & (NS - 1));
#pragma unroll
                for (int s = 0; s < R; s++) sm[smem_phys<LOG2L, PLANAR>(base + s * NS, q)] = v[u + s * NB];
            }
#ifndef NAPA_MUTANT_NO_EXCHANGE_BARRIER   /* tests/test_emul_orders.py builds a mutant without it to prove that its checks bite */
            Sync::sync();
#endif
            Next::run(v, sm, tw, t0, q0, t1, q1, hook, PREFETCH ? wn : nullptr);
        }
    }
};

// v[m] *= W_M^{(t + m*T) * col}, m < 16, from two table seeds and a depth-7 power tree:
//   B = W^{t*col}, D = W^{T*col};  P_m = B * D^m  built from D^2, D^4, D^8.
// 18 complex products instead of 32 dependent table look-ups; every seed load is issued at once.
template <int T>
__device__ __forceinline__ void apply_big_twiddle(cplx (&v)[16], const PassParams &p, int t, unsigned col, int ra) {
    const unsigned long long colx = (unsigned long long)(p.tw_col_offset + (p.tw_use_ra ? (long long)ra : (long long)(col >> p.tw_col_shift))) & p.tw_mask;
    const unsigned long long idx = ((unsigned long long)(p.tw_use_ra ? p.tw_mul_t * t : p.tw_mul_a * ra + p.tw_mul_t * t)) & p.tw_mask;
    const unsigned long long eb = (idx * colx) & p.tw_mask;
    const unsigned long long ed = ((((unsigned long long)(p.tw_mul_t * T)) & p.tw_mask) * colx) & p.tw_mask;
    const unsigned long long lomask = (1ull << p.tw_shift) - 1;
    const cplx blo = ldg_c(p.tw_lo + (eb & lomask)), bhi = ldg_c(p.tw_hi + (eb >> p.tw_shift));
    const cplx dlo = ldg_c(p.tw_lo + (ed & lomask)), dhi = ldg_c(p.tw_hi + (ed >> p.tw_shift));
    const cplx B = cmul(bhi, blo), D = cmul(dhi, dlo);
    const cplx D2 = cmul(D, D), D4 = cmul(D2, D2), D8 = cmul(D4, D4);
    cplx P[8];
    P[0] = B; P[1] = cmul(B, D); P[2] = cmul(B, D2); P[3] = cmul(P[1], D2);
    P[4] = cmul(B, D4); P[5] = cmul(P[1], D4); P[6] = cmul(P[2], D4); P[7] = cmul(P[3], D4);
#pragma unroll
    for (int m = 0; m < 8; m++) {
        v[m] = cmul(v[m], P[m]);
        v[m + 8] = cmul(v[m + 8], cmul(P[m], D8));
    }
}

// Where a thread's 16 points of a tile live on one side.  `tile` counts tiles over the whole
// launch (item-major).
struct TileCtx {
    int q, t;
    long long item;
    bool valid;
    long long s_xoff, d_xoff;
    unsigned col;                    // column offset inside the [L][B] sub-matrix (twiddle exponent)
    int ra;                          // outer index of the tile (twiddle exponent of distributed passes)
};
template <int LOG2L, int MAP>
__device__ __forceinline__ TileCtx tile_ctx(const PassParams &p, const long long tile, const unsigned flags) {
    constexpr int C = 1 << plan_log2c(LOG2L);
    TileCtx c;
    lane_ids<LOG2L, MAP>(threadIdx.x, c.t, c.q);
    const long long b = tile >> p.lg_tpi;
    const int r = (int)(tile & (p.tiles_per_item - 1));
    const int ra = r >> p.lg_G, rg = r & (p.G - 1);
    c.item = b;
    c.s_xoff = (long long)ra * p.s_SA + (long long)rg * p.s_SG;
    c.d_xoff = (long long)ra * p.d_SA + (long long)rg * p.d_SG;
    if (flags & PF_Q_IS_BATCH) {
        c.item = b * C + c.q;
    } else {
        c.s_xoff += c.q * p.s_q;
        c.d_xoff += c.q * p.d_q;
    }
    c.valid = c.item < p.batch;
    c.col = (unsigned)(rg * C + c.q);
    c.ra = ra;
    return c;
}

// One tile of one digit pass.
//   COHERENT: L2-only (ld.global.cg) data loads, required when another SM wrote the data earlier
//             in the SAME launch (group-fused kernel): a stale L1 line would be a bug.
//   LM / SM_: lane maps of the load side and of the store side.
// All 16 data loads of a thread are issued back to back into v[], and every optional fusion is a
// separate warp-uniform branch over the whole register array, so unused fusions cost nothing and
// used ones do not serialise the loads.
// ALLOWED: the flag bits this instantiation can ever see (the group-fused kernel knows the role of
// each pass at compile time; masking lets the compiler drop the other fusion paths = smaller code).
// DENSE (L < 128, single-pass row tiles of back-to-back transforms): the tile is one contiguous block
// of 2048 points.  With one or a few threads per transform a direct access touches 32 different
// lines per warp instruction, so the block is instead read and written fully coalesced and
// transposed through a padded shared-memory image [q][L + 1] (conflict-free both ways).
// STAGED: the tile's input was bulk-copied into `sm` itself (stage_issue below; rows of C points
// for a column tile, C rows of L points for a row tile) and its arrival is awaited on `mbar`;
// the exchange then reuses the same buffer (REUSE must be set).
template <int LOG2L, bool COHERENT, int LM, int SM_, class Sync, bool REUSE, class Hook, unsigned ALLOWED = 0xffffffffu, bool HINTED = true,
          bool STAGED = false, bool DENSE = false>
__device__ __forceinline__ void run_tile(const PassParams &p, const long long tile, cplx *sm,
                                         const long long src_item_delta, const long long dst_item_delta,
                                         const L2Policies &pol, Hook &hook,
                                         unsigned long long *mbar = nullptr, const unsigned mbar_parity = 0) {
    constexpr int L = 1 << LOG2L;
    constexpr int T = L / 16;
    constexpr bool PLANAR = !(LM == MAP_Q && SM_ == MAP_Q);
    constexpr bool REORDER_OK = LOG2L >= 7 && plan_nstages(LOG2L) > 1;   // swizzled reorder buffer
    static_assert(plan_nstages(LOG2L) > 1 || LM == SM_, "single-stage tiles cannot switch lane maps");
    const int src_hint = HINTED ? p.src_hint : -1, dst_hint = HINTED ? p.dst_hint : -1;
    const unsigned flags = p.flags & ALLOWED;
    const TileCtx c = tile_ctx<LOG2L, LM>(p, tile, flags);
    const int t = c.t, q = c.q;
    const long long s_xoff = c.s_xoff;
    // a bit-reversed access on a row side goes through shared memory so that global memory sees
    // whole contiguous lines; on a column side only the row index is permuted (free)
    const bool in_reorder = REORDER_OK && LM == MAP_T && (flags & PF_IN_REV);

    cplx v[16];
    constexpr int DPITCH = L + 1;                       // dense image: transform q at sm[q * (L + 1)]
    long long dense_valid = 0;                           // points of the tile that exist (ragged last tile)
    if constexpr (DENSE) {
        constexpr int C = 1 << plan_log2c(LOG2L), NT = plan_threads(LOG2L);
        static_assert(!STAGED && LM == MAP_Q && SM_ == MAP_Q, "dense rows use the q-fastest lane map");
        const long long item0 = tile * C;
        const long long left = p.batch - item0;
        dense_valid = (left < C ? left : (long long)C) << LOG2L;
        const cplx *g = p.src + (item0 + src_item_delta) * p.src_batch_stride;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int e = (int)threadIdx.x + NT * j;
            v[j] = e < dense_valid ? ld_data<COHERENT>(g + e, src_hint, pol) : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int e = (int)threadIdx.x + NT * j;
            sm[(e >> LOG2L) * DPITCH + (e & (L - 1))] = v[j];
        }
        Sync::sync();
    }
    if (c.valid) {
        const cplx *src = p.src + (c.item + src_item_delta) * p.src_batch_stride + s_xoff;
        const long long s_i = p.s_i;
        // memory position of register m on the load side
        auto lpos = [&](int m) -> int { return ((flags & PF_IN_REV) && !in_reorder) ? rev_bits(t + m * T, LOG2L) : t + m * T; };
        // real window values are fetched first: they do not depend on the data and their L2
        // latency then overlaps the data wait
        double wv[16];
        // window index = memory position; when the registers hold memory positions t + m*T (no permutation between
        // memory and registers on this side) it advances like the data pointer: one base + m * step
        const bool win_linear = !(flags & PF_WIN_REV) && (!(flags & PF_IN_REV) || in_reorder);
        if (flags & PF_WIN_REAL) {
            const double *w = reinterpret_cast<const double *>(p.window) + c.item * p.window_batch_stride;
            if (win_linear) {
                const double *w0 = w + s_xoff + (long long)t * s_i;
                const long long ws = (long long)T * s_i;
#pragma unroll
                for (int m = 0; m < 16; m++) wv[m] = __ldg(w0 + m * ws);
            } else {
#pragma unroll
                for (int m = 0; m < 16; m++) {
                    const long long off = s_xoff + (long long)lpos(m) * s_i;
                    wv[m] = __ldg(w + ((flags & PF_WIN_REV) ? rev_bits64(off, p.lgN) : off));
                }
            }
        }
        if constexpr (DENSE) {
#pragma unroll
            for (int m = 0; m < 16; m++) v[m] = sm[q * DPITCH + lpos(m)];
        } else if constexpr (STAGED) {
            static_assert(REUSE, "the staging buffer doubles as the exchange buffer");
            constexpr int LC = plan_log2c(LOG2L);
            mbar_wait(mbar, mbar_parity);
#pragma unroll
            for (int m = 0; m < 16; m++) v[m] = LM == MAP_Q ? sm[(lpos(m) << LC) + q] : sm[(q << LOG2L) + lpos(m)];
        } else if (flags & PF_ZIP2) {
            // %2rfft (real.lisp:33-48): two real vectors transformed as one complex vector
            const long long base = (c.item + src_item_delta) * p.src_batch_stride + s_xoff;
            const double *re = reinterpret_cast<const double *>(p.src) + base, *im = p.src_im + base;
#pragma unroll
            for (int m = 0; m < 16; m++) v[m].x = re[(long long)lpos(m) * s_i];
#pragma unroll
            for (int m = 0; m < 16; m++) v[m].y = im[(long long)lpos(m) * s_i];
        } else if (flags & PF_RIFFT_PRE) {
            // x = (a + b) + (a - b) * tw[off], b = src[off + N]   (real.lisp:170-177).  Two rounds of
            // eight elements: a and b of a round are all in flight together (16 loads), the twiddles
            // (L1/L2 hits) follow in fours -- two exposed DRAM latencies per tile and 112 live
            // registers at most, instead of one latency per element or two.
#pragma unroll
            for (int half = 0; half < 2; half++) {
                cplx y[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const long long pos = (long long)lpos(8 * half + j) * s_i;
                    v[8 * half + j] = ld_data<COHERENT>(src + pos, src_hint, pol);
                    y[j] = ld_data<COHERENT>(src + pos + p.aux_off, src_hint, pol);
                }
#pragma unroll
                for (int g4 = 0; g4 < 2; g4++) {
                    cplx w[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) w[j] = ldg_c(p.r2tw + s_xoff + (long long)lpos(8 * half + 4 * g4 + j) * s_i);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int m = 8 * half + 4 * g4 + j;
                        v[m] = cadd(cadd(v[m], y[4 * g4 + j]), cmul(csub(v[m], y[4 * g4 + j]), w[j]));
                    }
                }
            }
        } else if ((flags & PF_IN_REV) && !in_reorder) {
#pragma unroll
            for (int m = 0; m < 16; m++) v[m] = ld_data<COHERENT>(src + (long long)rev_bits(t + m * T, LOG2L) * s_i, src_hint, pol);
        } else if (flags & PF_SPLIT_IN) {
            const cplx *item0 = p.src + (c.item + src_item_delta) * p.src_batch_stride;
#pragma unroll
            for (int m = 0; m < 16; m++) v[m] = ld_data<COHERENT>(item0 + split_offset(p, s_xoff + (long long)(t + m * T) * s_i), src_hint, pol);
        } else {
            const cplx *s0 = src + (long long)t * s_i;
            const long long step = (long long)T * s_i;
#pragma unroll
            for (int m = 0; m < 16; m++) v[m] = ld_data<COHERENT>(s0 + m * step, src_hint, pol);
        }
        hook.loads_issued();
        if (flags & PF_SCALE) {
            const double sc = p.scale;
#pragma unroll
            for (int m = 0; m < 16; m++) v[m] = cscale(v[m], sc);
        }
        if (flags & PF_WIN_REAL) {
#pragma unroll
            for (int m = 0; m < 16; m++) v[m] = cscale(v[m], wv[m]);
        }
        if (flags & PF_WIN_CPLX) {
            const cplx *w = reinterpret_cast<const cplx *>(p.window) + c.item * p.window_batch_stride;
            const cplx *w0 = w + s_xoff + (long long)t * s_i;
            const long long ws = (long long)T * s_i;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                cplx wc[8];
                if (win_linear) {
#pragma unroll
                    for (int m = 0; m < 8; m++) wc[m] = ldg_c(w0 + (8 * h + m) * ws);
                } else {
#pragma unroll
                    for (int m = 0; m < 8; m++) {
                        const long long off = s_xoff + (long long)lpos(8 * h + m) * s_i;
                        wc[m] = ldg_c(w + ((flags & PF_WIN_REV) ? rev_bits64(off, p.lgN) : off));
                    }
                }
#pragma unroll
                for (int m = 0; m < 8; m++) v[8 * h + m] = cmul(v[8 * h + m], wc[m]);
            }
        }
        if (flags & PF_CONJ_IN) {
#pragma unroll
            for (int m = 0; m < 16; m++) v[m].y = -v[m].y;
        }
    } else {
        hook.loads_issued();
#pragma unroll
        for (int m = 0; m < 16; m++) v[m] = make_double2(0.0, 0.0);
    }
    if constexpr (REORDER_OK && LM == MAP_T) {
        if (in_reorder) {
            // registers hold memory positions t + m*T; fetch logical input k = t + m*T from rev(k)
            if constexpr (REUSE) Sync::sync();   // previous tile may still be reading sm
#pragma unroll
            for (int m = 0; m < 16; m++) sm[reorder_phys<LOG2L>(t + m * T, q)] = v[m];
            Sync::sync();
#pragma unroll
            for (int m = 0; m < 16; m++) v[m] = sm[reorder_phys<LOG2L>(rev_bits(t + m * T, LOG2L), q)];
        }
    }
    if (c.valid && (flags & PF_TW_LOAD)) apply_big_twiddle<T>(v, p, t, c.col, c.ra);

    const TileCtx d = tile_ctx<LOG2L, SM_>(p, tile, flags);
    if constexpr (REORDER_OK && LM == MAP_T) {
        if (in_reorder) Stages<LOG2L, 0, 0, PLANAR, true, Sync, Hook>::run(v, sm, p.stage_tw, t, q, d.t, d.q, hook);
        else Stages<LOG2L, 0, 0, PLANAR, REUSE, Sync, Hook>::run(v, sm, p.stage_tw, t, q, d.t, d.q, hook);
    } else {
        // DENSE: everyone must have read its points out of the dense image before the first exchange overwrites it
        Stages<LOG2L, 0, 0, PLANAR, REUSE || DENSE, Sync, Hook>::run(v, sm, p.stage_tw, t, q, d.t, d.q, hook);
    }

    const bool out_reorder = REORDER_OK && SM_ == MAP_T && (flags & PF_OUT_REV);
    if (d.valid) {
        if (flags & PF_TW_STORE) apply_big_twiddle<T>(v, p, d.t, d.col, d.ra);
        if (flags & PF_CONJ_OUT) {
#pragma unroll
            for (int m = 0; m < 16; m++) v[m].y = -v[m].y;
        }
    }
    if constexpr (REORDER_OK && SM_ == MAP_T) {
        if (out_reorder) {
            // registers hold outputs k = t + m*T; move the value of memory position t + m*T here
            Sync::sync();                      // everyone is done gathering the last exchange
#pragma unroll
            for (int m = 0; m < 16; m++) sm[reorder_phys<LOG2L>(rev_bits(d.t + m * T, LOG2L), d.q)] = v[m];
            Sync::sync();
#pragma unroll
            for (int m = 0; m < 16; m++) v[m] = sm[reorder_phys<LOG2L>(d.t + m * T, d.q)];
        }
    }
    if constexpr (LOG2L >= 7 && SM_ == MAP_T) {
        if (flags & PF_SPLIT2_POST) {
            // Z = FFT(v1 + i v2): E_i = FFT(v1)_i -> dst[i], O_i = FFT(v2)_i -> dst[L + i]; the same
            // operations as split2_post_kernel, each thread finishing its own bins
            Sync::sync();
#pragma unroll
            for (int m = 0; m < 16; m++) sm[(d.q << LOG2L) + d.t + m * T] = v[m];
            Sync::sync();
            if (d.valid) {
                cplx *dst = p.dst + (d.item + dst_item_delta) * p.dst_batch_stride + d.d_xoff;
#pragma unroll
                for (int m = 0; m < 16; m++) {
                    const int i = d.t + m * T;
                    const cplx x = v[m];
                    if (i == 0) {
                        st_data(dst, make_double2(x.x, 0.0), dst_hint, pol);
                        st_data(dst + L, make_double2(x.y, 0.0), dst_hint, pol);
                    } else {
                        const cplx y = sm[(d.q << LOG2L) + (L - i)];
                        const cplx oi = make_double2(0.5 * (x.y + y.y), 0.5 * (y.x - x.x));
                        st_data(dst + i, make_double2(x.x + oi.y, x.y - oi.x), dst_hint, pol);
                        st_data(dst + L + i, oi, dst_hint, pol);
                    }
                }
            }
            return;
        }
        if (flags & PF_RFFT_POST) {
            // Z = FFT_L(packed reals) is complete inside the tile (q = one transform).  Each thread
            // finishes its own bins i: X[i] = E_i + O_i w^i, X[i + L] = E_i - O_i w^i with
            // O_i = (conj(swap Z_i) + swap Z_{L-i}) / 2, E_i = Z_i - i O_i, the partner Z_{L-i}
            // fetched from a planar image of the outputs; same operations, in the same order, as
            // rfft_post_kernel, so the bits do not depend on which path ran.
            Sync::sync();                                // the last exchange has been read
#pragma unroll
            for (int m = 0; m < 16; m++) sm[(d.q << LOG2L) + d.t + m * T] = v[m];
            Sync::sync();
            if (d.valid) {
                cplx *dst = p.dst + (d.item + dst_item_delta) * p.dst_batch_stride + d.d_xoff;
#pragma unroll
                for (int m = 0; m < 16; m++) {
                    const int i = d.t + m * T;
                    const cplx x = v[m];
                    if (i == 0) {
                        st_data(dst, make_double2(x.x + x.y, 0.0), dst_hint, pol);
                        st_data(dst + L, make_double2(x.x - x.y, 0.0), dst_hint, pol);
                    } else {
                        const cplx y = sm[(d.q << LOG2L) + (L - i)];
                        const cplx oi = make_double2(0.5 * (x.y + y.y), 0.5 * (y.x - x.x));
                        const cplx ei = make_double2(x.x + oi.y, x.y - oi.x);
                        const cplx ti = cmul(oi, ldg_c(p.r2tw + i));
                        st_data(dst + i, cadd(ei, ti), dst_hint, pol);
                        st_data(dst + i + L, csub(ei, ti), dst_hint, pol);
                    }
                }
            }
            return;
        }
    }
    if constexpr (DENSE) {
        constexpr int C = 1 << plan_log2c(LOG2L), NT = plan_threads(LOG2L);
        Sync::sync();                                    // the last exchange (or the dense image) has been read
#pragma unroll
        for (int m = 0; m < 16; m++)
            sm[d.q * DPITCH + ((flags & PF_OUT_REV) ? rev_bits(d.t + m * T, LOG2L) : d.t + m * T)] = v[m];
        Sync::sync();
        cplx *g = p.dst + (tile * C + dst_item_delta) * p.dst_batch_stride;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int e = (int)threadIdx.x + NT * j;
            if (e < dense_valid) st_data(g + e, sm[(e >> LOG2L) * DPITCH + (e & (L - 1))], dst_hint, pol);
        }
        return;
    }
    if (d.valid) {
        cplx *dst = p.dst + (d.item + dst_item_delta) * p.dst_batch_stride + d.d_xoff;
        const long long d_i = p.d_i;
        if ((flags & PF_OUT_REV) && !out_reorder) {
#pragma unroll
            for (int m = 0; m < 16; m++) st_data(dst + (long long)rev_bits(d.t + m * T, LOG2L) * d_i, v[m], dst_hint, pol);
        } else if (flags & PF_SPLIT_OUT) {
            cplx *item0 = p.dst + (d.item + dst_item_delta) * p.dst_batch_stride;
#pragma unroll
            for (int m = 0; m < 16; m++) st_data(item0 + split_offset(p, d.d_xoff + (long long)(d.t + m * T) * d_i), v[m], dst_hint, pol);
        } else {
            cplx *d0 = dst + (long long)d.t * d_i;
            const long long step = (long long)T * d_i;
#pragma unroll
            for (int m = 0; m < 16; m++) st_data(d0 + m * step, v[m], dst_hint, pol);
        }
    }
}

// Lane maps by tile kind: 0 = column-like (Q,Q); 1 = row-like (T,T); 2 = row-like load,
// column-like (transposed) store (T,Q).  L < 128 always uses (Q,Q).
__host__ __device__ constexpr int kind_lm(int l, int kind) { return l >= 7 && kind != 0 ? MAP_T : MAP_Q; }
__host__ __device__ constexpr int kind_sm(int l, int kind) { return l >= 7 && kind == 1 ? MAP_T : MAP_Q; }
__host__ __device__ constexpr size_t dense_smem_bytes(int l) { return (size_t)(((1u << l) + 1) << plan_log2c(l)) * sizeof(cplx); }
__host__ __device__ constexpr size_t kind_smem_bytes(int l, int kind) {
    return (kind_lm(l, kind) == MAP_Q && kind_sm(l, kind) == MAP_Q)
               ? (l < 7 && kind == 1 && dense_smem_bytes(l) > plan_smem_bytes(l) ? dense_smem_bytes(l) : plan_smem_bytes(l))
               : plan_smem_bytes_planar(l);
}

// One tile per CTA (single-pass transforms and every pass of multi-pass ones by default).
// VARIANT: which fusion paths are compiled in -- 0 everything (real wrappers: zip, rifft pre-twiddle, rfft / split
// epilogues), 1 LEAN: none of the optional fusions (every non-first pass and un-windowed first passes), 2 lean + the
// window multiply.  Smaller code (the full variant of a 2^12 tile is ~100 KB of SASS, far beyond the instruction
// cache) and fewer flag tests on the launches that dominate the benchmarks.
constexpr unsigned LEAN_FLAGS = PF_Q_IS_BATCH | PF_IN_REV | PF_OUT_REV | PF_CONJ_IN | PF_CONJ_OUT | PF_TW_LOAD | PF_TW_STORE | PF_SCALE;
constexpr unsigned WINLEAN_FLAGS = LEAN_FLAGS | PF_WIN_REAL | PF_WIN_CPLX | PF_WIN_REV;
__host__ __device__ constexpr unsigned variant_flags(int v) { return v == 1 ? LEAN_FLAGS : (v == 2 ? WINLEAN_FLAGS : 0xffffffffu); }
template <int LOG2L, int KIND, int VARIANT>
__global__ void __launch_bounds__(plan_threads(LOG2L), plan_min_ctas(LOG2L))
fft_pass_kernel(const PassParams p) {
    NAPA_DYN_SMEM(cplx, sm);
    L2Policies pol{};                                 // plain loads/stores: no L2 policy operands
    NoHook nh;
    // L < 128: KIND 1 selects the dense-rows variant (single-pass batches of back-to-back transforms)
    run_tile<LOG2L, false, kind_lm(LOG2L, KIND), kind_sm(LOG2L, KIND), SyncAll, false, NoHook, variant_flags(VARIANT), false,
             false, (LOG2L < 7 && KIND == 1)>(p, (long long)blockIdx.x, sm, 0, 0, pol, nh);
}

// ------------------------------------------------------------------------------------------
// Fused two-pass kernel: ONE persistent cooperative launch runs both digit passes of every
// transform of a batch; the intermediate lives in L2 between them and every element crosses HBM
// once in and once out.  (The reference keeps large transforms cache-resident the same way, by
// recursing depth first: forward.lisp:76-112.)
//
// Work is cut into UNITS of `upg` consecutive transforms (about one tile per resident CTA and
// pass).  The whole grid works through the phases
//        A(0) A(1) | B(0) B(1) A(2) A(3) | B(2) B(3) A(4) A(5) | ...
// (A = first digit pass, B = second) as one flat list of tiles dealt round-robin to the CTAs, so
// that neighbouring CTAs hold neighbouring tiles (DRAM locality) and nobody owns a transform.
// Dependencies are counters, not barriers: a CTA publishes every finished tile with a release
// increment of its unit's counter and, before its first B(u) tile, checks that all A(u) tiles are
// in.  Between A(u) and B(u) lies a whole other phase of work, so that check almost never waits.
// At most two units of intermediate (<= ~40 MiB) are live at any time.
//   PK 0: natural order.  A = column digit, src -> ring slot (u & 1) as [k1][n2]; B = row digit,
//         ring -> dst stored transposed.  A(u) also waits for B(u - 2), the slot's last reader.
//   PK 1: forward, bit-reversed output.  A = column digit src -> dst, B = row digit in place.
//   PK 2: inverse of a bit-reversed input. A = row digit src -> dst, B = column digit in place.
// L2 eviction hints: caller's src / final dst are evict-first streams, the intermediate evict-last.
// ------------------------------------------------------------------------------------------
#ifdef NAPA_EMULATE
#define NAPA_GRID_CONSTANT
#else
#define NAPA_GRID_CONSTANT __grid_constant__
#endif
struct FusedParams {
    PassParams a, b;                 // pass A (first executed) and pass B, geometry as launch_pass would set it
    long long items;                 // batch
    int upg, nunits;                 // transforms per unit, units
    int lead;                        // D: pass A runs D units ahead of pass B; the ring (natural order) has D + 1 slots
    unsigned *done_a, *done_b;       // [nunits] finished-tile counters, zeroed before the launch
    int *error;                      // set to 1 if a wait timed out (never expected)
    int prefetch;                    // 1: pull each CTA's next pass-A tile into L2 while the current tile runs
};

__host__ __device__ constexpr unsigned fused_flags_a(int pk) {
    return pk == 0 ? (PF_SCALE | PF_WIN_REAL | PF_WIN_CPLX | PF_WIN_REV | PF_CONJ_IN | PF_TW_STORE | PF_ZIP2 | PF_RIFFT_PRE)
         : pk == 1 ? (PF_SCALE | PF_WIN_REAL | PF_WIN_CPLX | PF_TW_STORE | PF_OUT_REV)
                   : (PF_SCALE | PF_WIN_REAL | PF_WIN_CPLX | PF_CONJ_IN | PF_IN_REV);
}
__host__ __device__ constexpr unsigned fused_flags_b(int pk) {
    return pk == 0 ? PF_CONJ_OUT : pk == 1 ? PF_OUT_REV : (PF_IN_REV | PF_TW_LOAD | PF_CONJ_OUT);
}
// flags a pass carries whatever the call's options are (folded into the instantiation)
__host__ __device__ constexpr unsigned fused_fixed_a(int pk) {
    return pk == 0 ? PF_TW_STORE : pk == 1 ? (PF_TW_STORE | PF_OUT_REV) : (PF_CONJ_IN | PF_IN_REV);
}
__host__ __device__ constexpr unsigned fused_fixed_b(int pk) {
    return pk == 0 ? 0u : pk == 1 ? PF_OUT_REV : (PF_IN_REV | PF_TW_LOAD | PF_CONJ_OUT);
}
constexpr unsigned FUSED_WINDOW_FLAGS = PF_WIN_REAL | PF_WIN_CPLX | PF_WIN_REV | PF_ZIP2 | PF_RIFFT_PRE;
__host__ __device__ constexpr int fused_kind_a(int pk) { return pk == 2 ? 1 : 0; }
__host__ __device__ constexpr int fused_kind_b(int pk) { return pk == 0 ? 2 : (pk == 1 ? 1 : 0); }
__host__ __device__ constexpr size_t fused_smem_bytes(int la, int lb, int pk) {
    const size_t a = kind_smem_bytes(la, fused_kind_a(pk)), b = kind_smem_bytes(lb, fused_kind_b(pk));
    return ((a > b ? a : b) + 127) / 128 * 128;
}

// geometry of a column-like / row-like digit pass of a two-digit item [2^LCOL][2^LROW], as
// compile-time constants (the same values pass_geometry() computes on the host)
template <int LCOL, int LROW> __device__ __forceinline__ void fold_col_geometry(PassParams &q) {
    constexpr int LC = plan_log2c(LCOL);
    q.G = 1 << (LROW - LC); q.lg_G = LROW - LC; q.tiles_per_item = q.G; q.lg_tpi = q.lg_G;
    q.s_SA = 1LL << (LCOL + LROW); q.s_SG = 1 << LC; q.s_i = 1LL << LROW; q.s_q = 1;
    q.d_SA = q.s_SA; q.d_SG = q.s_SG; q.d_i = q.s_i; q.d_q = q.s_q;
}
template <int LCOL, int LROW> __device__ __forceinline__ void fold_row_geometry(PassParams &q) {
    constexpr int LC = plan_log2c(LROW);
    q.G = 1; q.lg_G = 0; q.tiles_per_item = 1 << (LCOL - LC); q.lg_tpi = LCOL - LC;
    q.s_SA = 1LL << (LC + LROW); q.s_SG = 0; q.s_i = 1; q.s_q = 1LL << LROW;
    q.d_SA = q.s_SA; q.d_SG = q.s_SG; q.d_i = q.s_i; q.d_q = q.s_q;
}

// This CTA's tiles of the flat phase list, in order.  With lead D = p.lead the list is
//      A(0) .. A(D-1) | B(0) A(D) | B(1) A(D+1) | B(2) A(D+2) | ...
// so that between A(u) and B(u) lie 2 (D - 1) other phases, and between B(u) and the pass that reuses its
// ring slot (A(u + D + 1), ring of D + 1 slots) lie 2 D - 1.  A unit past the end is an empty phase.
struct FlatTiles {
    long long G, items, tfull, base, s, tu;
    int upg, nunits, nphases, lgtpi, lead, ph, u;
    bool is_b;
    __device__ __forceinline__ void load_phase() {
        if (ph < lead) { is_b = false; u = ph; }
        else { const int j = ph - lead; is_b = !(j & 1); u = is_b ? (j >> 1) : lead + (j >> 1); }
        const long long left = items - (long long)u * upg;
        tu = u < nunits ? ((left < upg ? left : (long long)upg) << lgtpi) : 0;
    }
    __device__ __forceinline__ void init(const FusedParams &p, int lgtpi_) {
        G = (long long)gridDim.x; items = p.items; upg = p.upg; nunits = p.nunits; lead = p.lead; lgtpi = lgtpi_;
        nphases = p.lead + 2 * p.nunits;
        tfull = (long long)p.upg << lgtpi_;
        ph = 0; base = 0; s = (long long)blockIdx.x - G;
        load_phase();
    }
    // advances to the next tile; false when the list is exhausted
    __device__ __forceinline__ bool next(long long &tile) {
        s += G;
        while (s >= base + tu) {
            base += tu;
            if (++ph >= nphases) return false;
            load_phase();
        }
        tile = (long long)u * tfull + (s - base);
        return true;
    }
    // ring slot of unit u holds items [slot * upg, ...): what to add to an item index of unit u
    __device__ __forceinline__ long long ring_delta() const { return ((long long)(u % (lead + 1)) - u) * upg; }
};

__device__ __forceinline__ void publish_tile_relaxed(unsigned *ctr) {
#ifdef NAPA_EMULATE
    *ctr += 1;
    napa_emul::note_progress();
#else
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" :: "l"(ctr) : "memory");
#endif
}
__device__ __forceinline__ void publish_tile(unsigned *ctr) {
#ifdef NAPA_EMULATE
    *ctr += 1;
    napa_emul::note_progress();
#else
    // release at gpu scope: the CTA's stores (ordered before this thread by the preceding barrier) come first
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(ctr) : "memory");
#endif
}
// thread 0 waits until `target` tiles are in, then the CTA barrier carries the acquire to everyone
__device__ __forceinline__ void await_tiles(const unsigned *ctr, const unsigned target, int *error) {
    if (threadIdx.x == 0) {
        unsigned spins = 0;
        while (ld_acquire_u32(ctr) < target) {
            napa_nanosleep(40);
            if (++spins > (1u << 26)) { *error = 1; break; }      // seconds: a lost CTA, never expected
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void prefetch_l2(const void *p) {
#ifndef NAPA_EMULATE
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
#else
    (void)p;
#endif
}
// What a fused-kernel CTA does while a tile's loads are in flight: publish the previous tile (the release
// fence then waits in the shadow of the loads instead of holding up the CTA's next barrier) and pull
// the source lines of its NEXT pass-A tile from HBM into L2, so that tile's loads pay an L2 round trip.
// A tile is 2^LGPTS points = NL 128-byte lines: `runs` contiguous runs of `run_lines` lines, `run_stride` elements apart.
struct FusedHook {
    unsigned *pending = nullptr;
    const cplx *pf = nullptr;                // first element of the tile to prefetch (nullptr: nothing)
    long long pf_run_stride = 0;
    int pf_lg_run_lines = 0, pf_lines = 0;
    __device__ __forceinline__ void operator()() const {}
    // publish now: required before this CTA waits on a counter (its own unpublished tile may be what is missing)
    __device__ __forceinline__ void flush() {
        if (threadIdx.x == 0 && pending) publish_tile(pending);
        pending = nullptr;
    }
    __device__ __forceinline__ void loads_issued() {
        flush();
        if (pf) {
            for (int j = (int)threadIdx.x; j < pf_lines; j += (int)blockDim.x)
                prefetch_l2(pf + (long long)(j >> pf_lg_run_lines) * pf_run_stride + ((j & ((1 << pf_lg_run_lines) - 1)) << 3));
            pf = nullptr;
        }
    }
};

template <int LA, int LB, int PK, bool WIN>
__global__ void __launch_bounds__(plan_threads(LA), plan_min_ctas(LA)) fft_fused_kernel(const NAPA_GRID_CONSTANT FusedParams p) {
    static_assert(plan_threads(LA) == plan_threads(LB), "both passes of a fused pair use the same tile shape");
    constexpr int KA = fused_kind_a(PK), KB = fused_kind_b(PK);
    constexpr int LGTPI = LA + LB - plan_log2pts(LA);          // tiles per transform and pass
    constexpr bool RING = PK == 0;
    constexpr unsigned ALLOW_A = WIN ? fused_flags_a(PK) : (fused_flags_a(PK) & ~FUSED_WINDOW_FLAGS);
    NAPA_DYN_SMEM(cplx, sm);
    const L2Policies pol = make_policies();          // created once per thread, at kernel entry
    NoHook nh;
    // local copies of the pass descriptions with everything the instantiation knows folded in
    PassParams pa = p.a, pb = p.b;
    if (PK == 2) { fold_row_geometry<LB, LA>(pa); fold_col_geometry<LB, LA>(pb); }
    else         { fold_col_geometry<LA, LB>(pa); fold_row_geometry<LA, LB>(pb); }
    if (PK == 0) { pb.d_SA = 1 << plan_log2c(LB); pb.d_SG = 0; pb.d_i = 1LL << LA; pb.d_q = 1; }
    pa.flags = (p.a.flags & ~fused_fixed_a(PK) & ALLOW_A) | fused_fixed_a(PK);
    pb.flags = (p.b.flags & ~fused_fixed_b(PK) & fused_flags_b(PK)) | fused_fixed_b(PK);
    pa.src_hint = HINT_STREAM; pa.dst_hint = HINT_KEEP; pb.src_hint = HINT_KEEP; pb.dst_hint = HINT_STREAM;
    pa.lgN = LA + LB; pb.lgN = LA + LB;

    // source lines of a pass-A tile, for the L2 prefetch: runs of contiguous 128-byte lines
    constexpr int LGPTS = plan_log2pts(LA);
    constexpr bool A_IS_ROW = PK == 2;
    constexpr int A_LG_RUN_LINES = A_IS_ROW ? LA - 3 : plan_log2c(LA) - 3;       // a row of 2^LA points / C adjacent columns
    const bool can_prefetch = p.prefetch && !(pa.flags & (PF_ZIP2 | PF_RIFFT_PRE));
    int known_a = -1, known_b = -1;                               // newest unit whose pass is known to be complete
    FusedHook hook;
    FlatTiles it;
    it.init(p, LGTPI);
    long long tile;
    while (it.next(tile)) {
        __syncthreads();                             // the previous tile is done with sm and has issued its stores
        // this CTA's next tile: if it belongs to a pass A, its source lines are prefetched into L2 now
        if (can_prefetch) {
            FlatTiles nx = it;
            long long t2;
            if (nx.next(t2) && !nx.is_b) {
                const long long item2 = t2 >> LGTPI; const int r2 = (int)(t2 & ((1 << LGTPI) - 1));
                hook.pf = pa.src + item2 * pa.src_batch_stride + (A_IS_ROW ? (long long)r2 * pa.s_SA : (long long)r2 * pa.s_SG);
                hook.pf_run_stride = A_IS_ROW ? pa.s_q : pa.s_i;
                hook.pf_lg_run_lines = A_LG_RUN_LINES; hook.pf_lines = 1 << (LGPTS - 3);
            }
        }
        const int u = it.u;
        const long long ring_delta = RING ? it.ring_delta() : 0;
        if (!it.is_b) {
            if (RING && u > p.lead && u - p.lead - 1 > known_b) { hook.flush(); await_tiles(p.done_b + (u - p.lead - 1), (unsigned)it.tfull, p.error); known_b = u - p.lead - 1; }
            run_tile<LA, false, kind_lm(LA, KA), kind_sm(LA, KA), SyncAll, false, FusedHook, ALLOW_A>(pa, tile, sm, 0, ring_delta, pol, hook);
            hook.pending = p.done_a + u;
        } else {
            if (u > known_a) { hook.flush(); await_tiles(p.done_a + u, (unsigned)it.tu, p.error); known_a = u; }
            run_tile<LB, true, kind_lm(LB, KB), kind_sm(LB, KB), SyncAll, false, FusedHook, fused_flags_b(PK)>(pb, tile, sm, ring_delta, 0, pol, hook);
            hook.pending = RING ? p.done_b + u : nullptr;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && hook.pending) publish_tile(hook.pending);
}

// ------------------------------------------------------------------------------------------
// Warp-specialised variant of the fused kernel (TMA-fed): the same flat phase list and tile counters,
// but a CTA is `NC` compute threads plus ONE producer warp, and every tile's input is brought into
// shared memory by the TMA engine while the previous tile is still being computed:
//   * column tiles (L rows of C adjacent points): one cp.async.bulk.tensor.3d through a tensor map of the
//     operand ([item][row][2*col] doubles, box = one tile);
//   * row tiles (C consecutive rows = one contiguous block): one cp.async.bulk.
// Two buffers per CTA; a buffer holds a tile's staged input and then its exchanges.  Protocol per
// buffer: full[b] (producer's expect_tx + the copy's bytes) -> compute threads read, compute, store ->
// every compute thread arrives on done[b] -> the producer publishes the tile (release increment of its
// unit's counter) and may refill the buffer.  The producer alone touches the inter-CTA counters: it waits
// for a tile's dependency (all pass-A tiles of its unit in / the ring slot's last reader done) BEFORE
// issuing its copy, one to two tiles ahead of the compute threads, which therefore never wait on
// global memory or on another CTA -- only on their own mbarriers.
// Tiles whose load cannot be staged (PK 2 pass A: a bit-reversed row load goes through the swizzled reorder
// buffer) are loaded by the compute threads as in fft_fused_kernel; the producer only arms their barrier.
// ------------------------------------------------------------------------------------------
struct FusedTmaParams {
    alignas(64) unsigned long long tmap_a[16];     // CUtensorMap of pass A's source (PK 0 / 1)
    alignas(64) unsigned long long tmap_b[16];     // CUtensorMap of pass B's source (PK 2)
    FusedParams f;
};

__host__ __device__ constexpr bool fused_tma_stage_a(int pk) { return pk != 2; }
__host__ __device__ constexpr size_t fused_tma_smem_bytes(int la, int lb, int pk) { return 2 * fused_smem_bytes(la, lb, pk) + 64; }

template <int LA, int LB, int PK, bool WIN>
__global__ void __launch_bounds__(plan_threads(LA) + 32, plan_min_ctas(LA) > 1 ? plan_min_ctas(LA) - 1 : 1)
fft_fused_tma_kernel(const NAPA_GRID_CONSTANT FusedTmaParams P) {
    static_assert(plan_threads(LA) == plan_threads(LB), "both passes of a fused pair use the same tile shape");
    const FusedParams &p = P.f;
    constexpr int NC = plan_threads(LA);
    constexpr int KA = fused_kind_a(PK), KB = fused_kind_b(PK);
    constexpr int LGPTS = plan_log2pts(LA);
    constexpr int LGTPI = LA + LB - LGPTS;
    constexpr bool RING = PK == 0;
    constexpr bool STA = fused_tma_stage_a(PK);
    constexpr unsigned ALLOW_A = WIN ? fused_flags_a(PK) : (fused_flags_a(PK) & ~FUSED_WINDOW_FLAGS);
    constexpr size_t BUF = fused_smem_bytes(LA, LB, PK) / sizeof(cplx);
    constexpr unsigned TILE_BYTES = (unsigned)(sizeof(cplx) << LGPTS);
    NAPA_DYN_SMEM(cplx, sm);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(sm + 2 * BUF), *done = full + 2;
    const L2Policies pol = make_policies();
    PassParams pa = p.a, pb = p.b;
    if (PK == 2) { fold_row_geometry<LB, LA>(pa); fold_col_geometry<LB, LA>(pb); }
    else         { fold_col_geometry<LA, LB>(pa); fold_row_geometry<LA, LB>(pb); }
    if (PK == 0) { pb.d_SA = 1 << plan_log2c(LB); pb.d_SG = 0; pb.d_i = 1LL << LA; pb.d_q = 1; }
    pa.flags = (p.a.flags & ~fused_fixed_a(PK) & ALLOW_A) | fused_fixed_a(PK);
    pb.flags = (p.b.flags & ~fused_fixed_b(PK) & fused_flags_b(PK)) | fused_fixed_b(PK);
    pa.src_hint = HINT_STREAM; pa.dst_hint = HINT_KEEP; pb.src_hint = HINT_KEEP; pb.dst_hint = HINT_STREAM;
    pa.lgN = LA + LB; pb.lgN = LA + LB;

    if (threadIdx.x == 0) {
        mbar_init(full, 1); mbar_init(full + 1, 1); mbar_init(done, NC); mbar_init(done + 1, NC);
        fence_proxy_async();
    }
    __syncthreads();
    FlatTiles it;
    it.init(p, LGTPI);
    long long tile;

    if ((int)threadIdx.x >= NC) {
        // ---------------- producer: one elected thread ----------------
        if ((int)threadIdx.x != NC) return;
        unsigned *owed[2] = { nullptr, nullptr };        // counter the tile in buffer b will be published on
        unsigned *unpublished[2] = { nullptr, nullptr }; // ... and, once the tile is complete, until it has been
        bool busy[2] = { false, false };                 // buffer b holds a tile whose completion has not been seen
        unsigned par_done[2] = { 0, 0 };
        int known_a = -1, known_b = -1;
        // completion of the tile in buffer b (its compute threads have issued their stores and left the buffer)
        auto complete = [&](const int b, const bool blocking) {
            if (!busy[b]) return;
            if (blocking) mbar_wait(done + b, par_done[b]);
            else if (!mbar_test(done + b, par_done[b])) return;
            par_done[b] ^= 1u; busy[b] = false;
            unpublished[b] = owed[b]; owed[b] = nullptr;
        };
        // relaxed increment of the counters of every complete tile, older buffer first; the caller has fenced
        auto publish_fenced = [&](const int older) {
            if (unpublished[older]) { publish_tile_relaxed(unpublished[older]); unpublished[older] = nullptr; }
            if (unpublished[older ^ 1]) { publish_tile_relaxed(unpublished[older ^ 1]); unpublished[older ^ 1] = nullptr; }
        };
        // the dependency of a tile: the counter that must reach `need` before its copy may be issued
        auto dependency = [&](const FlatTiles &t, const unsigned *&dep, unsigned &need) {
            dep = nullptr; need = 0;
            if (t.is_b) { if (t.u > known_a) { dep = p.done_a + t.u; need = (unsigned)t.tu; } }
            else if (RING && t.u > p.lead && t.u - p.lead - 1 > known_b) { dep = p.done_b + (t.u - p.lead - 1); need = (unsigned)t.tfull; }
        };
        auto dependency_met = [&](const FlatTiles &t) { if (t.is_b) known_a = t.u; else known_b = t.u - p.lead - 1; };
        // One fence.acq_rel.gpu per tile does double duty.  RELEASE: the finished tiles' stores (ordered before this
        // thread by done[b]) come before the relaxed increments that publish them.  ACQUIRE: the relaxed read of the
        // NEXT tile's dependency counter, issued just before the fence, comes before that tile's copy.  The fence's
        // latency (the CTA's stores draining to L2) is paid after the current tile's copy has been issued and
        // overlaps the counter read.
        bool have_next = it.next(tile);
        for (int k = 0; have_next; k++) {
            const int b = k & 1;
            // 1. dependency of tile k: normally established by the previous iteration's fence; else wait here, and
            //    keep publishing finished tiles meanwhile (they may be exactly what the rest of the grid waits for)
            const unsigned *dep; unsigned need;
            dependency(it, dep, need);
            if (dep) {
                unsigned spins = 0;
                while (ld_relaxed_u32(dep) < need) {
                    complete(b, false); complete(b ^ 1, false);
                    if (unpublished[0] || unpublished[1]) { fence_acq_rel_gpu(); publish_fenced(b); }
                    napa_nanosleep(32);
                    if (++spins > (1u << 26)) { *p.error = 1; break; }
                }
                fence_acq_rel_gpu();
                dependency_met(it);
                // the other CTAs' generic-proxy stores, acquired above, come before this thread's TMA reads of them
                fence_proxy_async();
            }
            // 2. the buffer: tile k - 2 must have left it (the mbarrier orders its compute threads' accesses first)
            complete(b, true);
            // 3. issue the copy at once; 4. only then pay for the release fences of the finished tiles
            const long long item = tile >> LGTPI; const int r = (int)(tile & ((1 << LGTPI) - 1));
            const long long ring_delta = RING ? it.ring_delta() : 0;
            cplx *buf = sm + (size_t)b * BUF;
            if (!it.is_b) {
                if (STA) {
                    // column tile of pass A: columns [r*C, r*C + C) of item `item` of the source
                    mbar_expect_tx(full + b, TILE_BYTES);
#ifdef NAPA_EMULATE
                    const cplx *g = pa.src + item * pa.src_batch_stride + (long long)r * pa.s_SG;
                    for (int i = 0; i < (1 << LA); i++) bulk_g2s(buf + ((size_t)i << plan_log2c(LA)), g + (long long)i * pa.s_i, (unsigned)(sizeof(cplx) << plan_log2c(LA)), full + b, 0, pol);
                    mbar_emul_complete(full + b);
#else
                    tma_tile_g2s(buf, P.tmap_a, 2 * (r << plan_log2c(LA)), 0, (int)item, full + b, HINT_STREAM, pol);
#endif
                } else {
                    mbar_arrive(full + b);                // not staged: the compute threads load it themselves
                }
                owed[b] = p.done_a + it.u;
            } else {
                mbar_expect_tx(full + b, TILE_BYTES);
                if (PK == 2) {
                    // column tile of pass B, in place on dst
#ifdef NAPA_EMULATE
                    const cplx *g = pb.src + item * pb.src_batch_stride + (long long)r * pb.s_SG;
                    for (int i = 0; i < (1 << LB); i++) bulk_g2s(buf + ((size_t)i << plan_log2c(LB)), g + (long long)i * pb.s_i, (unsigned)(sizeof(cplx) << plan_log2c(LB)), full + b, 0, pol);
#else
                    tma_tile_g2s(buf, P.tmap_b, 2 * (r << plan_log2c(LB)), 0, (int)item, full + b, HINT_KEEP, pol);
#endif
                } else {
                    // row tile: C consecutive rows of the intermediate = one contiguous block
                    bulk_g2s(buf, pb.src + (item + ring_delta) * pb.src_batch_stride + (long long)r * pb.s_SA, TILE_BYTES, full + b, HINT_KEEP, pol);
                }
                mbar_emul_complete(full + b);
                owed[b] = RING ? p.done_b + it.u : nullptr;
            }
            busy[b] = true;
            // 4. next tile's dependency counter (relaxed, not waited for) + the fence + the publications
            have_next = it.next(tile);
            const unsigned *dep2 = nullptr; unsigned need2 = 0, seen2 = 0;
            if (have_next) { dependency(it, dep2, need2); if (dep2) seen2 = ld_relaxed_u32(dep2); }
            if (unpublished[0] || unpublished[1] || dep2) {
                fence_acq_rel_gpu();
                publish_fenced(b);
                if (dep2 && seen2 >= need2) { dependency_met(it); fence_proxy_async(); }
            }
        }
        complete(0, true); complete(1, true);
        fence_acq_rel_gpu(); publish_fenced(0);
        return;
    }

    // ---------------- compute threads ----------------
    NoHook nh;
    unsigned par_full[2] = { 0, 0 };
    for (int k = 0; it.next(tile); k++) {
        const int b = k & 1;
        cplx *buf = sm + (size_t)b * BUF;
        const long long ring_delta = RING ? it.ring_delta() : 0;
        if (!it.is_b) {
            if (STA) {
                run_tile<LA, false, kind_lm(LA, KA), kind_sm(LA, KA), SyncCompute<NC>, true, NoHook, ALLOW_A, true, true>(
                    pa, tile, buf, 0, ring_delta, pol, nh, full + b, par_full[b]);
            } else {
                mbar_wait(full + b, par_full[b]);
                run_tile<LA, false, kind_lm(LA, KA), kind_sm(LA, KA), SyncCompute<NC>, true, NoHook, ALLOW_A>(pa, tile, buf, 0, ring_delta, pol, nh);
            }
        } else {
            run_tile<LB, true, kind_lm(LB, KB), kind_sm(LB, KB), SyncCompute<NC>, true, NoHook, fused_flags_b(PK), true, true>(
                pb, tile, buf, ring_delta, 0, pol, nh, full + b, par_full[b]);
        }
        par_full[b] ^= 1u;
        mbar_arrive(done + b);        // stores issued, buffer free (the producer fences the proxies before it refills it)
    }
}

#ifndef NAPA_TABLES_ONLY   /* the small, non-template kernels below belong to napafft.cu alone */
// ------------------------------------------------------------------------------------------
// n <= 8: one thread per transform
// ------------------------------------------------------------------------------------------
struct TinyParams {
    const cplx *src; cplx *dst;
    long long src_batch_stride, dst_batch_stride, batch;
    int lgn;
    unsigned flags;
    double scale;
    const void *window;
    long long window_batch_stride;
    const cplx *r2tw;
};

__global__ void __launch_bounds__(256) fft_tiny_kernel(const TinyParams p) {
    const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= p.batch) return;
    const int n = 1 << p.lgn;
    const cplx *src = p.src + item * p.src_batch_stride;
    cplx *dst = p.dst + item * p.dst_batch_stride;
    cplx v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (i < n) {
            const int pos = (p.flags & PF_IN_REV) ? rev_bits(i, p.lgn) : i;
            cplx x = src[pos];
            if (p.flags & PF_RIFFT_PRE) {
                cplx y = src[pos + n];
                x = cadd(cadd(x, y), cmul(csub(x, y), ldg_c(p.r2tw + pos)));
            }
            if (p.flags & PF_SCALE) x = cscale(x, p.scale);
            if (p.flags & (PF_WIN_REAL | PF_WIN_CPLX)) {
                long long widx = pos;
                if (p.flags & PF_WIN_REV) widx = rev_bits(pos, p.lgn);
                widx += item * p.window_batch_stride;
                if (p.flags & PF_WIN_REAL) x = cscale(x, __ldg(reinterpret_cast<const double *>(p.window) + widx));
                else x = cmul(x, ldg_c(reinterpret_cast<const cplx *>(p.window) + widx));
            }
            if (p.flags & PF_CONJ_IN) x = cconj(x);
            v[i] = x;
        } else v[i] = make_double2(0.0, 0.0);
    }
    if (p.lgn == 1) Radix<2>::run(v);
    else if (p.lgn == 2) Radix<4>::run(v);
    else if (p.lgn == 3) Radix<8>::run(v);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (k < n) {
            cplx y = v[k];
            if (p.flags & PF_CONJ_OUT) y = cconj(y);
            const int pos = (p.flags & PF_OUT_REV) ? rev_bits(k, p.lgn) : k;
            dst[pos] = y;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Stand-alone bit reversal (bit-reversal.lisp:146-252): out[i] = in[rev_w(i)], in place allowed.
// Index = [hi:A][mid][lo:A]; a CTA exchanges the (2^A x 2^A) tiles of mid and rev(mid) through
// shared memory, so both its global reads and writes are 2^A contiguous elements.
// ------------------------------------------------------------------------------------------
template <typename E, int A>
__global__ void __launch_bounds__(256) bitrev_tiled_kernel(const E *src, E *dst, int w, long long src_stride,
                                                           long long dst_stride, int pairs_per_item,
                                                           const unsigned *pair_table) {
    constexpr int S = 1 << A;
    __shared__ E tileA[S][S + 1];
    __shared__ E tileB[S][S + 1];
    const long long item = blockIdx.x / pairs_per_item;
    const int pr = blockIdx.x - item * pairs_per_item;
    const unsigned mu = pair_table[2 * pr], mu2 = pair_table[2 * pr + 1];
    const E *s = src + item * src_stride;
    E *d = dst + item * dst_stride;
    const int hs = w - A;                   // shift of the hi field
    const int tx = threadIdx.x & (S - 1), ty0 = threadIdx.x >> A;
    constexpr int ROWS = 256 / S;
    // read tile(mu2) into A and (if distinct) tile(mu) into B: element (hi, lo)
    for (int hi = ty0; hi < S; hi += ROWS) {
        tileA[hi][tx] = s[((long long)hi << hs) + ((long long)mu2 << A) + tx];
        if (mu != mu2) tileB[hi][tx] = s[((long long)hi << hs) + ((long long)mu << A) + tx];
    }
    __syncthreads();
    // out[hi, mu, lo] = in[rev lo, mu2, rev hi]
    const int rlo = rev_bits(tx, A);
    for (int hi = ty0; hi < S; hi += ROWS) {
        const int rhi = rev_bits(hi, A);
        d[((long long)hi << hs) + ((long long)mu << A) + tx] = tileA[rlo][rhi];
        if (mu != mu2) d[((long long)hi << hs) + ((long long)mu2 << A) + tx] = tileB[rlo][rhi];
    }
}

// small sizes (n <= 1024 elements): a CTA takes `per_cta` items (2048 elements) through shared memory.
// An item is stored as rows of 32 elements with a pitch of 33, so that the bit-reversed gather
// (lane stride n/32 elements) spreads over the banks instead of hitting one.
template <typename E>
__global__ void __launch_bounds__(256) bitrev_small_kernel(const E *src, E *dst, int w, long long src_stride,
                                                           long long dst_stride, long long batch, int per_cta) {
    NAPA_DYN_SMEM(E, sm);
    const int n = 1 << w;
    const int item_pitch = n >= 32 ? (n >> 5) * 33 : n;          // whole rows only; shorter items are packed
    const long long item0 = (long long)blockIdx.x * per_cta;
    const int items = (int)(batch - item0 < per_cta ? batch - item0 : per_cta);
    const int total = items << w;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int it = e >> w, i = e & (n - 1);
        sm[it * item_pitch + (i >> 5) * 33 + (i & 31)] = src[(item0 + it) * src_stride + i];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int it = e >> w, i = e & (n - 1), r = rev_bits(i, w);
        dst[(item0 + it) * dst_stride + i] = sm[it * item_pitch + (r >> 5) * 33 + (r & 31)];
    }
}

// ------------------------------------------------------------------------------------------
// complex-samplify on the device (support.lisp:64-93): real double / complex single / real single
// input widened to complex double, so that the host ships 8 or 4 bytes per sample over PCIe
// instead of 16.  src_type: 1 = f64 real, 2 = complex f32, 3 = f32 real.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) widen_kernel(const void *__restrict__ src, int src_type, cplx *dst, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    cplx v;
    if (src_type == 1) v = make_double2(reinterpret_cast<const double *>(src)[i], 0.0);
    else if (src_type == 2) { const float *f = reinterpret_cast<const float *>(src) + 2 * i; v = make_double2((double)f[0], (double)f[1]); }
    else v = make_double2((double)reinterpret_cast<const float *>(src)[i], 0.0);
    dst[i] = v;
}

// ------------------------------------------------------------------------------------------
// Overlap-add streaming filter (example.lisp:81-114, filter-chunks): half-overlapping chunks of a
// real signal, each windowed-fft'ed, multiplied by a filter in the bit-reversed spectrum, inverse
// transformed, renormalised to its input energy and accumulated.  The chunks are independent, so
// the whole signal is one batch: gather (+ energy), two batched c2c launches, energy, overlap-add.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_256(double x) {
    __shared__ double red[256];
    red[threadIdx.x] = x;
    __syncthreads();
#pragma unroll
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
}
// chunk c = samples [c*half, c*half + chunk) zero-padded past `len`, widened to complex; energy[c] = sum |x|^2
__global__ void __launch_bounds__(256) oa_gather_kernel(const double *__restrict__ vec, long long len, cplx *X,
                                                        double *energy, int lg_chunk) {
    const long long chunk = 1LL << lg_chunk, half = chunk >> 1;
    const long long base = (long long)blockIdx.x * half;
    cplx *x = X + (long long)blockIdx.x * chunk;
    double acc = 0.0;
    for (long long j = threadIdx.x; j < chunk; j += 256) {
        const double v = base + j < len ? vec[base + j] : 0.0;
        x[j] = make_double2(v, 0.0);
        acc += v * v;
    }
    acc = block_sum_256(acc);
    if (threadIdx.x == 0) energy[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(256) oa_energy_kernel(const cplx *__restrict__ X, double *energy, int lg_chunk) {
    const long long chunk = 1LL << lg_chunk;
    const cplx *x = X + (long long)blockIdx.x * chunk;
    double acc = 0.0;
    for (long long j = threadIdx.x; j < chunk; j += 256) { const cplx v = x[j]; acc += v.x * v.x + v.y * v.y; }
    acc = block_sum_256(acc);
    if (threadIdx.x == 0) energy[blockIdx.x] = acc;
}
// dst[j] = (0 + s(c0) * X[c0][j - c0*half]) + s(c1) * X[c1][j - c1*half], c1 = j / half, c0 = c1 - 1:
// the accumulation order of the reference's chunk loop; s(c) = 0.5 * sqrt(e_old[c] / e_new[c])
__global__ void __launch_bounds__(256) oa_add_kernel(const cplx *__restrict__ X, const double *__restrict__ e_old,
                                                     const double *__restrict__ e_new, cplx *dst, long long len, int lg_chunk) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= len) return;
    const long long chunk = 1LL << lg_chunk, half = chunk >> 1;
    const long long c1 = j >> (lg_chunk - 1), c0 = c1 - 1;
    cplx acc = make_double2(0.0, 0.0);
    if (c0 >= 0) {
        const double s0 = 0.5 * sqrt(e_old[c0] / e_new[c0]);
        const cplx a = X[c0 * chunk + (j - c0 * half)];
        acc = make_double2(s0 * a.x, s0 * a.y);
    }
    const double s1 = 0.5 * sqrt(e_old[c1] / e_new[c1]);
    const cplx b = X[c1 * chunk + (j - c1 * half)];
    acc.x += s1 * b.x; acc.y += s1 * b.y;
    dst[j] = acc;
}

// ------------------------------------------------------------------------------------------
// real wrappers
// ------------------------------------------------------------------------------------------
// rfft epilogue = fft-swizzled-reals split (real.lisp:16-30) + final radix-2 (real.lisp:128-134),
// in place on dst (n complex per item) whose first half holds Z = FFT_{n/2}(packed reals).
// One thread owns the index pair (i, h-i) and the four cells it touches.
__global__ void __launch_bounds__(256) rfft_post_kernel(cplx *dst, long long dst_stride, long long batch, long long n,
                                                        const cplx *__restrict__ r2tw) {
    const long long h = n >> 1;
    const long long per = (h >> 1) + 1;                  // i = 0 .. h/2
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per * batch) return;
    const long long item = gid / per, i = gid - item * per;
    cplx *d = dst + item * dst_stride;
    if (i == 0) {
        cplx z = d[0];                                    // E0 = re, O0 = im, tw[0] = 1
        d[0] = make_double2(z.x + z.y, 0.0);
        d[h] = make_double2(z.x - z.y, 0.0);
        return;
    }
    const long long j = h - i;
    cplx x = d[i], y = d[j];
    // O_i = .5*(conj(swap x) + swap y), O_j = .5*(conj(swap y) + swap x); E = Z - i*O
    cplx oi = make_double2(0.5 * (x.y + y.y), 0.5 * (y.x - x.x));
    cplx oj = make_double2(0.5 * (y.y + x.y), 0.5 * (x.x - y.x));
    cplx ei = make_double2(x.x + oi.y, x.y - oi.x);
    cplx ej = make_double2(y.x + oj.y, y.y - oj.x);
    cplx ti = cmul(oi, ldg_c(r2tw + i));
    d[i] = cadd(ei, ti); d[i + h] = csub(ei, ti);
    if (j != i) {
        cplx tj = cmul(oj, ldg_c(r2tw + j));
        d[j] = cadd(ej, tj); d[j + h] = csub(ej, tj);
    }
}

// %2rfft epilogue (real.lisp:16-30 on a 2n array): dst[0..n) = FFT(v1), dst[n..2n) = FFT(v2)
__global__ void __launch_bounds__(256) split2_post_kernel(cplx *dst, long long dst_stride, long long batch, long long n) {
    const long long per = (n >> 1) + 1;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per * batch) return;
    const long long item = gid / per, i = gid - item * per;
    cplx *d = dst + item * dst_stride;
    if (i == 0) {
        cplx z = d[0];
        d[n] = make_double2(z.y, 0.0);
        d[0] = make_double2(z.x, 0.0);
        return;
    }
    const long long j = n - i;
    cplx x = d[i], y = d[j];
    cplx oi = make_double2(0.5 * (x.y + y.y), 0.5 * (y.x - x.x));
    cplx oj = make_double2(0.5 * (y.y + x.y), 0.5 * (x.x - y.x));
    d[n + i] = oi;
    d[2 * n - i] = oj;
    d[i] = make_double2(x.x + oi.y, x.y - oi.x);
    if (j != i) d[j] = make_double2(y.x + oj.y, y.y - oj.x);
    else d[i] = make_double2(y.x + oj.y, y.y - oj.x);     // reference writes vec[h-i] last
}

// pack two real vectors into one complex vector (real.lisp:45-47)
__global__ void __launch_bounds__(256) zip2_kernel(const double *v1, const double *v2, cplx *dst, long long n,
                                                   long long batch, long long v_stride, long long dst_stride) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * batch) return;
    const long long item = gid / n, i = gid - item * n;
    dst[item * dst_stride + i] = make_double2(v1[item * v_stride + i], v2[item * v_stride + i]);
}

// pointwise helpers used by device-resident pipelines (convolution building block, tests)
__global__ void __launch_bounds__(256) cmul_pointwise_kernel(cplx *a, const cplx *__restrict__ b, long long n,
                                                             long long batch, long long b_stride) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * batch) return;
    const long long item = gid / n, i = gid - item * n;
    a[gid] = cmul(a[gid], b[item * b_stride + i]);
}

#endif  // NAPA_TABLES_ONLY

}  // namespace napa
