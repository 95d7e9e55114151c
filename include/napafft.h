/*
 * napafft.h -- C ABI of libnapafft_b200.so: the B200 (sm_100a) replacement for Napa-FFT3's hot
 * path.  The reference (pkhuong/Napa-FFT3, Common Lisp) has no FFI layer of its own; its operator
 * boundary is the set of compiled closures of interface.lisp and the keyword front-ends built on
 * them.  Each entry point below names the reference interface it replaces (file:line); the
 * sb-alien forms a maintainer would add on the Lisp side are in INTEGRATION.md and lisp/.
 *
 * Conventions
 *   - plain C: pointers and sizes only; every function returns a napafft_status (0 = ok) and never
 *     throws; napafft_last_error() gives the thread-local message of the last failure.
 *   - complex data = interleaved (re, im) doubles, exactly the storage of an SBCL
 *     (simple-array (complex double-float) (*)); real data = doubles.
 *   - synchronous like the reference: a host-pointer call returns when dst is complete.
 *   - host-pointer entry points stage through the library's page-locked ring (or DMA directly if
 *     the caller's buffer is already page-locked, see napafft_host_register).
 *   - *_dev entry points take device pointers (as void*) valid on the current CUDA device plus a
 *     cudaStream_t (as void*, NULL = default stream) and are asynchronous on that stream.
 *   - there is no CPU fallback: without a usable CUDA device every compute call fails with
 *     NAPA_E_NO_DEVICE / NAPA_E_CUDA.
 */
#ifndef NAPAFFT_H
#define NAPAFFT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NAPAFFT_ABI_VERSION 2

typedef enum {
    NAPA_OK = 0,
    NAPA_E_NOT_POW2 = 1,      /* (assert (power-of-two-p n)), easy-interface.lisp:55 */
    NAPA_E_SHORT_BUFFER = 2,  /* (assert (>= (length vec) n)), easy-interface.lisp:56-57 */
    NAPA_E_NULL = 3,
    NAPA_E_BAD_ENUM = 4,      /* ecase failures of interface.lisp:12-31 */
    NAPA_E_CUDA = 5,
    NAPA_E_NCCL = 6,
    NAPA_E_UNSUPPORTED = 7,   /* n > 2^32 (interface.lisp:86) or a misaligned device pointer */
    NAPA_E_NO_DEVICE = 8
} napafft_status;

/* direction: interface.lisp:3-4  (1 :fwd) / (-1 :inv :bwd) */
#define NAPA_FWD 1
#define NAPA_INV (-1)
/* scaling: interface.lisp:6-7, factors of interface.lisp:41-48 */
#define NAPA_SCALE_NONE 0 /* nil, 1  */
#define NAPA_SCALE_INV 1  /* t, :inv : 1/n */
#define NAPA_SCALE_SQRT 2 /* sqrt, :sqrt : 1/sqrt n forward, (1/n)/(1/sqrt n) inverse */
/* windowing: interface.lisp:9-10 */
#define NAPA_WINDOW_NONE 0
#define NAPA_WINDOW_REAL 1    /* float, real-sample       */
#define NAPA_WINDOW_COMPLEX 2 /* complex, complex-sample  */
/* input sample types accepted by complex-samplify (support.lisp:64-93) on the device */
#define NAPA_SAMPLE_C64 0 /* (complex double-float): no conversion */
#define NAPA_SAMPLE_F64 1 /* double-float reals          */
#define NAPA_SAMPLE_C32 2 /* (complex single-float)      */
#define NAPA_SAMPLE_F32 3 /* single-float reals          */
/* element type of bit-reverse: interface.lisp:103-104 */
#define NAPA_ELT_COMPLEX 0
#define NAPA_ELT_REAL 1

/* ---- life cycle ------------------------------------------------------------------------- */
/* Binds the calling process to CUDA device `device` (one process per GPU) and builds nothing
 * eagerly; plans, twiddle tables and staging buffers are created on first use and memoised,
 * like %ensure-fft / %ensure-twiddles / %ensure-reverse (interface.lisp:81-163). */
int napafft_init(int device);
int napafft_shutdown(void);
int napafft_abi_version(void);
const char *napafft_last_error(void);
/* number of kernel launches issued by this library since napafft_init (bench.py gpu_launches) */
uint64_t napafft_launch_count(void);

/* ---- host-pointer entry points (what the Lisp easy interface binds) ------------------------ */
/* Replaces the closures of generate-fft / get-fft / get-windowed-fft (interface.lisp:33-74,
 * 165-217) as driven by fft / ifft (easy-interface.lisp:44-95).
 *   src, dst : batch transforms of n complex each, transform b at element b*stride (stride >= n;
 *              the `start` argument of the reference closures is pointer arithmetic on src/dst).
 *              src == dst is the in-place call; otherwise the ranges must not overlap.
 *   dir      : NAPA_FWD  = DIF  (natural in; out natural if in_order else bit-reversed)
 *              NAPA_INV  = DIT  (in natural if in_order else bit-reversed; natural out)
 *   window   : NULL or n doubles / n complex (window_type); multiplied into the input after the
 *              scale factor (gen-support.lisp:105-115).  For NAPA_INV the window is indexed by the
 *              memory position of the bit-reversed input, i.e. by rev(k) when in_order is set
 *              (inverse.lisp:9-18, easy-interface.lisp:83-95).
 *   window_per_batch : 0 = one window shared by all transforms, 1 = window b at b*n.          */
int napafft_c2c(const double *src, double *dst, size_t n, size_t batch, size_t stride, int dir, int scale,
                int in_order, const void *window, int window_type, int window_per_batch);

/* napafft_c2c for an input that the reference would first pass through complex-samplify
 * (support.lisp:64-93; fft / ifft call it on every argument, easy-interface.lisp:50, 75): src holds
 * batch x n samples of src_type, dense; the raw samples are what crosses PCIe and the widening to
 * complex double happens on the device.  dst: batch x n complex, must not alias src. */
int napafft_c2c_typed(const void *src, int src_type, double *dst, size_t n, size_t batch, int dir, int scale,
                      int in_order, const void *window, int window_type, int window_per_batch);

/* Replaces %ensure-reverse / get-reverse / bit-reverse (interface.lisp:103-133,
 * easy-interface.lisp:10-37): dst[b*stride + i] = src[b*stride + rev(i)], i < n. */
int napafft_bit_reverse(const void *src, void *dst, size_t n, size_t batch, size_t stride, int elt);

/* Replaces rfft (real.lisp:101-135): src = batch x n doubles (item b at b*n), dst = batch x n
 * complex (full spectrum, natural order).  The final radix-2 twiddles are the reference's
 * recurrence table (real.lisp:53-64), rebuilt bit-for-bit on the host by the library. n >= 2. */
int napafft_rfft(const double *src, double *dst, size_t n, size_t batch, int scale);
/* Replaces rifft (real.lisp:153-185): src = batch x n complex, dst = batch x n doubles.  Unlike the
 * reference the host copy of src is left intact (the reference documents it as destroyed). */
int napafft_rifft(const double *src, double *dst, size_t n, size_t batch, int scale);
/* Replaces %2rfft (real.lisp:33-48): dst = batch x 2n complex, [0,n) = FFT(v1), [n,2n) = FFT(v2) */
int napafft_2rfft(const double *v1, const double *v2, double *dst, size_t n, size_t batch, int scale);

/* The overlap-add streaming filter of example.lisp:81-114 / README:708-741 (filter-chunks) as one
 * device pipeline: half-overlapping chunks of `chunk` samples (zero-padded past
 * len), each windowed-fft'ed (:in-order nil) with `window` (chunk doubles; the reference uses
 * (window-vector 'triangle chunk)), multiplied by `filter_rev` (chunk doubles: the filter in
 * bit-reversed order, (bit-reverse filter)) inside ifft :in-order nil, renormalised to
 * 0.5 * sqrt(energy-in / energy-out) and accumulated.  vector: len doubles; dst: len complex. */
int napafft_filter_chunks(const double *vector, size_t len, const double *filter_rev, const double *window, size_t chunk,
                          double *dst);

/* Page-lock / unlock a caller-owned host range so that the entry points above DMA straight from
 * and to it (SBCL: a static vector or a pinned foreign block).  Optional. */
int napafft_host_register(void *ptr, size_t bytes);
int napafft_host_unregister(void *ptr);

/* ---- device-resident entry points (new exports for throughput paths; same semantics) ------- */
int napafft_c2c_dev(const void *src, void *dst, size_t n, size_t batch, size_t stride, int dir, int scale,
                    int in_order, const void *window, int window_type, int window_per_batch, void *stream);
int napafft_bit_reverse_dev(const void *src, void *dst, size_t n, size_t batch, size_t stride, int elt,
                            void *stream);
int napafft_rfft_dev(const void *src, void *dst, size_t n, size_t batch, int scale, void *stream);
int napafft_rifft_dev(const void *src, void *dst, size_t n, size_t batch, int scale, void *stream);
int napafft_2rfft_dev(const void *v1, const void *v2, void *dst, size_t n, size_t batch, int scale, void *stream);
int napafft_filter_chunks_dev(const void *vector, size_t len, const void *filter_rev, const void *window, size_t chunk,
                              void *dst, void *stream);
/* complex-samplify alone: dst[i] = (complex double) src[i], i < count */
int napafft_widen_dev(const void *src, int src_type, void *dst, size_t count, void *stream);
/* a[b*n + i] *= h[b*h_stride + i]  (pointwise spectrum product; h_stride 0 = shared filter) */
int napafft_cmul_dev(void *a, const void *h, size_t n, size_t batch, size_t h_stride, void *stream);

/* Building blocks of the distributed four-step (single transforms >= 2^28 over several GPUs,
 * SURVEY.md 8e): a transform along axis 0 of a row-major [n1][cols] device matrix, with the global
 * twiddle W_M^{+-(col_offset + c) k1} (M = 2^lg_twiddle_modulus, 0 = none) fused into its last /
 * first pass, and the [groups][rows][cols] <-> [rows][groups][cols] re-layout around the exchange. */
int napafft_c2c_cols_dev(const void *src, void *dst, size_t n1, size_t cols, int dir, double scale,
                         int lg_twiddle_modulus, long long col_offset, void *stream);
int napafft_relayout_dev(const void *src, void *dst, size_t groups, size_t rows, size_t cols, int inverse, void *stream);

/* ---- one transform over several GPUs (BASELINE config 5; replaces nothing in the reference, which is single-process:
 * it is what a caller of (fft vec) gets for a vector too large for one device; boundary = interface.lisp:165-189 get-fft) ----
 * One process per GPU.  N = 2^lgN = N1 * N2 with N1 = 2^(lgN/2); x[j1*N2 + j2] is an [N1][N2] matrix.
 *   time layout, rank g of P:  columns j2 in [g*B, (g+1)*B), B = N2/P, as a dense local [N1][B] array;
 *   freq layout, rank h:       rows k1 in [h*R, (h+1)*R), R = N1/P, of X[k1 + N1*k2], as a dense local [R][N2] array.
 * napafft_dist_c2c(_dev): dir = 1 maps the time layout to the freq layout (unscaled DFT), dir = -1 maps back and scales
 * by 1/N.  Out of place.  The library owns the NCCL communicator (bound at run time from libnccl.so.2; failures are
 * NAPA_E_NCCL): rank 0 calls napafft_dist_unique_id and hands the 128 bytes to the other ranks by any means, then every
 * rank calls napafft_dist_init(id, rank, nranks) after napafft_init(device).  nranks = 1 needs no NCCL. */
int napafft_dist_unique_id(void *out128);
int napafft_dist_init(const void *unique_id128, int rank, int nranks);
int napafft_dist_shutdown(void);
int napafft_dist_c2c_dev(const void *src_local, void *dst_local, int lgN, int dir, void *stream);
int napafft_dist_c2c(const double *src_local, double *dst_local, int lgN, int dir);
/* how the exchange travels: 0 not initialised, 1 one rank (device copy), 2 ncclSend / ncclRecv, 3 copy-engine
 * transfers over NVLink into the peers' receive buffers, mapped through CUDA IPC and fenced by tiny all-reduces
 * (the default when IPC mapping succeeds on every rank; NAPAFFT_DIST_PEER_COPY=0 forces 2).  Not a status code. */
int napafft_dist_exchange_mode(void);
/* geometry of the layouts above for 2^lgN points over nranks ranks (any output pointer may be NULL) */
int napafft_dist_geometry(int lgN, int nranks, size_t *n1, size_t *n2, size_t *cols_per_rank, size_t *rows_per_rank,
                          int *nchunks);
/* The two local stages, for hosts that carry the exchange with their own communicator: the exchange buffer xbuf is
 * [nchunks][N1][Bc] (Bc = B / nchunks): chunk c holds the columns [c*Bc, (c+1)*Bc) of every row; its P slabs of R rows
 * are the messages of an all-to-all (slab h goes to rank h), delivered into the same place of the receiver's buffer.
 *   cols: dir = 1: x_local [N1][B] -> chunk `chunk` of xbuf (column FFTs + global twiddle);  dir = -1: the inverse.
 *   rows: dir = 1: received xbuf -> X_local [R][N2] (row FFTs read the exchange layout directly);  dir = -1: the inverse. */
int napafft_dist_cols_dev(void *x_local, void *xbuf, int lgN, int nranks, int rank, int chunk, int dir, void *stream);
int napafft_dist_rows_dev(void *xbuf, void *X_local, int lgN, int nranks, int dir, void *stream);

/* plain device-memory helpers so that a non-CUDA host (SBCL) can hold device-resident vectors */
int napafft_dev_alloc(void **out, size_t bytes);
int napafft_dev_free(void *p);
int napafft_dev_upload(void *dst_dev, const void *src_host, size_t bytes);
int napafft_dev_download(void *dst_host, const void *src_dev, size_t bytes);
int napafft_dev_sync(void);

/* ---- introspection used by the tests ------------------------------------------------------- */
/* the numeric scale factor of interface.lisp:41-48 */
double napafft_scale_factor(size_t n, int dir, int scale);
/* the reference's radix-2 recurrence table (real.lisp:53-64): out = n/2 complex */
int napafft_radix2_twiddle(size_t n, int dir, double *out);

#ifdef __cplusplus
}
#endif
#endif /* NAPAFFT_H */
