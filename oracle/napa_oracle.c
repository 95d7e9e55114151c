/*
 * napa_oracle.c -- CPU ORACLE for the napa-fft3-b200 parity tests.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (napa_fft3_b200/ + libnapafft_b200.so) never links, imports or calls anything in here.
 *
 * What it is: an op-order-faithful restatement, in plain C, of the arithmetic that the
 * reference (pkhuong/Napa-FFT3, Common Lisp, SBCL-only) generates at run time.  SBCL is not
 * available in this image, so the reference itself cannot be executed; instead every rounding
 * step of the reference's generated code is reproduced here (separately rounded complex
 * products, no FMA contraction: build with -ffp-contract=off), and the result is pinned
 * bit-for-bit against every REPL transcript in the reference README (tests/test_oracle_kat.py).
 * Parity status: PINNED against README known answers (README.md:106-128, 174-193, 217-223,
 * 276-278, 307-308, 761-776, 794-804, 859-865).
 *
 * Each function cites the reference file:line it restates.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NAPA_ORACLE_API __attribute__((visibility("default")))

typedef struct { double re, im; } cplx;

/* ---- complex helpers: same rounding sequence as SBCL's (complex double-float) ops ---- */
static inline cplx c_add(cplx a, cplx b) { cplx r = { a.re + b.re, a.im + b.im }; return r; }
static inline cplx c_sub(cplx a, cplx b) { cplx r = { a.re - b.re, a.im - b.im }; return r; }
/* (* a b): re = ar*br - ai*bi ; im = ar*bi + ai*br, four products rounded separately */
static inline cplx c_mul(cplx a, cplx b) {
    double p0 = a.re * b.re, p1 = a.im * b.im, p2 = a.re * b.im, p3 = a.im * b.re;
    cplx r = { p0 - p1, p2 + p3 };
    return r;
}
static inline cplx c_mul_real(cplx a, double s) { cplx r = { a.re * s, a.im * s }; return r; }
static inline cplx c_conj(cplx a) { cplx r = { a.re, -a.im }; return r; }
static inline cplx c_swap(cplx a) { cplx r = { a.im, a.re }; return r; }           /* real.lisp:3-8 */
static inline cplx c_neg(cplx a) { cplx r = { -a.re, -a.im }; return r; }
/* gen-support.lisp:12-24 (complex-float-vops forms): +i*x = swap(conj x), -i*x = conj(swap x) */
static inline cplx mul_pi(cplx a) { cplx r = { -a.im, a.re }; return r; }
static inline cplx mul_mi(cplx a) { cplx r = { a.im, -a.re }; return r; }
/* gen-support.lisp:26-42 */
static inline cplx mul_sqrt_pi(cplx x, double s) { /* mul+/-sqrt+i */
    cplx w = c_mul_real(x, s); cplx r = { w.re - w.im, w.im + w.re }; return r;
}
static inline cplx mul_sqrt_mi(cplx x, double s) { /* mul+/-sqrt-i */
    cplx w = c_mul_real(x, s); cplx r = { w.re + w.im, w.im - w.re }; return r;
}

/* ---- support.lisp:30-36 ---- */
NAPA_ORACLE_API uint64_t napa_oracle_bit_reverse_integer(uint64_t x, int width) {
    uint64_t acc = 0;
    for (int i = 0; i < width; i++) { acc = (acc << 1) | (x & 1); x >>= 1; }
    return acc;
}
static int lb(size_t n) { int l = 0; while (((size_t)1 << l) < n) l++; return l; }   /* support.lisp:58-59 */
static int pow2p(size_t n) { return n != 0 && (n & (n - 1)) == 0; }                  /* support.lisp:61-62 */

/* ---- support.lisp:40-55: one table of n entries serves every size <= n ---- */
NAPA_ORACLE_API int napa_oracle_make_twiddle(size_t n, double dir, double *out /* 2n doubles */) {
    if (!pow2p(n)) return 1;
    cplx *tw = (cplx *)out;
    for (size_t i = 0; i < n; i++) { tw[i].re = 0.0; tw[i].im = 0.0; }
    const double pi = 3.141592653589793;
    for (size_t size = 4; size <= n; size *= 2) {
        size_t start = size / 2 - 1;                       /* +twiddle-offset+ = -1 */
        double base = ((dir * 2.0) * pi) / (double)size;   /* (/ (* dir 2 pi) size) */
        for (size_t i = 0; i < size / 4; i++) {
            double theta = (-1.0 * (double)i) * base;      /* (* -1 i base) */
            double theta3 = 3.0 * theta;                   /* (* 3 theta) */
            tw[start + 2 * i].re = cos(theta);      tw[start + 2 * i].im = sin(theta);
            tw[start + 2 * i + 1].re = cos(theta3); tw[start + 2 * i + 1].im = sin(theta3);
        }
    }
    return 0;
}

/* twiddle cache: interface.lisp:139-163 (forward table + element-wise conjugate) */
static cplx *g_tw_fwd = NULL, *g_tw_inv = NULL;
static size_t g_tw_n = 0;
static const cplx *ensure_twiddles(size_t n, int forwardp) {
    if (g_tw_n < n) {
        free(g_tw_fwd); free(g_tw_inv);
        size_t m = n < 4 ? 4 : n;
        g_tw_fwd = (cplx *)malloc(m * sizeof(cplx));
        g_tw_inv = (cplx *)malloc(m * sizeof(cplx));
        napa_oracle_make_twiddle(m, 1.0, (double *)g_tw_fwd);
        for (size_t i = 0; i < m; i++) g_tw_inv[i] = c_conj(g_tw_fwd[i]);
        g_tw_n = m;
    }
    return forwardp ? g_tw_fwd : g_tw_inv;
}

/* ---- transform context: scale factor + optional window (gen-support.lisp:44-60, 105-115) ---- */
typedef struct {
    double scale;            /* already the numeric factor, interface.lisp:41-48 */
    const double *window;    /* NULL, real (8 B/elt) or complex (16 B/elt) */
    int wtype;               /* 0 none, 1 real-sample, 2 complex-sample */
    const cplx *tw;
} ctx_t;

static inline int need_pre(const ctx_t *c) { return c->scale != 1.0 || c->wtype != 0; }

/* (%window (%scale x scale) window i): scale first, then window */
static inline cplx pre(const ctx_t *c, cplx x, size_t widx) {
    if (c->scale == 1.0) { /* no-op */ }
    else if (c->scale == -1.0) x = c_neg(x);
    else x = c_mul_real(x, c->scale);
    if (c->wtype == 1) x = c_mul_real(x, c->window[widx]);
    else if (c->wtype == 2) { cplx w = { c->window[2 * widx], c->window[2 * widx + 1] }; x = c_mul(x, w); }
    return x;
}

/* gen-support.lisp:62-86 mul-root: root = num/den (mod 1), special forms for multiples of 1/8 */
static inline cplx mul_root(cplx x, long num, long den, cplx deflt) {
    long r = num % den; if (r < 0) r += den;
    const double S = sqrt(2.0) / 2.0;                      /* (/ (sqrt 2d0) 2d0) */
    if ((8 * r) % den == 0) {
        switch ((8 * r) / den) {
        case 0: return x;
        case 4: return c_neg(x);
        case 2: return mul_pi(x);
        case 6: return mul_mi(x);
        case 1: return mul_sqrt_pi(x, S);
        case 5: return mul_sqrt_pi(x, -S);
        case 3: return mul_sqrt_mi(x, -S);
        case 7: return mul_sqrt_mi(x, S);
        }
    }
    return c_mul(x, deflt);
}

/* ---- forward.lisp:5-45 %gen-flat-dif: fully unrolled leaves, n <= 32 ---- */
static void flat_dif(const ctx_t *c, cplx *v, size_t start, size_t n, size_t last, size_t wstart) {
    if (n == 1) {
        if (n == last && need_pre(c)) v[start] = pre(c, v[start], wstart + start);
        return;
    }
    if (n == 2) {
        if (n == last && need_pre(c)) {
            v[start] = pre(c, v[start], wstart + start);
            v[start + 1] = pre(c, v[start + 1], wstart + start + 1);
        }
        cplx x = v[start], y = v[start + 1];
        v[start] = c_add(x, y); v[start + 1] = c_sub(x, y);
        return;
    }
    size_t h = n / 2, q = n / 4, s2 = start + h, s3 = s2 + q;
    for (size_t i = 0; i < h; i++) {
        cplx x = v[start + i], y = v[s2 + i];
        if (last == n && need_pre(c)) { x = pre(c, x, wstart + start + i); y = pre(c, y, wstart + s2 + i); }
        v[start + i] = c_add(x, y); v[s2 + i] = c_sub(x, y);
    }
    flat_dif(c, v, start, h, last, wstart);
    for (size_t cnt = 0; cnt < q; cnt++) {
        size_t i = s2 + cnt, j = s3 + cnt, k = h - 1 + 2 * cnt;
        cplx y = mul_mi(v[j]);                               /* rotate j nil 3/4 */
        cplx x = v[i];
        cplx a = c_add(x, y), b = c_sub(x, y);               /* butterfly i j */
        v[i] = mul_root(a, -(long)cnt, (long)n, c->tw[k]);          /* rotate i k (-count/n) */
        v[j] = mul_root(b, -3 * (long)cnt, (long)n, c->tw[k + 1]);  /* rotate j k+1 (-3count/n) */
    }
    flat_dif(c, v, s2, q, last, wstart);
    flat_dif(c, v, s3, q, last, wstart);
}

/* ---- forward.lisp:54-117 gen-dif: streaming levels above 32, flat leaves below ---- */
static void dif(const ctx_t *c, cplx *v, size_t n, int top, size_t wstart) {
    if (n <= 32) {                                          /* *fwd-base-case* forward.lisp:3 */
        ctx_t plain = *c;
        if (!top) { plain.scale = 1.0; plain.window = NULL; plain.wtype = 0; }
        flat_dif(&plain, v, 0, n, n, wstart);
        return;
    }
    size_t h = n / 2, q = n / 4;
    for (size_t i = 0; i < h; i++) {
        cplx x = v[i], y = v[i + h];
        if (top && need_pre(c)) { x = pre(c, x, wstart + i); y = pre(c, y, wstart + i + h); }
        v[i] = c_add(x, y); v[i + h] = c_sub(x, y);
    }
    dif(c, v, h, 0, 0);
    for (size_t cnt = 0; cnt < q; cnt++) {
        size_t k = h - 1 + 2 * cnt;
        cplx x = v[h + cnt], y = mul_mi(v[h + q + cnt]);
        v[h + cnt] = c_mul(c->tw[k], c_add(x, y));
        v[h + q + cnt] = c_mul(c->tw[k + 1], c_sub(x, y));
    }
    dif(c, v + h, q, 0, 0);
    dif(c, v + h + q, q, 0, 0);
}

/* ---- inverse.lisp:5-38 %gen-flat-dit: scale/window at EVERY n=1 / n=2 leaf ---- */
static void flat_dit(const ctx_t *c, cplx *v, size_t start, size_t n, size_t wstart) {
    if (n == 1) {
        if (need_pre(c)) v[start] = pre(c, v[start], wstart + start);
        return;
    }
    if (n == 2) {
        if (need_pre(c)) {
            v[start] = pre(c, v[start], wstart + start);
            v[start + 1] = pre(c, v[start + 1], wstart + start + 1);
        }
        cplx x = v[start], y = v[start + 1];
        v[start] = c_add(x, y); v[start + 1] = c_sub(x, y);
        return;
    }
    size_t h = n / 2, q = n / 4, s2 = start + h, s3 = s2 + q;
    flat_dit(c, v, s3, q, wstart);
    flat_dit(c, v, s2, q, wstart);
    for (size_t cnt = 0; cnt < q; cnt++) {
        size_t i = s2 + cnt, j = s3 + cnt, k = h - 1 + 2 * cnt;
        cplx x = mul_root(v[i], (long)cnt, (long)n, c->tw[k]);           /* rotate i k (count/n) */
        cplx y = mul_root(v[j], 3 * (long)cnt, (long)n, c->tw[k + 1]);   /* rotate j k+1 (3count/n) */
        v[i] = c_add(x, y);
        v[j] = mul_pi(c_sub(x, y));                                      /* rotate j nil -3/4 */
    }
    flat_dit(c, v, start, h, wstart);
    for (size_t i = 0; i < h; i++) {
        cplx x = v[start + i], y = v[s2 + i];
        v[start + i] = c_add(x, y); v[s2 + i] = c_sub(x, y);
    }
}

/* ---- inverse.lisp:47-112 gen-dit ---- */
static void dit(const ctx_t *c, cplx *v, size_t n, size_t wstart) {
    if (n <= 32) { flat_dit(c, v, 0, n, wstart); return; }   /* *inv-base-case* inverse.lisp:3 */
    size_t h = n / 2, q = n / 4;
    dit(c, v + h, q, wstart + h);
    dit(c, v + h + q, q, wstart + h + q);
    for (size_t cnt = 0; cnt < q; cnt++) {
        size_t k = h - 1 + 2 * cnt;
        cplx x = c_mul(c->tw[k], v[h + cnt]);
        cplx y = c_mul(c->tw[k + 1], v[h + q + cnt]);
        v[h + cnt] = c_add(x, y);
        v[h + q + cnt] = mul_pi(c_sub(x, y));
    }
    dit(c, v, h, wstart);
    for (size_t i = 0; i < h; i++) {
        cplx x = v[i], y = v[i + h];
        v[i] = c_add(x, y); v[i + h] = c_sub(x, y);
    }
}

/* ---- interface.lisp:41-48 scale factors.  mode: 0 nil/1, 1 t/:inv, 2 sqrt/:sqrt ---- */
NAPA_ORACLE_API double napa_oracle_scale_factor(size_t n, int forward, int mode) {
    double fn = (double)n;
    switch (mode) {
    case 0: return 1.0;
    case 1: return 1.0 / fn;
    case 2: return forward ? 1.0 / sqrt(fn) : (1.0 / fn) / (1.0 / sqrt(fn));
    }
    return NAN;
}

/* ---- low-level closures: interface.lisp:55-74 (vec start [window window-start] twiddle) ----
 * In place on v[start .. start+n).  dir: +1 DIF forward (natural in, bit-reversed out),
 * -1 DIT inverse (bit-reversed in, natural out).  scale is the numeric factor.          */
NAPA_ORACLE_API int napa_oracle_lowlevel(double *vec, size_t start, size_t n, int dir, double scale,
                                         const double *window, int wtype, size_t window_start) {
    if (!pow2p(n)) return 1;
    ctx_t c = { scale, wtype ? window : NULL, wtype, ensure_twiddles(n, dir > 0) };
    cplx *v = (cplx *)vec + start;
    if (dir > 0) dif(&c, v, n, 1, window_start);
    else dit(&c, v, n, window_start);
    return 0;
}

/* ---- bit reversal: bit-reversal.lisp:146-252 is an in-place swap network whose net effect is
 * vec[start+i] <-> vec[start+rev(i)] (support.lisp:30-36 defines rev); only the permutation is
 * observable, so it is restated as the plain involution.  elt: 0 complex (16 B), 1 real (8 B) */
NAPA_ORACLE_API int napa_oracle_bit_reverse_inplace(void *vec, size_t start, size_t n, int elt) {
    if (!pow2p(n)) return 1;
    int w = lb(n);
    if (elt == 0) {
        cplx *v = (cplx *)vec + start;
        for (size_t i = 0; i < n; i++) {
            size_t r = (size_t)napa_oracle_bit_reverse_integer(i, w);
            if (i < r) { cplx t = v[i]; v[i] = v[r]; v[r] = t; }
        }
    } else {
        double *v = (double *)vec + start;
        for (size_t i = 0; i < n; i++) {
            size_t r = (size_t)napa_oracle_bit_reverse_integer(i, w);
            if (i < r) { double t = v[i]; v[i] = v[r]; v[r] = t; }
        }
    }
    return 0;
}

/* ---- easy-interface.lisp:44-67 fft / :69-95 ifft on an already-resolved dst (the copy/alias
 * rule of easy-interface.lisp:3-8 is applied by the caller: dst holds the samples).         */
NAPA_ORACLE_API int napa_oracle_fft(double *dst, size_t n, int in_order, int scale_mode,
                                    const double *window, int wtype) {
    if (!pow2p(n)) return 1;
    double s = napa_oracle_scale_factor(n, 1, scale_mode);
    napa_oracle_lowlevel(dst, 0, n, +1, s, window, wtype, 0);
    if (in_order) napa_oracle_bit_reverse_inplace(dst, 0, n, 0);
    return 0;
}
NAPA_ORACLE_API int napa_oracle_ifft(double *dst, size_t n, int in_order, int scale_mode,
                                     const double *window, int wtype) {
    if (!pow2p(n)) return 1;
    double s = napa_oracle_scale_factor(n, 0, scale_mode);
    if (in_order) napa_oracle_bit_reverse_inplace(dst, 0, n, 0);
    napa_oracle_lowlevel(dst, 0, n, -1, s, window, wtype, 0);
    return 0;
}

/* ---- real.lisp:53-64 %get-radix-2-twiddle: built by RECURRENCE (error grows with k) ---- */
NAPA_ORACLE_API int napa_oracle_radix2_twiddle(size_t n, double direction, double *out /* n doubles */) {
    const double pi = 3.141592653589793;
    cplx *tw = (cplx *)out;
    cplx root = direction > 0 ? (cplx){ 1.0, 0.0 } : (cplx){ 0.0, 1.0 };
    /* (exp (* -2 pi (complex 0d0 direction) (/ 1d0 n))) = cis((-2 pi direction) * (1/n)) */
    double a = ((-2.0 * pi) * direction) * (1.0 / (double)n);
    cplx mul = { cos(a), sin(a) };
    for (size_t k = 0; k < n / 2; k++) { tw[k] = root; root = c_mul(root, mul); }
    return 0;
}

/* with-scale, real.lisp:83-99 */
static double real_scale(int mode) { return mode == 0 ? 1.0 : mode == 1 ? 0.5 : 1.0 / sqrt(2.0); }

/* ---- real.lisp:10-31 fft-swizzled-reals: len = (length vec), NOT the user's size ---- */
NAPA_ORACLE_API int napa_oracle_swizzle(double *vec, size_t len, int scale_mode) {
    if (!pow2p(len) || len < 2) return 1;
    cplx *d = (cplx *)vec;
    size_t h = len / 2;
    napa_oracle_fft(vec, h, 1, scale_mode, NULL, 0);
    cplx z = d[0];
    d[h] = (cplx){ z.im, 0.0 };
    d[0] = (cplx){ z.re, 0.0 };
    for (size_t i = 1; i <= h / 2; i++) {
        cplx x = d[i], xs = c_swap(x);
        cplx y = d[h - i], ys = c_swap(y);
        cplx xj = c_mul_real(c_add(c_conj(xs), ys), 0.5);
        cplx mxj = c_mul_real(c_add(c_conj(ys), xs), 0.5);
        d[h + i] = xj;
        d[len - i] = mxj;
        d[i] = c_sub(x, mul_pi(xj));
        d[h - i] = c_sub(y, mul_pi(mxj));
    }
    return 0;
}

/* ---- real.lisp:101-135 rfft: src n doubles -> dst n complex (full spectrum, natural order) ---- */
NAPA_ORACLE_API int napa_oracle_rfft(const double *src, double *dst, size_t n, int scale_mode) {
    if (!pow2p(n)) return 1;
    cplx *d = (cplx *)dst;
    if (n == 1) { /* n/2 = 0: loops are empty, swizzle on a length-1 array is degenerate */
        d[0] = (cplx){ src[0], 0.0 }; return 0;
    }
    size_t h = n / 2;
    double s = real_scale(scale_mode);
    for (size_t j = 0; j < h; j++) {
        cplx x = { src[2 * j], src[2 * j + 1] };
        d[j] = scale_mode == 0 ? x : c_mul_real(x, s);
    }
    napa_oracle_swizzle(dst, n, scale_mode);
    cplx *tw = (cplx *)malloc(h * sizeof(cplx));
    napa_oracle_radix2_twiddle(n, 1.0, (double *)tw);
    for (size_t i = 0; i < h; i++) {
        cplx x = d[i], y = c_mul(d[i + h], tw[i]);
        d[i] = c_add(x, y); d[i + h] = c_sub(x, y);
    }
    free(tw);
    return 0;
}

/* ---- real.lisp:153-185 rifft: vec (n complex) is DESTROYED, dst gets n doubles ---- */
NAPA_ORACLE_API int napa_oracle_rifft(double *vec, double *dst, size_t n, int scale_mode) {
    if (!pow2p(n) || n < 2) return 1;
    cplx *v = (cplx *)vec;
    size_t h = n / 2;
    cplx *tw = (cplx *)malloc(h * sizeof(cplx));
    napa_oracle_radix2_twiddle(n, -1.0, (double *)tw);
    for (size_t i = 0; i < h; i++) {
        cplx x = v[i], y = v[i + h];
        cplx sum = c_add(x, y), sub = c_mul(c_sub(x, y), tw[i]);
        v[i] = c_add(sum, sub);
    }
    free(tw);
    napa_oracle_ifft(vec, h, 1, scale_mode, NULL, 0);
    double s = real_scale(scale_mode);
    for (size_t i = 0; i < h; i++) {
        cplx x = scale_mode == 0 ? v[i] : c_mul_real(v[i], s);
        dst[2 * i] = x.re; dst[2 * i + 1] = x.im;
    }
    return 0;
}

/* ---- real.lisp:33-48 %2rfft: dst has 2n complex; dst[0..n) = FFT(v1), dst[n..2n) = FFT(v2) ---- */
NAPA_ORACLE_API int napa_oracle_2rfft(const double *v1, const double *v2, double *dst, size_t n, int scale_mode) {
    if (!pow2p(n)) return 1;
    cplx *d = (cplx *)dst;
    for (size_t j = 0; j < n; j++) d[j] = (cplx){ v1[j], v2[j] };
    for (size_t j = n; j < 2 * n; j++) d[j] = (cplx){ 0.0, 0.0 };
    return napa_oracle_swizzle(dst, 2 * n, scale_mode);
}

/* ---- windowing.lisp:6-51 window functions, with the reference's single-float quirks.
 * id: 0 rectangular, 1 hann, 2 blackman, 3 triangle, 4 bartlett, 5 blackman-harris,
 *     6 blackman*(p0 = alpha as single), 7 gauss*(p0 = sigma as single)               */
NAPA_ORACLE_API double napa_oracle_window_fn(int id, long i, long n, double p0) {
    const double pi = 3.141592653589793;
    switch (id) {
    case 0: return (double)1.0f;                                             /* :6-8 */
    case 1: return 0.5 * (1.0 - cos(((2.0 * pi) * (double)i) / (double)(n - 1)));   /* :10 */
    case 2: case 6: {                                                        /* :12-20 */
        float alpha = id == 2 ? 0.16f : (float)p0;
        float a0 = (1.0f - alpha) / 2.0f, a1 = 0.5f, a2 = alpha / 2.0f;
        double c1 = cos(((2.0 * pi) * (double)i) / (double)(n - 1));
        double c2 = cos(((4.0 * pi) * (double)i) / (double)(n - 1));
        return ((double)a0 + (-((double)a1 * c1))) + ((double)a2 * c2);
    }
    case 3: {                                                                /* :22-23, single */
        float t = (float)n * 0.5f - fabsf((float)i - 0.5f * (float)(n - 1));
        return (double)((float)(2.0 / (double)n) * t);
    }
    case 4: {                                                                /* :25-26, single */
        float m = (float)(n - 1);
        float t = m * 0.5f - fabsf((float)i - 0.5f * m);
        return (double)((float)(2.0 / (double)(n - 1)) * t);
    }
    case 5: {                                                                /* :46-51 */
        double a0 = (double)0.35875f, a1 = (double)0.48829f, a2 = (double)0.14128f, a3 = (double)0.01168f;
        double f1 = a1 * cos(((2.0 * pi) * (double)i) / (double)(n - 1));
        double f2 = a2 * cos(((4.0 * pi) * (double)i) / (double)(n - 1));
        double f3 = a3 * cos(((6.0 * pi) * (double)i) / (double)(n - 1));
        return ((a0 + (-f1)) + f2) + (-f3);
    }
    case 7: {                                                                /* :28-30, single */
        float sigma = (float)p0;
        float m2 = 0.5f * (float)(n - 1);
        float r = ((float)i - m2) / (sigma * m2);
        float e = -0.5f * (r * r);
        return (double)(float)exp((double)e);
    }
    }
    return NAN;
}

/* windowing.lisp:43-45 cosine-series with caller-supplied coefficients (all arithmetic in double: PI is a
 * double, so (* scale (cos ...)) is double whatever the coefficient's float type) */
NAPA_ORACLE_API double napa_oracle_cosine_series(long i, long n, double a0, double a1, double a2, double a3) {
    const double pi = 3.141592653589793;
    double f1 = a1 * cos(((2.0 * pi) * (double)i) / (double)(n - 1));
    double f2 = a2 * cos(((4.0 * pi) * (double)i) / (double)(n - 1));
    double f3 = a3 * cos(((6.0 * pi) * (double)i) / (double)(n - 1));
    return ((a0 + (-f1)) + f2) + (-f3);
}

/* windowing.lisp:36-41 gaussian*bartlett^x for a single-float sigma and a non-negative INTEGER exponent:
 * (expt single integer) is SBCL's INTEXP, repeated squaring in single precision; the product with gauss*
 * (single) is single; window-vector then widens it */
NAPA_ORACLE_API double napa_oracle_gaussian_bartlett_x(long i, long n, double sigma, long power) {
    float base = (float)napa_oracle_window_fn(4, i, n, 0.0);
    float total = (power & 1) ? base : 1.0f;
    for (long nextn = power >> 1; nextn != 0; nextn >>= 1) {
        base = base * base;
        if (nextn & 1) total = base * total;
    }
    float g = (float)napa_oracle_window_fn(7, i, n, sigma);
    return (double)(total * g);
}

/* windowing.lisp:53-65 window-vector (the cache is the caller's business) */
NAPA_ORACLE_API int napa_oracle_window_vector(int id, size_t n, double p0, int bit_reverse, double *out) {
    for (size_t i = 0; i < n; i++) out[i] = napa_oracle_window_fn(id, (long)i, (long)n, p0);
    if (bit_reverse) return napa_oracle_bit_reverse_inplace(out, 0, n, 1);
    return 0;
}

/* windowing.lisp:67-111 extract-centered-window on complex samples (zero padded) */
NAPA_ORACLE_API int napa_oracle_extract_centered_window(const double *signal, size_t siglen, long center,
                                                        size_t size, double *out, int elt) {
    long start = center - (long)(size / 2);               /* (floor size 2) */
    size_t w = elt == 0 ? 2 : 1;
    long s = start < 0 ? 0 : (start > (long)siglen ? (long)siglen : start);
    long e = start + (long)size;
    e = e < 0 ? 0 : (e > (long)siglen ? (long)siglen : e);
    memset(out, 0, size * w * sizeof(double));
    long off = s - start;
    if (off > -1 && off < (long)size && e > s)
        memcpy(out + (size_t)off * w, signal + (size_t)s * w, (size_t)(e - s) * w * sizeof(double));
    return 0;
}

/* ---- bounded batch drivers used as the timed CPU baseline (bench.py cpu_baseline) ---- */
NAPA_ORACLE_API int napa_oracle_c2c_batch(double *data, size_t n, size_t batch, int dir, int in_order,
                                          int scale_mode, const double *window, int wtype) {
    for (size_t b = 0; b < batch; b++) {
        double *d = data + 2 * n * b;
        int rc = dir > 0 ? napa_oracle_fft(d, n, in_order, scale_mode, window, wtype)
                         : napa_oracle_ifft(d, n, in_order, scale_mode, window, wtype);
        if (rc) return rc;
    }
    return 0;
}
/* make sure the shared twiddle cache is built before worker threads race on it */
NAPA_ORACLE_API void napa_oracle_warm(size_t n) { ensure_twiddles(n, 1); }
