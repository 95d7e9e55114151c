// napafft_tables.cu -- where the kernel templates of napafft_kernels.cuh are instantiated.  Compiled once per slice
// (-DNAPA_TABLE_PASS=0|1|2, -DNAPA_TABLE_FUSED, -DNAPA_TABLE_FUSED_TMA; napa_fft3_b200/build.py runs the five
// compilations in parallel) and linked with napafft.cu, which only sees the look-up functions declared below.
#define NAPA_TABLES_ONLY
#include "napafft_kernels.cuh"

namespace napa {
typedef void (*pass_fn)(const PassParams);
typedef void (*fused_fn)(const FusedParams);
typedef void (*fused_tma_fn)(const FusedTmaParams);

#ifdef NAPA_TABLE_PASS
template <int K> static pass_fn pass_kernel_k(int l) {
    constexpr int V = NAPA_TABLE_PASS;
    switch (l) {
    case 4: return fft_pass_kernel<4, K, V>;
    case 5: return fft_pass_kernel<5, K, V>;
    case 6: return fft_pass_kernel<6, K, V>;
    case 7: return fft_pass_kernel<7, K, V>;
    case 8: return fft_pass_kernel<8, K, V>;
    case 9: return fft_pass_kernel<9, K, V>;
    case 10: return fft_pass_kernel<10, K, V>;
    case 11: return fft_pass_kernel<11, K, V>;
    case 12: return fft_pass_kernel<12, K, V>;
    case 13: return fft_pass_kernel<13, K, V>;
    }
    return nullptr;
}
#define NAPA_CAT2(a, b) a##b
#define NAPA_CAT(a, b) NAPA_CAT2(a, b)
// kind: 0 column-like, 1 row-like, 2 row-like load + transposed store
pass_fn NAPA_CAT(pass_kernel_table_v, NAPA_TABLE_PASS)(int l, int kind) {
    return kind == 0 ? pass_kernel_k<0>(l) : (kind == 1 ? pass_kernel_k<1>(l) : pass_kernel_k<2>(l));
}
#endif

#ifdef NAPA_TABLE_FUSED
#define NAPA_FUSED_CASE(a, b)                                                                     \
    if (la == a && lb == b) {                                                                     \
        if (win) return pk == 0 ? fft_fused_kernel<a, b, 0, true> : (pk == 1 ? fft_fused_kernel<a, b, 1, true> : fft_fused_kernel<a, b, 2, true>); \
        return pk == 0 ? fft_fused_kernel<a, b, 0, false> : (pk == 1 ? fft_fused_kernel<a, b, 1, false> : fft_fused_kernel<a, b, 2, false>); \
    }
fused_fn fused_kernel_table(int la, int lb, int pk, bool win) {
    NAPA_FUSED_CASE(7, 7) NAPA_FUSED_CASE(7, 8) NAPA_FUSED_CASE(8, 7) NAPA_FUSED_CASE(8, 8)
    NAPA_FUSED_CASE(9, 9) NAPA_FUSED_CASE(9, 10) NAPA_FUSED_CASE(10, 9) NAPA_FUSED_CASE(10, 10)
    return nullptr;
}
#endif

#ifdef NAPA_TABLE_FUSED_TMA
#define NAPA_FUSED_TMA_CASE(a, b)                                                                 \
    if (la == a && lb == b) {                                                                     \
        if (win) return pk == 0 ? fft_fused_tma_kernel<a, b, 0, true> : (pk == 1 ? fft_fused_tma_kernel<a, b, 1, true> : fft_fused_tma_kernel<a, b, 2, true>); \
        return pk == 0 ? fft_fused_tma_kernel<a, b, 0, false> : (pk == 1 ? fft_fused_tma_kernel<a, b, 1, false> : fft_fused_tma_kernel<a, b, 2, false>); \
    }
// TMA-fed variant: 2048-point tiles only (a column tile is one tensor-map box of <= 256 rows)
fused_tma_fn fused_tma_kernel_table(int la, int lb, int pk, bool win) {
    NAPA_FUSED_TMA_CASE(7, 7) NAPA_FUSED_TMA_CASE(7, 8) NAPA_FUSED_TMA_CASE(8, 7) NAPA_FUSED_TMA_CASE(8, 8)
    return nullptr;
}
#endif
}  // namespace napa
