// Microbenchmark: HBM throughput of "column tile" traffic -- every CTA copies ROWS segments of
// SEG bytes at a row pitch of PITCH bytes (what pass A of the two-pass FFT reads), for several
// SEG.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o strided_copy strided_copy.cu
#include <cstdio>
#include <cuda_runtime.h>

// tile t: item = t / tiles_per_item, column block = t % tiles_per_item; ROWS x SEG bytes
template <int SEG16>   // segment length in 16-byte units
__global__ void __launch_bounds__(128, 4) copy_tiles(const double2 *src, double2 *dst, int rows, long long pitch16,
                                                     int tiles_per_item, long long item16, int contiguous_dst) {
    const long long t = blockIdx.x;
    const long long item = t / tiles_per_item, cb = t % tiles_per_item;
    const double2 *s = src + item * item16 + cb * SEG16;
    double2 *d = contiguous_dst ? dst + t * (long long)rows * SEG16 : dst + item * item16 + cb * SEG16;
    const int lane_in_seg = threadIdx.x % SEG16, row0 = threadIdx.x / SEG16;
    constexpr int RPI = 128 / SEG16;                 // rows covered per iteration
    double2 v[16];
    const int iters = rows / RPI;                    // <= 16 * k
    for (int base = 0; base < iters; base += 16) {
#pragma unroll
        for (int m = 0; m < 16; m++) if (base + m < iters) v[m] = __ldcs(s + (long long)((base + m) * RPI + row0) * pitch16 + lane_in_seg);
#pragma unroll
        for (int m = 0; m < 16; m++) {
            long long o = contiguous_dst ? (long long)((base + m) * RPI + row0) * SEG16 + lane_in_seg
                                         : (long long)((base + m) * RPI + row0) * pitch16 + lane_in_seg;
            if (base + m < iters) __stcs(d + o, v[m]);
        }
    }
}

template <int SEG16>
static void run(const double2 *src, double2 *dst, long long total16, int n_log2, int contiguous_dst) {
    // view memory as items of n = 2^n_log2 points, each a [L1][L2] matrix with L2 = pitch; tile = L1 rows x SEG
    const long long n = 1LL << n_log2;
    const int l1 = n_log2 / 2, l2 = n_log2 - l1;
    const int rows = 1 << l1;
    const long long pitch16 = 1LL << l2;
    const int tiles_per_item = (int)(pitch16 / SEG16);
    const long long items = total16 / n;
    const long long tiles = items * tiles_per_item;
    // tile must hold rows*SEG16 = multiple of 2048 elements per CTA pass: loop handles it
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) copy_tiles<SEG16><<<(unsigned)tiles, 128>>>(src, dst, rows, pitch16, tiles_per_item, n, contiguous_dst);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; i++) copy_tiles<SEG16><<<(unsigned)tiles, 128>>>(src, dst, rows, pitch16, tiles_per_item, n, contiguous_dst);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("n=2^%d rows=%d pitch=%lld B seg=%4d B dst=%s : %.3f ms  %.0f GB/s (%s)\n", n_log2, rows, pitch16 * 16, SEG16 * 16,
           contiguous_dst ? "contig" : "strided", ms, 2.0 * total16 * 16 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const long long total16 = (2048LL << 20) / 16;
    double2 *a, *b;
    cudaMalloc(&a, total16 * 16); cudaMalloc(&b, total16 * 16);
    cudaMemset(a, 1, total16 * 16); cudaMemset(b, 0, total16 * 16);
    for (int nl : {16, 20}) {
        for (int cd : {0, 1}) {
            run<4>(a, b, total16, nl, cd);
            run<8>(a, b, total16, nl, cd);
            run<16>(a, b, total16, nl, cd);
            run<32>(a, b, total16, nl, cd);
            run<64>(a, b, total16, nl, cd);
        }
    }
    return 0;
}
