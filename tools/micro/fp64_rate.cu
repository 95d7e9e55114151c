// FP64 pipe throughput of one SM on sm_100a (DFMA / DADD / DMUL warp-instructions per clock): the denominator of the
// "FP64-pipe time of a transform" figures in DESIGN.md section 9.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/fp64_rate tools/micro/fp64_rate.cu && tools/micro/fp64_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(1024, 1) spin(int iters, double *out, double seed) {
    double a[8];
    for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x + i;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == 0) a[i] = fma(a[i], m, c);
                else if (OP == 1) a[i] = a[i] + c;
                else a[i] = a[i] * m;
            }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 12345.678) out[0] = s;
}

template <int OP> static void run(const char *name) {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double *out; cudaMalloc(&out, 8);
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    spin<OP><<<sms, 1024>>>(100, out, 1.0);
    cudaEventRecord(e0);
    spin<OP><<<sms, 1024>>>(iters, out, 1.0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double thread_instr_per_sm = (double)iters * 16 * 8 * 1024;
    const double per_clk_nominal = thread_instr_per_sm / (ms * 1e-3) / (clk * 1e3);
    printf("%s: %.3f ms, %.1f thread-instructions per clock per SM at the nominal %d MHz (%.2f T instr/s chip-wide)\n", name, ms,
           per_clk_nominal, clk / 1000, thread_instr_per_sm * sms / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

int main() {
    run<0>("DFMA"); run<1>("DADD"); run<2>("DMUL");
    return 0;
}
