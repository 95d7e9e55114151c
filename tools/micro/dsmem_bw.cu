// Distributed-shared-memory bandwidth between the CTAs of a thread-block cluster on sm_100a: the number that decides
// whether a 2^14 .. 2^17 transform can live in the shared memory of 8-16 SMs (open item 1).
// Every thread of every CTA stores 16-byte values into the shared memory of the other CTAs of its cluster (the
// all-to-all transposition of a four-step FFT inside a cluster), `iters` rounds, cluster barrier per round.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/dsmem_bw tools/micro/dsmem_bw.cu && tools/micro/dsmem_bw
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

template <bool REMOTE>
__global__ void __launch_bounds__(512, 1) xchg(int iters, int per_thread, double2 *sink) {
    extern __shared__ __align__(16) unsigned char raw[];
    double2 *sm = reinterpret_cast<double2 *>(raw);
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank(), csz = cluster.num_blocks();
    double2 v = make_double2(threadIdx.x, rank);
    for (int it = 0; it < iters; it++) {
        for (int j = 0; j < per_thread; j++) {
            const unsigned peer = REMOTE ? (rank + 1 + (j % (csz - 1))) % csz : rank;      // never myself when REMOTE
            double2 *dst = cluster.map_shared_rank(sm, peer);
            dst[j * 512 + threadIdx.x] = v;
            v.x += 1.0;
        }
        cluster.sync();
    }
    if (sm[threadIdx.x].x == -1.0) sink[0] = sm[threadIdx.x];
}

template <bool REMOTE>
static void run(int csz, int iters) {
    const int per_thread = 16, threads = 512;
    const size_t smem = (size_t)per_thread * threads * sizeof(double2);     // 128 KiB: the tile of a 2^16 transform on 8 CTAs
    cudaFuncSetAttribute(xchg<REMOTE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (csz > 8) cudaFuncSetAttribute(xchg<REMOTE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.attrs = attr; cfg.numAttrs = 1;
    int nclusters = 0;
    cfg.gridDim = dim3(csz);
    if (cudaOccupancyMaxActiveClusters(&nclusters, xchg<REMOTE>, &cfg) != cudaSuccess || nclusters < 1) {
        printf("cluster size %2d: not launchable (%s)\n", csz, cudaGetErrorString(cudaGetLastError()));
        return;
    }
    cfg.gridDim = dim3(csz * nclusters);
    double2 *sink; cudaMalloc(&sink, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaLaunchKernelEx(&cfg, xchg<REMOTE>, 10, per_thread, sink);
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, xchg<REMOTE>, iters, per_thread, sink);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double bytes_per_cta = (double)iters * per_thread * threads * 16.0;
    const double gbs_per_sm = bytes_per_cta / (ms * 1e-3) / 1e9;
    printf("%s cluster size %2d: %3d co-resident clusters (%3d SMs), %.3f ms, %.1f GB/s per SM, %.1f B/clk per SM at %d MHz, %.2f TB/s chip-wide%s\n",
           REMOTE ? "remote" : "local ", csz, nclusters, csz * nclusters, ms, gbs_per_sm, gbs_per_sm * 1e9 / (clk * 1e3), clk / 1000,
           gbs_per_sm * csz * nclusters / 1e3, err == cudaSuccess ? "" : " (FAILED)");
    cudaFree(sink);
}

int main() {
    for (int csz : { 2, 4, 8, 16 }) run<true>(csz, 400);
    run<false>(8, 400);
    return 0;
}
