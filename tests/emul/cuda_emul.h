// cuda_emul.h -- TEST-ONLY functional emulation of the small CUDA subset used by
// napa_fft3_b200/csrc, so that the *shipped* kernel source and planner can be executed on a CPU
// in the `-m "not gpu"` tests (this container has no GPU).  It is compiled into
// tests/emul/libnapafft_emul.so by tests/emul/build_emul.py and is never built, loaded or linked
// by the product package: the product library is nvcc-compiled sm_100a code with no CPU path.
//
// Threads of a CTA are ucontext fibers run round-robin on one OS thread; __syncthreads() yields to
// the scheduler until every live fiber of the CTA has arrived.  CTAs run one after another, except in
// launch_cooperative(), which keeps the fibers of the whole grid alive so that CTAs can wait on each other.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(x)
#define __shared__ static

struct alignas(16) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }

struct uint3_ { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace napa_emul {
extern uint3_ threadIdx_, blockIdx_;
extern dim3 blockDim_, gridDim_;
extern unsigned char *dyn_smem;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
// every CTA alive at once (round-robin over all fibers of the grid): kernels whose CTAs wait on each other
void launch_cooperative(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
void yield_spin();                           // called from a spin-wait on another CTA's progress
void yield_any();                            // spin-wait on another THREAD's progress (same CTA or not), any launch mode
void note_progress();                        // a waited-for event happened (mbarrier arrival, counter increment)
void syncthreads();
void bar_sync(int id, unsigned count);      // named barrier (bar.sync id, count)
void bar_arrive(int id, unsigned count);    // bar.arrive id, count
}  // namespace napa_emul

#define threadIdx (napa_emul::threadIdx_)
#define blockIdx (napa_emul::blockIdx_)
#define blockDim (napa_emul::blockDim_)
#define gridDim (napa_emul::gridDim_)
static inline void __syncthreads() { napa_emul::syncthreads(); }

template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned __brev(unsigned x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
}
static inline unsigned long long __brevll(unsigned long long x) {
    return ((unsigned long long)__brev((unsigned)x) << 32) | __brev((unsigned)(x >> 32));
}

#define NAPA_DYN_SMEM(T, name) T *name = reinterpret_cast<T *>(napa_emul::dyn_smem)
#define NAPA_LAUNCH(kern, grid, block, smem, stream, ...) \
    napa_emul::launch(dim3(grid), dim3(block), (smem), [=]() { kern(__VA_ARGS__); })
#define NAPA_LAUNCH_COOP(kern, grid, block, smem, stream, ...) \
    napa_emul::launch_cooperative(dim3(grid), dim3(block), (smem), [=]() { kern(__VA_ARGS__); })

// ---- minimal runtime -------------------------------------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorNoDevice = 100, cudaErrorInsufficientDriver = 35, cudaErrorEmul = 999 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaHostRegisterDefault = 0, cudaHostAllocMapped = 2 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaPointerAttributes { int type; };
struct cudaDeviceProp { int major, minor, multiProcessorCount; };

static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *c) { *c = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { p->major = 10; p->minor = 0; p->multiProcessorCount = 2; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) {
    void *q = nullptr;
    if (posix_memalign(&q, 256, bytes ? bytes : 16)) return cudaErrorEmul;
    *p = (T *)q; return cudaSuccess;
}
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned) { return cudaMalloc(p, bytes); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t i = 0; i < h; i++) memmove((char *)d + i * dp, (const char *)s + i * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = 0; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { a->type = cudaMemoryTypeUnregistered; return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
static inline cudaError_t cudaFuncSetAttribute(const void *, cudaFuncAttribute, int) { return cudaSuccess; }

// atomics / fences (CTAs and fibers run sequentially on one OS thread)
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline void __threadfence() {}
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, const void *, int, size_t) { *n = 2; return cudaSuccess; }
// the control warp of the pipeline kernel broadcasts a flag from lane 0; fibers run one at a
// time, so emulate with a per-warp slot written by lane 0 (lane 0 always runs first in a round)
namespace napa_emul { extern int shfl_slot[64]; }
static inline int __shfl_sync(unsigned, int v, int src_lane) {
    const unsigned tid = napa_emul::threadIdx_.x;
    const unsigned warp = tid / 32, lane = tid % 32;
    if ((int)lane == src_lane) napa_emul::shfl_slot[warp] = v;
    return napa_emul::shfl_slot[warp];
}
