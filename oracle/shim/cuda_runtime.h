/*
 * CPU shim for the CUDA runtime — TEST INFRASTRUCTURE ONLY (oracle/_ref build).
 * Lets the reference's .cu sources (gsw4869/CUDA_LDPC, bldpc_实习/) compile with g++ and run
 * serially on the host: kernels become plain functions, <<<g,b>>> launches (rewritten to
 * SHIM_LAUNCH by oracle/build_ref.sh) become a double loop over blocks and threads in
 * ascending order, device memory is host memory.  Sequential execution makes the
 * reference's duplicate-slot write race deterministic: ascending thread id,
 * last writer wins.
 */
#ifndef SHIM_CUDA_RUNTIME_H
#define SHIM_CUDA_RUNTIME_H
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define __global__
#define __device__
#define __host__
#define __shared__ static

struct shim_dim3 {
    unsigned x, y, z;
};
extern shim_dim3 threadIdx, blockIdx, blockDim, gridDim;

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind {
    cudaMemcpyHostToHost = 0,
    cudaMemcpyHostToDevice = 1,
    cudaMemcpyDeviceToHost = 2,
    cudaMemcpyDeviceToDevice = 3
};
struct cudaDeviceProp {
    char name[256];
    size_t totalGlobalMem;
    int maxThreadsPerBlock;
    int clockRate;
    int multiProcessorCount;
    int maxThreadsPerMultiProcessor;
};
typedef void *cudaEvent_t;

static inline cudaError_t cudaMalloc(void **p, size_t n)
{
    *p = malloc(n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void *p)
{
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind)
{
    memcpy(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void *d, int v, size_t n)
{
    memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceCount(int *n)
{
    *n = 1;
    return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof(*p));
    strcpy(p->name, "cpu-shim");
    p->maxThreadsPerBlock = 1024;
    return cudaSuccess;
}
static inline cudaError_t cudaThreadSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaThreadExit() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e)
{
    *e = 0;
    return cudaSuccess;
}
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, int) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t)
{
    *ms = 0;
    return cudaSuccess;
}
static inline void __syncthreads() {}

#define SHIM_LAUNCH(K, G, B, ...)                                                  \
    do {                                                                           \
        gridDim.x = (unsigned)(G);                                                 \
        blockDim.x = (unsigned)(B);                                                \
        for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)                 \
            for (threadIdx.x = 0; threadIdx.x < blockDim.x; ++threadIdx.x)         \
                K(__VA_ARGS__);                                                    \
    } while (0)

#endif
