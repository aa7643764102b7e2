// Developer microbenchmark: issue interval (cycles per warp instruction per SM sub-partition) of the
// instructions the layered kernel is made of, alone and in ALU/FMA mixes.  nvcc -arch=sm_100a pipes.cu
#include <cstdio>
#include <cuda_fp16.h>
#define OPS_PER_ITER 64
#define ITERS 2000
enum { HFMA2_R, HFMA2_I, HADD2_R, HADD2_I, HFMA2_SAT, HSET2, VMIN, VMIN3, VADDMNMX, LOP3, PRMT, IMADSHL, HMNMX2,
       MIX_HFMA_VMIN, MIX_HFMA_LOP3, MIX_HADD_PRMT, MIX_HFMA_HADD, MIX_VMIN_LOP3, MIX3, FFMA_R, MIX_HFMA_IMAD, NOPS };
const char *names[] = {"HFMA2 rrr", "HFMA2 r,imm,r", "HADD2 rr", "HADD2 r,imm", "HFMA2.SAT", "HSET2.EQ", "VIMNMX.U16x2",
                       "VIMNMX3.U16x2", "VIADDMNMX.S16x2.RELU", "LOP3", "PRMT", "IMAD.SHL", "HMNMX2",
                       "mix HFMA2+VIMNMX", "mix HFMA2+LOP3", "mix HADD2+PRMT", "mix HFMA2+HADD2", "mix VIMNMX+LOP3",
                       "mix HFMA2+HADD2+VIMNMX+LOP3", "FFMA rrr", "mix HFMA2+IMAD.SHL"};
template <int OP>
__device__ __forceinline__ void op(unsigned &a, unsigned b, unsigned c, int j)
{
    if (OP == HFMA2_R) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == HFMA2_I) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(0x40004000u), "r"(c));
    if (OP == HADD2_R) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == HADD2_I) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(a) : "r"(0x64006400u));
    if (OP == HFMA2_SAT) asm volatile("fma.rn.sat.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(0xBC00BC00u), "r"(0u));
    if (OP == HSET2) asm volatile("set.eq.f16x2.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == VMIN) a = (j & 1) ? __vminu2(a, b) : __vmaxu2(a, c);
    if (OP == VMIN3) a = (j & 1) ? __vimin3_u16x2(a, b, c) : __vimax3_u16x2(a, c, b);
    if (OP == VADDMNMX) a = __viaddmin_s16x2_relu(a, b, c);
    if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(0x4140u));
    if (OP == IMADSHL) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == HMNMX2) { if (j & 1) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b)); else asm volatile("max.f16x2 %0, %0, %1;" : "+r"(a) : "r"(c)); }
    if (OP == FFMA_R) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == MIX_HFMA_VMIN) { if (j & 1) op<VMIN>(a, b, c, j); else op<HFMA2_R>(a, b, c, j); }
    if (OP == MIX_HFMA_LOP3) { if (j & 1) op<LOP3>(a, b, c, j); else op<HFMA2_R>(a, b, c, j); }
    if (OP == MIX_HADD_PRMT) { if (j & 1) op<PRMT>(a, b, c, j); else op<HADD2_R>(a, b, c, j); }
    if (OP == MIX_HFMA_HADD) { if (j & 1) op<HADD2_R>(a, b, c, j); else op<HFMA2_R>(a, b, c, j); }
    if (OP == MIX_VMIN_LOP3) { if (j & 1) op<LOP3>(a, b, c, j); else op<VMIN>(a, b, c, j); }
    if (OP == MIX_HFMA_IMAD) { if (j & 1) op<IMADSHL>(a, b, c, j); else op<HFMA2_R>(a, b, c, j); }
    if (OP == MIX3) {
        if ((j & 3) == 0) op<HFMA2_R>(a, b, c, j);
        if ((j & 3) == 1) op<VMIN>(a, b, c, j);
        if ((j & 3) == 2) op<HADD2_R>(a, b, c, j);
        if ((j & 3) == 3) op<LOP3>(a, b, c, j);
    }
}
template <int OP>
__global__ void k(unsigned *out, long long *cyc, unsigned b, unsigned c)
{
    unsigned a[8];
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int j = 0; j < OPS_PER_ITER; j++) op<OP>(a[j & 7], b, c, j >> 3);
    }
    long long t1 = clock64();
    __syncthreads();
    unsigned s = 0;
    for (int i = 0; i < 8; i++) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(unsigned *out, long long *cyc)
{
    for (int w : {1, 2, 5, 8}) {
        k<OP><<<148, 128 * w>>>(out, cyc, 0x3c003c00u, 0x00010001u);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double m = 0;
        for (int i = 0; i < 148; i++) m += h[i];
        m /= 148;
        printf("%-30s warps/SMSP=%d  cycles per warp-inst per SMSP = %.3f\n", names[OP], w,
               m / ((double)ITERS * OPS_PER_ITER * w));
    }
}
template <int OP>
struct RunAll {
    static void go(unsigned *o, long long *c) { run<OP>(o, c); RunAll<OP + 1>::go(o, c); }
};
template <>
struct RunAll<NOPS> {
    static void go(unsigned *, long long *) {}
};
int main()
{
    unsigned *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    RunAll<0>::go(out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
