// A/B for the FFT-BP transform at q = 256 (BASELINE.json north star: "Tensor cores are used only if the GF(2^p) FFT-BP
// Walsh-Hadamard stage is written as a dense q by q contraction"):
//   arm A  butterfly: one warp per row, 8 elements per lane, 3 stages in registers + 5 stages by __shfl_xor — no
//          shared memory, 64 FADD + 40 SHFL per lane and row (fp32, the decoder's arithmetic);
//   arm B  dense contraction [R, 256] x H_256 (+-1) on the tensor cores through cuBLAS (fp16 operands, fp32
//          accumulate; cublasGemmEx is the stand-in for a hand-written tcgen05.mma tile loop: it bounds what such a
//          loop could reach).
// Both arms are timed (1) once over device-resident rows (HBM in / out) and (2) with the transform applied REP times
// per load (compute rate with the data on chip: the situation inside the decoder, where the rows sit in shared memory).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o wht_ab wht_ab.cu -lcublas
#include <cublas_v2.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <int REP>
__global__ void __launch_bounds__(256) wht256_rows(const float *__restrict__ in, float *__restrict__ out, int rows)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    float v[8];
    const float4 *src = reinterpret_cast<const float4 *>(in + (size_t)warp * 256 + lane * 8);
    const float4 a = src[0], b = src[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
#pragma unroll 1
    for (int rep = 0; rep < REP; rep++) {
#pragma unroll
        for (int len = 1; len < 8; len <<= 1)  // element index = lane * 8 + t: the three low stages stay in the lane
#pragma unroll
            for (int t = 0; t < 8; t++)
                if (!(t & len)) {
                    const float x = v[t], y = v[t | len];
                    v[t] = x + y;
                    v[t | len] = x - y;
                }
#pragma unroll
        for (int m = 1; m < 32; m <<= 1)       // the five high stages cross lanes
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const float o = __shfl_xor_sync(0xffffffffu, v[t], m);
                v[t] = (lane & m) ? o - v[t] : v[t] + o;
            }
        if (REP > 1) {
#pragma unroll
            for (int t = 0; t < 8; t++) v[t] *= 0.0625f;  // keep the values bounded over the repetitions
        }
    }
    float4 *dst = reinterpret_cast<float4 *>(out + (size_t)warp * 256 + lane * 8);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
}

int main()
{
    const int R = 1 << 20, q = 256, REP = 16;  // 1 Mi rows = 87 381 checks of degree 12
    float *x, *y;
    __half *xh, *hh;
    float *yh;
    CK(cudaMalloc(&x, (size_t)R * q * 4)); CK(cudaMalloc(&y, (size_t)R * q * 4));
    CK(cudaMalloc(&xh, (size_t)R * q * 2)); CK(cudaMalloc(&hh, (size_t)q * q * 2)); CK(cudaMalloc(&yh, (size_t)R * q * 4));
    std::vector<float> hx((size_t)R * q);
    for (size_t i = 0; i < hx.size(); i++) hx[i] = (float)((i * 2654435761u >> 8) & 255) / 256.0f;
    std::vector<__half> hxh(hx.size()), hH((size_t)q * q);
    for (size_t i = 0; i < hx.size(); i++) hxh[i] = __float2half(hx[i]);
    for (int i = 0; i < q; i++) for (int j = 0; j < q; j++) hH[(size_t)i * q + j] = __float2half((__builtin_popcount(i & j) & 1) ? -1.0f : 1.0f);
    CK(cudaMemcpy(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(xh, hxh.data(), hxh.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(hh, hH.data(), hH.size() * 2, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    // ---- arm A
    const int blocks = (R * 32 + 255) / 256;
    wht256_rows<1><<<blocks, 256>>>(x, y, R);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < 5; i++) wht256_rows<1><<<blocks, 256>>>(x, y, R);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    const double a1 = ms / 5;
    cudaEventRecord(e0);
    for (int i = 0; i < 3; i++) wht256_rows<REP><<<blocks, 256>>>(x, y, R);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    const double aR = ms / 3;
    // check arm A against a host WHT of row 5
    std::vector<float> row(q), got(q);
    wht256_rows<1><<<blocks, 256>>>(x, y, R);
    CK(cudaMemcpy(got.data(), y + 5 * q, q * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int k = 0; k < q; k++) { double s = 0; for (int j = 0; j < q; j++) s += ((__builtin_popcount(k & j) & 1) ? -1.0 : 1.0) * hx[5 * q + j]; if (fabs(s - got[k]) > maxerr) maxerr = fabs(s - got[k]); }
    // ---- arm B: Y[R,256] = X[R,256] * H (row-major) == column-major Y^T = H^T X^T
    cublasHandle_t h; cublasCreate(&h);
    const float one = 1.0f, zero = 0.0f;
    auto gemm = [&]() { return cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_N, q, R, q, &one, hh, CUDA_R_16F, q, xh, CUDA_R_16F, q, &zero, yh, CUDA_R_32F, q, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT_TENSOR_OP); };
    if (gemm() != CUBLAS_STATUS_SUCCESS) { printf("cublasGemmEx failed\n"); return 1; }
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < 5; i++) gemm();
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    const double b1 = ms / 5;
    CK(cudaMemcpy(got.data(), yh + 5 * q, q * 4, cudaMemcpyDeviceToHost));
    double maxerr_b = 0;
    for (int k = 0; k < q; k++) { double s = 0; for (int j = 0; j < q; j++) s += ((__builtin_popcount(k & j) & 1) ? -1.0 : 1.0) * hx[5 * q + j]; if (fabs(s - got[k]) > maxerr_b) maxerr_b = fabs(s - got[k]); }
    const double flop = 2.0 * R * q * q;
    printf("rows %d, q %d\n", R, q);
    printf("A butterfly (registers + shuffles, fp32): %.3f ms per pass over HBM-resident rows = %.2f G rows/s (%.0f GB/s);  %d transforms per load: %.3f ms = %.2f G transforms/s on chip;  max |err| vs host %.2e\n",
           a1, R / a1 / 1e6, 2.0 * R * q * 4 / a1 / 1e6, REP, aR, (double)R * REP / aR / 1e6, maxerr);
    printf("B dense [R,256] x H_256 on tensor cores (cuBLAS, fp16 in / fp32 out): %.3f ms = %.2f G rows/s, %.0f TFLOP/s;  max |err| vs host %.2e (fp16 operand rounding)\n",
           b1, R / b1 / 1e6, flop / b1 / 1e9, maxerr_b);
    return 0;
}
