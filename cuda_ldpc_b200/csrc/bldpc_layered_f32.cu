// Layered (row-block) min-sum in fp32 — the float twin of the throughput mode, for accuracy studies.
// State streams through HBM like the reference's layout (frames interleaved, frame index fastest):
// APP app[n*F+f] and the explicit check-to-variable messages c2v[(m*dc_max+k)*F+f]; one launch per
// layer, thread = (check row, 4 consecutive frames), 128-bit coalesced accesses, circulant shift =
// address arithmetic.  Update rule per B/LDPC_Decoder.cu:279-314 on t = APP - c2v_old, alpha applied
// to min1/min2 (1.0 = the reference's un-normalised rule); operation order = oracle
// (orc_layered_fp32), so the results are bit-identical.  HBM bound: 4*E*4 bytes per codeword-iteration
// plus the record-free c2v traffic — this mode is not the headline (bldpc_layered.cu is).
#include "common.h"

namespace ldpcb {

template <int VEC>
struct VecF;
template <>
struct VecF<4> {
    using T = float4;
};
template <>
struct VecF<1> {
    using T = float;
};
__device__ __forceinline__ float vgetf(const float4 &v, int j) { return (&v.x)[j]; }
__device__ __forceinline__ float vgetf(const float &v, int) { return v; }
__device__ __forceinline__ void vsetf(float4 &v, int j, float x) { (&v.x)[j] = x; }
__device__ __forceinline__ void vsetf(float &v, int, float x) { v = x; }

// Two passes over the row's edges, each in batches of kBatch edges whose APP and message vectors (128-bit for
// VEC = 4) are all requested before the first is used; the second pass re-reads what the first just loaded
// (L1/L2 hits).  The first version loaded scalars one edge at a time and ran 5x below the HBM bound.
template <int VEC>
__global__ void __launch_bounds__(256)
layered_f32_layer_kernel(const __grid_constant__ LayerTables lt, float *__restrict__ app, float *__restrict__ c2v,
                         const int *__restrict__ done, int r, int Z, int F, int dcmax, float alpha)
{
    using V = typename VecF<VEC>::T;
    constexpr int kBatch = 4;
    const int FV = F / VEC;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)Z * FV) return;
    const int i = (int)(tid / FV), f = (int)(tid % FV) * VEC;
    const int dc = lt.dc[r], off = lt.off[r];
    float *cm = c2v + ((size_t)(r * Z + i) * dcmax) * F + f;
    auto app_of = [&](int k) -> float * {
        int col = i + (int)lt.shift[off + k];
        col -= (col >= Z) ? Z : 0;
        return app + (size_t)((int)lt.col[off + k] * Z + col) * F + f;
    };
    float min1[VEC], min2[VEC];
    int idx[VEC];
    unsigned neg[VEC];
    bool act[VEC];
#pragma unroll
    for (int l = 0; l < VEC; l++) {
        min1[l] = __int_as_float(0x7f800000);
        min2[l] = __int_as_float(0x7f800000);
        idx[l] = 0;
        neg[l] = 0;
        act[l] = !done || !done[f + l];
    }
    for (int k0 = 0; k0 < dc; k0 += kBatch) {
        V av[kBatch], cv[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; u++)
            if (k0 + u < dc) {
                av[u] = *reinterpret_cast<const V *>(app_of(k0 + u));
                cv[u] = *reinterpret_cast<const V *>(cm + (size_t)(k0 + u) * F);
            }
#pragma unroll
        for (int u = 0; u < kBatch; u++) {
            const int k = k0 + u;
            if (k < dc) {
#pragma unroll
                for (int l = 0; l < VEC; l++) {
                    const float t = __fsub_rn(vgetf(av[u], l), vgetf(cv[u], l));
                    const float m = (t < 0.0f) ? -t : t;
                    neg[l] |= (t < 0.0f ? 1u : 0u) << k;
                    if (m < min1[l]) {
                        min2[l] = min1[l];
                        min1[l] = m;
                        idx[l] = k;
                    } else if (m < min2[l])
                        min2[l] = m;
                }
            }
        }
    }
    for (int k0 = 0; k0 < dc; k0 += kBatch) {
        V av[kBatch], cv[kBatch];
        float *ap[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; u++)
            if (k0 + u < dc) {
                ap[u] = app_of(k0 + u);
                av[u] = *reinterpret_cast<const V *>(ap[u]);
                cv[u] = *reinterpret_cast<const V *>(cm + (size_t)(k0 + u) * F);
            }
#pragma unroll
        for (int u = 0; u < kBatch; u++) {
            const int k = k0 + u;
            if (k < dc) {
                V na = av[u], nc = cv[u];
#pragma unroll
                for (int l = 0; l < VEC; l++) {
                    if (!act[l]) continue;
                    const float t = __fsub_rn(vgetf(av[u], l), vgetf(cv[u], l));
                    const float m = __fmul_rn(alpha, (k == idx[l]) ? min2[l] : min1[l]);
                    const bool s = ((__popc(neg[l]) & 1) != 0) != (((neg[l] >> k) & 1u) != 0);
                    const float nw = s ? -m : m;
                    vsetf(nc, l, nw);
                    vsetf(na, l, __fadd_rn(t, nw));
                }
                *reinterpret_cast<V *>(cm + (size_t)k * F) = nc;
                *reinterpret_cast<V *>(ap[u]) = na;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
hard_from_app_kernel(const float *__restrict__ app, unsigned char *__restrict__ hard, const int *__restrict__ done,
                     long long n, int F)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    if (done && done[(int)(tid % F)]) return;
    hard[tid] = app[tid] < 0.0f;
}

// kernels shared with the flooding path (bldpc_flooding.cu)
__global__ void syndrome_nf_kernel(const __grid_constant__ LayerTables lt, const unsigned char *__restrict__ hard,
                                   int *__restrict__ bad, int M, int Z, int F);
__global__ void flood_finalize_kernel(const int *__restrict__ bad, int *__restrict__ done, int *__restrict__ iters,
                                      int *__restrict__ ok, int *__restrict__ counter, const int *__restrict__ stop, int it,
                                      int F, int latch);

// y: device fp32 [N][F]; app: device fp32 [N][F] scratch (also the debug dump); msgs: [M*dc_max][F]
int launch_layered_f32_nf(const ldpc_code *c, const float *y, int F, int iters, int exit_mode, float alpha, float *app,
                          float *msgs, unsigned char *hard, int *iters_dev, int *ok_dev, int *flag_scratch,
                          cudaStream_t st, int *launches, bool may_block)
{
    int *bad = flag_scratch, *done = flag_scratch + F, *counter = flag_scratch + 2 * F;
    LDPC_CUDA_TRY(cudaMemcpyAsync(app, y, (size_t)c->N * F * sizeof(float), cudaMemcpyDeviceToDevice, st));
    LDPC_CUDA_TRY(cudaMemsetAsync(msgs, 0, (size_t)c->M * c->dc_max * F * sizeof(float), st));
    LDPC_CUDA_TRY(cudaMemsetAsync(flag_scratch, 0, (size_t)(2 * F + 2) * sizeof(int), st));
    const bool vec4 = (F % 4 == 0);
    const int latch = (exit_mode == LDPC_EXIT_SYNDROME);
    const long long nt = (long long)c->Z * (vec4 ? F / 4 : F), nf = (long long)c->N * F;
    int n = 0, it = 0;
    while (it < iters) {
        it++;
        for (int r = 0; r < c->J; r++) {
            if (vec4)
                layered_f32_layer_kernel<4><<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(c->lt, app, msgs, latch ? done : nullptr,
                                                                                         r, c->Z, F, c->dc_max, alpha);
            else
                layered_f32_layer_kernel<1><<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(c->lt, app, msgs, latch ? done : nullptr,
                                                                                         r, c->Z, F, c->dc_max, alpha);
        }
        n += c->J;
        const bool last = (it == iters);
        if (latch || last) {
            hard_from_app_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(app, hard, latch ? done : nullptr, nf, F);
            LDPC_CUDA_TRY(cudaMemsetAsync(bad, 0, (size_t)F * sizeof(int), st));
            LDPC_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(int), st));
            syndrome_nf_kernel<<<(unsigned)(((long long)c->M * F + 255) / 256), 256, 0, st>>>(c->lt, hard, bad, c->M, c->Z, F);
            flood_finalize_kernel<<<(F + 255) / 256, 256, 0, st>>>(bad, done, iters_dev, ok_dev, counter, nullptr, it, F, latch);
            n += 3;
            // device buffers: never block — converged frames are frozen one by one through done[], the remaining
            // iterations run as no-ops; host buffers: the caller synchronises anyway, so stop enqueuing early
            if (may_block && latch && !last) {
                int running = 0;
                LDPC_CUDA_TRY(cudaMemcpyAsync(&running, counter, sizeof(int), cudaMemcpyDeviceToHost, st));
                LDPC_CUDA_TRY(cudaStreamSynchronize(st));
                if (running == 0) break;
            }
        }
    }
    LDPC_CUDA_TRY(cudaGetLastError());
    *launches += n;
    return LDPC_OK;
}

}  // namespace ldpcb
