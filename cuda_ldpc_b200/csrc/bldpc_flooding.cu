// Flooding un-normalised min-sum, fp32 — the strict-parity mode.
//
// Same update rules, message memory layout and iteration structure as the reference
// (B/LDPC_Decoder.cu:94-131 host loop, :172-261 VN kernels, :262-372 CN kernels), on the true
// circulant graph.  Messages live in msgs[(m*dc_max + p)*F + f] exactly like Memory_RQ
// (B/LDPC_Decoder.cu:38, slot = check*Wc_max + position), frames fastest, so every access of a
// warp is a coalesced 128-bit load/store over 4 consecutive frames.  The circulant shift is
// address arithmetic (no Address_Variablenode table in memory).  fp32 adds are issued in the
// reference's order, so results are bit-identical to the oracle (oracle/bldpc_oracle.c).
//
// HBM bound: (4E + 2N) * 4 bytes per codeword-iteration (SURVEY §8d "reference layout").
#include "common.h"

namespace ldpcb {

template <int VEC>
struct Vec;
template <>
struct Vec<4> {
    using T = float4;
};
template <>
struct Vec<1> {
    using T = float;
};

__device__ __forceinline__ float vget(const float4 &v, int j) { return (&v.x)[j]; }
__device__ __forceinline__ float vget(const float &v, int) { return v; }
__device__ __forceinline__ void vset(float4 &v, int j, float x) { (&v.x)[j] = x; }
__device__ __forceinline__ void vset(float &v, int, float x) { v = x; }

// Variablenode_Kernel (B/LDPC_Decoder.cu:172-216): S = sum_i R_i + y; D = S < 0; Q_i = S - R_i.
// Two passes over the node's messages instead of holding all of them in registers (dv up to 15 would need
// ~150 registers and leave one CTA per SM — ncu: 7.6 warps per SM, 28 % of the HBM peak): pass 1 sums in the
// reference's order with kBatch loads in flight, pass 2 re-reads the same vectors — L1 hits, a CTA touches
// 256 x dv x 16 B — and writes S - R_i in place.
template <int VEC>
__global__ void __launch_bounds__(256)
flood_vn_kernel(const __grid_constant__ ColumnTables ct, float *__restrict__ msgs, const float *__restrict__ y,
                unsigned char *__restrict__ hard, const int *__restrict__ done, const int *__restrict__ stop, int N, int Z,
                int F, int dcmax)
{
    using V = typename Vec<VEC>::T;
    constexpr int kBatch = 4;
    const int FV = F / VEC;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)N * FV) return;
    if (stop && *stop) return;  // genie stop reached in an earlier iteration: the batch is frozen (device-side exit)
    const int n = (int)(tid / FV);
    const int f = (int)(tid % FV) * VEC;
    const int c = n / Z, j = n - c * Z;
    const int dv = ct.dv[c], voff = ct.voff[c];
    auto addr = [&](int i) -> size_t {
        const int e = voff + i;
        int row = j - (int)ct.shift[e];
        row += (row < 0) ? Z : 0;
        return ((size_t)((int)ct.row[e] * Z + row) * dcmax + ct.pos[e]) * F + f;
    };
    float s[VEC];
#pragma unroll
    for (int l = 0; l < VEC; l++) s[l] = 0.0f;  // the reference leaves Add_result uninitialised; 0 is the intent
    for (int i0 = 0; i0 < dv; i0 += kBatch) {
        V R[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; u++)
            if (i0 + u < dv) R[u] = *reinterpret_cast<const V *>(msgs + addr(i0 + u));
#pragma unroll
        for (int u = 0; u < kBatch; u++)
            if (i0 + u < dv)
#pragma unroll
                for (int l = 0; l < VEC; l++) s[l] = __fadd_rn(s[l], vget(R[u], l));
    }
    const V yy = *reinterpret_cast<const V *>(y + (size_t)n * F + f);
    unsigned hbits = 0;
    bool wr = true;
#pragma unroll
    for (int l = 0; l < VEC; l++) {
        s[l] = __fadd_rn(s[l], vget(yy, l));
        hbits |= (s[l] < 0.0f ? 1u : 0u) << (8 * l);
        if (done && done[f + l]) wr = false;
    }
    if (VEC == 4 && (!done || wr)) {
        *reinterpret_cast<unsigned *>(hard + (size_t)n * F + f) = hbits;  // 4 frames, one store
    } else {
#pragma unroll
        for (int l = 0; l < VEC; l++)
            if (!done || !done[f + l]) hard[(size_t)n * F + f + l] = (hbits >> (8 * l)) & 1u;
    }
    for (int i0 = 0; i0 < dv; i0 += kBatch) {
        V R[kBatch];
        size_t ad[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; u++)
            if (i0 + u < dv) {
                ad[u] = addr(i0 + u);
                R[u] = *reinterpret_cast<const V *>(msgs + ad[u]);
            }
#pragma unroll
        for (int u = 0; u < kBatch; u++)
            if (i0 + u < dv) {
                V q;
#pragma unroll
                for (int l = 0; l < VEC; l++) vset(q, l, __fsub_rn(s[l], vget(R[u], l)));
                *reinterpret_cast<V *>(msgs + ad[u]) = q;
            }
    }
}

// Checknode_Kernel (B/LDPC_Decoder.cu:262-314) + sortQ (:374-398): two smallest magnitudes,
// first index of the smallest, product of signs (zero counts as +), no scaling.  The dc loads of a check are
// issued kBatch at a time before their first use (ncu: with one load in flight per thread the kernel sat at
// 54 % of the HBM peak, 8 warps out of 9 waiting on the long scoreboard).
template <int VEC>
__global__ void __launch_bounds__(256)
flood_cn_kernel(const __grid_constant__ LayerTables lt, float *__restrict__ msgs, const int *__restrict__ stop, int M,
                int Z, int F, int dcmax)
{
    using V = typename Vec<VEC>::T;
    constexpr int kBatch = 8;
    const int FV = F / VEC;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)M * FV) return;
    if (stop && *stop) return;
    const int m = (int)(tid / FV);
    const int f = (int)(tid % FV) * VEC;
    const int dc = lt.dc[m / Z];
    float *base = msgs + ((size_t)m * dcmax) * F + f;
    float min1[VEC], min2[VEC];
    int idx[VEC];
    unsigned neg[VEC];
#pragma unroll
    for (int l = 0; l < VEC; l++) {
        min1[l] = __int_as_float(0x7f800000);
        min2[l] = __int_as_float(0x7f800000);
        idx[l] = 0;
        neg[l] = 0;
    }
    for (int i0 = 0; i0 < dc; i0 += kBatch) {
        V q[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; u++)
            if (i0 + u < dc) q[u] = *reinterpret_cast<const V *>(base + (size_t)(i0 + u) * F);
#pragma unroll
        for (int u = 0; u < kBatch; u++) {
            const int i = i0 + u;
            if (i < dc) {
#pragma unroll
                for (int l = 0; l < VEC; l++) {
                    const float v = vget(q[u], l);
                    const float a = (v < 0.0f) ? -v : v;
                    neg[l] |= (v < 0.0f ? 1u : 0u) << i;
                    if (a < min1[l]) {
                        min2[l] = min1[l];
                        min1[l] = a;
                        idx[l] = i;
                    } else if (a < min2[l])
                        min2[l] = a;
                }
            }
        }
    }
    for (int i = 0; i < dc; i++) {
        V r;
#pragma unroll
        for (int l = 0; l < VEC; l++) {
            const int sign_all = (__popc(neg[l]) & 1) ? -1 : 1;
            const int sign_i = ((neg[l] >> i) & 1u) ? -1 : 1;
            const float mag = (i != idx[l]) ? min1[l] : min2[l];
            vset(r, l, __fmul_rn((float)(sign_all * sign_i), mag));
        }
        *reinterpret_cast<V *>(base + (size_t)i * F) = r;
    }
}

// bad[f] |= 1 if any check of frame f has odd parity on hard[N][F]
__global__ void
syndrome_nf_kernel(const __grid_constant__ LayerTables lt, const unsigned char *__restrict__ hard,
                   int *__restrict__ bad, int M, int Z, int F)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)M * F) return;
    const int m = (int)(tid / F), f = (int)(tid % F);
    const int r = m / Z, i = m - r * Z;
    const int dc = lt.dc[r], off = lt.off[r];
    int p = 0;
    for (int k = 0; k < dc; k++) {
        int col = i + (int)lt.shift[off + k];
        col -= (col >= Z) ? Z : 0;
        p ^= hard[(size_t)((int)lt.col[off + k] * Z + col) * F + f];
    }
    if (p & 1) bad[f] = 1;
}

// genie test (B/LDPC_Decoder.cu:134-149): bad[f] = 1 if any of the first `length` bits is 1
__global__ void __launch_bounds__(256)
genie_nf_kernel(const unsigned char *__restrict__ hard, int *__restrict__ bad, int length, int F)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)length * F) return;
    if (hard[tid]) bad[(int)(tid % F)] = 1;
}

// per-frame bookkeeping after iteration `it`; counter[0] += frames still running
__global__ void
flood_finalize_kernel(const int *__restrict__ bad, int *__restrict__ done, int *__restrict__ iters,
                      int *__restrict__ ok, int *__restrict__ counter, const int *__restrict__ stop, int it, int F,
                      int latch)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    if (stop && *stop) return;
    if (latch) {
        if (done[f]) return;
        iters[f] = it;
        if (!bad[f]) {
            done[f] = 1;
            ok[f] = 1;
        } else {
            ok[f] = 0;
            atomicAdd(counter, 1);
        }
    } else {
        iters[f] = it;
        ok[f] = bad[f] ? 0 : 1;
        if (bad[f]) atomicAdd(counter, 1);
    }
}

// batch-wide stop (genie rule, B/LDPC_Decoder.cu:150-153: "all frames decoded" ends the loop): evaluated ON THE
// DEVICE — later iterations are enqueued anyway and return at their first instruction, so ldpc_decode_batch with
// device buffers never blocks the host (the reference copies N*F ints to the host and tests them there, every iteration)
__global__ void flood_stop_kernel(const int *__restrict__ counter, int *__restrict__ stop)
{
    if (*counter == 0) *stop = 1;
}

static inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

template <int VEC>
static void launch_vn(const ldpc_code *c, float *msgs, const float *y, unsigned char *hard, const int *done,
                      const int *stop, int F, cudaStream_t st)
{
    const long long n = (long long)c->N * (F / VEC);
    flood_vn_kernel<VEC><<<blocks_for(n, 256), 256, 0, st>>>(c->ct, msgs, y, hard, done, stop, c->N, c->Z, F, c->dc_max);
}

// y_nf: device fp32 [N][F]; hard_nf: device u8 [N][F]; iters_dev/ok_dev: device int [F];
// msgs: device fp32 [M*dc_max][F]; flag_scratch: device int [2F + 2] (bad, done, counter, stop).
// may_block: the caller synchronises anyway (host buffers), so the loop may read the 4-byte counter back and stop
// enqueuing; otherwise (device buffers) every iteration is enqueued and the exit happens on the device.
int launch_flooding_fp32(const ldpc_code *c, const float *y, int F, int iters, int exit_mode,
                         unsigned char *hard, int *iters_dev, int *ok_dev, float *msgs, int *flag_scratch,
                         cudaStream_t st, int *launches, bool may_block)
{
    int *bad = flag_scratch, *done = flag_scratch + F, *counter = flag_scratch + 2 * F, *stop = counter + 1;
    const size_t msg_bytes = (size_t)c->M * c->dc_max * F * sizeof(float);
    LDPC_CUDA_TRY(cudaMemsetAsync(msgs, 0, msg_bytes, st));  // B/LDPC_Decoder.cu:82
    LDPC_CUDA_TRY(cudaMemsetAsync(flag_scratch, 0, (size_t)(2 * F + 2) * sizeof(int), st));
    const bool vec4 = (F % 4 == 0);
    const int latch = (exit_mode == LDPC_EXIT_SYNDROME);
    const int *gstop = (exit_mode == LDPC_EXIT_GENIE) ? stop : nullptr;  // syndrome mode freezes frames one by one (done[])
    int n = 0, it = 0;
    while (it < iters) {
        it++;
        if (vec4) {
            launch_vn<4>(c, msgs, y, hard, latch ? done : nullptr, gstop, F, st);
            flood_cn_kernel<4><<<blocks_for((long long)c->M * (F / 4), 256), 256, 0, st>>>(c->lt, msgs, gstop, c->M,
                                                                                          c->Z, F, c->dc_max);
        } else {
            launch_vn<1>(c, msgs, y, hard, latch ? done : nullptr, gstop, F, st);
            flood_cn_kernel<1><<<blocks_for((long long)c->M * F, 256), 256, 0, st>>>(c->lt, msgs, gstop, c->M, c->Z, F,
                                                                                    c->dc_max);
        }
        n += 2;
        const bool last = (it == iters);
        if (exit_mode != LDPC_EXIT_NONE || last) {
            LDPC_CUDA_TRY(cudaMemsetAsync(bad, 0, (size_t)F * sizeof(int), st));
            LDPC_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(int), st));
            if (exit_mode == LDPC_EXIT_GENIE)
                genie_nf_kernel<<<blocks_for((long long)c->K * F, 256), 256, 0, st>>>(hard, bad, c->K, F);
            else
                syndrome_nf_kernel<<<blocks_for((long long)c->M * F, 256), 256, 0, st>>>(c->lt, hard, bad, c->M,
                                                                                        c->Z, F);
            flood_finalize_kernel<<<blocks_for(F, 256), 256, 0, st>>>(bad, done, iters_dev, ok_dev, counter, gstop, it, F,
                                                                     latch);
            n += 2;
            if (gstop && !last) {
                flood_stop_kernel<<<1, 1, 0, st>>>(counter, stop);
                n++;
            }
            if (may_block && exit_mode != LDPC_EXIT_NONE && !last) {
                int running = 0;  // the reference copies N*F ints per iteration here (:135); we copy 4 bytes
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
