// Layered (row-block) normalised min-sum with int8 state — the throughput mode.
//
// B200 mapping:
//  * one CTA decodes a GROUP of 4 codewords at a time; their APP values live in shared memory
//    for the whole decode as one 32-bit word per code bit (4 x biased uint8), so HBM sees
//    only the channel values in and the hard decisions out;
//  * thread = check row i of the current layer (block row of H); the circulant shift is the
//    shared-memory address (i + s) mod Z, consecutive threads hit consecutive banks;
//  * arithmetic is 2-way SIMD on half2 lanes holding small integers (|x| <= 254 is exact in
//    fp16): HADD2 / HMNMX2 / HSET2.BM / LOP3 / PRMT — sm_100a has no native 8x4 SIMD
//    integer min/compare (the __v*4 intrinsics expand to 5-10 instructions, profiles/r01_simd_sass.txt);
//  * the compressed check record {min1, min2, idx, sign bits} is streamed through a per-CTA
//    slice of a scratch buffer that stays L2-resident (148 CTAs x M x 32 B << 126 MB);
//  * the syndrome (early exit / ok flag) is a cheap extra pass: XOR of the sign bits.
//
// Update rules: exactly oracle/bldpc_oracle.c "int8 layered rules" (t = APP - c2v_old,
// min1/min2/first-index on min(|t|, msg_max), m' = m - ((m*beta_num) >> beta_shift), APP =
// sat127(t + c2v_new)); check-node rule per B/LDPC_Decoder.cu:279-314, schedule ours.
#include <cuda_fp16.h>

#include "common.h"

namespace ldpcb {

struct LayeredParams {
    const void *llr;
    void *out;
    int *iters_out, *ok_out;
    signed char *dbg_app;
    unsigned *dbg_rec;
    uint4 *rec;             // scratch: per CTA, M records of REC_U4 uint4
    int llr_dtype, layout, out_format;
    int F, N, Z, J, M;
    int iters, exit_mode, num_groups;
    float scale;
    float msg_max;
    float beta_mul;         // beta_num / 2^beta_shift
    float beta_bias;        // 0.5 - 2^-(beta_shift+1)
    LayerTables lt;
};

__device__ __forceinline__ unsigned h2u(__half2 h) { return *reinterpret_cast<unsigned *>(&h); }
__device__ __forceinline__ __half2 u2h(unsigned u) { return *reinterpret_cast<__half2 *>(&u); }
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned s)
{
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}
__device__ __forceinline__ unsigned sel(unsigned mask, unsigned a, unsigned b) { return (a & mask) | (b & ~mask); }

// biased uint8 x4  ->  half2 lanes (frames 0,1) / (frames 2,3): 0x6400|b is 1024+b in fp16
__device__ __forceinline__ __half2 unpack_lo(unsigned w)
{
    return __hsub2(u2h(prmt(w, 0x64646464u, 0x4140u)), __float2half2_rn(1152.0f));
}
__device__ __forceinline__ __half2 unpack_hi(unsigned w)
{
    return __hsub2(u2h(prmt(w, 0x64646464u, 0x4342u)), __float2half2_rn(1152.0f));
}
__device__ __forceinline__ unsigned pack4(__half2 a, __half2 b)
{
    const __half2 k = __float2half2_rn(1152.0f);
    return prmt(h2u(__hadd2(a, k)), h2u(__hadd2(b, k)), 0x6420u);
}

__device__ __forceinline__ int quant(float y, float scale)
{
    int q = __float2int_rn(__fmul_rn(y, scale));
    return max(-127, min(127, q));
}

// m' = m - floor(m * beta_num / 2^beta_shift), exact in fp16 for m <= 127, beta_num <= 8:
// u = m*mul - (0.5 - 2^-(sh+1)) has at most 11 significant bits; u + 1025 rounds to floor + 1025.
__device__ __forceinline__ __half2 beta_scale(__half2 m, __half2 mul, __half2 nbias)
{
    const __half2 k = __float2half2_rn(1025.0f);  // keeps u + k inside [1024, 2048): ulp 1, never a tie
    __half2 u = __hfma2(m, mul, nbias);
    __half2 fl = __hsub2(__hadd2(u, k), k);
    return __hsub2(m, fl);
}

template <int DCMAX>
struct RecLayout {
    static constexpr int SW = (DCMAX + 15) / 16;  // sign words per half-group
    static constexpr int WORDS = 6 + 2 * SW;      // m1 x2, m2 x2, idx x2, signs
    static constexpr int U4 = (WORDS + 3) / 4;
};

// One check row of one layer for the 4 codewords of the group.
template <int DCMAX, bool FIRST, bool SCALE>
__device__ __forceinline__ void process_row(unsigned char *app, const LayeredParams &p, int off, int dc, int i4,
                                            int Z4, uint4 *recp, __half2 amax, __half2 bmul, __half2 nbias)
{
    constexpr int SW = RecLayout<DCMAX>::SW;
    constexpr int U4 = RecLayout<DCMAX>::U4;
    unsigned rw[U4 * 4];
    if (!FIRST) {
#pragma unroll
        for (int q = 0; q < U4; q++) {
            const uint4 v = recp[q];
            rw[4 * q] = v.x;
            rw[4 * q + 1] = v.y;
            rw[4 * q + 2] = v.z;
            rw[4 * q + 3] = v.w;
        }
    }
    // record words: [0,1] m1 (lo,hi)  [2,3] m2  [4,5] idx  [6 + h*SW + w] sign bits, edge k at
    // bit 15-(k%16) of each 16-bit lane of word k/16
    __half2 t[2][DCMAX];
    int addr[DCMAX];
    __half2 min1[2] = {amax, amax}, min2[2] = {amax, amax};
    unsigned idx[2] = {0u, 0u};
    unsigned sacc[2][SW];
    unsigned pacc[2] = {0u, 0u};
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int w = 0; w < SW; w++) sacc[h][w] = 0u;

#pragma unroll
    for (int k = 0; k < DCMAX; k++) {
        if (k < dc) {
            const int e = off + k;
            int col4 = i4 + 4 * (int)p.lt.shift[e];
            col4 -= (col4 >= Z4) ? Z4 : 0;
            addr[k] = (int)p.lt.col[e] * Z4 + col4;
            const unsigned w = *reinterpret_cast<const unsigned *>(app + addr[k]);
            const unsigned kh = h2u(__float2half2_rn((float)k));
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const __half2 a = h ? unpack_hi(w) : unpack_lo(w);
                __half2 tt;
                if (FIRST) {
                    tt = a;
                } else {
                    const unsigned msk = __heq2_mask(u2h(rw[4 + h]), u2h(kh));
                    const unsigned mag = sel(msk, rw[2 + h], rw[h]);
                    const unsigned sg = (rw[6 + h * SW + (k >> 4)] << (k & 15)) & 0x80008000u;
                    tt = __hsub2(a, u2h(mag ^ sg));
                }
                t[h][k] = tt;
                const unsigned tu = h2u(tt);
                const __half2 ab = __hmin2(__habs2(tt), amax);
                pacc[h] ^= tu;
                sacc[h][k >> 4] |= (tu >> (k & 15)) & (0x80008000u >> (k & 15));
                const unsigned lt = __hlt2_mask(ab, min1[h]);
                idx[h] = sel(lt, kh, idx[h]);
                min2[h] = __hmin2(min2[h], __hmax2(min1[h], ab));
                min1[h] = __hmin2(min1[h], ab);
            }
        }
    }
    unsigned nsg[2][SW];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        if (SCALE) {
            min1[h] = beta_scale(min1[h], bmul, nbias);
            min2[h] = beta_scale(min2[h], bmul, nbias);
        }
        const unsigned pm = prmt(pacc[h], 0u, 0xBB99u);  // 0xFFFF per lane whose parity is odd
#pragma unroll
        for (int w = 0; w < SW; w++) nsg[h][w] = sacc[h][w] ^ pm;
        rw[h] = h2u(min1[h]);
        rw[2 + h] = h2u(min2[h]);
        rw[4 + h] = idx[h];
#pragma unroll
        for (int w = 0; w < SW; w++) rw[6 + h * SW + w] = nsg[h][w];
    }
#pragma unroll
    for (int q = 0; q < U4; q++) recp[q] = make_uint4(rw[4 * q], rw[4 * q + 1], rw[4 * q + 2], rw[4 * q + 3]);

    const __half2 lim = __float2half2_rn(127.0f), nlim = __float2half2_rn(-127.0f);
#pragma unroll
    for (int k = 0; k < DCMAX; k++) {
        if (k < dc) {
            const unsigned kh = h2u(__float2half2_rn((float)k));
            __half2 v[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const unsigned msk = __heq2_mask(u2h(idx[h]), u2h(kh));
                const unsigned mag = sel(msk, h2u(min2[h]), h2u(min1[h]));
                const unsigned sg = (nsg[h][k >> 4] << (k & 15)) & 0x80008000u;
                __half2 x = __hadd2(t[h][k], u2h(mag ^ sg));
                v[h] = __hmin2(__hmax2(x, nlim), lim);
            }
            *reinterpret_cast<unsigned *>(app + addr[k]) = pack4(v[0], v[1]);
        }
    }
}

// fail bits (0x80 per frame byte) of the checks this thread owns in layer r
__device__ __forceinline__ unsigned syndrome_row(const unsigned char *app, const LayeredParams &p, int off, int dc,
                                                 int i4, int Z4)
{
    unsigned x = (dc & 1) ? 0x80808080u : 0u;  // biased byte: negative <=> bit7 clear
    for (int k = 0; k < dc; k++) {
        const int e = off + k;
        int col4 = i4 + 4 * (int)p.lt.shift[e];
        col4 -= (col4 >= Z4) ? Z4 : 0;
        x ^= *reinterpret_cast<const unsigned *>(app + (int)p.lt.col[e] * Z4 + col4);
    }
    return x & 0x80808080u;
}

// hard decisions of frames selected by `fmask` (bit j = frame 4g+j) -> global memory
__device__ void write_outputs(const unsigned *appw, const LayeredParams &p, int g, unsigned fmask)
{
    const int F = p.F, N = p.N, f0 = 4 * g;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const unsigned w = appw[n];
        // bit j of `bits` = hard decision of frame j (biased byte < 128)
        const unsigned nb = ~w;
        const unsigned bits = ((nb >> 7) & 1u) | ((nb >> 14) & 2u) | ((nb >> 21) & 4u) | ((nb >> 28) & 8u);
        if (p.out_format == LDPC_OUT_INT32_REF) {
            int *D = reinterpret_cast<int *>(p.out);
#pragma unroll
            for (int j = 0; j < 4; j++)
                if ((fmask >> j) & 1u) D[(size_t)n * F + f0 + j] = (bits >> j) & 1u;
        } else if (p.out_format == LDPC_OUT_U8) {
            unsigned char *D = reinterpret_cast<unsigned char *>(p.out);
#pragma unroll
            for (int j = 0; j < 4; j++)
                if ((fmask >> j) & 1u) {
                    const size_t o = (p.layout == LDPC_LAYOUT_NF) ? (size_t)n * F + f0 + j : (size_t)(f0 + j) * N + n;
                    D[o] = (bits >> j) & 1u;
                }
        }
        if (p.dbg_app) {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if ((fmask >> j) & 1u) p.dbg_app[(size_t)n * F + f0 + j] = (signed char)((int)((w >> (8 * j)) & 255u) - 128);
        }
    }
    if (p.out_format == LDPC_OUT_BITPACK) {
        unsigned *D = reinterpret_cast<unsigned *>(p.out);
        const int W = (N + 31) / 32;
        const int lane = threadIdx.x & 31;
        const int nround = (N + 31) & ~31;
        for (int n = threadIdx.x; n < nround; n += blockDim.x) {  // blockDim is a multiple of 32
            const unsigned w = (n < N) ? appw[n] : 0x80808080u;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const unsigned b = __ballot_sync(0xffffffffu, ((w >> (8 * j + 7)) & 1u) == 0u);
                if (lane == 0 && ((fmask >> j) & 1u)) D[(size_t)(f0 + j) * W + (n >> 5)] = b;
            }
        }
    }
}

template <int DCMAX>
__device__ void dump_records(const LayeredParams &p, const uint4 *rec, int g, unsigned fmask)
{
    constexpr int SW = RecLayout<DCMAX>::SW;
    constexpr int U4 = RecLayout<DCMAX>::U4;
    const unsigned *rw = reinterpret_cast<const unsigned *>(rec);
    for (int m = threadIdx.x; m < p.M; m += blockDim.x) {
        const unsigned *r = rw + (size_t)m * U4 * 4;
        for (int j = 0; j < 4; j++) {
            if (!((fmask >> j) & 1u)) continue;
            const int h = j >> 1, sh = (j & 1) * 16;
            const __half m1 = __ushort_as_half((unsigned short)(r[h] >> sh));
            const __half m2 = __ushort_as_half((unsigned short)(r[2 + h] >> sh));
            const __half ix = __ushort_as_half((unsigned short)(r[4 + h] >> sh));
            unsigned sg = 0;
            for (int k = 0; k < DCMAX; k++)
                sg |= ((r[6 + h * SW + (k >> 4)] >> (sh + 15 - (k & 15))) & 1u) << k;
            const int dc = p.lt.dc[m / p.Z];
            sg &= (dc >= 32) ? 0xffffffffu : ((1u << dc) - 1u);
            const size_t f = (size_t)4 * g + j;
            p.dbg_rec[((size_t)m * 4 + 0) * p.F + f] = (unsigned)__half2int_rn(m1);
            p.dbg_rec[((size_t)m * 4 + 1) * p.F + f] = (unsigned)__half2int_rn(m2);
            p.dbg_rec[((size_t)m * 4 + 2) * p.F + f] = (unsigned)__half2int_rn(ix);
            p.dbg_rec[((size_t)m * 4 + 3) * p.F + f] = sg;
        }
    }
}

// threads per CTA are capped so that the per-row register arrays (2*DCMAX + DCMAX) never spill
template <int DCMAX>
struct ThreadCap {
    static constexpr int value = DCMAX <= 8 ? 640 : (DCMAX <= 16 ? 512 : 384);
};

template <int DCMAX, bool SCALE>
__global__ void __launch_bounds__(ThreadCap<DCMAX>::value, 1)
ldpc_layered_i8_kernel(const __grid_constant__ LayeredParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned *appw = reinterpret_cast<unsigned *>(smem);
    __shared__ unsigned s_fail;
    const int tid = threadIdx.x, T = blockDim.x;
    const int N = p.N, Z = p.Z, F = p.F, Z4 = 4 * Z;
    constexpr int U4 = RecLayout<DCMAX>::U4;
    uint4 *rec = p.rec + (size_t)blockIdx.x * p.M * U4;
    const __half2 amax = __float2half2_rn(p.msg_max);
    const __half2 bmul = __float2half2_rn(p.beta_mul);
    const __half2 nbias = __float2half2_rn(-p.beta_bias);

    for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x) {
        const int f0 = 4 * g;
        unsigned valid = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) valid |= (f0 + j < F) ? (1u << j) : 0u;
        // ---- load + quantise: q = sat127(rint(y * scale)), stored biased by 128
        for (int n = tid; n < N; n += T) {
            int q[4];
            if (p.llr_dtype == LDPC_DTYPE_FP32) {
                const float *y = reinterpret_cast<const float *>(p.llr);
                if (p.layout == LDPC_LAYOUT_NF && valid == 0xFu && (F & 3) == 0) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(y + (size_t)n * F + f0));
                    q[0] = quant(v.x, p.scale);
                    q[1] = quant(v.y, p.scale);
                    q[2] = quant(v.z, p.scale);
                    q[3] = quant(v.w, p.scale);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const size_t o = (p.layout == LDPC_LAYOUT_NF) ? (size_t)n * F + f0 + j : (size_t)(f0 + j) * N + n;
                        q[j] = ((valid >> j) & 1u) ? quant(__ldg(y + o), p.scale) : 127;
                    }
                }
            } else if (p.llr_dtype == LDPC_DTYPE_FP16) {
                const __half *y = reinterpret_cast<const __half *>(p.llr);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const size_t o = (p.layout == LDPC_LAYOUT_NF) ? (size_t)n * F + f0 + j : (size_t)(f0 + j) * N + n;
                    q[j] = ((valid >> j) & 1u) ? quant(__half2float(y[o]), p.scale) : 127;
                }
            } else {
                const signed char *y = reinterpret_cast<const signed char *>(p.llr);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const size_t o = (p.layout == LDPC_LAYOUT_NF) ? (size_t)n * F + f0 + j : (size_t)(f0 + j) * N + n;
                    q[j] = ((valid >> j) & 1u) ? max(-127, (int)y[o]) : 127;
                }
            }
            appw[n] = (unsigned)(q[0] + 128) | ((unsigned)(q[1] + 128) << 8) | ((unsigned)(q[2] + 128) << 16) |
                      ((unsigned)(q[3] + 128) << 24);
        }
        if (tid == 0) s_fail = 0u;
        __syncthreads();

        unsigned running = valid;  // frames not yet latched
        int it = 0;
        while (it < p.iters) {
            it++;
            for (int r = 0; r < p.J; r++) {
                const int dc = p.lt.dc[r], off = p.lt.off[r];
                for (int i = tid; i < Z; i += T) {
                    uint4 *recp = rec + (size_t)(r * Z + i) * U4;
                    if (it == 1)
                        process_row<DCMAX, true, SCALE>(smem, p, off, dc, 4 * i, Z4, recp, amax, bmul, nbias);
                    else
                        process_row<DCMAX, false, SCALE>(smem, p, off, dc, 4 * i, Z4, recp, amax, bmul, nbias);
                }
                __syncthreads();
            }
            if (p.exit_mode == LDPC_EXIT_SYNDROME || it == p.iters) {
                unsigned fail = 0u;
                for (int r = 0; r < p.J; r++) {
                    const int dc = p.lt.dc[r], off = p.lt.off[r];
                    for (int i = tid; i < Z; i += T) fail |= syndrome_row(smem, p, off, dc, 4 * i, Z4);
                }
                fail = __reduce_or_sync(0xffffffffu, fail);
                if ((tid & 31) == 0 && fail) atomicOr(&s_fail, fail);
                __syncthreads();
                const unsigned fl = s_fail;
                __syncthreads();
                if (tid == 0) s_fail = 0u;
                // okmask bit j = frame j satisfies all checks
                const unsigned okmask = (((~fl) >> 7) & 1u) | (((~fl) >> 14) & 2u) | (((~fl) >> 21) & 4u) |
                                        (((~fl) >> 28) & 8u);
                const unsigned finish = (it == p.iters) ? running : (running & okmask);
                if (finish) {
                    write_outputs(appw, p, g, finish);
                    if (p.dbg_rec) dump_records<DCMAX>(p, rec, g, finish);
                    if (tid < 4 && ((finish >> tid) & 1u)) {
                        if (p.iters_out) p.iters_out[f0 + tid] = it;
                        if (p.ok_out) p.ok_out[f0 + tid] = (okmask >> tid) & 1u;
                        if (p.out_format == LDPC_OUT_INT32_REF)
                            reinterpret_cast<int *>(p.out)[(size_t)N * F + f0 + tid] = (okmask >> tid) & 1u;
                    }
                    running &= ~finish;
                }
                if (!running) break;
            }
        }
        __syncthreads();  // everyone is done with appw before the next group's load
    }
}

template <int DCMAX>
static int launch_i8(const ldpc_code *c, LayeredParams &p, bool scale, int threads, int grid, size_t smem,
                     cudaStream_t st)
{
    if (scale) {
        LDPC_CUDA_TRY(cudaFuncSetAttribute(ldpc_layered_i8_kernel<DCMAX, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ldpc_layered_i8_kernel<DCMAX, true><<<grid, threads, smem, st>>>(p);
    } else {
        LDPC_CUDA_TRY(cudaFuncSetAttribute(ldpc_layered_i8_kernel<DCMAX, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ldpc_layered_i8_kernel<DCMAX, false><<<grid, threads, smem, st>>>(p);
    }
    LDPC_CUDA_TRY(cudaGetLastError());
    return LDPC_OK;
}

template <int DCMAX>
static int occupancy_i8(bool scale, int threads, size_t smem, int *out)
{
    if (scale) {
        LDPC_CUDA_TRY(cudaFuncSetAttribute(ldpc_layered_i8_kernel<DCMAX, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LDPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, ldpc_layered_i8_kernel<DCMAX, true>, threads,
                                                                    smem));
    } else {
        LDPC_CUDA_TRY(cudaFuncSetAttribute(ldpc_layered_i8_kernel<DCMAX, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LDPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, ldpc_layered_i8_kernel<DCMAX, false>,
                                                                    threads, smem));
    }
    return LDPC_OK;
}

struct I8Plan {
    int dcb, threads, grid, u4;
    size_t smem;
};

static int plan_i8(const ldpc_code *c, int F, bool scale, I8Plan *pl)
{
    pl->smem = (size_t)c->N * 4;
    if (pl->smem > 227 * 1024) return LDPC_ERR_UNSUPPORTED;  // N > 58112: one group does not fit one SM
    pl->dcb = c->dc_max <= 8 ? 8 : (c->dc_max <= 16 ? 16 : (c->dc_max <= 24 ? 24 : 32));
    // threads: the Z rows of a layer spread over whole warps, at most ThreadCap<DCMAX> threads
    const int cap = pl->dcb == 8 ? ThreadCap<8>::value : pl->dcb == 16 ? ThreadCap<16>::value : ThreadCap<24>::value;
    const int rows_per_thread = (c->Z + cap - 1) / cap;
    int threads = (c->Z + rows_per_thread - 1) / rows_per_thread;
    pl->threads = (threads + 31) & ~31;
    int occ = 0, rc;
    switch (pl->dcb) {
        case 8: rc = occupancy_i8<8>(scale, pl->threads, pl->smem, &occ); break;
        case 16: rc = occupancy_i8<16>(scale, pl->threads, pl->smem, &occ); break;
        case 24: rc = occupancy_i8<24>(scale, pl->threads, pl->smem, &occ); break;
        default: rc = occupancy_i8<32>(scale, pl->threads, pl->smem, &occ); break;
    }
    if (rc != LDPC_OK) return rc;
    if (occ < 1) return LDPC_ERR_UNSUPPORTED;
    const int groups = (F + 3) / 4;
    pl->grid = c->num_sms * occ;
    if (pl->grid > groups) pl->grid = groups;
    pl->u4 = pl->dcb == 8 ? RecLayout<8>::U4 : pl->dcb == 16 ? RecLayout<16>::U4 : pl->dcb == 24 ? RecLayout<24>::U4 : RecLayout<32>::U4;
    return LDPC_OK;
}

int layered_i8_scratch_bytes(const ldpc_code *c, int F, int beta_num, size_t *bytes)
{
    I8Plan pl;
    int rc = plan_i8(c, F, beta_num != 0, &pl);
    if (rc != LDPC_OK) return rc;
    *bytes = (size_t)pl.grid * c->M * pl.u4 * sizeof(uint4);
    return LDPC_OK;
}

int launch_layered_i8(const ldpc_code *c, const LayeredArgs &a, cudaStream_t st, int *launches)
{
    if (a.msg_max < 1 || a.msg_max > 127 || a.beta_num < 0 || a.beta_num > 8 || a.beta_shift < 0 ||
        a.beta_shift > 7 || (a.beta_num != 0 && a.beta_num >= (1 << a.beta_shift)))
        return LDPC_ERR_ARG;
    const bool scale = a.beta_num != 0;
    I8Plan pl;
    int rc = plan_i8(c, a.F, scale, &pl);
    if (rc != LDPC_OK) return rc;
    if ((size_t)pl.grid * c->M * pl.u4 * sizeof(uint4) > a.scratch_bytes) return LDPC_ERR_NOMEM;
    LayeredParams p;
    memset(&p, 0, sizeof(p));
    p.llr = a.llr;
    p.out = a.out;
    p.iters_out = a.iters_out;
    p.ok_out = a.ok_out;
    p.dbg_app = reinterpret_cast<signed char *>(a.dbg_app);
    p.dbg_rec = reinterpret_cast<unsigned *>(a.dbg_rec);
    p.rec = reinterpret_cast<uint4 *>(a.scratch);
    p.llr_dtype = a.llr_dtype;
    p.layout = a.layout;
    p.out_format = a.out_format;
    p.F = a.F;
    p.N = c->N;
    p.Z = c->Z;
    p.J = c->J;
    p.M = c->M;
    p.iters = a.iters;
    p.exit_mode = a.exit_mode;
    p.num_groups = (a.F + 3) / 4;
    p.scale = a.scale;
    p.msg_max = (float)a.msg_max;
    p.beta_mul = (float)a.beta_num / (float)(1 << a.beta_shift);
    p.beta_bias = 0.5f - 1.0f / (float)(2 << a.beta_shift);
    p.lt = c->lt;
    switch (pl.dcb) {
        case 8: rc = launch_i8<8>(c, p, scale, pl.threads, pl.grid, pl.smem, st); break;
        case 16: rc = launch_i8<16>(c, p, scale, pl.threads, pl.grid, pl.smem, st); break;
        case 24: rc = launch_i8<24>(c, p, scale, pl.threads, pl.grid, pl.smem, st); break;
        default: rc = launch_i8<32>(c, p, scale, pl.threads, pl.grid, pl.smem, st); break;
    }
    if (rc == LDPC_OK) *launches += 1;
    return rc;
}

int launch_layered_fp32(const ldpc_code *, const LayeredArgs &, cudaStream_t, int *) { return LDPC_ERR_UNSUPPORTED; }

}  // namespace ldpcb
