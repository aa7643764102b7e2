// Layered (row-block) normalised min-sum with fp16 state and fp16 messages — the "packed fp16" message mode
// (ldpc_decode_opts_t::msg_dtype = LDPC_DTYPE_FP16, schedule LDPC_SCHED_LAYERED).
//
// Same B200 mapping as the int8 mode (bldpc_layered.cu) with half the codewords per CTA:
//  * one CTA decodes a GROUP of 2 codewords; their APP values live in shared memory for the whole decode as one
//    half2 word per code bit (lane = codeword) — the same 4 bytes per code bit as the int8 mode's 4 biased bytes;
//  * thread = check row i of the current layer; the circulant shift is the shared-memory address (i + s) mod Z;
//  * every arithmetic step is ONE correctly rounded fp16 operation on both codewords (HADD2 / HMUL2 / HMNMX2) or a
//    bit operation on the two patterns (sign, magnitude, equality masks), so the oracle's rule — the same operations
//    on IEEE binary16, oracle/bldpc_oracle.c "fp16 layered rules" — is matched bit for bit;
//  * the check-to-variable messages are NOT stored per edge (4 bytes per edge and group would make the mode HBM-bound:
//    2.3 MB per group and iteration for J15_L30_Z1280).  A check row keeps a 32-byte record {m1', m2', idx, signs}
//    per group, and an old message is rebuilt in four instructions: HSET2.BM (k == idx), LOP3 (select m2' / m1'),
//    SHL + LOP3 (sign bit k of both codewords, stored pre-shifted so that one shift serves both lanes).
//
// Versus the int8 mode: no integer rounding of channel values, messages or APP (10-bit significand instead of
// integers up to 127) at half the codewords per instruction.  Check-node rule per B/LDPC_Decoder.cu:279-314,
// schedule / normalisation / clamps ours.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.h"
#include "philox.cuh"
#include "stress.cuh"

namespace ldpcb {

namespace {

struct F16Params {
    const void *llr;
    void *out;
    int *iters_out, *ok_out;
    __half *dbg_app;
    __half *dbg_msg;
    int dc_max;
    uint4 *rec;         // scratch: per CTA, M records of RW words
    int *work_counter;  // dynamic group scheduling
    int llr_dtype, layout, out_format;
    int F, N, Z, J, M;
    int iters, exit_mode, num_groups;
    float scale;
    float msg_max;
    float beta_c;       // 1 - beta_num / 2^beta_shift
    unsigned one;       // 1 (LDPC_F16_MAD_ADDR)
    int scale_on;
    float ch_sigma;     // fused channel
    unsigned ch_k0, ch_k1;
    unsigned long long ch_first;
    const unsigned char *ch_cw;
    // entry e = {byte offset (c*Z + s)*4 of row 0's bit, wrap threshold (Z - s)*4}
    int2 tab[kMaxBlocks];
    unsigned short off[kMaxLayers];
    unsigned char dc[kMaxLayers];
};

__device__ __forceinline__ unsigned lds32(unsigned a)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts32(unsigned a, unsigned v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned h2u(__half2 h) { return *reinterpret_cast<unsigned *>(&h); }
__device__ __forceinline__ __half2 u2h(unsigned u) { return *reinterpret_cast<__half2 *>(&u); }
__device__ __forceinline__ unsigned sel(unsigned mask, unsigned a, unsigned b) { return (a & mask) | (b & ~mask); }

__device__ __forceinline__ unsigned mad_u32(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

constexpr unsigned kSign2 = 0x80008000u;

// binary16 pattern of the integer k (0..31) in both lanes: the edge index as the record stores it
__host__ __device__ constexpr unsigned kh(int k)
{
    const int e = (k >= 2) + (k >= 4) + (k >= 8) + (k >= 16);  // floor(log2 k)
    const unsigned h = k == 0 ? 0u : (((unsigned)(e + 15) << 10) | ((unsigned)(k - (1 << e)) << (10 - e)));
    return h | (h << 16);
}
static_assert(kh(1) == 0x3C003C00u && kh(2) == 0x40004000u && kh(3) == 0x42004200u && kh(31) == 0x4FC04FC0u, "kh");

// words per record: {m1', m2', idx, signs 0-15[, signs 16-31, pad x3]}
template <int DCMAX>
struct RecLayout {
    static constexpr int RW = DCMAX <= 16 ? 4 : 8;
    static constexpr int U4 = RW / 4;
};

template <int DCMAX>
__device__ __forceinline__ void rec_load(const uint4 *p, unsigned *w)
{
    const uint4 a = p[0];
    w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w;
    if (DCMAX > 16) w[4] = reinterpret_cast<const unsigned *>(p)[4];
}
template <int DCMAX>
__device__ __forceinline__ void rec_store(uint4 *p, unsigned m1, unsigned m2, unsigned idx, unsigned A, unsigned B)
{
    p[0] = make_uint4(m1, m2, idx, A);
    if (DCMAX > 16) reinterpret_cast<unsigned *>(p)[4] = B;
}
// the message of edge k of both codewords from the row's record (run-time k: debug dump; process_row inlines the
// same four operations with a compile-time k)
__device__ __forceinline__ unsigned rec_message_rt(const unsigned *w, int k)
{
    const __half kk = __int2half_rn(k);
    const unsigned eq = __heq2_mask(u2h(w[2]), __halves2half2(kk, kk));
    const unsigned mag = sel(eq, w[1], w[0]);
    const unsigned sg = (w[k < 16 ? 3 : 4] << (k & 15)) & kSign2;
    return mag | sg;
}

// Edges whose shared-memory address is kept in a register between the two passes of a row (exact-degree paths);
// above this degree the address is rebuilt from the table (4 instructions) to stay inside the register cap.
#ifndef LDPC_F16_SAVE_ADDR_MAX
#define LDPC_F16_SAVE_ADDR_MAX 32
#endif
// 1: the address additions of the first pass are written as multiply-adds (FMA pipe: 8 % busy) instead of IADD3 (ALU
// pipe: 75 % busy, the bound) and both candidates (wrapped / not) are formed before the compare selects one
#ifndef LDPC_F16_MAD_ADDR
#define LDPC_F16_MAD_ADDR 1
#endif
// 1: a row's record is requested one row ahead (process_row); 0: at the row's first instruction.  Measured: -1.5 to
// -4 % (the five extra live registers cost more than the wait, which other warps cover)
#ifndef LDPC_F16_PRELOAD
#define LDPC_F16_PRELOAD 0
#endif

// One check row of one layer for the 2 codewords of the group.  isb = shared-memory address of row 0's word + 4 i.
// EXACT: the check degree is the template constant DC (no per-edge branches, every table entry is a constant-bank
// load at an immediate offset).  !EXACT: DC is the bucket's maximum and edges k >= dc are predicated off.
// w: this row's record, already in registers (!FIRST); nx / ld_next: the record of the thread's NEXT row is requested
// at the start of this row and replaces w at its end, so the 20 warps of a 640-thread CTA never wait for a record
// at a row's first instruction (records are thread-private: row i of every layer belongs to the same thread).
template <int DC, int DCHI, bool FIRST, bool EXACT>
__device__ __forceinline__ void process_row(unsigned isb, int i4, const F16Params &p, int off, int dc, int Z4, uint4 *recp,
                                            __half2 amaxh, __half2 betac, bool store_rec, unsigned (&w)[5],
                                            const uint4 *nx, bool ld_next)
{
    constexpr bool kSaveAddr = EXACT && DC <= LDPC_F16_SAVE_ADDR_MAX;
    unsigned t[DC], addr[kSaveAddr ? DC : 1], wn[5];
    if (ld_next) rec_load<DCHI>(nx, wn);
    __half2 m1 = u2h(0x7C007C00u), m2 = u2h(0x7C007C00u), held = u2h(0u);  // +inf
    unsigned par = 0u, A = 0u, B = 0u;
    const unsigned dw = FIRST ? 0u : (w[0] ^ w[1]);
    const unsigned one = p.one;  // a run-time 1: a literal would be folded back into an add
#pragma unroll
    for (int k = 0; k < DC; k++) {
        if (EXACT || k < dc) {
            const int2 e = p.tab[off + k];
            unsigned a;
            if (LDPC_F16_MAD_ADDR) {
                const unsigned a0 = mad_u32((unsigned)e.x, one, isb), a1 = mad_u32((unsigned)e.x, one, isb - (unsigned)Z4);
                a = (i4 >= e.y) ? a1 : a0;
            } else {
                a = isb + (unsigned)e.x;
                a -= (i4 >= e.y) ? (unsigned)Z4 : 0u;
            }
            if (kSaveAddr) addr[k] = a;
            const unsigned ap = lds32(a);
            unsigned tk = ap;
            if (!FIRST) {  // the old message from the record: k == idx ? m2' : m1', sign bit k of both codewords.
                // Two LOP3: m1' ^ (eq & (m1' ^ m2')), then | (signs << k & mask) — written as a select it took three
                const unsigned eq = __heq2_mask(u2h(w[2]), u2h(kh(k)));
                const unsigned mag = w[0] ^ (eq & dw);
                const unsigned sg = (w[k < 16 ? 3 : 4] << (k & 15)) & kSign2;
                tk = h2u(__hsub2(u2h(ap), u2h(mag | sg)));
            }
            t[k] = tk;
            par ^= tk;
            // two smallest magnitudes with multiplicity.  HMNMX2 takes |t| as an operand modifier, so the magnitude is
            // never materialised (the integer min/max on the patterns would need it in a register: one more LOP3 per
            // edge on the ALU pipe, which bounds this kernel — profiles/r02_ncu_layered_f16_C2.txt).  Exact paths take
            // the edges in pairs: 5 min/max per 2 edges instead of 6
            const __half2 ak = __habs2(u2h(tk));
            if (EXACT && (k & 1)) {
                const __half2 lo = __hmin2(held, ak), hi = __hmax2(held, ak);
                m2 = __hmin2(__hmin2(m2, hi), __hmax2(m1, lo));
                m1 = __hmin2(m1, lo);
            } else if (EXACT && k + 1 < DC) {
                held = ak;
            } else {
                m2 = __hmin2(m2, __hmax2(m1, ak));
                m1 = __hmin2(m1, ak);
            }
            if (k < 16)
                A |= (tk & kSign2) >> (k & 15);
            else
                B |= (tk & kSign2) >> (k & 15);
        }
    }
    // m' = min(m, amax) * c
    __half2 m1s = __hmin2(m1, amaxh), m2s = __hmin2(m2, amaxh);
    if (p.scale_on) {
        m1s = __hmul2(m1s, betac);
        m2s = __hmul2(m2s, betac);
    }
    const unsigned pm = par & kSign2;
    const unsigned m1x = h2u(m1s) ^ pm, dx = h2u(m1s) ^ h2u(m2s);  // m2x = m1x ^ dx
    unsigned idx = 0u;
#pragma unroll
    for (int k = 0; k < DC; k++) {
        if (EXACT || k < dc) {
            unsigned a;
            if (kSaveAddr) {
                a = addr[k];
            } else {
                const int2 e = p.tab[off + k];
                a = isb + (unsigned)e.x;
                a -= (i4 >= e.y) ? (unsigned)Z4 : 0u;
            }
            const unsigned tk = t[k];
            const unsigned eq = __heq2_mask(__habs2(u2h(tk)), m1);  // |t_k| == min1 (magnitudes: no NaN, no -0)
            idx = sel(eq, kh(k), idx);  // any index of a tied minimum serves: ties have m2' == m1'
            const unsigned nw = (m1x ^ (eq & dx)) ^ (tk & kSign2);  // two LOP3
            // no APP clamp: |APP| <= |channel| + dv * msg_max (+ rounding drift) <= 127 + 16 * 127, far inside binary16
            sts32(a, h2u(__hadd2(u2h(tk), u2h(nw))));
        }
    }
    if (store_rec) {
        const unsigned flip = (pm >> 15) * 0xFFFFu;  // a lane's parity flips all of its sign bits
        rec_store<DCHI>(recp, h2u(m1s), h2u(m2s), idx, A ^ flip, B ^ flip);
    }
    if (ld_next) {
#pragma unroll
        for (int j = 0; j < 5; j++) w[j] = wn[j];
    }
}

// Rows of a layer whose degree is outside the exact range of the kernel's bucket: predicated, out of line so that
// its register needs do not leak into the allocation of the exact-degree paths.
template <int DCHI, bool FIRST>
__device__ __noinline__ void generic_rows(unsigned sbase, const F16Params &p, int off, int dc, uint4 *recl, __half2 amaxh,
                                          __half2 betac, bool store_rec)
{
    const int Z = p.Z, Z4 = 4 * Z, T = blockDim.x;
    for (int i = threadIdx.x; i < Z; i += T) {
        unsigned w[5];  // out of line: this path loads its records itself
        if (!FIRST) rec_load<DCHI>(recl + (size_t)i * RecLayout<DCHI>::U4, w);
        process_row<DCHI, DCHI, FIRST, false>(sbase + 4u * i, 4 * i, p, off, dc, Z4, recl + (size_t)i * RecLayout<DCHI>::U4,
                                              amaxh, betac, store_rec, w, nullptr, false);
    }
}

// One full iteration: all layers in order, the Z rows of a layer spread over the CTA.  Six exact degrees per bucket
// (DCHI .. DCHI - 5), like the int8 mode.
template <int DCHI, bool FIRST>
__device__ __forceinline__ void sweep_layers(unsigned sbase, const F16Params &p, uint4 *rec, __half2 amaxh, __half2 betac,
                                             bool store_rec)
{
    constexpr int U4 = RecLayout<DCHI>::U4;
    const int Z = p.Z, Z4 = 4 * Z, T = blockDim.x, tid = threadIdx.x;
    unsigned w[5];
    if (LDPC_F16_PRELOAD && !FIRST && tid < Z)
        rec_load<DCHI>(rec + (size_t)tid * U4, w);  // the one record per sweep that is waited for
    for (int r = 0; r < p.J; r++) {
        const int dc = p.dc[r], off = p.off[r];
        uint4 *recl = rec + (size_t)r * Z * U4;
        const uint4 *nxl = rec + (size_t)(r + 1) * Z * U4;  // this thread's first row of the next layer
        const bool more_layers = r + 1 < p.J;
#define LDPC_ROWS(DCX)                                                                                               \
    for (int i = tid; i < Z; i += T) {                                                                               \
        const bool same_layer = i + T < Z;                                                                           \
        const uint4 *nx = same_layer ? recl + (size_t)(i + T) * U4 : nxl + (size_t)tid * U4;                         \
        if (!LDPC_F16_PRELOAD && !FIRST) rec_load<DCHI>(recl + (size_t)i * U4, w);                                   \
        process_row<DCX, DCHI, FIRST, true>(sbase + 4u * i, 4 * i, p, off, DCX, Z4, recl + (size_t)i * U4, amaxh, betac, \
                                            store_rec, w, nx,                                                        \
                                            LDPC_F16_PRELOAD && !FIRST && (same_layer || more_layers));              \
    }
        if (dc == DCHI) {
            LDPC_ROWS(DCHI)
        } else if (DCHI >= 2 && dc == DCHI - 1) {
            LDPC_ROWS((DCHI >= 2 ? DCHI - 1 : 1))
        } else if (DCHI >= 3 && dc == DCHI - 2) {
            LDPC_ROWS((DCHI >= 3 ? DCHI - 2 : 1))
        } else if (DCHI >= 4 && dc == DCHI - 3) {
            LDPC_ROWS((DCHI >= 4 ? DCHI - 3 : 1))
        } else if (DCHI >= 5 && dc == DCHI - 4) {
            LDPC_ROWS((DCHI >= 5 ? DCHI - 4 : 1))
        } else if (DCHI >= 6 && dc == DCHI - 5) {
            LDPC_ROWS((DCHI >= 6 ? DCHI - 5 : 1))
        } else {
            generic_rows<DCHI, FIRST>(sbase, p, off, dc, recl, amaxh, betac, store_rec);
            if (LDPC_F16_PRELOAD && !FIRST && more_layers && tid < Z) rec_load<DCHI>(nxl + (size_t)tid * U4, w);
        }
#undef LDPC_ROWS
        __syncthreads();  // rows of one layer touch disjoint code bits; the next layer reads what this one wrote
        LDPC_STRESS_POINT(1);
    }
}

// f16(clamp(y * scale, -127, 127)); NaN -> -127 (fmaxf returns the other operand)
__device__ __forceinline__ __half quant_h(float y, float scale)
{
    return __float2half_rn(fminf(fmaxf(__fmul_rn(y, scale), -127.0f), 127.0f));
}

__device__ __forceinline__ float load_value(const F16Params &p, int n, int f)
{
    const size_t o = (p.layout == LDPC_LAYOUT_NF) ? (size_t)n * p.F + f : (size_t)f * p.N + n;
    if (p.llr_dtype == LDPC_DTYPE_FP32) return __ldg(reinterpret_cast<const float *>(p.llr) + o);
    if (p.llr_dtype == LDPC_DTYPE_FP16) return __half2float(reinterpret_cast<const __half *>(p.llr)[o]);
    return (float)reinterpret_cast<const signed char *>(p.llr)[o];
}

// hard decisions (sign bit) of the frames selected by fmask (bit j = frame 2g + j) -> global memory
__device__ void write_outputs(const unsigned *appw, const F16Params &p, int g, unsigned fmask)
{
    const int F = p.F, N = p.N, f0 = 2 * g;
    const bool per_bit_pass = p.out_format != LDPC_OUT_BITPACK || p.dbg_app != nullptr;
    for (int n = threadIdx.x; per_bit_pass && n < N; n += blockDim.x) {
        const unsigned w = appw[n];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            if (!((fmask >> j) & 1u)) continue;
            const unsigned bit = (w >> (16 * j + 15)) & 1u;
            if (p.out_format == LDPC_OUT_INT32_REF)
                reinterpret_cast<int *>(p.out)[(size_t)n * F + f0 + j] = (int)bit;
            else if (p.out_format == LDPC_OUT_U8) {
                const size_t o = (p.layout == LDPC_LAYOUT_NF) ? (size_t)n * F + f0 + j : (size_t)(f0 + j) * N + n;
                reinterpret_cast<unsigned char *>(p.out)[o] = (unsigned char)bit;
            }
            if (p.dbg_app) reinterpret_cast<unsigned short *>(p.dbg_app)[(size_t)n * F + f0 + j] = (unsigned short)(w >> (16 * j));
        }
    }
    if (p.out_format == LDPC_OUT_BITPACK) {
        unsigned *D = reinterpret_cast<unsigned *>(p.out);
        const int W = (N + 31) / 32, lane = threadIdx.x & 31;
        const int nround = (N + 31) & ~31;
        for (int n = threadIdx.x; n < nround; n += blockDim.x) {  // blockDim is a multiple of 32
            const unsigned w = (n < N) ? appw[n] : 0u;
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const unsigned b = __ballot_sync(0xffffffffu, ((w >> (16 * j + 15)) & 1u) != 0u);
                if (lane == 0 && ((fmask >> j) & 1u)) D[(size_t)(f0 + j) * W + (n >> 5)] = b;
            }
        }
    }
}

// debug: the c2v messages of frames selected by fmask -> binary16 [M][dc_max][F] (0 for absent edges)
template <int DCMAX>
__device__ void dump_messages(const F16Params &p, const uint4 *rec, int g, unsigned fmask)
{
    unsigned short *out = reinterpret_cast<unsigned short *>(p.dbg_msg);
    for (int m = threadIdx.x; m < p.M; m += blockDim.x) {
        unsigned w[5] = {0u, 0u, 0u, 0u, 0u};
        rec_load<DCMAX>(rec + (size_t)m * RecLayout<DCMAX>::U4, w);
        const int dc = p.dc[m / p.Z];
        for (int k = 0; k < p.dc_max; k++) {
            const unsigned v = (k < dc) ? rec_message_rt(w, k) : 0u;
            for (int j = 0; j < 2; j++)
                if ((fmask >> j) & 1u) out[((size_t)m * p.dc_max + k) * p.F + 2 * g + j] = (unsigned short)(v >> (16 * j));
        }
    }
}

// CTA width cap = register cap of the kernel (640 threads -> 96 registers, 512 -> 128, 384 -> 168) of the buckets above
// 16, whose per-row arrays (t[] and addr[], 2 * DC registers) want more than 96.  Measured (profiles/r02_fp16_mode.txt):
// 128 registers PON +0.7 %, J10_L60_Z160 -9 % (fewer resident CTAs); 168 registers PON -12 %: 640 stays.
#ifndef LDPC_F16_CAP_HI
#define LDPC_F16_CAP_HI 640
#endif
template <int DCMAX>
struct F16ThreadCap {
    static constexpr int value = DCMAX > 16 ? LDPC_F16_CAP_HI : 640;
};

template <int DCMAX>
__global__ void __launch_bounds__(F16ThreadCap<DCMAX>::value, 1) ldpc_layered_f16_kernel(const __grid_constant__ F16Params p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned *appw = reinterpret_cast<unsigned *>(smem);
    const int tid = threadIdx.x, T = blockDim.x;
    const int N = p.N, Z = p.Z, F = p.F, Z4 = 4 * Z;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    uint4 *rec = p.rec + (size_t)blockIdx.x * p.M * RecLayout<DCMAX>::U4;
    const __half2 amaxh = __float2half2_rn(p.msg_max);
    const __half2 betac = __float2half2_rn(p.beta_c);
    __shared__ int s_next;
    for (;;) {
        if (tid == 0) s_next = atomicAdd(p.work_counter, 1);
        __syncthreads();
        LDPC_STRESS_POINT(6);
        const int g = s_next;
        if (g >= p.num_groups) break;
        const int f0 = 2 * g;
        const unsigned valid = 1u | ((f0 + 1 < F) ? 2u : 0u);
        const unsigned pad = h2u(__float2half2_rn(127.0f)) & 0xFFFFu;  // absent frame: every bit a confident 0
        if (p.llr_dtype == LDPC_DTYPE_CHANNEL) {
            // fused channel: y = 1 - 2c + sigma * n generated here (one Philox call = 4 consecutive bits of a frame)
            for (int nb = tid; nb < (N + 3) / 4; nb += T) {
                float gs[2][4];
#pragma unroll
                for (int j = 0; j < 2; j++) awgn_normals4(p.ch_first + (unsigned long long)(f0 + j), nb, p.ch_k0, p.ch_k1, gs[j]);
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int n = 4 * nb + b;
                    if (n >= N) break;
                    const int bit = p.ch_cw ? (p.ch_cw[n] & 1) : 0;
                    unsigned w = 0u;
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const unsigned h = ((valid >> j) & 1u)
                                               ? (unsigned)__half_as_ushort(quant_h(awgn_bpsk_sample(bit, p.ch_sigma, gs[j][b]), p.scale))
                                               : pad;
                        w |= h << (16 * j);
                    }
                    appw[n] = w;
                }
            }
        } else if (p.llr_dtype == LDPC_DTYPE_FP32 && p.layout == LDPC_LAYOUT_NF && valid == 3u && (F & 1) == 0) {
            // fp32 [N][F], whole group: one aligned 8-byte vector per code bit, 8 loads in flight per thread
            constexpr int kDepth = 8;
            const float2 *y2 = reinterpret_cast<const float2 *>(reinterpret_cast<const float *>(p.llr) + f0);
            const size_t st2 = (size_t)(F >> 1);
            for (int n0 = tid; n0 < N; n0 += kDepth * T) {
                float2 v[kDepth];
#pragma unroll
                for (int u = 0; u < kDepth; u++) v[u] = __ldg(y2 + (size_t)min(n0 + u * T, N - 1) * st2);
#pragma unroll
                for (int u = 0; u < kDepth; u++) {
                    const int n = n0 + u * T;
                    if (n < N) appw[n] = h2u(__halves2half2(quant_h(v[u].x, p.scale), quant_h(v[u].y, p.scale)));
                }
            }
        } else {
            for (int n = tid; n < N; n += T) {
                unsigned w = 0u;
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const unsigned h = ((valid >> j) & 1u)
                                           ? (unsigned)__half_as_ushort(quant_h(load_value(p, n, f0 + j), p.scale))
                                           : pad;
                    w |= h << (16 * j);
                }
                appw[n] = w;
            }
        }
        __syncthreads();
        LDPC_STRESS_POINT(5);

        unsigned running = valid;
        int it = 0;
        while (it < p.iters) {
            it++;
            // the records of the last fixed iteration are dead unless somebody dumps them
            const bool store_rec = !(it == p.iters && p.exit_mode != LDPC_EXIT_SYNDROME && p.dbg_msg == nullptr);
            if (it == 1)
                sweep_layers<DCMAX, true>(sbase, p, rec, amaxh, betac, store_rec);
            else
                sweep_layers<DCMAX, false>(sbase, p, rec, amaxh, betac, store_rec);
            if (p.exit_mode == LDPC_EXIT_SYNDROME || it == p.iters) {
                // XOR of the sign bits of every check row; the reduction needs no shared flag word (the int8 mode's
                // ping-pong, bldpc_layered.cu): __syncthreads_or is barrier and reduction in one
                unsigned fail = 0u;
                for (int r = 0; r < p.J; r++) {
                    const int dc = p.dc[r], off = p.off[r];
                    for (int i = tid; i < Z; i += T) {
                        unsigned x = 0u;
#pragma unroll 4
                        for (int k = 0; k < dc; k++) {
                            const int2 e = p.tab[off + k];
                            unsigned a = sbase + 4u * i + (unsigned)e.x;
                            a -= (4 * i >= e.y) ? (unsigned)Z4 : 0u;
                            x ^= lds32(a);
                        }
                        fail |= x & kSign2;
                    }
                }
                const unsigned bad0 = __syncthreads_or((int)(fail & 0x8000u)) ? 1u : 0u;
                LDPC_STRESS_POINT(2);
                const unsigned bad1 = __syncthreads_or((int)(fail >> 31)) ? 1u : 0u;
                const unsigned okmask = (bad0 ^ 1u) | ((bad1 ^ 1u) << 1);
                const unsigned finish = (it == p.iters) ? running : (running & okmask);
                if (finish) {
                    LDPC_STRESS_POINT(3);
                    write_outputs(appw, p, g, finish);
                    if (p.dbg_msg) dump_messages<DCMAX>(p, rec, g, finish);
                    if (tid < 2 && ((finish >> tid) & 1u)) {
                        if (p.iters_out) p.iters_out[f0 + tid] = it;
                        if (p.ok_out) p.ok_out[f0 + tid] = (okmask >> tid) & 1u;
                        if (p.out_format == LDPC_OUT_INT32_REF)
                            reinterpret_cast<int *>(p.out)[(size_t)N * F + f0 + tid] = (okmask >> tid) & 1u;
                    }
                    running &= ~finish;
                    // the frame that keeps running goes straight into the next sweep, which rewrites the APP words and
                    // records that slower warps are still emitting for the latched frame
                    if (running) __syncthreads();
                }
                if (!running) break;
            }
        }
        __syncthreads();  // everyone is done with appw and s_next before the next group
        LDPC_STRESS_POINT(4);
    }
}

template <int DCMAX>
struct F16Kernel {
    static int occupancy(int threads, size_t smem, int *out)
    {
        LDPC_CUDA_TRY(cudaFuncSetAttribute(ldpc_layered_f16_kernel<DCMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        LDPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, ldpc_layered_f16_kernel<DCMAX>, threads, smem));
        return LDPC_OK;
    }
    static int launch(const F16Params &p, int threads, int grid, size_t smem, cudaStream_t st)
    {
        ldpc_layered_f16_kernel<DCMAX><<<grid, threads, smem, st>>>(p);
        LDPC_CUDA_TRY(cudaGetLastError());
        return LDPC_OK;
    }
};

#define LDPC_F16_BUCKETS(X) X(4) X(8) X(12) X(16) X(20) X(24) X(28) X(32)

struct F16Plan {
    int dcb, threads, grid, u4;
    size_t smem;
};

int plan_f16(const ldpc_code *c, int F, F16Plan *pl)
{
    pl->smem = (size_t)c->N * 4;
    if (pl->smem > 227 * 1024) return LDPC_ERR_UNSUPPORTED;  // N > 58112: one group does not fit one SM
    pl->dcb = (c->dc_max + 3) & ~3;
    int occ = 0, rc = LDPC_ERR_UNSUPPORTED;
    switch (pl->dcb) {
#define X(D) case D: pl->u4 = RecLayout<D>::U4; break;
        LDPC_F16_BUCKETS(X)
#undef X
        default: return LDPC_ERR_UNSUPPORTED;
    }
    int cap = 640;
    switch (pl->dcb) {
#define X(D) case D: cap = F16ThreadCap<D>::value; break;
        LDPC_F16_BUCKETS(X)
#undef X
    }
    const int rows_per_thread = (c->Z + cap - 1) / cap;
    const int threads = (c->Z + rows_per_thread - 1) / rows_per_thread;
    pl->threads = (threads + 31) & ~31;
    switch (pl->dcb) {
#define X(D) case D: rc = F16Kernel<D>::occupancy(pl->threads, pl->smem, &occ); break;
        LDPC_F16_BUCKETS(X)
#undef X
    }
    if (rc != LDPC_OK) return rc;
    if (occ < 1) return LDPC_ERR_UNSUPPORTED;
    const int groups = (F + 1) / 2;
    pl->grid = c->num_sms * occ;
    if (pl->grid > groups) pl->grid = groups;
    return LDPC_OK;
}

}  // namespace

int layered_f16_wave_frames(const ldpc_code *c, int *frames)
{
    F16Plan pl;
    int rc = plan_f16(c, 1 << 30, &pl);
    if (rc != LDPC_OK) return rc;
    *frames = pl.grid * 2;  // resident CTAs x the 2 codewords of a group
    return LDPC_OK;
}

int layered_f16_scratch_bytes(const ldpc_code *c, int F, size_t *bytes)
{
    F16Plan pl;
    int rc = plan_f16(c, F, &pl);
    if (rc != LDPC_OK) return rc;
    *bytes = (size_t)pl.grid * c->M * pl.u4 * sizeof(uint4) + 256;  // + the work counter
    return LDPC_OK;
}

int launch_layered_f16(const ldpc_code *c, const LayeredArgs &a, cudaStream_t st, int *launches)
{
    if (a.msg_max < 1 || a.msg_max > 127 || a.beta_num < 0 || a.beta_num > 8 || a.beta_shift < 0 || a.beta_shift > 7 ||
        (a.beta_num != 0 && a.beta_num >= (1 << a.beta_shift)))
        return LDPC_ERR_ARG;
    F16Plan pl;
    int rc = plan_f16(c, a.F, &pl);
    if (rc != LDPC_OK) return rc;
    const size_t rec_bytes = (size_t)pl.grid * c->M * pl.u4 * sizeof(uint4);
    if (rec_bytes + 256 > a.scratch_bytes) return LDPC_ERR_NOMEM;
    F16Params p;
    memset(&p, 0, sizeof(p));
    p.llr = a.llr;
    p.out = a.out;
    p.iters_out = a.iters_out;
    p.ok_out = a.ok_out;
    p.dbg_app = reinterpret_cast<__half *>(a.dbg_app);
    p.dbg_msg = reinterpret_cast<__half *>(a.dbg_rec);
    p.dc_max = c->dc_max;
    p.rec = reinterpret_cast<uint4 *>(a.scratch);
    p.work_counter = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(a.scratch) + rec_bytes);
    LDPC_CUDA_TRY(cudaMemsetAsync(p.work_counter, 0, sizeof(int), st));
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
    p.num_groups = (a.F + 1) / 2;
    p.scale = a.scale;
    p.msg_max = (float)a.msg_max;
    p.beta_c = 1.0f - (float)a.beta_num / (float)(1 << a.beta_shift);
    p.scale_on = a.beta_num != 0;
    p.one = 1u;
    p.ch_sigma = a.ch_sigma;
    p.ch_k0 = (unsigned)a.ch_seed;
    p.ch_k1 = (unsigned)(a.ch_seed >> 32);
    p.ch_first = a.ch_first;
    p.ch_cw = a.ch_cw;
    for (int r = 0; r < c->J; r++) {
        p.off[r] = c->lt.off[r];
        p.dc[r] = c->lt.dc[r];
        for (int k = 0; k < c->lt.dc[r]; k++) {
            const int e = c->lt.off[r] + k;
            p.tab[e] = make_int2(((int)c->lt.col[e] * c->Z + (int)c->lt.shift[e]) * 4, (c->Z - (int)c->lt.shift[e]) * 4);
        }
    }
    switch (pl.dcb) {
#define X(D) case D: rc = F16Kernel<D>::launch(p, pl.threads, pl.grid, pl.smem, st); break;
        LDPC_F16_BUCKETS(X)
#undef X
    }
    if (rc == LDPC_OK) *launches += 1;
    return rc;
}

}  // namespace ldpcb
