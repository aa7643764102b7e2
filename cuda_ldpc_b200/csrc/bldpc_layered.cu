// Layered (row-block) normalised min-sum with int8 state — the throughput mode.
//
// B200 mapping:
//  * one CTA decodes a GROUP of 4 codewords at a time; their APP values live in shared memory
//    for the whole decode as one 32-bit word per code bit (4 x biased uint8);
//  * thread = check row i of the current layer (block row of H); the circulant shift is the
//    shared-memory address (i + s) mod Z, consecutive threads hit consecutive banks;
//  * arithmetic is 2-way SIMD on half2 lanes holding small integers (exact in fp16 below 2048):
//    HADD2 / HFMA2 / HSET2 / PRMT / LOP3 plus the 16x2 integer min/max family on fp16 bit patterns
//    (VIMNMX.U16x2, VIMNMX3.U16x2, VIADDMNMX.S16x2.RELU) — sm_100a has no native 8x4 SIMD integer
//    min/compare (the __v*4 intrinsics expand to 5-10 instructions, profiles/r01_simd_sass.txt);
//  * the check-to-variable messages are one byte (c2v + 127) per edge and codeword = one 32-bit word per
//    edge and group, 256-bit blocks per check row in a per-CTA slice of a scratch buffer (L2 evict-last;
//    for J15_L30_Z1280 the 148 x 19200 x 32 B = 91 MB do not stay in L2 beside the streaming input, about
//    half of the reads come from HBM — ~2 MB of DRAM traffic per frame, 30 % of the HBM peak); each thread
//    loads the block of its NEXT row into registers while it finishes the current one (dc <= 12), which
//    hides that latency.  Round 1 stored compressed records {min1, min2, idx, sign bits} instead and paid
//    five instructions per edge to rebuild every old message (process_row_x's comment);
//  * the syndrome (early exit / ok flag) is a cheap extra pass: XOR of the sign bits.
//
// Update rules: exactly oracle/bldpc_oracle.c "int8 layered rules" (t = APP - c2v_old,
// min1/min2/first-index on min(|t|, msg_max), m' = m - ((m*beta_num) >> beta_shift), APP =
// sat127(t + c2v_new)); check-node rule per B/LDPC_Decoder.cu:279-314, schedule ours.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.h"
#include "philox.cuh"
#include "stress.cuh"

// Tuning switches; the defaults are the winners of A/B runs on B200 (tools/ab, J15_L30_Z1280, 10 it):
//   LDPC_LOAD_DEPTH    8  channel-value vectors in flight per thread in the load phase
//   LDPC_L2_PREFETCH   0  L2 prefetch of the CTA's next group of channel values: 1 = at the start of this group's
//                         decode (evicts messages, +4 % time), 2 = before the LAST iteration, when the messages are
//                         dead (+6 %: 60 prefetch instructions per thread and group cost more than the wait they save)
//   LDPC_SKIP_LAST_ST  1  the last iteration of a fixed-iteration decode does not store its messages (nobody reads
//                         them): J15_L30_Z1280 7.44 -> 7.37 ms; only for buckets <= 12 — in the register-bound
//                         high-degree buckets the extra predicate costs more (PON 5.81 -> 6.23 ms)
#ifndef LDPC_L2_PREFETCH
#define LDPC_L2_PREFETCH 0
#endif
#ifndef LDPC_SKIP_LAST_ST
#define LDPC_SKIP_LAST_ST 1
#endif
//   LDPC_HI_PRELOAD    0  degree buckets above 12 load a row's old messages at the start of the row instead of one
//                         step ahead: 16-32 fewer live registers in phase 2 (B200, PON: 39.4 -> 41.6 Gbit/s)
#ifndef LDPC_HI_PRELOAD
#define LDPC_HI_PRELOAD 0
#endif
#ifndef LDPC_LOAD_DEPTH
#define LDPC_LOAD_DEPTH 8
#endif

// Developer build (make lib VARIANT=_phase EXTRA=-DLDPC_PHASE_TIMERS=1; tools/ab/phases.py): thread 0 of every CTA
// accumulates clock64() deltas per phase (0 load + quantise, 1 first sweep, 2 later sweeps, 3 syndrome pass, 4 outputs).
#ifndef LDPC_PHASE_TIMERS
#define LDPC_PHASE_TIMERS 0
#endif
#if LDPC_PHASE_TIMERS
__device__ unsigned long long g_phase_cycles[8];
extern "C" void ldpcb_debug_phase_cycles(unsigned long long *out, int reset)
{
    cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * 8);
    if (reset) {
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z);
    }
}
#define LDPC_PHASE_MARK(k)                                                            \
    do {                                                                              \
        if (threadIdx.x == 0) {                                                       \
            const long long now_ = clock64();                                         \
            atomicAdd(&g_phase_cycles[k], (unsigned long long)(now_ - phase_t0_));    \
            phase_t0_ = now_;                                                         \
        }                                                                             \
    } while (0)
#else
#define LDPC_PHASE_MARK(k) ((void)0)
#endif

namespace ldpcb {

struct LayeredParams {
    const void *llr;
    void *out;
    int *iters_out, *ok_out;
    signed char *dbg_app;
    signed char *dbg_msg;
    int dc_max;
    uint4 *rec;             // scratch: per CTA, M message blocks of MsgLayout::STRIDE uint4
    int *work_counter;      // dynamic group scheduling (early exit), or nullptr
    int llr_dtype, layout, out_format;
    int F, N, Z, J, M;
    int iters, exit_mode, num_groups;
    float scale;
    float msg_max;
    float beta_mul;         // beta_num / 2^beta_shift
    float beta_bias;        // 0.5 - 2^-(beta_shift+1)
    float ch_sigma;         // fused channel: sigma, Philox key, first global frame, codeword bits
    unsigned ch_k0, ch_k1;
    unsigned long long ch_first;
    const unsigned char *ch_cw;
    int scale_on;           // beta_num != 0
    unsigned c65;           // 0x65656565: kept in a register so PRMT can take the selector as its immediate
    unsigned c9b;           // 0x9B009B00 = -0x6500 per lane: VIADDMNMX's addend (as a literal it is re-materialised
                            // by a move before every use: +2.6 % on B200)
    // kernel-ready layer tables: entry e = {byte offset (c*Z + s)*4 of row 0's bit, wrap threshold (Z - s)*4}
    int2 tab[kMaxBlocks];
    unsigned short off[kMaxLayers];
    unsigned char dc[kMaxLayers];
    LayerTables lt;         // compact form, used by the debug dump only
};

// Check-to-variable messages: per check row a block of CH x 8 words, word k = the c2v bytes (message + 127) of
// edge k for the 4 codewords of the group.  256-bit global accesses with the L2 evict-last policy
// (LDG.E.ELL2.256 / STG.E.ELL2.256 on sm_100a) so that the per-CTA message slices are the last thing L2 gives
// up while the channel values stream through.
template <int CH>
__device__ __forceinline__ void msg_load(const uint4 *p, unsigned *w)
{
#pragma unroll
    for (int c = 0; c < CH; c++)
        asm volatile("ld.global.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(w[8 * c + 0]), "=r"(w[8 * c + 1]), "=r"(w[8 * c + 2]), "=r"(w[8 * c + 3]),
                       "=r"(w[8 * c + 4]), "=r"(w[8 * c + 5]), "=r"(w[8 * c + 6]), "=r"(w[8 * c + 7])
                     : "l"(p + 2 * c));
}
__device__ __forceinline__ void msg_store8(uint4 *p, const unsigned *w)
{
    asm volatile("st.global.L2::evict_last.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]),
                 "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

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
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned s)
{
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}
__device__ __forceinline__ unsigned sel(unsigned mask, unsigned a, unsigned b) { return (a & mask) | (b & ~mask); }

// q = sat127(rint(y * scale)) without a float -> int conversion (F2I runs on the quarter-rate XU pipe): clamp in float
// (the bounds are integers, so clamping commutes with the rounding; NaN -> -127 like the oracle's (int)rintf), then
// add 1.5 * 2^23 — the sum's ulp is 1, so the FADD itself rounds to nearest-even and the low mantissa byte is q in two's
// complement.
__device__ __forceinline__ unsigned quant_bits(float y, float scale)
{
    const float v = fminf(fmaxf(__fmul_rn(y, scale), -127.0f), 127.0f);
    return __float_as_uint(__fadd_rn(v, 12582912.0f));
}
__device__ __forceinline__ int quant(float y, float scale)
{
    return (int)(signed char)(quant_bits(y, scale) & 0xFFu);
}
// four values -> the biased APP word (byte j = q_j + 127): two's-complement bytes ^ 0x80 = q + 128, minus 1 per byte
// (no borrow: every byte is >= 1)
__device__ __forceinline__ unsigned quant4(float a, float b, float c, float d, float scale)
{
    const unsigned lo = prmt(quant_bits(a, scale), quant_bits(b, scale), 0x0040u);
    const unsigned hi = prmt(quant_bits(c, scale), quant_bits(d, scale), 0x0040u);
    return (prmt(lo, hi, 0x5410u) ^ 0x80808080u) - 0x01010101u;
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

__device__ __forceinline__ __half2 umin2(__half2 a, __half2 b) { return u2h(__vminu2(h2u(a), h2u(b))); }
__device__ __forceinline__ __half2 umax2(__half2 a, __half2 b) { return u2h(__vmaxu2(h2u(a), h2u(b))); }
__device__ __forceinline__ __half2 umin3(__half2 a, __half2 b, __half2 c)
{
    return u2h(__vimin3_u16x2(h2u(a), h2u(b), h2u(c)));
}

// Message block of a check row in the kernel's degree bucket
template <int DCMAX>
struct MsgLayout {
    static constexpr int CH = (DCMAX + 7) / 8;  // 256-bit chunks per row
    static constexpr int WORDS = CH * 8;
    static constexpr int STRIDE = CH * 2;        // uint4 between rows: every row 32-byte aligned
};

// One check row of one layer for the 4 codewords of the group.
// EXACT: the check degree is the template constant DC (no per-edge branches, every table entry is
// a constant-bank load at an immediate offset).  !EXACT: DC is the bucket's maximum and edges k >= dc are
// predicated off (rare degrees only).
//
// Number formats (all exact):
//  * smem byte b = APP + 127 (0..254) and message byte c = c2v + 127.  PRMT with the constant byte 0x65 turns a
//    byte into the fp16 pattern 0x6500|b = 1280 + b = value + 1407, so t = APP - c2v_old is ONE HADD2 of the two
//    patterns (the biases cancel), two frames per instruction.
//  * min1 / min2 are tracked on the patterns of |t| with VIMNMX(3).U16x2 (positive fp16 patterns order like
//    unsigned integers): 2.5 min/max per edge; the amax clamp commutes with the minimum and is applied once per row.
//  * The oracle's "k == first index of the minimum" needs no index at all for the NEW message: |t_k| <= min1 holds
//    for the first minimum and otherwise only when min2 == min1, where both choices give the same value.  The
//    select is fma.sat(|t_k|, -1, min1 + 1) in {0, 1} on the FMA pipe, then mag = sel * (m2' - m1') + m1'.
//  * sign of the new message = parity ^ sign(t_k): parity = xor of the t patterns (bit 15 of a lane); +-1.0 with
//    that sign is one LOP3, the signed biased message nb = mag * (+-1) + 1407 one HFMA2, nb + t = APP_new + 1407
//    stays inside the binade [1024, 2048) where the pattern is 0x6400 + (value - 1024): "clamp to [-127, 127],
//    re-bias" is ONE integer SIMD instruction, VIADDMNMX.S16x2.RELU: max(min(pattern - 0x6500, 254), 0).
//    The low byte of nb's pattern IS the message byte to store.
// Round 1 kept compressed records {min1, min2, idx, sign bits} per check instead (same bytes for dc = 8, fewer for
// dc >= 12) and rebuilt the old message of every edge from them: 5 more instructions per edge and pair of frames
// in phase 1, plus sign collection and index tracking (keyed minima, decode) — 57 instead of ~42 instructions per
// edge and 4 codewords.
//
// Message flow: `mw` holds this row's old messages on entry (!FIRST) and those of the thread's NEXT step on exit
// (`ld_next`: the load is issued between the two phases, so its latency hides under phase 2).
template <int DC, int DCHI, bool FIRST, bool EXACT>
__device__ __forceinline__ void process_row_x(unsigned isb, int i4, const LayeredParams &p, int off, int dc_rt,
                                              int Z4, uint4 *msgp, const uint4 *nx, bool ld_next, bool st_msgs,
                                              unsigned (&mw)[MsgLayout<DCHI>::WORDS], __half2 amaxp, __half2 bmul,
                                              __half2 nbias, unsigned c9b)
{
    constexpr int CH = MsgLayout<DCHI>::CH;
    constexpr bool PRELOAD = LDPC_HI_PRELOAD || DCHI <= 12;
    const int dc = EXACT ? DC : dc_rt;
    if (!PRELOAD && !FIRST && EXACT) msg_load<CH>(msgp, mw);
    const __half2 kbias = __float2half2_rn(1407.0f), one = __float2half2_rn(1.0f);
    const __half2 sent = __float2half2_rn(2047.0f);
    unsigned addr[DC];
    __half2 tt[2][DC];
    __half2 k1[2] = {sent, sent}, k2[2] = {sent, sent}, kprev[2];
    unsigned par[2] = {0u, 0u};  // xor of the t patterns, bit 15 of a lane = sign parity
#pragma unroll
    for (int k = 0; k < DC; k++) {
        if (EXACT || k < dc) {
            const int2 e = p.tab[off + k];  // {column-block base + shift (bytes), wrap threshold (Z - s) * 4}
            unsigned a = isb + (unsigned)e.x;
            a -= (i4 >= e.y) ? (unsigned)Z4 : 0u;
            addr[k] = a;
            const unsigned wk = lds32(a);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const __half2 ap = u2h(prmt(wk, p.c65, h ? 0x4342u : 0x4140u));
                const __half2 t = FIRST ? __hsub2(ap, kbias) : __hsub2(ap, u2h(prmt(mw[k], p.c65, h ? 0x4342u : 0x4140u)));
                tt[h][k] = t;
                par[h] ^= h2u(t);
                const __half2 av = __habs2(t);
                if (!EXACT) {
                    k2[h] = umin2(k2[h], umax2(k1[h], av));
                    k1[h] = umin2(k1[h], av);
                } else if ((k & 1) == 0) {
                    if (k == DC - 1) {  // unpaired last edge
                        k2[h] = umin2(k2[h], umax2(k1[h], av));
                        k1[h] = umin2(k1[h], av);
                    } else {
                        kprev[h] = av;
                    }
                } else {  // pair (k-1, k): 5 min/max for two edges, 2 for the first pair
                    const __half2 lo = umin2(kprev[h], av), hi = umax2(kprev[h], av);
                    if (k == 1) {
                        k1[h] = lo;
                        k2[h] = hi;
                    } else {
                        k2[h] = umin3(k2[h], hi, umax2(k1[h], lo));
                        k1[h] = umin2(k1[h], lo);
                    }
                }
            }
        }
    }
    // the old messages are dead: start the load of the next step's block, the row finalisation and the whole
    // phase 2 before its first use
    if (PRELOAD && ld_next) msg_load<CH>(nx, mw);
    __half2 min1[2], dnew[2], m1p1[2];
    unsigned pm1[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        __half2 m1 = umin2(k1[h], amaxp), m2 = umin2(k2[h], amaxp);  // dc == 1: k2 is the sentinel -> amax
        m1p1[h] = __hadd2(m1, one);
        if (p.scale_on) {
            m1 = beta_scale(m1, bmul, nbias);
            m2 = beta_scale(m2, bmul, nbias);
        }
        min1[h] = m1;
        dnew[h] = __hsub2(m2, m1);
        pm1[h] = (par[h] & 0x80008000u) ^ 0x3C003C00u;  // +-1.0 carrying the parity of the negative signs
    }
    unsigned mn[8];
#pragma unroll
    for (int k = 0; k < CH * 8; k++) {
        if (k < DC && (EXACT || k < dc)) {
            unsigned v[2], nbw[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const __half2 t = tt[h][k];
                unsigned sl;  // 1.0 where |t| <= min1 (this edge holds the minimum), else 0
                asm("fma.rn.sat.f16x2 %0, %1, %2, %3;" : "=r"(sl) : "r"(h2u(__habs2(t))), "r"(0xBC00BC00u), "r"(h2u(m1p1[h])));
                const __half2 mag = __hfma2(u2h(sl), dnew[h], min1[h]);
                const unsigned sg = (h2u(t) & 0x80008000u) ^ pm1[h];
                const __half2 nb = __hfma2(mag, u2h(sg), kbias);
                const __half2 x = __hadd2(nb, t);
                // pattern(x) = 0x6400 + (APP_new + 383): subtract 0x6500, clamp to [0, 254]
                v[h] = __viaddmin_s16x2_relu(h2u(x), c9b, 0x00FE00FEu);
                nbw[h] = h2u(nb);
            }
            sts32(addr[k], prmt(v[0], v[1], 0x6420u));
            mn[k & 7] = prmt(nbw[0], nbw[1], 0x6420u);
        } else {
            mn[k & 7] = 0x7F7F7F7Fu;  // padding / absent edges: message 0
        }
        if ((k & 7) == 7 && st_msgs) msg_store8(msgp + 2 * (k >> 3), mn);
    }
}

// Address of the message block a thread needs in its NEXT step (its next row of this layer, else its first row
// of the next layer / next iteration) and whether that step reads old messages at all.  History (round 1, compressed
// records): prefetching a whole layer ahead into L1 (80 KB in flight for J15_L30_Z1280 against ~64 KB of L1) lost most
// lines before their use — ncu showed 26 % of the stall samples as long-scoreboard waits on the load; one step
// ahead was better, the register preload (process_row_x) better still; an L1 / L2 prefetch on top of it loses 2 %.
template <int DCHI, bool FIRST>
__device__ __forceinline__ bool next_block(const uint4 *msgl, const uint4 *nxl, int i, int T, int Z, bool next_layer_reads,
                                           const uint4 *&nx)
{
    constexpr int RS = MsgLayout<DCHI>::STRIDE;
    const bool same_layer = i + T < Z;
    nx = same_layer ? msgl + (size_t)(i + T) * RS : nxl + (size_t)threadIdx.x * RS;
    return same_layer ? !FIRST : next_layer_reads;
}

// Rows of a layer whose degree is outside the exact range of the kernel's bucket.  Out of line so
// that its register needs do not leak into the allocation of the hot exact-degree paths.
template <int DCHI, bool FIRST>
__device__ __noinline__ void generic_rows(unsigned sbase, const LayeredParams &p, int off, int dc, uint4 *msgl,
                                          __half2 amaxp, __half2 bmul, __half2 nbias, unsigned c9b)
{
    constexpr int RS = MsgLayout<DCHI>::STRIDE;
    const int Z = p.Z, Z4 = 4 * Z, T = blockDim.x;
    for (int i = threadIdx.x; i < Z; i += T) {
        unsigned mw[MsgLayout<DCHI>::WORDS];  // out of line: this path loads its messages itself
        if (!FIRST) msg_load<MsgLayout<DCHI>::CH>(msgl + (size_t)i * RS, mw);
        process_row_x<DCHI, DCHI, FIRST, false>(sbase + 4u * i, 4 * i, p, off, dc, Z4, msgl + (size_t)i * RS, nullptr,
                                                false, true, mw, amaxp, bmul, nbias, c9b);
    }
}

// One full iteration: all layers in order, the Z rows of a layer spread over the CTA.
template <int DCHI, bool FIRST>
__device__ __forceinline__ void sweep_layers(unsigned sbase, const LayeredParams &p, uint4 *msgs, bool last,
                                             unsigned (&mw)[MsgLayout<DCHI>::WORDS], __half2 amaxp, __half2 bmul,
                                             __half2 nbias, unsigned c9b)
{
    constexpr int RS = MsgLayout<DCHI>::STRIDE;
    const int tid = threadIdx.x, T = blockDim.x, Z = p.Z, Z4 = 4 * Z;
    for (int r = 0; r < p.J; r++) {
        const int dc = p.dc[r], off = p.off[r];
        uint4 *msgl = msgs + (size_t)r * Z * RS;
        // this thread's first row of the next layer (of the next iteration after the last layer); it is
        // preloaded only if a sweep that reads messages follows
        const uint4 *nxl = msgs + (size_t)((r + 1 == p.J) ? 0 : r + 1) * Z * RS;
        const bool next_layer_reads = (!FIRST || r == p.J - 1) && !(last && r == p.J - 1);
        // the messages of the last sweep are read by nobody (debug dumps excepted)
        const bool st_msgs = !(LDPC_SKIP_LAST_ST && DCHI <= 12 && last) || p.dbg_msg != nullptr;
#define LDPC_ROWS(DCX)                                                                                        \
    for (int i = tid; i < Z; i += T) {                                                                        \
        const uint4 *nx;                                                                                      \
        const bool ldn = next_block<DCHI, FIRST>(msgl, nxl, i, T, Z, next_layer_reads, nx);                   \
        process_row_x<DCX, DCHI, FIRST, true>(sbase + 4u * i, 4 * i, p, off, DCX, Z4, msgl + (size_t)i * RS, nx, ldn, \
                                              st_msgs, mw, amaxp, bmul, nbias, c9b);                          \
    }
        if (dc == DCHI) {
            LDPC_ROWS(DCHI)
        } else if (DCHI >= 2 && dc == DCHI - 1) {
            LDPC_ROWS((DCHI >= 2 ? DCHI - 1 : 1))
        } else if (DCHI >= 3 && dc == DCHI - 2) {
            LDPC_ROWS((DCHI >= 3 ? DCHI - 2 : 1))
        } else if (DCHI >= 4 && dc == DCHI - 3) {
            LDPC_ROWS((DCHI >= 4 ? DCHI - 3 : 1))
        } else if (DCHI >= 5 && dc == DCHI - 4) {  // six exact degrees per bucket: every shipped code whose
            LDPC_ROWS((DCHI >= 5 ? DCHI - 4 : 1))  // degrees spread over two values stays on the exact paths
        } else if (DCHI >= 6 && dc == DCHI - 5) {  // (J20 / J40 / J48_L60_Z160 ran 2x slower through the generic one)
            LDPC_ROWS((DCHI >= 6 ? DCHI - 5 : 1))
        } else {  // degree outside the bucket's exact range: predicated generic path (rare, kept out of line)
            generic_rows<DCHI, FIRST>(sbase, p, off, dc, msgl, amaxp, bmul, nbias, c9b);
            if ((LDPC_HI_PRELOAD || DCHI <= 12) && next_layer_reads && tid < Z)
                msg_load<MsgLayout<DCHI>::CH>(nxl + (size_t)tid * RS, mw);
        }
#undef LDPC_ROWS
        __syncthreads();
        LDPC_STRESS_POINT(1);
    }
}

// fail bits (0x80 per frame byte) of check row i (isb = shared-memory address of row 0's word + 4 i) of one layer
__device__ __forceinline__ unsigned syndrome_row(unsigned isb, const LayeredParams &p, int off, int dc, int i4, int Z4)
{
    unsigned x = (dc & 1) ? 0x80808080u : 0u;  // byte = APP + 127 (<= 254): negative <=> bit 7 of byte + 1 clear
#pragma unroll 4
    for (int k = 0; k < dc; k++) {
        const int2 e = p.tab[off + k];  // the sweep's own table: {byte offset of row 0's bit, wrap threshold}
        unsigned a = isb + (unsigned)e.x;
        a -= (i4 >= e.y) ? (unsigned)Z4 : 0u;
        x ^= lds32(a) + 0x01010101u;
    }
    return x & 0x80808080u;
}

// hard decisions of frames selected by `fmask` (bit j = frame 4g+j) -> global memory
__device__ void write_outputs(const unsigned *appw, const LayeredParams &p, int g, unsigned fmask)
{
    const int F = p.F, N = p.N, f0 = 4 * g;
    const bool per_bit_pass = p.out_format != LDPC_OUT_BITPACK || p.dbg_app != nullptr;
    for (int n = threadIdx.x; per_bit_pass && n < N; n += blockDim.x) {
        const unsigned w = appw[n];
        // bit j of `bits` = hard decision of frame j (APP < 0 <=> byte + 1 < 128; bytes never exceed 254)
        const unsigned nb = ~(w + 0x01010101u);
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
                if ((fmask >> j) & 1u) p.dbg_app[(size_t)n * F + f0 + j] = (signed char)((int)((w >> (8 * j)) & 255u) - 127);
        }
    }
    if (p.out_format == LDPC_OUT_BITPACK) {
        unsigned *D = reinterpret_cast<unsigned *>(p.out);
        const int W = (N + 31) / 32;
        if ((N & 31) == 0) {
            // one thread per output word: 32 code bits x 4 frames.  (~w - 0x01..)'s bit 7 per byte lane = hard decision of
            // that frame; shifted to bit k of its lane for code bit k of an octet, the byte lanes collect 8 decisions
            // each; a 4 x 4 byte transpose (PRMT) turns the four octet words into one 32-bit word per frame.
            // (The ballot version spent 7.5 instructions per decision, this one ~1: 4.9 % -> 0.7 % of the kernel's
            // instructions for J15_L30_Z1280.)
            const uint4 *a4 = reinterpret_cast<const uint4 *>(appw);
            for (int wi = threadIdx.x; wi < W; wi += blockDim.x) {
                unsigned acc[4];
#pragma unroll
                for (int o = 0; o < 4; o++) {
                    const uint4 x = a4[8 * wi + 2 * o], y = a4[8 * wi + 2 * o + 1];
                    const unsigned w8[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
                    unsigned a = 0u;
#pragma unroll
                    for (int k = 0; k < 8; k++) a |= ((~(w8[k] + 0x01010101u) & 0x80808080u) >> 7) << k;
                    acc[o] = a;
                }
                const unsigned t0 = prmt(acc[0], acc[1], 0x5140u), t1 = prmt(acc[2], acc[3], 0x5140u);
                const unsigned t2 = prmt(acc[0], acc[1], 0x7362u), t3 = prmt(acc[2], acc[3], 0x7362u);
                const unsigned o4[4] = {prmt(t0, t1, 0x5410u), prmt(t0, t1, 0x7632u), prmt(t2, t3, 0x5410u),
                                        prmt(t2, t3, 0x7632u)};
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if ((fmask >> j) & 1u) D[(size_t)(f0 + j) * W + wi] = o4[j];
            }
        } else {
            const int lane = threadIdx.x & 31;
            const int nround = (N + 31) & ~31;
            for (int n = threadIdx.x; n < nround; n += blockDim.x) {  // blockDim is a multiple of 32
                const unsigned w = (n < N) ? appw[n] + 0x01010101u : 0x80808080u;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const unsigned b = __ballot_sync(0xffffffffu, ((w >> (8 * j + 7)) & 1u) == 0u);
                    if (lane == 0 && ((fmask >> j) & 1u)) D[(size_t)(f0 + j) * W + (n >> 5)] = b;
                }
            }
        }
    }
}

// debug: the c2v messages of frames selected by `fmask` -> int8 [M][dc_max][F] (0 for absent edges)
template <int DCMAX>
__device__ void dump_messages(const LayeredParams &p, const uint4 *msgs, int g, unsigned fmask, int dc_max)
{
    constexpr int RS = MsgLayout<DCMAX>::STRIDE;
    const unsigned *mw = reinterpret_cast<const unsigned *>(msgs);
    for (int m = threadIdx.x; m < p.M; m += blockDim.x) {
        const unsigned *r = mw + (size_t)m * RS * 4;
        const int dc = p.dc[m / p.Z];
        for (int k = 0; k < dc_max; k++) {
            const unsigned w = (k < dc) ? r[k] : 0x7F7F7F7Fu;
            for (int j = 0; j < 4; j++)
                if ((fmask >> j) & 1u)
                    p.dbg_msg[((size_t)m * dc_max + k) * p.F + 4 * g + j] = (signed char)((int)((w >> (8 * j)) & 255u) - 127);
        }
    }
}

// threads per CTA are capped so that the per-row register arrays (2*DCMAX + DCMAX) never spill
template <int DCMAX>
struct ThreadCap {
    static constexpr int value = DCMAX <= 12 ? 640 : (DCMAX <= 24 ? 512 : 384);
};
// __launch_bounds__ of the kernel = its register cap (640 -> 96, 512 -> 128, 384 -> 168 registers).  Buckets 16 and
// 20 take 96 registers although their arrays want more: the shipped codes there have small Z (J4_L24_Z96: 96
// threads per CTA), where two more resident CTAs per SM beat the ~0.3 KB of spills (B200: 45.6 -> 47.5 Gbit/s);
// bucket 24 (PON, Z = 256, two CTAs per SM by shared memory either way) keeps 128 (39.4 vs 38.0 at 96).
template <int DCMAX>
struct RegCapThreads {
    static constexpr int value = (DCMAX > 12 && DCMAX <= 20) ? 640 : ThreadCap<DCMAX>::value;
};

template <int DCMAX>
__global__ void __launch_bounds__(RegCapThreads<DCMAX>::value, 1)
ldpc_layered_i8_kernel(const __grid_constant__ LayeredParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned *appw = reinterpret_cast<unsigned *>(smem);
    // Syndrome fail bits of the CTA, ping-ponged by layer parity: layer r ORs into s_fail[r & 1] before the
    // layer's barrier and every thread reads that word after it; the word is next written in layer r + 2, i.e.
    // behind barrier r + 1, which no thread passes before all have read (a single word raced: a fast warp's
    // atomicOr of layer r + 1 could reach a slow warp's read of layer r and split the CTA at the break below).
    __shared__ unsigned s_fail[2];
    const int tid = threadIdx.x, T = blockDim.x;
    const int N = p.N, Z = p.Z, F = p.F, Z4 = 4 * Z;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    uint4 *rec = p.rec + (size_t)blockIdx.x * p.M * MsgLayout<DCMAX>::STRIDE;
    const __half2 amaxp = __float2half2_rn(p.msg_max);  // magnitude clip of min1 / min2
    const unsigned c9b = p.c9b;
    const __half2 bmul = __float2half2_rn(p.beta_mul);
    const __half2 nbias = __float2half2_rn(-p.beta_bias);

    // Group scheduling: with early exit the decode time of a group depends on its slowest frame, so groups
    // are handed out through an atomic counter (first come, first served); with fixed iterations every
    // group costs the same and a static stride keeps the next group known (L2 prefetch below).
    __shared__ int s_next;
    const bool dynamic = p.work_counter != nullptr;
    int g = blockIdx.x;
    for (;; g += gridDim.x) {
        if (dynamic) {
            if (tid == 0) s_next = atomicAdd(p.work_counter, 1);
            __syncthreads();
            LDPC_STRESS_POINT(6);
            g = s_next;
        }
        if (g >= p.num_groups) break;
#if LDPC_PHASE_TIMERS
        long long phase_t0_ = clock64();
#endif
        const int f0 = 4 * g;
        unsigned valid = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) valid |= (f0 + j < F) ? (1u << j) : 0u;
        // ---- fused channel: generate y = 1 - 2c + sigma*n for the 4 frames of the group right here (one
        // Philox call = 4 consecutive bits of one frame), quantise, store — no channel buffer in HBM at all
        if (p.llr_dtype == LDPC_DTYPE_CHANNEL) {
            for (int nb = tid; nb < (N + 3) / 4; nb += T) {
                float g[4][4];
#pragma unroll
                for (int j = 0; j < 4; j++) awgn_normals4(p.ch_first + (unsigned long long)(f0 + j), nb, p.ch_k0, p.ch_k1, g[j]);
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int n = 4 * nb + b;
                    if (n >= N) break;
                    const int bit = p.ch_cw ? (p.ch_cw[n] & 1) : 0;
                    unsigned w = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int q = ((valid >> j) & 1u) ? quant(awgn_bpsk_sample(bit, p.ch_sigma, g[j][b]), p.scale) : 127;
                        w |= (unsigned)(q + 127) << (8 * j);
                    }
                    appw[n] = w;
                }
            }
        }
        // ---- load + quantise: q = sat127(rint(y * scale)), stored biased by 127
        // fp32 [N][F], whole group: one 16-byte vector per code bit, LDPC_LOAD_DEPTH loads in flight per thread
        const bool fast_load = LDPC_LOAD_DEPTH > 1 && p.llr_dtype == LDPC_DTYPE_FP32 && p.layout == LDPC_LAYOUT_NF &&
                               valid == 0xFu && (F & 3) == 0;
        if (fast_load) {
            constexpr int kLoadDepth = LDPC_LOAD_DEPTH;
            const float4 *y4 = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(p.llr) + f0);
            const size_t st4 = (size_t)(F >> 2);
            for (int n0 = tid; n0 < N; n0 += kLoadDepth * T) {
                float4 v[kLoadDepth];
#pragma unroll
                for (int u = 0; u < kLoadDepth; u++) {  // tail: re-read the last bit instead of predicating
                    const int n = min(n0 + u * T, N - 1);
                    v[u] = __ldg(y4 + (size_t)n * st4);
                }
#pragma unroll
                for (int u = 0; u < kLoadDepth; u++) {
                    const int n = n0 + u * T;
                    if (n < N)
                        appw[n] = quant4(v[u].x, v[u].y, v[u].z, v[u].w, p.scale);
                }
            }
        }
        // fp16 [N][F], whole group: one aligned 8-byte vector per code bit (same rule as fp32 after the exact
        // half -> float conversion)
        const bool fast_load16 = LDPC_LOAD_DEPTH > 1 && p.llr_dtype == LDPC_DTYPE_FP16 && p.layout == LDPC_LAYOUT_NF &&
                                 valid == 0xFu && (F & 3) == 0 && (reinterpret_cast<size_t>(p.llr) & 7) == 0;
        if (fast_load16) {
            constexpr int kLoadDepth = LDPC_LOAD_DEPTH;
            const uint2 *y2 = reinterpret_cast<const uint2 *>(reinterpret_cast<const __half *>(p.llr) + f0);
            const size_t st2 = (size_t)(F >> 2);
            for (int n0 = tid; n0 < N; n0 += kLoadDepth * T) {
                uint2 v[kLoadDepth];
#pragma unroll
                for (int u = 0; u < kLoadDepth; u++) v[u] = __ldg(y2 + (size_t)min(n0 + u * T, N - 1) * st2);
#pragma unroll
                for (int u = 0; u < kLoadDepth; u++) {
                    const int n = n0 + u * T;
                    if (n < N) {
                        const float2 a = __half22float2(u2h(v[u].x)), b = __half22float2(u2h(v[u].y));
                        appw[n] = quant4(a.x, a.y, b.x, b.y, p.scale);
                    }
                }
            }
        }
        // int8 [N][F], whole group: the 4 frames of a code bit are one aligned 32-bit word (the host-pack path of
        // ldpc_decode_batch feeds this); byte s -> max(s, -127) + 127 without unpacking
        const bool fast_load8 = LDPC_LOAD_DEPTH > 1 && p.llr_dtype == LDPC_DTYPE_INT8 && p.layout == LDPC_LAYOUT_NF &&
                                valid == 0xFu && (F & 3) == 0 && (reinterpret_cast<size_t>(p.llr) & 3) == 0;
        if (fast_load8) {
            constexpr int kLoadDepth = LDPC_LOAD_DEPTH;
            const unsigned *y1 = reinterpret_cast<const unsigned *>(reinterpret_cast<const signed char *>(p.llr) + f0);
            const size_t st1 = (size_t)(F >> 2);
            for (int n0 = tid; n0 < N; n0 += kLoadDepth * T) {
                unsigned v[kLoadDepth];
#pragma unroll
                for (int u = 0; u < kLoadDepth; u++) v[u] = __ldg(y1 + (size_t)min(n0 + u * T, N - 1) * st1);
#pragma unroll
                for (int u = 0; u < kLoadDepth; u++) {
                    const int n = n0 + u * T;
                    if (n < N) appw[n] = __vmaxu4(v[u] ^ 0x80808080u, 0x01010101u) - 0x01010101u;
                }
            }
        }
#pragma unroll 2
        for (int n = (p.llr_dtype == LDPC_DTYPE_CHANNEL || fast_load || fast_load8 || fast_load16) ? N : tid; n < N; n += T) {
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
            appw[n] = (unsigned)(q[0] + 127) | ((unsigned)(q[1] + 127) << 8) | ((unsigned)(q[2] + 127) << 16) |
                      ((unsigned)(q[3] + 127) << 24);
        }
        if (tid == 0) s_fail[0] = s_fail[1] = 0u;
        __syncthreads();
        LDPC_STRESS_POINT(5);
        // pull the channel values of this CTA's next group towards L2 while this group is decoded
        const bool can_prefetch = !dynamic && g + (int)gridDim.x < p.num_groups && p.llr_dtype == LDPC_DTYPE_FP32 &&
                                  p.layout == LDPC_LAYOUT_NF;
        if (LDPC_L2_PREFETCH == 1 && can_prefetch) {
            const float *y = reinterpret_cast<const float *>(p.llr) + 4 * (size_t)(g + gridDim.x);
            for (int n = tid; n < N; n += T) asm volatile("prefetch.global.L2 [%0];" ::"l"(y + (size_t)n * F));
        }

        LDPC_PHASE_MARK(0);
        unsigned running = valid;  // frames not yet latched
        int it = 0;
        unsigned rw[MsgLayout<DCMAX>::WORDS];  // the old messages of the thread's next step
        while (it < p.iters) {
            it++;
            if (LDPC_L2_PREFETCH == 2 && can_prefetch && it == p.iters) {
                // the next group's 16 bytes per code bit: by the time this group's last sweep and output are done
                // they wait in L2 (the load phase at the start of a group otherwise waits for HBM with nothing to
                // overlap: one CTA per SM for the big codes)
                const float *y = reinterpret_cast<const float *>(p.llr) + 4 * (size_t)(g + gridDim.x);
                for (int n = tid; n < N; n += T) asm volatile("prefetch.global.L2 [%0];" ::"l"(y + (size_t)n * F));
            }
            if (it == 1)
                sweep_layers<DCMAX, true>(sbase, p, rec, it == p.iters, rw, amaxp, bmul, nbias, c9b);
            else
                sweep_layers<DCMAX, false>(sbase, p, rec, it == p.iters, rw, amaxp, bmul, nbias, c9b);
            LDPC_PHASE_MARK(it == 1 ? 1 : 2);
            if (p.exit_mode == LDPC_EXIT_SYNDROME || it == p.iters) {
                // The pass stops after the first layer that leaves every running frame with a failed check (the
                // usual case until the last iterations: one layer of J is read instead of all) — the outcome
                // "not converged" is then already exact for all of them.
                unsigned rbytes = 0u;  // 0x80 in the byte of every running frame
#pragma unroll
                for (int j = 0; j < 4; j++) rbytes |= ((running >> j) & 1u) ? (0x80u << (8 * j)) : 0u;
                unsigned fl = 0u;
                for (int r = 0; r < p.J; r++) {
                    const int dc = p.dc[r], off = p.off[r];
                    unsigned fail = 0u;
                    for (int i = tid; i < Z; i += T) fail |= syndrome_row(sbase + 4u * i, p, off, dc, 4 * i, Z4);
                    fail = __reduce_or_sync(0xffffffffu, fail);
                    constexpr int kFlip = LDPC_R1_RACES ? 0 : 1;
                    if ((tid & 31) == 0 && fail) atomicOr(&s_fail[r & kFlip], fail);
                    __syncthreads();
                    LDPC_STRESS_POINT(2);
                    fl |= s_fail[r & kFlip];             // words accumulate over layers r, r-2, ...: OR is idempotent
                    if ((fl & rbytes) == rbytes) break;  // uniform: nobody writes this word before the next barrier
                }
                __syncthreads();  // every read of s_fail is done
                if (tid == 0) s_fail[0] = s_fail[1] = 0u;  // next written behind the barriers of the next sweep
                // okmask bit j = frame j satisfies all checks
                const unsigned okmask = (((~fl) >> 7) & 1u) | (((~fl) >> 14) & 2u) | (((~fl) >> 21) & 4u) |
                                        (((~fl) >> 28) & 8u);
                const unsigned finish = (it == p.iters) ? running : (running & okmask);
                LDPC_PHASE_MARK(3);
                if (finish) {
                    LDPC_STRESS_POINT(3);
                    write_outputs(appw, p, g, finish);
                    if (p.dbg_msg) dump_messages<DCMAX>(p, rec, g, finish, p.dc_max);
                    if (tid < 4 && ((finish >> tid) & 1u)) {
                        if (p.iters_out) p.iters_out[f0 + tid] = it;
                        if (p.ok_out) p.ok_out[f0 + tid] = (okmask >> tid) & 1u;
                        if (p.out_format == LDPC_OUT_INT32_REF)
                            reinterpret_cast<int *>(p.out)[(size_t)N * F + f0 + tid] = (okmask >> tid) & 1u;
                    }
                    running &= ~finish;
                    // `finish` is CTA-uniform.  The frames that keep running go straight into the next sweep, which
                    // rewrites the APP words and records that write_outputs / dump_records of slower warps are
                    // still reading for the latched frames.
                    if (running && !LDPC_R1_RACES) __syncthreads();
                }
                if (!running) break;
            }
        }
        __syncthreads();  // everyone is done with appw before the next group's load
        LDPC_PHASE_MARK(4);
        LDPC_STRESS_POINT(4);
    }
}

// host-side dispatch over the degree buckets (DCHI = dc_max rounded up to a multiple of 4)
template <int DCHI>
struct I8Kernel {
    static int occupancy(int threads, size_t smem, int *out)
    {
        LDPC_CUDA_TRY(cudaFuncSetAttribute(ldpc_layered_i8_kernel<DCHI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        LDPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, ldpc_layered_i8_kernel<DCHI>, threads, smem));
        return LDPC_OK;
    }
    static int launch(const LayeredParams &p, int threads, int grid, size_t smem, cudaStream_t st)
    {
        ldpc_layered_i8_kernel<DCHI><<<grid, threads, smem, st>>>(p);
        LDPC_CUDA_TRY(cudaGetLastError());
        return LDPC_OK;
    }
};

#ifndef LDPC_BUCKETS
#define LDPC_BUCKETS(X) X(4) X(8) X(12) X(16) X(20) X(24) X(28) X(32)
#endif

struct I8Plan {
    int dcb, threads, grid, u4;
    size_t smem;
};

static int plan_i8(const ldpc_code *c, int F, I8Plan *pl)
{
    pl->smem = (size_t)c->N * 4;
    if (pl->smem > 227 * 1024) return LDPC_ERR_UNSUPPORTED;  // N > 58112: one group does not fit one SM
    pl->dcb = (c->dc_max + 3) & ~3;
    int cap = 0, occ = 0, rc = LDPC_ERR_UNSUPPORTED;
    switch (pl->dcb) {
#define X(D) case D: cap = ThreadCap<D>::value; pl->u4 = MsgLayout<D>::STRIDE; break;
        LDPC_BUCKETS(X)
#undef X
        default: return LDPC_ERR_UNSUPPORTED;
    }
    // threads: the Z rows of a layer spread over whole warps, at most ThreadCap<DCHI> threads
    if (const char *e = getenv("LDPC_B200_THREADS"))  // tuning knob (tools/ab): CTA width cap
        if (atoi(e) >= 32 && atoi(e) < cap) cap = atoi(e);
    const int rows_per_thread = (c->Z + cap - 1) / cap;
    const int threads = (c->Z + rows_per_thread - 1) / rows_per_thread;
    pl->threads = (threads + 31) & ~31;
    switch (pl->dcb) {
#define X(D) case D: rc = I8Kernel<D>::occupancy(pl->threads, pl->smem, &occ); break;
        LDPC_BUCKETS(X)
#undef X
    }
    if (rc != LDPC_OK) return rc;
    if (occ < 1) return LDPC_ERR_UNSUPPORTED;
    const int groups = (F + 3) / 4;
    pl->grid = c->num_sms * occ;
    if (pl->grid > groups) pl->grid = groups;
    return LDPC_OK;
}

int layered_i8_wave_frames(const ldpc_code *c, int *frames)
{
    I8Plan pl;
    int rc = plan_i8(c, 1 << 30, &pl);
    if (rc != LDPC_OK) return rc;
    *frames = pl.grid * 4;  // resident CTAs x the 4 codewords of a group
    return LDPC_OK;
}

int layered_i8_scratch_bytes(const ldpc_code *c, int F, int, size_t *bytes)
{
    I8Plan pl;
    int rc = plan_i8(c, F, &pl);
    if (rc != LDPC_OK) return rc;
    *bytes = (size_t)pl.grid * c->M * pl.u4 * sizeof(uint4) + 256;  // + the work counter
    return LDPC_OK;
}

int launch_layered_i8(const ldpc_code *c, const LayeredArgs &a, cudaStream_t st, int *launches)
{
    if (a.msg_max < 1 || a.msg_max > 127 || a.beta_num < 0 || a.beta_num > 8 || a.beta_shift < 0 ||
        a.beta_shift > 7 || (a.beta_num != 0 && a.beta_num >= (1 << a.beta_shift)))
        return LDPC_ERR_ARG;
    I8Plan pl;
    int rc = plan_i8(c, a.F, &pl);
    if (rc != LDPC_OK) return rc;
    const size_t rec_bytes = (size_t)pl.grid * c->M * pl.u4 * sizeof(uint4);
    if (rec_bytes + 256 > a.scratch_bytes) return LDPC_ERR_NOMEM;
    LayeredParams p;
    memset(&p, 0, sizeof(p));
    p.llr = a.llr;
    p.out = a.out;
    p.iters_out = a.iters_out;
    p.ok_out = a.ok_out;
    p.dbg_app = reinterpret_cast<signed char *>(a.dbg_app);
    p.dbg_msg = reinterpret_cast<signed char *>(a.dbg_rec);
    p.dc_max = c->dc_max;
    p.rec = reinterpret_cast<uint4 *>(a.scratch);
    if (a.exit_mode == LDPC_EXIT_SYNDROME) {
        p.work_counter = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(a.scratch) + rec_bytes);
        LDPC_CUDA_TRY(cudaMemsetAsync(p.work_counter, 0, sizeof(int), st));
    }
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
    p.ch_sigma = a.ch_sigma;
    p.ch_k0 = (unsigned)a.ch_seed;
    p.ch_k1 = (unsigned)(a.ch_seed >> 32);
    p.ch_first = a.ch_first;
    p.ch_cw = a.ch_cw;
    p.scale_on = a.beta_num != 0;
    p.c65 = 0x65656565u;
    p.c9b = 0x9B009B00u;
    p.lt = c->lt;
    for (int r = 0; r < c->J; r++) {
        p.off[r] = c->lt.off[r];
        p.dc[r] = c->lt.dc[r];
        for (int k = 0; k < c->lt.dc[r]; k++) {
            const int e = c->lt.off[r] + k;
            p.tab[e] = make_int2(((int)c->lt.col[e] * c->Z + (int)c->lt.shift[e]) * 4, (c->Z - (int)c->lt.shift[e]) * 4);
        }
    }
    switch (pl.dcb) {
#define X(D) case D: rc = I8Kernel<D>::launch(p, pl.threads, pl.grid, pl.smem, st); break;
        LDPC_BUCKETS(X)
#undef X
    }
    if (rc == LDPC_OK) *launches += 1;
    return rc;
}


}  // namespace ldpcb
