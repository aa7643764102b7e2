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
//  * the compressed check record {min1, min2, idx, sign bits} is streamed through a per-CTA
//    slice of a scratch buffer (L2 evict-last; for J15_L30_Z1280 the 148 x 19200 x 32 B = 91 MB do
//    not stay in L2 beside the streaming input, about half of the record reads come from HBM —
//    ~2 MB of DRAM traffic per frame, a quarter of the HBM peak); each thread loads the record of
//    its NEXT row into registers while it finishes the current one, which hides that latency;
//  * the syndrome (early exit / ok flag) is a cheap extra pass: XOR of the sign bits.
//
// Update rules: exactly oracle/bldpc_oracle.c "int8 layered rules" (t = APP - c2v_old,
// min1/min2/first-index on min(|t|, msg_max), m' = m - ((m*beta_num) >> beta_shift), APP =
// sat127(t + c2v_new)); check-node rule per B/LDPC_Decoder.cu:279-314, schedule ours.
#include <cuda_fp16.h>

#include "common.h"
#include "philox.cuh"
#include "stress.cuh"

// Tuning switches; the defaults are the winners of A/B runs on B200 (tools/ab, J15_L30_Z1280, 10 it):
//   LDPC_REC_PRELOAD   2  the record of a thread's next step is loaded into registers right after the edge loop
//                         of the current row (1 = after the record store): 9.07 -> 8.70 -> 8.45 ms; an L1 or L2
//                         prefetch on top of it LOSES 2 %
//   LDPC_REC_PREFETCH  0  1 = prefetch.global.L1, 2 = prefetch.global.L2 of the next step's record at row start
//   LDPC_PARITY_XOR    1  sign parity as xor of the t patterns (LOP3) instead of an fp16 count (9.30 -> 9.06 ms)
//   LDPC_NEG_ALU       1  (t < 0) as HSET2 on the ALU pipe instead of fma.sat on the FMA pipe: after the rework the
//                         fp16 FMA pipe is the fuller one (8.76 -> 8.63 ms); moving the sign shifts to the ALU pipe
//                         (SHF instead of IMAD.SHL) or phase 2's sign product to the FMA pipe both LOSE
//   LDPC_LOAD_DEPTH    8  channel-value vectors in flight per thread in the load phase
//   LDPC_L2_PREFETCH   0  L2 prefetch of the CTA's next group of channel values (on: evicts records, +4 %)
#ifndef LDPC_L2_PREFETCH
#define LDPC_L2_PREFETCH 0
#endif
#ifndef LDPC_REC_PREFETCH
#define LDPC_REC_PREFETCH 0
#endif
#ifndef LDPC_PARITY_XOR
#define LDPC_PARITY_XOR 1
#endif
#ifndef LDPC_REC_PRELOAD
#define LDPC_REC_PRELOAD 2
#endif
#ifndef LDPC_NEG_ALU
#define LDPC_NEG_ALU 1
#endif
#ifndef LDPC_LOAD_DEPTH
#define LDPC_LOAD_DEPTH 8
#endif

namespace ldpcb {

struct LayeredParams {
    const void *llr;
    void *out;
    int *iters_out, *ok_out;
    signed char *dbg_app;
    unsigned *dbg_rec;
    uint4 *rec;             // scratch: per CTA, M records of REC_U4 uint4
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
    // kernel-ready layer tables: entry e = {byte offset (c*Z + s)*4 of row 0's bit, wrap threshold (Z - s)*4}
    int2 tab[kMaxBlocks];
    unsigned short off[kMaxLayers];
    unsigned char dc[kMaxLayers];
    LayerTables lt;         // compact form, used by the debug dump only
};

// Check records: 256-bit global accesses with the L2 evict-last policy (LDG.E.ELL2.256 /
// STG.E.ELL2.256 on sm_100a) so that the per-CTA record slices are the last thing L2 gives up while the
// channel values stream through; the trailing 128 bits of wide records use a plain access.
template <int U4>
__device__ __forceinline__ void rec_load(const uint4 *p, unsigned *rw)
{
    asm volatile("ld.global.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(rw[0]), "=r"(rw[1]), "=r"(rw[2]), "=r"(rw[3]), "=r"(rw[4]), "=r"(rw[5]), "=r"(rw[6]), "=r"(rw[7])
                 : "l"(p));
    if (U4 == 3) {
        asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(rw[8]), "=r"(rw[9]), "=r"(rw[10]), "=r"(rw[11])
                     : "l"(p + 2));
    } else if (U4 == 4) {
        asm volatile("ld.global.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(rw[8]), "=r"(rw[9]), "=r"(rw[10]), "=r"(rw[11]), "=r"(rw[12]), "=r"(rw[13]), "=r"(rw[14]),
                       "=r"(rw[15])
                     : "l"(p + 2));
    }
}
template <int U4>
__device__ __forceinline__ void rec_store(uint4 *p, const unsigned *rw)
{
    asm volatile("st.global.L2::evict_last.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(rw[0]), "r"(rw[1]),
                 "r"(rw[2]), "r"(rw[3]), "r"(rw[4]), "r"(rw[5]), "r"(rw[6]), "r"(rw[7])
                 : "memory");
    if (U4 == 3) {
        asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p + 2), "r"(rw[8]), "r"(rw[9]), "r"(rw[10]),
                     "r"(rw[11])
                     : "memory");
    } else if (U4 == 4) {
        asm volatile("st.global.L2::evict_last.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p + 2), "r"(rw[8]),
                     "r"(rw[9]), "r"(rw[10]), "r"(rw[11]), "r"(rw[12]), "r"(rw[13]), "r"(rw[14]), "r"(rw[15])
                     : "memory");
    }
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

// Sign bits are collected on the fp16 pipe (acc = 2*acc + (t<0)), 10 edges per accumulator so the
// value stays below 1024 and its bit pattern can be read off 1024 + acc.  Edge k of a degree-dc
// check lives in sign word g = k/10 at bit (min(10, dc-10g) - 1 - (k-10g)) of each 16-bit lane.
constexpr int kSignGroup = 10;
template <int DCMAX>
struct RecLayout {
    static constexpr int SW = (DCMAX + kSignGroup - 1) / kSignGroup;  // sign words per half-group
    static constexpr int WORDS = 6 + 2 * SW;                          // m1 x2, m2 x2, idx x2, signs
    static constexpr int U4 = (WORDS + 3) / 4;                        // uint4 moved per record (2..4)
    static constexpr int STRIDE = U4 <= 2 ? 2 : 4;  // uint4 between records: every record 32-byte aligned
};
__host__ __device__ constexpr int sign_bit_pos(int dc, int k)
{
    const int g = k / kSignGroup, j = k - g * kSignGroup;
    const int gs = (dc - g * kSignGroup) < kSignGroup ? (dc - g * kSignGroup) : kSignGroup;
    return gs - 1 - j;
}

// One check row of one layer for the 4 codewords of the group.
// EXACT: the check degree is the template constant DC (no per-edge branches, every table entry is
// a constant-bank load at an immediate offset, every shift amount is an immediate).  !EXACT: DC is
// the bucket's maximum and edges k >= dc are predicated off (rare degrees only).
//
// Number formats (all exact):
//  * smem byte b = APP + 127 (0..254).  PRMT with the constant byte 0x65 turns it into the fp16 pattern
//    0x6500|b = 1280 + b = APP + 1407; t' = t + 1407 then stays inside the binade [1024, 2048) for every
//    reachable t (|t| <= 254) and so does t' + c2v_new, where an fp16 pattern is 0x6400 + (value - 1024):
//    the final "add, clamp to [-127,127], re-bias" is ONE integer SIMD instruction on the pattern,
//    VIADDMNMX.S16x2.RELU: max(min(pattern - 0x6500, 254), 0) = clamp(APP_new) + 127.
//  * min1 / min2 / first index are tracked on keys |t|*8 + (k mod 8) (< 2048, exact in fp16; positive
//    fp16 patterns order like unsigned integers, so VIMNMX(3).U16x2 does the comparisons), one
//    accumulator pair per set of 8 edges; the amax clamp commutes with the minimum and is applied to
//    the keys once per row (key1 -> min(key1, 8*amax) also yields the oracle's "first index" when every
//    |t| saturates).  Keys are decoded (floor(key/8), key mod 8) on the FMA pipe.
//  * sign bits are collected on the fp16 pipe (acc = 2*acc + (t<0)); (t<0) is an HSET2 (LDPC_NEG_ALU) or fma.sat(t, -1, 0).
// Pipe balance (profiles/r01_pipe_ubench.txt: every instruction here issues once per 2 cycles per sub-partition,
// HFMA2 / HADD2 / IMAD on one pipe, HSET2 / VIMNMX / LOP3 / PRMT on the other, ~0.8 IPC at best when mixed): per
// edge and pair of frames the ALU pipe sees PRMT, HSET2.EQ, LOP3, HSET2.LT, 2.5 x VIMNMX, 0.5 x LOP3 (parity) in
// phase 1 and HSET2.EQ, LOP3, VIADDMNMX in phase 2; the FMA pipe the select HFMA2, IMAD.SHL, two HADD2, the sign
// and key HFMA2 in phase 1 and HFMA2, IMAD.SHL, HADD2 in phase 2, plus decode / beta scaling per row.  A/B runs
//  put the optimum at this split: moving one more op either way loses.
__device__ __forceinline__ __half2 neg01(__half2 t)
{
    unsigned d;
    asm("fma.rn.sat.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(h2u(t)), "r"(0xBC00BC00u), "r"(0u));
    return u2h(d);
}
__device__ __forceinline__ __half2 umin2(__half2 a, __half2 b) { return u2h(__vminu2(h2u(a), h2u(b))); }
__device__ __forceinline__ __half2 umax2(__half2 a, __half2 b) { return u2h(__vmaxu2(h2u(a), h2u(b))); }
__device__ __forceinline__ __half2 umin3(__half2 a, __half2 b, __half2 c)
{
    return u2h(__vimin3_u16x2(h2u(a), h2u(b), h2u(c)));
}
// floor(c / 8) for integer-valued 0 <= c <= 1023: c/8 - 7/16 has at most 11 significant bits, and adding
// 1025 (ulp 1, never a tie) rounds it to 1025 + floor
__device__ __forceinline__ __half2 floor8(__half2 c)
{
    const __half2 k = __float2half2_rn(1025.0f);
    const __half2 u = __hfma2(c, __float2half2_rn(0.125f), __float2half2_rn(-0.4375f));
    return __hsub2(__hadd2(u, k), k);
}

// Record flow: `rw` holds this row's record on entry (!FIRST) and the record of the thread's NEXT step on
// exit (`ld_next`: the load is issued between the two phases, so its latency hides under phase 2; the
// registers of the old record are dead by then).
template <int DC, int DCHI, bool FIRST, bool EXACT>
__device__ __forceinline__ void process_row_x(unsigned isb, int i4, const LayeredParams &p, int off, int dc_rt,
                                              int Z4, uint4 *recp, const uint4 *nx, bool ld_next,
                                              unsigned (&rw)[RecLayout<DCHI>::U4 * 4], __half2 amax8,
                                              __half2 amax8p7, __half2 bmul, __half2 nbias)
{
    constexpr int SW = RecLayout<DCHI>::SW;  // record layout of the kernel's degree bucket
    constexpr int U4 = RecLayout<DCHI>::U4;
    constexpr int NS = (DC + 7) / 8;         // key sets
    const int dc = EXACT ? DC : dc_rt;
    if (!FIRST && !LDPC_REC_PRELOAD) rec_load<U4>(recp, rw);
    // record words: [0,1] m1 (frames 01, 23)  [2,3] m2  [4,5] idx  [6 + h*SW + g] sign words
    const __half2 kbias = __float2half2_rn(1407.0f), k1024 = __float2half2_rn(1024.0f);
    const __half2 zero = __float2half2_rn(0.0f), two = __float2half2_rn(2.0f), eight = __float2half2_rn(8.0f);
    const __half2 sent = __float2half2_rn(2047.0f);
    __half2 dold[2];
    if (!FIRST) {
        dold[0] = __hsub2(u2h(rw[2]), u2h(rw[0]));
        dold[1] = __hsub2(u2h(rw[3]), u2h(rw[1]));
    }
    unsigned addr[DC];
    __half2 tp[2][DC];
    __half2 k1[2][NS], k2[2][NS], kprev[2];
    __half2 cnt[2] = {zero, zero};  // !LDPC_PARITY_XOR: number of negative t per lane
    unsigned par[2] = {0u, 0u};     // LDPC_PARITY_XOR: xor of the t patterns, bit 15 of a lane = sign parity
    __half2 sacc[2][SW];
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll
        for (int g = 0; g < SW; g++) sacc[h][g] = zero;
#pragma unroll
        for (int s = 0; s < NS; s++) k1[h][s] = k2[h][s] = sent;
    }
#pragma unroll
    for (int k = 0; k < DC; k++) {
        if (EXACT || k < dc) {
            const int s = k >> 3, kl = k & 7;
            const __half2 kh = __float2half2_rn((float)k), klh = __float2half2_rn((float)kl);
            const int2 e = p.tab[off + k];  // {column-block base + shift (bytes), wrap threshold (Z - s) * 4}
            unsigned a = isb + (unsigned)e.x;
            a -= (i4 >= e.y) ? (unsigned)Z4 : 0u;
            addr[k] = a;
            const unsigned wk = lds32(a);
            const int sh = 15 - sign_bit_pos(dc, k);  // brings edge k's sign bit to bit 15 of each lane
#pragma unroll
            for (int h = 0; h < 2; h++) {
                __half2 t1 = u2h(prmt(wk, p.c65, h ? 0x4342u : 0x4140u));
                if (!FIRST) {
                    const __half2 mag = __hfma2(__heq2(u2h(rw[4 + h]), kh), dold[h], u2h(rw[h]));
                    const unsigned sg = (rw[6 + h * SW + k / kSignGroup] << sh) & 0x80008000u;
                    t1 = __hsub2(t1, u2h(h2u(mag) ^ sg));
                }
                tp[h][k] = t1;
                const __half2 tt = __hsub2(t1, kbias);
                const __half2 neg = LDPC_NEG_ALU ? __hlt2(tt, zero) : neg01(tt);  // 1.0 where t < 0
                sacc[h][k / kSignGroup] = __hfma2(sacc[h][k / kSignGroup], two, neg);
                if (LDPC_PARITY_XOR)
                    par[h] ^= h2u(tt);
                else
                    cnt[h] = __hadd2(cnt[h], neg);
                const __half2 key = __hfma2(__habs2(tt), eight, klh);
                if (!EXACT) {
                    k2[h][s] = umin2(k2[h][s], umax2(k1[h][s], key));
                    k1[h][s] = umin2(k1[h][s], key);
                } else if ((kl & 1) == 0) {
                    if (k == DC - 1) {  // unpaired last edge
                        k2[h][s] = umin2(k2[h][s], umax2(k1[h][s], key));
                        k1[h][s] = umin2(k1[h][s], key);
                    } else {
                        kprev[h] = key;
                    }
                } else {  // pair (k-1, k): 5 min/max for two edges, 2 for the first pair of a set
                    const __half2 lo = umin2(kprev[h], key), hi = umax2(kprev[h], key);
                    if (kl == 1) {
                        k1[h][s] = lo;
                        k2[h][s] = hi;
                    } else {
                        k2[h][s] = umin3(k2[h][s], hi, umax2(k1[h][s], lo));
                        k1[h][s] = umin2(k1[h][s], lo);
                    }
                }
            }
        }
    }
    // the old record is dead: start the load of the next step's record, ~150 instructions of row finalisation
    // and the whole phase 2 before its first use
    if (LDPC_REC_PRELOAD == 2 && ld_next) rec_load<U4>(nx, rw);
    unsigned nsg[2][SW], rwn[U4 * 4];
#pragma unroll
    for (int w = 6 + 2 * SW; w < U4 * 4; w++) rwn[w] = 0u;  // padding words of the record
    __half2 min1[2], idx[2], dnew[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        __half2 m1 = zero, m2 = zero, ix = zero;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const __half2 c1 = umin2(k1[h][s], amax8), c2 = umin2(k2[h][s], amax8p7);
            const __half2 a1 = floor8(c1), a2 = floor8(c2);
            const __half2 ia = __hfma2(a1, __float2half2_rn(-8.0f), c1);
            if (s == 0) {
                m1 = a1;
                m2 = a2;
                ix = ia;
            } else {  // strict "<": the earlier set keeps the index on equal minima (first index)
                const __half2 lt = __hlt2(a1, m1);
                ix = __hfma2(lt, __hsub2(__hadd2(ia, __float2half2_rn(8.0f * s)), ix), ix);
                m2 = umin2(umax2(m1, a1), umin2(m2, a2));
                m1 = umin2(m1, a1);
            }
        }
        if (p.scale_on) {
            m1 = beta_scale(m1, bmul, nbias);
            m2 = beta_scale(m2, bmul, nbias);
        }
        min1[h] = m1;
        idx[h] = ix;
        dnew[h] = __hsub2(m2, m1);
        // parity of the negative-sign count -> 0xFFFF per lane; sign bits of all dc edges flip with it
        const unsigned pm = LDPC_PARITY_XOR ? ((par[h] >> 15) & 0x00010001u) * 0xFFFFu
                                            : (h2u(__hadd2(cnt[h], k1024)) & 0x00010001u) * 0xFFFFu;
#pragma unroll
        for (int g = 0; g < SW; g++) {
            int gs = dc - g * kSignGroup;
            gs = gs < 0 ? 0 : (gs > kSignGroup ? kSignGroup : gs);
            const unsigned gm = ((1u << gs) - 1u) * 0x00010001u;
            nsg[h][g] = (h2u(__hadd2(sacc[h][g], k1024)) ^ pm) & gm;
        }
        rwn[h] = h2u(m1);
        rwn[2 + h] = h2u(m2);
        rwn[4 + h] = h2u(ix);
#pragma unroll
        for (int g = 0; g < SW; g++) rwn[6 + h * SW + g] = nsg[h][g];
    }
    rec_store<U4>(recp, rwn);
    if (LDPC_REC_PRELOAD == 1 && ld_next) rec_load<U4>(nx, rw);

#pragma unroll
    for (int k = 0; k < DC; k++) {
        if (EXACT || k < dc) {
            const __half2 kh = __float2half2_rn((float)k);
            const int sh = 15 - sign_bit_pos(dc, k);
            unsigned v[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const __half2 mag = __hfma2(__heq2(idx[h], kh), dnew[h], min1[h]);
                const unsigned sg = (nsg[h][k / kSignGroup] << sh) & 0x80008000u;
                const __half2 x = __hadd2(tp[h][k], u2h(h2u(mag) ^ sg));
                // pattern(x) = 0x6400 + (APP_new + 383): subtract 0x6500, clamp to [0, 254]
                v[h] = __viaddmin_s16x2_relu(h2u(x), 0x9B009B00u, 0x00FE00FEu);
            }
            sts32(addr[k], prmt(v[0], v[1], 0x6420u));
        }
    }
}

// Address of the record a thread needs in its NEXT step (its next row of this layer, else its first row of
// the next layer / next iteration) and whether that step reads a record at all; optional L1/L2 prefetch of
// it (LDPC_REC_PREFETCH).  History: prefetching a whole layer ahead into L1 (80 KB in flight for
// J15_L30_Z1280 against ~64 KB of L1) lost most lines before their use — ncu showed 26 % of the stall
// samples as long-scoreboard waits on the record load; one step ahead was better, the register preload
// (process_row_x) better still.
template <int DCHI, bool FIRST>
__device__ __forceinline__ bool prefetch_next_record(const uint4 *recl, const uint4 *nxl, int i, int T, int Z,
                                                     bool pf_next_layer, const uint4 *&nx)
{
    constexpr int RS = RecLayout<DCHI>::STRIDE;
    const bool same_layer = i + T < Z;
    nx = same_layer ? recl + (size_t)(i + T) * RS : nxl + (size_t)threadIdx.x * RS;
    const bool want = same_layer ? !FIRST : pf_next_layer;
    if (LDPC_REC_PREFETCH == 1 && want) {
        prefetch_l1(nx);
        if (RecLayout<DCHI>::U4 > 2) prefetch_l1(nx + 2);
    }
    if (LDPC_REC_PREFETCH == 2 && want) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
    return want;
}

// Rows of a layer whose degree is outside the exact range of the kernel's bucket.  Out of line so
// that its register needs do not leak into the allocation of the hot exact-degree paths.
template <int DCHI, bool FIRST>
__device__ __noinline__ void generic_rows(unsigned sbase, const LayeredParams &p, int off, int dc, uint4 *recl,
                                          const uint4 *nxl, bool pf_next_layer, __half2 amax8, __half2 amax8p7,
                                          __half2 bmul, __half2 nbias)
{
    constexpr int RS = RecLayout<DCHI>::STRIDE;
    const int Z = p.Z, Z4 = 4 * Z, T = blockDim.x;
    for (int i = threadIdx.x; i < Z; i += T) {
        const uint4 *nx;
        prefetch_next_record<DCHI, FIRST>(recl, nxl, i, T, Z, pf_next_layer, nx);
        unsigned rw[RecLayout<DCHI>::U4 * 4];  // out of line: this path loads its record itself
        if (!FIRST && LDPC_REC_PRELOAD) rec_load<RecLayout<DCHI>::U4>(recl + (size_t)i * RS, rw);
        process_row_x<DCHI, DCHI, FIRST, false>(sbase + 4u * i, 4 * i, p, off, dc, Z4, recl + (size_t)i * RS, nx, false,
                                                rw, amax8, amax8p7, bmul, nbias);
    }
}

// One full iteration: all layers in order, the Z rows of a layer spread over the CTA.
template <int DCHI, bool FIRST>
__device__ __forceinline__ void sweep_layers(unsigned sbase, const LayeredParams &p, uint4 *rec, bool last,
                                             unsigned (&rw)[RecLayout<DCHI>::U4 * 4], __half2 amax8,
                                             __half2 amax8p7, __half2 bmul, __half2 nbias)
{
    constexpr int RS = RecLayout<DCHI>::STRIDE;
    const int tid = threadIdx.x, T = blockDim.x, Z = p.Z, Z4 = 4 * Z;
    for (int r = 0; r < p.J; r++) {
        const int dc = p.dc[r], off = p.off[r];
        uint4 *recl = rec + (size_t)r * Z * RS;
        // this thread's first row of the next layer (of the next iteration after the last layer); it is
        // prefetched only if a sweep that reads records follows
        const uint4 *nxl = rec + (size_t)((r + 1 == p.J) ? 0 : r + 1) * Z * RS;
        const bool pf_next_layer = (!FIRST || r == p.J - 1) && !(last && r == p.J - 1);
#define LDPC_ROWS(DCX)                                                                                        \
    for (int i = tid; i < Z; i += T) {                                                                        \
        const uint4 *nx;                                                                                      \
        const bool ldn = prefetch_next_record<DCHI, FIRST>(recl, nxl, i, T, Z, pf_next_layer, nx);            \
        process_row_x<DCX, DCHI, FIRST, true>(sbase + 4u * i, 4 * i, p, off, DCX, Z4, recl + (size_t)i * RS, nx, ldn, \
                                              rw, amax8, amax8p7, bmul, nbias);                               \
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
            generic_rows<DCHI, FIRST>(sbase, p, off, dc, recl, nxl, pf_next_layer, amax8, amax8p7, bmul, nbias);
            if (LDPC_REC_PRELOAD && pf_next_layer && tid < Z) rec_load<RecLayout<DCHI>::U4>(nxl + (size_t)tid * RS, rw);
        }
#undef LDPC_ROWS
        __syncthreads();
        LDPC_STRESS_POINT(1);
    }
}

// fail bits (0x80 per frame byte) of the checks this thread owns in layer r
__device__ __forceinline__ unsigned syndrome_row(const unsigned char *app, const LayeredParams &p, int off, int dc,
                                                 int i4, int Z4)
{
    unsigned x = (dc & 1) ? 0x80808080u : 0u;  // byte = APP + 127 (<= 254): negative <=> bit 7 of byte + 1 clear
    for (int k = 0; k < dc; k++) {
        const int e = off + k;
        int col4 = i4 + 4 * (int)p.lt.shift[e];
        col4 -= (col4 >= Z4) ? Z4 : 0;
        x ^= *reinterpret_cast<const unsigned *>(app + (int)p.lt.col[e] * Z4 + col4) + 0x01010101u;
    }
    return x & 0x80808080u;
}

// hard decisions of frames selected by `fmask` (bit j = frame 4g+j) -> global memory
__device__ void write_outputs(const unsigned *appw, const LayeredParams &p, int g, unsigned fmask)
{
    const int F = p.F, N = p.N, f0 = 4 * g;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
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

template <int DCMAX>
__device__ void dump_records(const LayeredParams &p, const uint4 *rec, int g, unsigned fmask)
{
    constexpr int SW = RecLayout<DCMAX>::SW;
    constexpr int RS = RecLayout<DCMAX>::STRIDE;
    const unsigned *rw = reinterpret_cast<const unsigned *>(rec);
    for (int m = threadIdx.x; m < p.M; m += blockDim.x) {
        const unsigned *r = rw + (size_t)m * RS * 4;
        const int dc = p.dc[m / p.Z];
        for (int j = 0; j < 4; j++) {
            if (!((fmask >> j) & 1u)) continue;
            const int h = j >> 1, sh = (j & 1) * 16;
            const __half m1 = __ushort_as_half((unsigned short)(r[h] >> sh));
            const __half m2 = __ushort_as_half((unsigned short)(r[2 + h] >> sh));
            const __half ix = __ushort_as_half((unsigned short)(r[4 + h] >> sh));
            unsigned sg = 0;
            for (int k = 0; k < dc; k++)
                sg |= ((r[6 + h * SW + k / kSignGroup] >> (sh + sign_bit_pos(dc, k))) & 1u) << k;
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
    static constexpr int value = DCMAX <= 12 ? 640 : (DCMAX <= 24 ? 512 : 384);
};

template <int DCMAX>
__global__ void __launch_bounds__(ThreadCap<DCMAX>::value, 1)
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
    uint4 *rec = p.rec + (size_t)blockIdx.x * p.M * RecLayout<DCMAX>::STRIDE;
    const __half2 amax8 = __float2half2_rn(8.0f * p.msg_max);  // key clamps: min1 -> 8*amax, min2 -> 8*amax + 7
    const __half2 amax8p7 = __float2half2_rn(8.0f * p.msg_max + 7.0f);
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
                        appw[n] = (unsigned)(quant(v[u].x, p.scale) + 127) | ((unsigned)(quant(v[u].y, p.scale) + 127) << 8) |
                                  ((unsigned)(quant(v[u].z, p.scale) + 127) << 16) |
                                  ((unsigned)(quant(v[u].w, p.scale) + 127) << 24);
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
                        appw[n] = (unsigned)(quant(a.x, p.scale) + 127) | ((unsigned)(quant(a.y, p.scale) + 127) << 8) |
                                  ((unsigned)(quant(b.x, p.scale) + 127) << 16) | ((unsigned)(quant(b.y, p.scale) + 127) << 24);
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
        if (LDPC_L2_PREFETCH && !dynamic && g + (int)gridDim.x < p.num_groups && p.llr_dtype == LDPC_DTYPE_FP32 &&
            p.layout == LDPC_LAYOUT_NF) {
            const float *y = reinterpret_cast<const float *>(p.llr) + 4 * (size_t)(g + gridDim.x);
            for (int n = tid; n < N; n += T) asm volatile("prefetch.global.L2 [%0];" ::"l"(y + (size_t)n * F));
        }

        unsigned running = valid;  // frames not yet latched
        int it = 0;
        unsigned rw[RecLayout<DCMAX>::U4 * 4];  // the record of the thread's next step (LDPC_REC_PRELOAD)
        while (it < p.iters) {
            it++;
            if (it == 1)
                sweep_layers<DCMAX, true>(sbase, p, rec, it == p.iters, rw, amax8, amax8p7, bmul, nbias);
            else
                sweep_layers<DCMAX, false>(sbase, p, rec, it == p.iters, rw, amax8, amax8p7, bmul, nbias);
            if (p.exit_mode == LDPC_EXIT_SYNDROME || it == p.iters) {
                // The pass stops after the first layer that leaves every running frame with a failed check (the
                // usual case until the last iterations: one layer of J is read instead of all) — the outcome
                // "not converged" is then already exact for all of them.
                unsigned rbytes = 0u;  // 0x80 in the byte of every running frame
#pragma unroll
                for (int j = 0; j < 4; j++) rbytes |= ((running >> j) & 1u) ? (0x80u << (8 * j)) : 0u;
                unsigned fl = 0u;
                for (int r = 0; r < p.J; r++) {
                    const int dc = p.lt.dc[r], off = p.lt.off[r];
                    unsigned fail = 0u;
                    for (int i = tid; i < Z; i += T) fail |= syndrome_row(smem, p, off, dc, 4 * i, Z4);
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
                if (finish) {
                    LDPC_STRESS_POINT(3);
                    write_outputs(appw, p, g, finish);
                    if (p.dbg_rec) dump_records<DCMAX>(p, rec, g, finish);
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
#define X(D) case D: cap = ThreadCap<D>::value; pl->u4 = RecLayout<D>::STRIDE; break;
        LDPC_BUCKETS(X)
#undef X
        default: return LDPC_ERR_UNSUPPORTED;
    }
    // threads: the Z rows of a layer spread over whole warps, at most ThreadCap<DCHI> threads
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
    p.dbg_rec = reinterpret_cast<unsigned *>(a.dbg_rec);
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
