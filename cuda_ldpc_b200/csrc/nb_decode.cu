// Batched GF(q) LDPC decoding: EMS, trellis min-max (TMM) and layered TMM, one CTA per frame,
// one launch per batch.  Replaces Demodulate + Decoding_EMS(_GPU) / Decoding_TMM(_GPU) /
// Decoding_layered_TMM (NB/src/LDPC_Decoder.cpp:132-702, NB/src/Decode_GPU.cu:138-1070), which
// decode ONE frame per call with a kernel launch per node phase, cudaMalloc/cudaFree, in-kernel
// malloc and a D2H copy + host syndrome check per iteration.
//
// Numerics: fp32, same operation order as the oracle (oracle/nb_oracle.c, `fresh` EMS sums;
// TMM bit-for-bit the reference's order), --fmad=false, the 1/1.2 and 0.8 scalings in double
// like the reference (NB/src/LDPC_Decoder.cpp:309,519).
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.h"
#include "nb_common.h"
#include "stress.cuh"

namespace ldpcb {

// every CTA barrier of this file; race-hunting builds (-DLDPC_STRESS=1, stress.cuh) skew the warps behind it
__device__ __forceinline__ void cta_sync()
{
    __syncthreads();
    LDPC_STRESS_POINT(0);
}

constexpr int kNbThreads = 256;      // CTA width when the frame state lives in the global scratch slot
constexpr int kNbThreadsMax = 1024;  // ... and the widest CTA tried when it fits in shared memory
constexpr int kNmMax = 4;  // EMS_NM supported on the device path

struct NbParams {
    const void *in;
    uint16_t *out;
    int *iters_out, *ok_out;
    float *scratch;        // per CTA slot (global memory), unused when slot_in_smem
    size_t slot_floats;
    int slot_in_smem;      // the frame state sits in shared memory behind the algorithm's work arrays
    size_t work_floats;    // size of those work arrays
    int F, N, M, q, p, dv_max, dc_max, n_const;
    int algo, in_kind, maxit, nm, nc;
    float sigma;
    int ems_chunk;         // EMS check tasks per shared-memory chunk
    int ems_full;          // log-QSPA (NB/src/Simulation.cpp:64-67): every configuration, Nm = q, Nc = dc - 1
    const uint16_t *mul, *inv;
    const int *vw, *cw, *v_cn, *v_pos, *c_vn, *c_gf, *c_pos;
    const float *cre, *cim;
    int *work_counter;     // dynamic frame scheduling: frames past the first gridDim.x are handed out by an atomic counter
};

// GF(q) multiply table.  Default: the read-only path (LDG.CONSTANT; 8 KB for GF(64), 128 KB for GF(256) stay
// L1/L2-resident).  NB_MUL_SMEM=1 keeps a copy in shared memory for q <= 64 (what BASELINE.json's north star
// asks for) — measured on B200 it is SLOWER (C4 EMS 16.2 -> 17.4 ms, TMM 7.5 -> 7.7, layered TMM 4.1 -> 4.6 ms per
// 8192 frames): L1 and shared memory are the same SRAM, the LDS saves nothing over an L1 hit, and 8 KB per CTA
// costs the q-thread layered CTAs their 32-per-SM occupancy.
#ifndef NB_MUL_SMEM
#define NB_MUL_SMEM 0
#endif
constexpr int kMulSmemQ = 64;
__shared__ uint16_t g_smul[NB_MUL_SMEM ? kMulSmemQ * kMulSmemQ : 1];
constexpr size_t kNbDynSmemMax = (size_t)227 * 1024 - (NB_MUL_SMEM ? sizeof(uint16_t) * kMulSmemQ * kMulSmemQ : 0) - 64;  // minus the static part
__device__ __forceinline__ int gmul(const NbParams &p, int a, int b)
{
    if (NB_MUL_SMEM && p.q <= kMulSmemQ) return g_smul[a * p.q + b];
    return __ldg(p.mul + a * p.q + b);
}

// ---- Demodulate, NB/src/LDPC_Decoder.cpp:132-171 -------------------------------------------------
// Thread = (symbol slot, field element) when the CTA width is a multiple of q (every launch shape of this file): no
// division per element.  BPSK: the p per-bit terms -2 y_b / sigma^2 of a symbol are computed once into `tmp` (N * p
// floats, the not yet initialised LLR area) instead of once per field element that contains the bit — the same terms
// added in the same ascending-bit order, so L_ch is bit-identical (255 x 4 divisions per GF(256) symbol became 8).
__device__ void demodulate(const NbParams &p, int f, float *lch, float *tmp)
{
    const int q = p.q, N = p.N, tid = threadIdx.x, T = blockDim.x;
    const bool grid2d = (T % q) == 0;
    const int x0 = grid2d ? tid % q : 0, s0 = grid2d ? tid / q : 0, sstep = grid2d ? T / q : 1;
    if (p.in_kind == NB_IN_SYMBOL_LLR) {
        const float *src = reinterpret_cast<const float *>(p.in) + (size_t)f * N * (q - 1);
        for (int i = tid; i < N * (q - 1); i += T) lch[i] = src[i];
    } else if (p.in_kind == NB_IN_BPSK) {
        const float *rx = reinterpret_cast<const float *>(p.in) + (size_t)f * N * p.p;
        const float s2 = __fmul_rn(p.sigma, p.sigma);
        for (int i = tid; i < N * p.p; i += T) tmp[i] = __fdiv_rn(__fmul_rn(-2.0f, rx[i]), s2);
        cta_sync();
        if (grid2d) {
            if (x0 >= 1)
                for (int s = s0; s < N; s += sstep) {
                    float v = 0.0f;
                    for (int b = 0; b < p.p; b++)
                        if (x0 & (1 << b)) v = __fadd_rn(v, tmp[s * p.p + b]);
                    lch[s * (q - 1) + x0 - 1] = v;
                }
        } else {
            for (int i = tid; i < N * (q - 1); i += T) {
                const int s = i / (q - 1), a = i - s * (q - 1) + 1;
                float v = 0.0f;
                for (int b = 0; b < p.p; b++)
                    if (a & (1 << b)) v = __fadd_rn(v, tmp[s * p.p + b]);
                lch[i] = v;
            }
        }
        cta_sync();  // tmp (= the LLR area) is free again
    } else {
        const float *rx = reinterpret_cast<const float *>(p.in) + (size_t)f * N * 2;
        const float den = __fmul_rn(__fmul_rn(2.0f, p.sigma), p.sigma);
        const float c0r = p.cre[0], c0i = p.cim[0];
        auto one = [&](int s, int a) {
            const float yr = rx[2 * s], yi = rx[2 * s + 1], car = p.cre[a], cai = p.cim[a];
            const float tr = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(2.0f, yr), c0r), car), __fsub_rn(car, c0r));
            const float ti = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(2.0f, yi), c0i), cai), __fsub_rn(cai, c0i));
            lch[s * (q - 1) + a - 1] = __fdiv_rn(__fadd_rn(tr, ti), den);
        };
        if (grid2d) {
            if (x0 >= 1)
                for (int s = s0; s < N; s += sstep) one(s, x0);
        } else {
            for (int i = tid; i < N * (q - 1); i += T) one(i / (q - 1), i % (q - 1) + 1);
        }
    }
}

// syndrome of the decided symbols; sets *fail != 0 when any check is violated
__device__ void syndrome(const NbParams &p, const uint16_t *sym, int *fail)
{
    for (int row = threadIdx.x; row < p.M; row += blockDim.x) {
        int s = 0;
        for (int i = 0; i < p.cw[row]; i++)
            s ^= gmul(p, sym[p.c_vn[row * p.dc_max + i]], p.c_gf[row * p.dc_max + i]);
        if (s) *fail = 1;
    }
}

// ---- EMS, NB/src/LDPC_Decoder.cpp:172-359 -------------------------------------------------------
// Warp-centric: a warp owns a variable node (lanes = symbols) in the node phase and a check row (lanes =
// syndromes) in the check phase, so every access to the q-vectors of the frame state is one coalesced
// warp-wide load/store and no thread ever loops serially over q.
//
// Check node: E_d[s] = max over conf(q,1) U conf(Nm,Nc) of the fresh sum (inputs summed from 0.0f in
// ascending edge position, the oracle's `fresh` rule) for every output edge d of the row.  Lane l owns the
// syndromes s = l, l+32, ...: for a fixed deviating input j the map a -> s is a bijection (h_j != 0), so
// every lane finds "its" symbol a = h_j^-1 (s ^ s_rest) directly and no two lanes write the same E[s].
// The v2c vectors of the row's edges are staged in shared memory once per task.
struct EmsWarpShared {   // per warp, in shared memory
    float *E;             // [q]
    float *v;             // [dc_max][q] v2c vectors of the row's edges
    int *ih;              // [32] coefficient of edge b
    int *cb, *hi;         // [32] h_b * best_b and h_b^-1
    int *ds;              // [32][kNmMax] syndrome change when edge b moves from its best to its k-th symbol
    float *tval;          // [32][kNmMax] best values of every edge
    int *tsym;            // [32][kNmMax] and their symbols
};
__host__ __device__ constexpr size_t ems_warp_floats(int q, int dc_max)
{
    return (size_t)q + (size_t)dc_max * q + 3 * 32 + 32 * kNmMax * 3;
}

// E[s] = max(E[s], v) on a float in shared memory (any signs, -inf included): non-negative floats order like
// signed ints, negative floats like reversed unsigned ints
__device__ __forceinline__ void atomic_max_float(float *addr, float v)
{
    if (v >= 0.0f)
        atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned *>(addr), __float_as_uint(v));
}

// output edges [d0, d1) of check `row`, one warp.  W = the row's degree as a compile-time constant (every loop
// over edges unrolls, the output edge d becomes a constant and "i != d" disappears), 0 = any degree.
template <int W>
__device__ void ems_check_row_warp(const NbParams &p, int row, int d0, int d1, const float *v2c, const uint16_t *topsym,
                                   const float *topval, float *c2v, const EmsWarpShared &ws, int lane)
{
    const int q = p.q, w = W ? W : p.cw[row], n = w - 1;
    if (lane < w) {
        const int e = p.c_vn[row * p.dc_max + lane] * p.dv_max + p.c_pos[row * p.dc_max + lane];
        const int h = p.c_gf[row * p.dc_max + lane];
        const int cb = gmul(p, topsym[e * kNmMax], h);
        ws.ih[lane] = h;
        ws.cb[lane] = cb;
        ws.hi[lane] = __ldg(p.inv + h);
        for (int k = 0; k < p.nm; k++) {
            ws.tval[lane * kNmMax + k] = topval[e * kNmMax + k];
            ws.tsym[lane * kNmMax + k] = topsym[e * kNmMax + k];
            ws.ds[lane * kNmMax + k] = cb ^ gmul(p, topsym[e * kNmMax + k], h);
        }
    }
    for (int b = 0; b < w; b++) {  // stage the v2c vectors (coalesced)
        const float *src = v2c + ((size_t)p.c_vn[row * p.dc_max + b] * p.dv_max + p.c_pos[row * p.dc_max + b]) * q;
        for (int a = lane; a < q; a += 32) ws.v[b * q + a] = src[a];
    }
    __syncwarp();
    int s_all = 0;
#pragma unroll
    for (int b = 0; b < w; b++) s_all ^= ws.cb[b];
    const int Nc = (p.nc == p.dc_max - 1) ? w - 1 : p.nc;  // :297-304
#pragma unroll
    for (int d = W ? 0 : d0; d < (W ? W : d1); d++) {
        if (W && (d < d0 || d >= d1)) continue;
        // inputs = the other edges in ascending position
        const int s0 = s_all ^ ws.cb[d];
        float sum_top = 0.0f;
#pragma unroll
        for (int b = 0; b < w; b++)
            if (b != d) sum_top = __fadd_rn(sum_top, ws.tval[b * kNmMax]);
        for (int s = lane; s < q; s += 32) {
            float e = (s == s0) ? sum_top : -INFINITY;
            // conf(q,1): input j at the symbol a that makes the syndrome s, everybody else at their best.  a == best
            // only happens for s == s0 and then reproduces sum_top exactly, so it needs no branch.
#pragma unroll
            for (int j = 0; j < w; j++) {
                if (j == d) continue;
                const int hj = ws.ih[j];
                if (hj != 0) {
                    const int a = gmul(p, ws.hi[j], s ^ s0 ^ ws.cb[j]);
                    const float va = ws.v[j * q + a];
                    float sum = 0.0f;
#pragma unroll
                    for (int i = 0; i < w; i++)
                        if (i != d) sum = __fadd_rn(sum, (i == j) ? va : ws.tval[i * kNmMax]);
                    e = fmaxf(e, sum);
                } else if (s == s0) {  // coefficient 0 (raw *_exp files): every symbol of input j lands on s0
                    const int best = ws.tsym[j * kNmMax];
                    for (int a = 0; a < q; a++) {
                        if (a == best) continue;
                        float sum = 0.0f;
                        for (int i = 0; i < w; i++)
                            if (i != d) sum = __fadd_rn(sum, (i == j) ? ws.v[j * q + a] : ws.tval[i * kNmMax]);
                        e = fmaxf(e, sum);
                    }
                }
            }
            ws.E[s] = e;
        }
        __syncwarp();
        // conf(Nm,Nc): at most Nc inputs at sorted index 1..Nm-1.  The all-best leaf and every single deviation
        // are already in conf(q,1) (same fresh sums), so for Nc <= 2 only the PAIRS of deviating inputs remain:
        // C(n,2) (Nm-1)^2 leaves, dealt round-robin to the lanes; leaves of different lanes that land on the
        // same syndrome meet in an atomic max (max is order-independent, the result stays exact).
        if (Nc <= 2) {
            const int nk = p.nm - 1, per_pair = nk * nk;
            const int leaves = (Nc == 2) ? (n * (n - 1) / 2) * per_pair : 0;
            for (int L = lane; L < leaves; L += 32) {
                // Nm = 2 (the reference's EMS_NM): one leaf per pair, no index arithmetic — the two runtime divisions
                // below were 8 % of the kernel's instructions for 3 leaves per output edge
                int P = (nk == 1) ? L : L / per_pair;
                const int kk = (nk == 1) ? 0 : L - P * per_pair;
                const int k1 = (nk == 1) ? 1 : 1 + kk / nk, k2 = (nk == 1) ? 1 : 1 + kk - (kk / nk) * nk;
                int j1 = 0;
                for (int cnt = n - 1; P >= cnt; cnt--) {
                    P -= cnt;
                    j1++;
                }
                int j2 = j1 + 1 + P;
                j1 += (j1 >= d);  // input index -> edge position
                j2 += (j2 >= d);
                const int sy = s0 ^ ws.ds[j1 * kNmMax + k1] ^ ws.ds[j2 * kNmMax + k2];
                float sum = 0.0f;
#pragma unroll
                for (int i = 0; i < w; i++)
                    if (i != d) sum = __fadd_rn(sum, ws.tval[i * kNmMax + ((i == j1) ? k1 : ((i == j2) ? k2 : 0))]);
                atomic_max_float(ws.E + sy, sum);
            }
        } else {  // general budget: every lane walks the same odometer (digits packed 2 bits per input) and
                  // evaluates every 32nd leaf
            unsigned long long ks = 0;
            int diff = 0, leaf = 0;
            while (true) {
                if ((leaf & 31) == lane) {
                    int sy = s0;
                    float sum = 0.0f;
                    for (int j = 0; j < n; j++) {
                        const int k = (int)((ks >> (2 * j)) & 3ull), bj = j + (j >= d);
                        sy ^= ws.ds[bj * kNmMax + k];
                        sum = __fadd_rn(sum, ws.tval[bj * kNmMax + k]);
                    }
                    atomic_max_float(ws.E + sy, sum);
                }
                leaf++;
                int j = n - 1;
                for (; j >= 0; j--) {  // increment digit j; reset and carry when it overflows or breaks the budget
                    const int k = (int)((ks >> (2 * j)) & 3ull);
                    if (k == 0) diff++;
                    if (k + 1 < p.nm && diff <= Nc) {
                        ks += 1ull << (2 * j);
                        break;
                    }
                    ks &= ~(3ull << (2 * j));
                    diff--;
                }
                if (j < 0) break;
            }
        }
        __syncwarp();
        const int h = ws.ih[d];
        float *m = c2v + ((size_t)row * p.dc_max + d) * q;
        const float e0 = ws.E[0];
        for (int k = 1 + lane; k < q; k += 32) {
            const float df = __fsub_rn(ws.E[gmul(p, k, h)], e0);
            m[k - 1] = (float)((double)df / 1.2);  // float difference, double division (:309)
        }
        __syncwarp();
    }
}

// log-QSPA, the reference's decoder_method 2 = Decoding_EMS(..., GFQ, maxdc - 1, ...) (NB/src/Simulation.cpp:64-67):
// every one of the q^(dc-1) configurations.  The fresh sum of a configuration is a left-to-right sum over the
// inputs, and floating-point addition is monotone in its first argument, so the maximum over all
// configurations with syndrome s equals — bit for bit — the forward (max,+) recursion
//   F_0[s] = (s == 0 ? 0.0f : -inf),   F_j[s] = max_t ( F_{j-1}[s ^ t] + v_j[h_j^-1 t] ),
// O(dc q^2) per output edge instead of q^(dc-1).  One warp per row slice; lane l owns the syndromes l, l+32, ...
// (s ^ t keeps the lanes on distinct banks); the permuted vectors v_j[h_j^-1 t] are staged once per row.
__device__ void ems_check_row_full_warp(const NbParams &p, int row, int d0, int d1, const float *v2c, float *c2v,
                                        const EmsWarpShared &ws, int lane)
{
    const int q = p.q, w = p.cw[row];
    float *Fa = ws.E, *Fb = reinterpret_cast<float *>(ws.ds);  // 3 * 32 * kNmMax >= q words
    if (lane < w) ws.ih[lane] = p.c_gf[row * p.dc_max + lane];
    __syncwarp();
    for (int b = 0; b < w; b++) {
        const float *src = v2c + ((size_t)p.c_vn[row * p.dc_max + b] * p.dv_max + p.c_pos[row * p.dc_max + b]) * q;
        const int h = ws.ih[b];
        if (h != 0) {
            const int hinv = __ldg(p.inv + h);
            for (int t = lane; t < q; t += 32) ws.v[b * q + t] = src[gmul(p, hinv, t)];
        } else {  // coefficient 0 (raw *_exp files): every symbol contributes syndrome 0
            float mx = -INFINITY;
            for (int a = lane; a < q; a += 32) mx = fmaxf(mx, src[a]);
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            for (int t = lane; t < q; t += 32) ws.v[b * q + t] = (t == 0) ? mx : -INFINITY;
        }
    }
    __syncwarp();
    for (int d = d0; d < d1; d++) {
        float *F = Fa, *Fn = Fb;
        for (int s = lane; s < q; s += 32) F[s] = (s == 0) ? 0.0f : -INFINITY;
        __syncwarp();
        for (int j = 0; j < w; j++) {
            if (j == d) continue;
            const float *vp = ws.v + j * q;
            for (int s = lane; s < q; s += 32) {
                float e = -INFINITY;
#pragma unroll 4
                for (int t = 0; t < q; t++) e = fmaxf(e, __fadd_rn(F[s ^ t], vp[t]));
                Fn[s] = e;
            }
            __syncwarp();
            float *tmp = F;
            F = Fn;
            Fn = tmp;
        }
        const int h = ws.ih[d];
        float *m = c2v + ((size_t)row * p.dc_max + d) * q;
        const float e0 = F[0];
        for (int k = 1 + lane; k < q; k += 32) {
            const float df = __fsub_rn(F[gmul(p, k, h)], e0);
            m[k - 1] = (float)((double)df / 1.2);  // float difference, double division (:309)
        }
        __syncwarp();
    }
}

// order-preserving float -> uint map (x + 0.0f folds -0 into +0 first), for warp reductions with redux.sync
__device__ __forceinline__ unsigned ord_key(float x)
{
    const unsigned u = __float_as_uint(__fadd_rn(x, 0.0f));
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// (value, list position) order of the stable descending sort of [1, 2, ..., q-1, 0] (BubleSort :17-36)
__device__ __forceinline__ bool ems_before(float xa, int ia, float xb, int ib) { return xa > xb || (xa == xb && ia < ib); }

// Variable node `col`, one warp: LLR = L_ch + sum_d c2v_d, decision (DecideLLRVector :71-91), and — the frame
// not having converged is only known after the syndrome, so this work is wasted once per frame — v2c = LLR - c2v
// per edge with the first Nm entries of its sorted list.
// PER = symbols per lane (q <= 32 PER), DV = the node's degree as a constant (0 = any)
template <int PER, int DV>
__device__ void ems_var_node_warp(const NbParams &p, int col, const float *lch, const float *c2v, float *v2c,
                                  float *topval, uint16_t *topsym, uint16_t *sym, int lane)
{
    constexpr int kMaxPerLane = PER, per = PER;
    const int q = p.q, dv = DV ? DV : p.vw[col];
    float llr[kMaxPerLane];  // symbol a = lane + 32 t (a = 0: the implicit 0)
    const float *mv[kMaxDv];
#pragma unroll
    for (int d = 0; d < dv; d++)
        mv[d] = c2v + ((size_t)p.v_cn[col * p.dv_max + d] * p.dc_max + p.v_pos[col * p.dv_max + d]) * q;
    float mx = 0.0f;
    int best = 0;
#pragma unroll
    for (int t = 0; t < kMaxPerLane; t++) {
        const int a = lane + 32 * t;
        if (t < per && a >= 1 && a < q) {
            float v = lch[col * (q - 1) + a - 1];
#pragma unroll
            for (int d = 0; d < dv; d++) v = __fadd_rn(v, mv[d][a - 1]);
            llr[t] = v;
            if (v > mx) {  // ascending a within the lane: the first maximum stays
                mx = v;
                best = a;
            }
        } else {
            llr[t] = 0.0f;
        }
    }
    // warp argmax, lowest symbol on ties (the serial scan keeps the first strict maximum): two redux.sync
    {
        const unsigned key = ord_key(mx), wk = __reduce_max_sync(0xffffffffu, key);
        best = (int)__reduce_min_sync(0xffffffffu, key == wk ? (unsigned)best : 0xffffffffu);
        if (lane == 0) sym[col] = (uint16_t)((wk <= ord_key(0.0f)) ? 0 : best);
    }
#pragma unroll
    for (int d = 0; d < dv; d++) {
        const int e = col * p.dv_max + d;
        float x[kMaxPerLane];
        bool used[kMaxPerLane];
#pragma unroll
        for (int t = 0; t < kMaxPerLane; t++) {
            const int a = lane + 32 * t;
            used[t] = !(t < per && a < q);
            x[t] = (t < per && a >= 1 && a < q) ? __fsub_rn(llr[t], mv[d][a - 1]) : 0.0f;
            if (!used[t]) v2c[(size_t)e * q + a] = x[t];
        }
        for (int k = 0; k < p.nm; k++) {
            float bx = -INFINITY;
            int bi = 0x7fffffff, bt = -1;  // list position i: a-1 for a >= 1, q-1 for a = 0
#pragma unroll
            for (int t = 0; t < kMaxPerLane; t++) {
                const int a = lane + 32 * t, i = (a == 0) ? q - 1 : a - 1;
                if (!used[t] && (bt < 0 || ems_before(x[t], i, bx, bi))) {
                    bx = x[t];
                    bi = i;
                    bt = t;
                }
            }
            // warp winner: largest value, smallest list position among equals — one redux.sync each
            const unsigned key = bt >= 0 ? ord_key(bx) : 0u;
            const unsigned wk = __reduce_max_sync(0xffffffffu, key);
            const int wi = (int)__reduce_min_sync(0xffffffffu, (bt >= 0 && key == wk) ? (unsigned)bi : 0x7fffffffu);
            const float wx = bx;
            if (bt >= 0 && bi == wi) {  // this lane holds the winner: retire it and publish
#pragma unroll
                for (int t = 0; t < kMaxPerLane; t++)
                    if (t == bt) used[t] = true;
                topval[e * kNmMax + k] = wx;
                topsym[e * kNmMax + k] = (uint16_t)((wi == q - 1) ? 0 : wi + 1);
            }
        }
    }
}

__device__ void decode_ems(const NbParams &p, int f, float *lch, float *LLR, float *c2v, float *v2c, float *topval,
                           uint16_t *topsym, uint16_t *sym, float *smemE, int *s_fail)
{
    const int q = p.q, N = p.N, M = p.M, tid = threadIdx.x, T = blockDim.x;
    const int warp = tid >> 5, nwarps = T >> 5, lane = tid & 31;
    for (int i = tid; i < M * p.dc_max * q; i += T) c2v[i] = 0.0f;
    if (tid == 0) *s_fail = 0;
    cta_sync();
    // output edges of a row are split over several warps when there are fewer rows than 2 x warps (C5: M = 12)
    int split = (2 * nwarps + M - 1) / M;
    split = split < 1 ? 1 : (split > p.dc_max ? p.dc_max : split);
    int it = 0, ok = 0;
    while (it < p.maxit) {
        it++;
        for (int col = warp; col < N; col += nwarps) {
            const int dv = p.vw[col];
#define VN(PER)                                                                                                \
    if (dv == 2)                                                                                               \
        ems_var_node_warp<PER, 2>(p, col, lch, c2v, v2c, topval, topsym, sym, lane);                           \
    else if (dv == 3)                                                                                          \
        ems_var_node_warp<PER, 3>(p, col, lch, c2v, v2c, topval, topsym, sym, lane);                           \
    else                                                                                                       \
        ems_var_node_warp<PER, 0>(p, col, lch, c2v, v2c, topval, topsym, sym, lane);
            if (q <= 32) {
                VN(1)
            } else if (q <= 64) {
                VN(2)
            } else if (q <= 128) {
                VN(4)
            } else {
                VN(8)
            }
#undef VN
        }
        cta_sync();
        syndrome(p, sym, s_fail);
        cta_sync();
        const int fail = *s_fail;
        cta_sync();
        if (tid == 0) *s_fail = 0;
        if (fail == 0) {
            ok = 1;
            it--;  // the reference returns iterations-1 on success (:236)
            break;
        }
        float *base = smemE + (size_t)warp * ems_warp_floats(q, p.dc_max);
        EmsWarpShared ws;
        ws.E = base;
        ws.v = base + q;
        ws.ih = reinterpret_cast<int *>(ws.v + (size_t)p.dc_max * q);
        ws.cb = ws.ih + 32;
        ws.hi = ws.cb + 32;
        ws.ds = ws.hi + 32;
        ws.tval = reinterpret_cast<float *>(ws.ds + 32 * kNmMax);
        ws.tsym = reinterpret_cast<int *>(ws.tval + 32 * kNmMax);
        for (int t = warp; t < M * split; t += nwarps) {
            const int row = t / split, u = t - row * split, w = p.cw[row];
            const int d0 = (u * w) / split, d1 = ((u + 1) * w) / split;
            if (d0 >= d1) continue;
            if (p.ems_full) {
                ems_check_row_full_warp(p, row, d0, d1, v2c, c2v, ws, lane);
                continue;
            }
            switch (w) {
#define X(W) case W: ems_check_row_warp<W>(p, row, d0, d1, v2c, topsym, topval, c2v, ws, lane); break;
                X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#undef X
                default: ems_check_row_warp<0>(p, row, d0, d1, v2c, topsym, topval, c2v, ws, lane);
            }
        }
        cta_sync();
    }
    if (tid == 0) {
        if (p.iters_out) p.iters_out[f] = it;
        if (p.ok_out) p.ok_out[f] = ok;
    }
}

// ---- TMM / layered TMM, NB/src/LDPC_Decoder.cpp:361-817 -----------------------------------------
struct TmmShared {
    float *vv, *dU, *Min1, *Min2, *I, *Ev;  // per group: [dc_max*q], [dc_max*q], [q] x4
    int2 *MC;                               // [q] {Min1 bits, MinCol}
    int *MinCol, *Path, *Zn;                // [q], [2q], [dc_max + 1] (last = syndrome)
    int *tvn, *tgf, *tinv;                  // [dc_max] x3: the row's edge table (variable, coefficient, its inverse)
    float *mn;                              // [dc_max] minimum of every edge's v2c vector
};
__host__ __device__ constexpr size_t tmm_group_floats(int q, int dc_max)
{
    return ((size_t)2 * dc_max * q + 9 * q + dc_max + 1 + 4 * dc_max + 3) & ~(size_t)3;
}

// the q threads of a group process check `row`; a = thread's symbol index.  Every thread of the CTA
// calls this (same barriers); threads with act == false touch no memory.
//
// Latency (ncu, profiles/r02_ncu_nb_*.txt: long-scoreboard stalls 3-4 warps per issue cycle): the frame state
// lives in an L2-resident slot, so every dependent global load costs ~300 cycles.  The row's edge table
// (variable, coefficient, inverse coefficient — read again in every phase below) is therefore staged in
// shared memory once per row, and the LLR / c2v vectors of the row's edges are loaded four edges at a time
// before the first of them is used (a loop of dependent load -> subtract -> store pairs serialised them): layered TMM
// C5 126 -> 140, BDS 159 -> 170, C4 660 -> 688 info Mbit/s.  An L1 prefetch of the next row's vectors on top LOSES 4 %.
__device__ void tmm_check(const NbParams &p, int row, int a, bool act, const TmmShared &s, float *LLR, float *c2v,
                          bool write_llr)
{
    const int q = p.q, w = act ? p.cw[row] : 0;
    if (act)
        for (int d = a; d < w; d += q) {
            const int h = p.c_gf[row * p.dc_max + d];
            s.tvn[d] = p.c_vn[row * p.dc_max + d];
            s.tgf[d] = h;
            s.tinv[d] = __ldg(p.inv + h);
        }
    cta_sync();
    // v2c = LLR - c2v (:472-475 / :644)
    constexpr int kBatch = 4;
    for (int d0 = 0; d0 < w; d0 += kBatch) {
        float l[kBatch], c[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; u++) {
            const int d = min(d0 + u, w - 1);  // tail: re-read the last edge instead of predicating
            l[u] = LLR[s.tvn[d] * q + a];
            c[u] = c2v[((size_t)row * p.dc_max + d) * q + a];
        }
#pragma unroll
        for (int u = 0; u < kBatch; u++)
            if (d0 + u < w) s.vv[(d0 + u) * q + a] = __fsub_rn(l[u], c[u]);
    }
    cta_sync();
    // d_TMM_Get_Zn :704-723: hard symbol of every edge = first minimum of its v2c vector.  q >= 32: one warp of
    // the group per edge (lanes scan x = lane, lane+32, ... then a warp argmin, lowest x on ties); dc may exceed q
    if (q >= 32) {
        const int lane = a & 31;
        for (int d = a >> 5; d < w; d += q >> 5) {
            float mn = INFINITY;
            int mx = 0;
            for (int x = lane; x < q; x += 32)
                if (s.vv[d * q + x] < mn) {
                    mn = s.vv[d * q + x];
                    mx = x;
                }
            for (int o = 16; o > 0; o >>= 1) {
                const float on = __shfl_xor_sync(0xffffffffu, mn, o);
                const int ox = __shfl_xor_sync(0xffffffffu, mx, o);
                if (on < mn || (on == mn && ox < mx)) {
                    mn = on;
                    mx = ox;
                }
            }
            if (lane == 0) {
                s.Zn[d] = gmul(p, mx, s.tgf[d]);
                s.mn[d] = mn;
            }
        }
    } else {
        for (int d = a; d < w; d += q) {
            float mn = INFINITY;
            int me = 0;
            const int h = s.tgf[d];
            for (int x = 0; x < q; x++)
                if (s.vv[d * q + x] < mn) {
                    mn = s.vv[d * q + x];
                    me = gmul(p, x, h);
                }
            s.Zn[d] = me;
            s.mn[d] = mn;
        }
    }
    cta_sync();
    if (act && a == 0) {
        int syn = 0;
        for (int d = 0; d < w; d++) syn ^= s.Zn[d];
        s.Zn[p.dc_max] = syn;
    }
    for (int d = 0; d < w; d++)  // d_TMM_Get_deltaU :725-743; vv[h^-1 Zn] is the minimum found above
        s.dU[d * q + (a ^ s.Zn[d])] = __fsub_rn(s.vv[d * q + gmul(p, s.tinv[d], a)], s.mn[d]);
    cta_sync();
    if (act) {  // TMM_Get_Min :745-770
        float m1 = INFINITY, m2 = INFINITY;
        int col = 0;
        for (int d = 0; d < w; d++) {
            const float x = s.dU[d * q + a];
            if (x < m1) {
                m2 = m1;
                m1 = x;
                col = d;
            } else if (x < m2)
                m2 = x;
        }
        s.Min1[a] = m1;
        s.Min2[a] = m2;
        s.MinCol[a] = col;
        s.MC[a] = make_int2(__float_as_int(m1), col);
    }
    cta_sync();
    if (act) {  // TMM_ConstructConf :772-817.  dU[MinCol[j]][j] IS Min1[j], so the two-deviation search reads one
                // {Min1, MinCol} pair per symbol (a broadcast for j, a permutation for a ^ j) and selects without branches:
                // "(d1 > d2 && d1 < I) or (d1 < d2 && d2 < I)"  ==  "d1 != d2 && max(d1, d2) < I"
        float Ii = 0.0f, Ei = 0.0f;
        int p0 = -1, p1 = -1;
        if (a != 0) {
            const float m1a = s.Min1[a];
            Ii = m1a;
            p0 = p1 = s.MinCol[a];
            bool two = false;
            // The pairs {j, a ^ j} and {a ^ j, j} carry the same max(d1, d2); the serial scan meets the smaller
            // index first and the later visit can never pass the strict "<", so only j < (a ^ j) is walked: the
            // numbers whose bit msb(a) is clear, in ascending order.  (j = a is the partner of j = 0 and skipped.)
            const int hb = 31 - __clz(a), lowm = (1 << hb) - 1;
#pragma unroll 4
            for (int t = 0; t < q / 2; t++) {
                const int j = ((t >> hb) << (hb + 1)) | (t & lowm);
                const int2 ej = s.MC[j], ek = s.MC[a ^ j];
                const float d1 = __int_as_float(ej.x), d2 = __int_as_float(ek.x);
                const float m = fmaxf(d1, d2);
                if (ej.y != ek.y && d1 != d2 && m < Ii) {
                    Ii = m;
                    p0 = ej.y;
                    p1 = ek.y;
                    two = true;
                }
            }
            Ei = two ? m1a : s.Min2[a];
        }
        s.I[a] = Ii;
        s.Ev[a] = Ei;
        s.Path[2 * a] = p0;
        s.Path[2 * a + 1] = p1;
    }
    cta_sync();
    const int syn = act ? s.Zn[p.dc_max] : 0;
    for (int d = 0; d < w; d++) {  // :496-521, thread = eta
        const float l = (a == 0) ? 0.0f : ((d != s.Path[2 * a] && d != s.Path[2 * a + 1]) ? s.I[a] : s.Ev[a]);
        const int beta = gmul(p, s.tinv[d], a ^ syn ^ s.Zn[d]);
        const float m = (float)((double)l * 0.8);
        c2v[((size_t)row * p.dc_max + d) * q + beta] = m;
        if (write_llr)  // layered: LLR = v2c + c2v_new (:684-689)
            LLR[s.tvn[d] * q + beta] = __fadd_rn(s.vv[d * q + beta], m);
    }
    cta_sync();
}

__device__ void decode_tmm(const NbParams &p, int f, float *lch, float *LLR, float *c2v, uint16_t *sym, float *smem,
                           int *s_fail, bool layered)
{
    const int q = p.q, N = p.N, M = p.M, tid = threadIdx.x, T = blockDim.x;
    // init :363-402: LLR[0] = max_a L_ch, LLR[a] = max - L_ch[a-1]; c2v = 0
    for (int col = tid; col < N; col += T) {
        float mx = -INFINITY;
        for (int a = 0; a < q - 1; a++) mx = fmaxf(mx, lch[col * (q - 1) + a]);
        LLR[col * q] = mx;
        for (int a = 1; a < q; a++) LLR[col * q + a] = __fsub_rn(mx, lch[col * (q - 1) + a - 1]);
    }
    for (int i = tid; i < M * p.dc_max * q; i += T) c2v[i] = 0.0f;
    cta_sync();
    // thread groups of q threads: one check per group (layered: a single group keeps the row order)
    const int groups = layered ? 1 : max(1, T / q);
    const int g = tid / q, a = tid - g * q;
    const size_t per_group = tmm_group_floats(q, p.dc_max);
    TmmShared s;
    {
        float *base = smem + (size_t)(g < groups ? g : 0) * per_group;
        s.vv = base;
        s.dU = s.vv + p.dc_max * q;
        s.MC = reinterpret_cast<int2 *>(s.dU + p.dc_max * q);
        s.Min1 = s.dU + p.dc_max * q + 2 * q;
        s.Min2 = s.Min1 + q;
        s.I = s.Min2 + q;
        s.Ev = s.I + q;
        s.MinCol = reinterpret_cast<int *>(s.Ev + q);
        s.Path = s.MinCol + q;
        s.Zn = s.Path + 2 * q;
        s.tvn = s.Zn + p.dc_max + 1;
        s.tgf = s.tvn + p.dc_max;
        s.tinv = s.tgf + p.dc_max;
        s.mn = reinterpret_cast<float *>(s.tinv + p.dc_max);
    }
    int it = 0, ok = 0;
    while (it < p.maxit) {
        it++;
        if (!layered) {  // :423-434: LLR += sum_d c2v_d — cumulative over the iterations, as the reference
            if (T % q == 0) {  // thread = (column slot, symbol): no division per element, the column's edge list is uniform
                const int x = tid % q;
                for (int col = tid / q; col < N; col += T / q) {
                    float v = LLR[col * q + x];
                    for (int d = 0; d < p.vw[col]; d++)
                        v = __fadd_rn(v, c2v[((size_t)p.v_cn[col * p.dv_max + d] * p.dc_max + p.v_pos[col * p.dv_max + d]) * q + x]);
                    LLR[col * q + x] = v;
                }
            } else {
                for (int i = tid; i < N * q; i += T) {
                    const int col = i / q, x = i - col * q;
                    float v = LLR[i];
                    for (int d = 0; d < p.vw[col]; d++)
                        v = __fadd_rn(v, c2v[((size_t)p.v_cn[col * p.dv_max + d] * p.dc_max + p.v_pos[col * p.dv_max + d]) * q + x]);
                    LLR[i] = v;
                }
            }
        }
        if (tid == 0) *s_fail = 0;
        cta_sync();
        // d_DecideLLRVector :92-105 (first minimum).  Wide CTAs: one warp per column — lanes scan x = lane, lane + 32, ...
        // ascending, then a warp argmin with the lowest x on ties (one thread per column walked q dependent loads
        // serially: TMM C5 117 -> 125 info Mbit/s).  The 2-warp CTAs of the layered decoder keep one thread per column
        // (the warp form costs them 5 %), and so do small fields (q = 16 with N = 9472 columns: a warp per column leaves
        // half its lanes idle and serialises 1184 columns per warp — Tanner GF(16) TMM 1039 -> 732 info Mbit/s).
        if (T >= 256 && q >= 64) {
            for (int col = tid >> 5; col < N; col += T >> 5) {
                const int lane = tid & 31;
                float mn = INFINITY;
                int best = 0;
                for (int x = lane; x < q; x += 32) {
                    const float v = LLR[col * q + x];
                    if (v < mn) {
                        mn = v;
                        best = x;
                    }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const float on = __shfl_xor_sync(0xffffffffu, mn, o);
                    const int ox = __shfl_xor_sync(0xffffffffu, best, o);
                    if (on < mn || (on == mn && ox < best)) {
                        mn = on;
                        best = ox;
                    }
                }
                if (lane == 0) sym[col] = (uint16_t)best;
            }
        } else {
            for (int col = tid; col < N; col += T) {
                float mn = INFINITY;
                int best = 0;
                for (int x = 0; x < q; x++)
                    if (LLR[col * q + x] < mn) {
                        mn = LLR[col * q + x];
                        best = x;
                    }
                sym[col] = (uint16_t)best;
            }
        }
        cta_sync();
        syndrome(p, sym, s_fail);
        cta_sync();
        if (*s_fail == 0) {
            ok = 1;
            it--;
            break;
        }
        const int rounds = (M + groups - 1) / groups;
        for (int r = 0; r < rounds; r++) {
            const int row = r * groups + g;
            tmm_check(p, row, a, g < groups && row < M, s, LLR, c2v, layered);
        }
    }
    if (tid == 0) {
        if (p.iters_out) p.iters_out[f] = it;
        if (p.ok_out) p.ok_out[f] = ok;
    }
}


// ---- FFT-BP (not in the reference; rules = oracle/nb_oracle.c decode_fftbp) ------------------------
// Check node in the Walsh-Hadamard domain: GF(2^p) addition is XOR, so the convolution of the (permuted)
// input distributions is a pointwise product of their WHTs.  q threads per check; ALL output edges of the
// row go through the inverse transform and the normalising sum together (one barrier per butterfly / tree
// stage for the whole row instead of one per stage and output edge).  Sums are pairwise trees
// (t[x] += t[x + len], len = q/2 ... 1 — the oracle's tree_sum) so that they parallelise without changing a
// bit.  A dense [dc x q] x H_q contraction on the tensor cores is the alternative the north star mentions; at
// q <= 256 and dc <= 12 the transform is 8 butterfly stages on 12 KB and is not the bottleneck of this path.
__device__ __forceinline__ void wht_stage_all(float *F, int q, int w, int a)
{
    for (int len = 1; len < q; len <<= 1) {
        if (a < q / 2) {
            const int j = (a / len) * 2 * len + (a % len);
            for (int d = 0; d < w; d++) {
                const float u = F[d * q + j], v = F[d * q + j + len];
                F[d * q + j] = __fadd_rn(u, v);
                F[d * q + j + len] = __fsub_rn(u, v);
            }
        }
        cta_sync();
    }
}

// The same transform without a single barrier: one warp per vector, PER = q / 32 consecutive elements per lane — the
// stages below PER stay inside the lane's registers, the others cross lanes by __shfl_xor.  Stage order and the operands
// of every add / subtract are those of wht_stage_all (x = the element with the stage's bit clear: x + y, x - y), so
// the results are bit-identical.  Measured in isolation (tools/ubench/wht_ab.cu, profiles/r02_wht_ab.txt): 6.7 G
// transforms/s with the data on chip, HBM-bound at 3.1 G rows/s otherwise.
template <int PER>
__device__ __forceinline__ void wht_vectors_warp(float *F, int q, int w, int warp_in_group, int warps_per_group, int lane)
{
    for (int d = warp_in_group; d < w; d += warps_per_group) {
        float v[PER];
        float *row = F + d * q + lane * PER;
#pragma unroll
        for (int t = 0; t < PER; t++) v[t] = row[t];
#pragma unroll
        for (int len = 1; len < PER; len <<= 1)
#pragma unroll
            for (int t = 0; t < PER; t++)
                if (!(t & len)) {
                    const float x = v[t], y = v[t | len];
                    v[t] = __fadd_rn(x, y);
                    v[t | len] = __fsub_rn(x, y);
                }
#pragma unroll
        for (int m = 1; m < 32; m <<= 1)
#pragma unroll
            for (int t = 0; t < PER; t++) {
                const float o = __shfl_xor_sync(0xffffffffu, v[t], m);
                v[t] = (lane & m) ? __fsub_rn(o, v[t]) : __fadd_rn(v[t], o);
            }
#pragma unroll
        for (int t = 0; t < PER; t++) row[t] = v[t];
    }
}
// all vectors of a row: warp form for q = 32 PER (ends with one barrier), staged form otherwise
__device__ __forceinline__ void wht_all(float *F, int q, int w, int a)
{
    const int lane = a & 31, wg = a >> 5, nwg = q >> 5;
    if (q == 256)
        wht_vectors_warp<8>(F, q, w, wg, nwg, lane);
    else if (q == 128)
        wht_vectors_warp<4>(F, q, w, wg, nwg, lane);
    else if (q == 64)
        wht_vectors_warp<2>(F, q, w, wg, nwg, lane);
    else if (q == 32)
        wht_vectors_warp<1>(F, q, w, wg, nwg, lane);
    else {
        wht_stage_all(F, q, w, a);
        return;
    }
    cta_sync();
}

// tree sum of q values spread over a warp (element a = lane + 32 t), result in every lane
template <int PER>
__device__ __forceinline__ float warp_tree_sum(float (&x)[PER], int q)
{
#pragma unroll
    for (int lt = PER / 2; lt >= 1; lt >>= 1)
#pragma unroll
        for (int t = 0; t < lt; t++) x[t] = __fadd_rn(x[t], x[t + lt]);
    float v = x[0];
    for (int len = (q < 32 ? q : 32) / 2; len >= 1; len >>= 1) v = __fadd_rn(v, __shfl_down_sync(0xffffffffu, v, len));
    return __shfl_sync(0xffffffffu, v, 0);
}

// variable node `col`, one warp: decision on pch * prod_d c2v_d, then v2c_d = normalised pch * prod_{d2 != d} c2v_d2
template <int PER>
__device__ void fft_var_node_warp(const NbParams &p, int col, const float *pch, const float *c2v, float *v2c,
                                  uint16_t *sym, int lane)
{
    const int q = p.q, dv = p.vw[col];
    float pc[PER];
    float best = -1.0f;
    int bi = 0;
#pragma unroll
    for (int t = 0; t < PER; t++) {
        const int a = lane + 32 * t;
        pc[t] = (a < q) ? pch[col * q + a] : 0.0f;
        if (a < q) {
            float v = pc[t];
            for (int d = 0; d < dv; d++)
                v = __fmul_rn(v, c2v[((size_t)p.v_cn[col * p.dv_max + d] * p.dc_max + p.v_pos[col * p.dv_max + d]) * q + a]);
            if (v > best) {
                best = v;
                bi = a;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {  // first maximum = lowest symbol on ties
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) {
            best = ob;
            bi = oi;
        }
    }
    if (lane == 0) sym[col] = (uint16_t)bi;
    for (int d = 0; d < dv; d++) {
        float x[PER], xs[PER];
#pragma unroll
        for (int t = 0; t < PER; t++) {
            const int a = lane + 32 * t;
            float v = pc[t];
            if (a < q)
                for (int d2 = 0; d2 < dv; d2++)
                    if (d2 != d)
                        v = __fmul_rn(v, c2v[((size_t)p.v_cn[col * p.dv_max + d2] * p.dc_max + p.v_pos[col * p.dv_max + d2]) * q + a]);
            x[t] = xs[t] = v;
        }
        const float sum = warp_tree_sum<PER>(xs, q);
#pragma unroll
        for (int t = 0; t < PER; t++) {
            const int a = lane + 32 * t;
            if (a < q) v2c[((size_t)col * p.dv_max + d) * q + a] = __fdiv_rn(x[t], sum);
        }
    }
}

// channel probabilities: pch = normalised exp(L - max(0, max L)), one warp per column
template <int PER>
__device__ void fft_init_warp(const NbParams &p, int col, const float *lch, float *pch, int lane)
{
    const int q = p.q;
    float l[PER], e[PER], es[PER];
    float mx = 0.0f;
#pragma unroll
    for (int t = 0; t < PER; t++) {
        const int a = lane + 32 * t;
        l[t] = (a >= 1 && a < q) ? lch[col * (q - 1) + a - 1] : 0.0f;
        mx = fmaxf(mx, l[t]);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
#pragma unroll
    for (int t = 0; t < PER; t++) {
        const int a = lane + 32 * t;
        e[t] = es[t] = (a < q) ? expf(__fsub_rn(l[t], mx)) : 0.0f;
    }
    const float sum = warp_tree_sum<PER>(es, q);
#pragma unroll
    for (int t = 0; t < PER; t++) {
        const int a = lane + 32 * t;
        if (a < q) pch[col * q + a] = __fdiv_rn(e[t], sum);
    }
}

__device__ void decode_fftbp(const NbParams &p, int f, float *lch, float *pch, float *c2v, float *v2c, uint16_t *sym,
                             float *smem, int *s_fail)
{
    const int q = p.q, N = p.N, M = p.M, tid = threadIdx.x, T = blockDim.x;
    const int warp = tid >> 5, nwarps = T >> 5, lane = tid & 31;
#define FFT_PER(CALL)          \
    if (q <= 32) {             \
        CALL(1)                \
    } else if (q <= 64) {      \
        CALL(2)                \
    } else if (q <= 128) {     \
        CALL(4)                \
    } else {                   \
        CALL(8)                \
    }
    for (int col = warp; col < N; col += nwarps) {
#define CALL(PER) fft_init_warp<PER>(p, col, lch, pch, lane);
        FFT_PER(CALL)
#undef CALL
    }
    for (int i = tid; i < M * p.dc_max * q; i += T) c2v[i] = 1.0f;
    if (tid == 0) *s_fail = 0;
    cta_sync();
    const int groups = max(1, T / q), g = tid / q, a = tid - g * q;
    const size_t per_group = (size_t)2 * p.dc_max * q;
    float *F = smem + (size_t)(g < groups ? g : 0) * per_group, *G = F + (size_t)p.dc_max * q;
    int it = 0, ok = 0;
    while (it < p.maxit) {
        it++;
        for (int col = warp; col < N; col += nwarps) {
#define CALL(PER) fft_var_node_warp<PER>(p, col, pch, c2v, v2c, sym, lane);
            FFT_PER(CALL)
#undef CALL
        }
        cta_sync();
        syndrome(p, sym, s_fail);
        cta_sync();
        const int fail = *s_fail;
        cta_sync();
        if (tid == 0) *s_fail = 0;
        if (fail == 0) {
            ok = 1;
            it--;
            break;
        }
        const int rounds = (M + groups - 1) / groups;
        for (int r = 0; r < rounds; r++) {
            const int row = r * groups + g;
            const bool act = g < groups && row < M;
            const int w = act ? p.cw[row] : 0;
            for (int d = 0; d < w; d++) {
                const float *v = v2c + ((size_t)p.c_vn[row * p.dc_max + d] * p.dv_max + p.c_pos[row * p.dc_max + d]) * q;
                F[d * q + gmul(p, a, p.c_gf[row * p.dc_max + d])] = v[a];
            }
            cta_sync();
            wht_all(F, q, w, a);
            for (int d = 0; d < w; d++) {  // products of the other edges' transforms, ascending edge position
                float gy = 1.0f;
                bool first = true;
                for (int d2 = 0; d2 < w; d2++) {
                    if (d2 == d) continue;
                    gy = first ? F[d2 * q + a] : __fmul_rn(gy, F[d2 * q + a]);
                    first = false;
                }
                G[d * q + a] = gy;
            }
            cta_sync();
            wht_all(G, q, w, a);
            for (int d = 0; d < w; d++) {  // permute back, clamp; F is dead and takes the clamped values
                const float gv = G[d * q + gmul(p, a, p.c_gf[row * p.dc_max + d])];
                F[d * q + a] = gv > 1e-30f ? gv : 1e-30f;
            }
            cta_sync();
            if (q >= 32) {  // tree sums of all output edges: one warp per vector, no barrier inside (same pairing:
                            // element a = lane + 32 t, strides q/2 ... 32 in the lane's registers, 16 ... 1 by shuffles)
                const int lane_g = a & 31;
                for (int d = a >> 5; d < w; d += q >> 5) {
                    float sum;
#define CALL(PER)                                                              \
    {                                                                          \
        float x[PER];                                                          \
        _Pragma("unroll") for (int t = 0; t < PER; t++) x[t] = F[d * q + lane_g + 32 * t]; \
        sum = warp_tree_sum<PER>(x, q);                                        \
    }
                    FFT_PER(CALL)
#undef CALL
                    if (lane_g == 0) G[d * q] = sum;  // G was consumed by the permutation above (behind the barrier)
                }
                cta_sync();
            } else {
                for (int d = 0; d < w; d++) G[d * q + a] = F[d * q + a];
                cta_sync();
                for (int len = q / 2; len >= 1; len >>= 1) {
                    if (a < len)
                        for (int d = 0; d < w; d++) G[d * q + a] = __fadd_rn(G[d * q + a], G[d * q + a + len]);
                    cta_sync();
                }
            }
            for (int d = 0; d < w; d++) c2v[((size_t)row * p.dc_max + d) * q + a] = __fdiv_rn(F[d * q + a], G[d * q]);
            cta_sync();
        }
    }
#undef FFT_PER
    if (tid == 0) {
        if (p.iters_out) p.iters_out[f] = it;
        if (p.ok_out) p.ok_out[f] = ok;
    }
}

// Frame state (channel LLRs, a-posteriori vectors, both message arrays, EMS candidate lists): one slot per
// CTA.  When the slot fits in shared memory beside the algorithm's work arrays it lives there (C4: 150 KB)
// and the CTA is as wide as fits; otherwise (GF(256) EMS/FFT-BP: > 227 KB in fp32) it is a slice of the
// handle's global scratch that stays L2-resident.  The decoders only see pointers.
__host__ __device__ inline size_t nb_slot_floats(int algo, int N, int M, int q, int dv_max, int dc_max)
{
    size_t n = (size_t)N * (q - 1) + (size_t)N * q + (size_t)M * dc_max * q + (size_t)(N + 1) / 2 + 8;  // lch LLR c2v sym
    if (algo == NB_ALGO_EMS || algo == NB_ALGO_FFT_BP) n += (size_t)N * dv_max * q;                      // v2c
    if (algo == NB_ALGO_EMS) n += (size_t)N * dv_max * kNmMax + ((size_t)N * dv_max * kNmMax + 1) / 2;   // top lists
    return (n + 63) & ~(size_t)63;
}

__global__ void __launch_bounds__(kNbThreadsMax)
nb_decode_kernel(const __grid_constant__ NbParams p)
{
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_fail;
    float *slot = p.slot_in_smem ? smem + p.work_floats : p.scratch + (size_t)blockIdx.x * p.slot_floats;
    const int q = p.q, N = p.N, M = p.M;
    if (NB_MUL_SMEM && q <= kMulSmemQ) {
        for (int i = threadIdx.x; i < q * q; i += blockDim.x) g_smul[i] = p.mul[i];
        cta_sync();
    }
    const bool has_v2c = p.algo == NB_ALGO_EMS || p.algo == NB_ALGO_FFT_BP;
    float *lch = slot;
    float *LLR = lch + (size_t)N * (q - 1);
    float *c2v = LLR + (size_t)N * q;
    float *v2c = c2v + (size_t)M * p.dc_max * q;
    float *topval = v2c + (has_v2c ? (size_t)N * p.dv_max * q : 0);
    const size_t ntop = p.algo == NB_ALGO_EMS ? (size_t)N * p.dv_max * kNmMax : 0;
    uint16_t *topsym = reinterpret_cast<uint16_t *>(topval + ntop);
    uint16_t *sym = topsym + ((ntop + 1) & ~(size_t)1);
    // Frame scheduling: with the syndrome stop a frame costs 1 .. maxit iterations, so a static stride leaves most CTAs
    // idle while the unluckiest one finishes its share; CTAs draw their next frame from a counter instead (results
    // do not depend on which CTA decodes a frame: the slot is re-initialised by demodulate()).
    __shared__ int s_next;
    for (int f = blockIdx.x; f < p.F;) {
        demodulate(p, f, lch, LLR);
        cta_sync();
        if (p.algo == NB_ALGO_EMS)
            decode_ems(p, f, lch, LLR, c2v, v2c, topval, topsym, sym, smem, &s_fail);
        else if (p.algo == NB_ALGO_FFT_BP)
            decode_fftbp(p, f, lch, LLR, c2v, v2c, sym, smem, &s_fail);
        else
            decode_tmm(p, f, lch, LLR, c2v, sym, smem, &s_fail, p.algo == NB_ALGO_LAYERED_TMM);
        cta_sync();
        for (int col = threadIdx.x; col < N; col += blockDim.x) p.out[(size_t)f * N + col] = sym[col];
        if (threadIdx.x == 0) s_next = (int)gridDim.x + atomicAdd(p.work_counter, 1);
        cta_sync();
        f = s_next;
    }
}

static std::mutex g_nb_mu;

int nb_upload_tables(nb_ldpc_code *c)
{
    std::lock_guard<std::mutex> lk(g_nb_mu);
    if (c->device >= 0) return LDPC_OK;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        return LDPC_ERR_NO_DEVICE;
    }
    int dev = 0, major = 0, sms = 0;
    LDPC_CUDA_TRY(cudaGetDevice(&dev));
    LDPC_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    LDPC_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (major != 10) return LDPC_ERR_NO_DEVICE;
#define UP(dst, vec, T)                                                                             \
    LDPC_CUDA_TRY(cudaMalloc(&c->dst, (vec).size() * sizeof(T) + 16));                              \
    LDPC_CUDA_TRY(cudaMemcpy(c->dst, (vec).data(), (vec).size() * sizeof(T), cudaMemcpyHostToDevice));
    UP(d_mul, c->mul, uint16_t)
    UP(d_inv, c->inv, uint16_t)
    UP(d_vw, c->vw, int)
    UP(d_cw, c->cw, int)
    UP(d_v_cn, c->v_cn, int)
    UP(d_v_pos, c->v_pos, int)
    UP(d_c_vn, c->c_vn, int)
    UP(d_c_gf, c->c_gf, int)
    UP(d_c_pos, c->c_pos, int)
    UP(d_cre, c->cre, float)
    UP(d_cim, c->cim, float)
#undef UP
    c->num_sms = sms;
    c->device = dev;
    return LDPC_OK;
}

}  // namespace ldpcb

using namespace ldpcb;

static int nb_decode_batch_locked(const nb_ldpc_code_t *cc, const void *in, uint16_t *hard_syms, int iters,
                                  const nb_decode_opts_t *o);

extern "C" int nb_ldpc_decode_batch(const nb_ldpc_code_t *cc, const void *in, uint16_t *hard_syms, int iters,
                                    const nb_decode_opts_t *o)
{
    if (!cc || !in || !hard_syms || !o || iters <= 0) return LDPC_ERR_ARG;
    nb_ldpc_code *c = const_cast<nb_ldpc_code *>(cc);
    // calls on one handle are safe from any host thread / stream and serialise on its scratch arena
    std::lock_guard<std::mutex> lk(c->call_mu);
    int prev = -1;
    const bool bound = c->device >= 0;
    if (bound) {  // run on the device the handle was bound to by its first decode, whatever the caller's current one is
        LDPC_CUDA_TRY(cudaGetDevice(&prev));
        if (prev != c->device) LDPC_CUDA_TRY(cudaSetDevice(c->device));
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(o->stream);
    int rc = LDPC_OK;
    if (c->last_use_valid && cudaStreamWaitEvent(st, c->last_use, 0) != cudaSuccess) rc = LDPC_ERR_CUDA;
    if (rc == LDPC_OK) rc = nb_decode_batch_locked(cc, in, hard_syms, iters, o);
    if (c->device >= 0) {
        if (!c->last_use && cudaEventCreateWithFlags(&c->last_use, cudaEventDisableTiming) != cudaSuccess) c->last_use = nullptr;
        if (c->last_use && cudaEventRecord(c->last_use, st) == cudaSuccess) c->last_use_valid = true;
    }
    if (bound && prev != c->device) (void)cudaSetDevice(prev);
    return rc;
}

static int nb_decode_batch_locked(const nb_ldpc_code_t *cc, const void *in, uint16_t *hard_syms, int iters,
                                  const nb_decode_opts_t *o)
{
    if (!cc || !in || !hard_syms || !o || iters <= 0) return LDPC_ERR_ARG;
    if (o->struct_size != (int)sizeof(nb_decode_opts_t) || o->batch <= 0) return LDPC_ERR_ARG;
    if (o->algo != NB_ALGO_EMS && o->algo != NB_ALGO_TMM && o->algo != NB_ALGO_LAYERED_TMM &&
        o->algo != NB_ALGO_FFT_BP)
        return LDPC_ERR_ARG;
    if (o->in_kind < 0 || o->in_kind > 2) return LDPC_ERR_ARG;
    nb_ldpc_code *c = const_cast<nb_ldpc_code *>(cc);
    // EMS budgets: Nm <= 4 sorted entries per input with any Nc, or the reference's log-QSPA setting
    // (decoder_method 2: Nm = q, Nc = dc_max - 1 = every configuration)
    const bool ems_full = o->algo == NB_ALGO_EMS && o->ems_nm >= c->q && o->ems_nc >= c->dc_max - 1;
    if (o->algo == NB_ALGO_EMS && !ems_full && (o->ems_nm < 1 || o->ems_nm > kNmMax || o->ems_nc < 0))
        return LDPC_ERR_UNSUPPORTED;
    if (o->in_kind == NB_IN_BPSK && c->n_const != 2) return LDPC_ERR_ARG;
    if (o->in_kind == NB_IN_QAM && c->n_const != c->q) return LDPC_ERR_ARG;
    if (o->in_kind != NB_IN_SYMBOL_LLR && !(o->sigma > 0.0f)) return LDPC_ERR_ARG;
    if (o->algo == NB_ALGO_TMM || o->algo == NB_ALGO_LAYERED_TMM)
        for (int g : c->c_gf)
            if (g == 0) return LDPC_ERR_UNSUPPORTED;  // h^-1 does not exist (the reference exits with "Div 0 Error!")
    if (o->in_kind == NB_IN_BPSK && (c->cre[0] != 1.0f || c->cre[1] != -1.0f)) return LDPC_ERR_UNSUPPORTED;
    int rc = nb_upload_tables(c);
    if (rc != LDPC_OK) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(o->stream);
    const int F = o->batch, q = c->q, N = c->N, M = c->M;
    const bool host = o->mem_space == LDPC_MEM_HOST;
    const size_t in_elems = (o->in_kind == NB_IN_SYMBOL_LLR) ? (size_t)N * (q - 1)
                            : (o->in_kind == NB_IN_BPSK)     ? (size_t)N * c->p
                                                             : (size_t)N * 2;
    const size_t in_bytes = in_elems * F * sizeof(float), out_bytes = (size_t)F * N * sizeof(uint16_t);
    if (o->algo == NB_ALGO_FFT_BP)
        for (int g : c->c_gf)
            if (g == 0) return LDPC_ERR_UNSUPPORTED;  // multiplication by 0 is not a permutation
    if (q > kNbThreads) return LDPC_ERR_UNSUPPORTED;  // q = 512 needs a wider CTA (no shipped code uses GF(512))
    // shared-memory work arrays of a CTA of T threads: EMS E[q] + input lists per warp, FFT-BP / TMM group arrays
    auto work_floats_of = [&](int T) -> size_t {
        size_t w;
        if (o->algo == NB_ALGO_EMS)
            w = (size_t)(T / 32) * ems_warp_floats(q, c->dc_max);
        else if (o->algo == NB_ALGO_FFT_BP)
            w = (size_t)(T / q) * ((size_t)2 * c->dc_max * q);
        else
            w = (size_t)((o->algo == NB_ALGO_LAYERED_TMM) ? 1 : T / q) *
                tmm_group_floats(q, c->dc_max);
        return (w + 3) & ~(size_t)3;
    };
    const size_t slot_floats = nb_slot_floats(o->algo, N, M, q, c->dv_max, c->dc_max);
    // CTA shape (measured on B200, profiles/r01_nb_bench.txt): frames in flight per SM matter more than where the
    // state lives, so EMS / TMM keep 256-thread CTAs (4-8 per SM) with the state in a global-scratch slot (L2); the
    // row-serial layered TMM runs one q-thread group per CTA (up to 32 CTAs per SM); FFT-BP, a chain of short
    // barrier-separated stages, is fastest as one wide CTA with the whole frame state in shared memory.
    int threads = kNbThreads, slot_in_smem = 0;
    if (o->algo == NB_ALGO_LAYERED_TMM) threads = (q + 31) & ~31;  // one q-thread group, whole warps
    int t_first = kNbThreadsMax;
    if (const char *e = getenv("LDPC_B200_NB_THREADS"))
        if (atoi(e) >= kNbThreads && atoi(e) <= kNbThreadsMax && (atoi(e) & (atoi(e) - 1)) == 0) t_first = atoi(e);
    if (o->algo == NB_ALGO_FFT_BP)
        for (int T = t_first; T >= kNbThreads; T >>= 1)
            if ((work_floats_of(T) + slot_floats) * sizeof(float) <= kNbDynSmemMax) {
                threads = T;
                slot_in_smem = 1;
                break;
            }
    // global-slot CTA widths from a sweep on B200 (ms per batch at T = 256 / 512 / 1024): C5 FFT-BP 15.0 / 12.3 / 11.8,
    // C5 TMM 21.1 / 17.0 / 18.5, C4 TMM 6.4 / 6.6 / 7.0, C4 EMS 13.9 / 13.2 / 15.2
    if (!slot_in_smem) {
        if (o->algo == NB_ALGO_FFT_BP) threads = 4 * q > kNbThreadsMax ? kNbThreadsMax : (4 * q < kNbThreads ? kNbThreads : 4 * q);
        if (o->algo == NB_ALGO_TMM && 2 * q > kNbThreads) threads = 2 * q > kNbThreadsMax ? kNbThreadsMax : 2 * q;
        if (o->algo == NB_ALGO_EMS) threads = 512;
        if ((work_floats_of(threads)) * sizeof(float) > kNbDynSmemMax) threads = kNbThreads;
        if (const char *e = getenv("LDPC_B200_NB_THREADS")) {  // tuning knob (tools/nb_bench.py sweeps)
            const int T = atoi(e), unit = q > 32 ? q : 32;
            if (o->algo != NB_ALGO_LAYERED_TMM && T >= unit && T <= kNbThreadsMax && T % unit == 0 &&
                work_floats_of(T) * sizeof(float) <= kNbDynSmemMax)
                threads = T;
        }
    }
    const size_t work_floats = work_floats_of(threads);
    const size_t smem = (work_floats + (slot_in_smem ? slot_floats : 0)) * sizeof(float);
    const int ems_chunk = threads / 32;
    if (smem > kNbDynSmemMax) return LDPC_ERR_UNSUPPORTED;
    LDPC_CUDA_TRY(cudaFuncSetAttribute(nb_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    LDPC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, nb_decode_kernel, threads, smem));
    if (occ < 1) return LDPC_ERR_UNSUPPORTED;
    int grid = c->num_sms * occ;
    if (grid > F) grid = F;
    const size_t a256 = 255;
    size_t need = slot_in_smem ? 256 : (size_t)grid * slot_floats * sizeof(float);
    const size_t o_in = need = (need + a256) & ~a256;
    need += host ? ((in_bytes + a256) & ~a256) : 0;
    const size_t o_out = need;
    need += host ? ((out_bytes + a256) & ~a256) : 0;
    const size_t o_it = need;
    need += ((size_t)F * 4 + a256) & ~a256;
    const size_t o_ok = need;
    need += ((size_t)F * 4 + a256) & ~a256;
    const size_t o_ctr = need;
    need += 256;
    {
        std::lock_guard<std::mutex> lk(g_nb_mu);
        if (c->scratch_bytes < need) {
            if (c->scratch) {
                LDPC_CUDA_TRY(cudaDeviceSynchronize());
                LDPC_CUDA_TRY(cudaFree(c->scratch));
                c->scratch = nullptr;
                c->scratch_bytes = 0;
            }
            LDPC_CUDA_TRY(cudaMalloc(&c->scratch, need));
            c->scratch_bytes = need;
        }
    }
    unsigned char *base = reinterpret_cast<unsigned char *>(c->scratch);
    NbParams p;
    memset(&p, 0, sizeof(p));
    p.in = in;
    p.out = hard_syms;
    p.iters_out = o->iters_out;
    p.ok_out = o->ok_out;
    if (host) {
        LDPC_CUDA_TRY(cudaMemcpyAsync(base + o_in, in, in_bytes, cudaMemcpyHostToDevice, st));
        p.in = base + o_in;
        p.out = reinterpret_cast<uint16_t *>(base + o_out);
        p.iters_out = reinterpret_cast<int *>(base + o_it);
        p.ok_out = reinterpret_cast<int *>(base + o_ok);
    }
    p.scratch = reinterpret_cast<float *>(base);
    p.slot_floats = slot_floats;
    p.slot_in_smem = slot_in_smem;
    p.work_floats = work_floats;
    p.F = F;
    p.N = N;
    p.M = M;
    p.q = q;
    p.p = c->p;
    p.dv_max = c->dv_max;
    p.dc_max = c->dc_max;
    p.n_const = c->n_const;
    p.algo = o->algo;
    p.in_kind = o->in_kind;
    p.maxit = iters;
    p.nm = ems_full ? 0 : o->ems_nm;
    p.ems_full = ems_full ? 1 : 0;
    p.nc = o->ems_nc;
    p.sigma = o->sigma;
    p.ems_chunk = ems_chunk;
    p.mul = c->d_mul;
    p.inv = c->d_inv;
    p.vw = c->d_vw;
    p.cw = c->d_cw;
    p.v_cn = c->d_v_cn;
    p.v_pos = c->d_v_pos;
    p.c_vn = c->d_c_vn;
    p.c_gf = c->d_c_gf;
    p.c_pos = c->d_c_pos;
    p.cre = c->d_cre;
    p.cim = c->d_cim;
    p.work_counter = reinterpret_cast<int *>(base + o_ctr);
    LDPC_CUDA_TRY(cudaMemsetAsync(p.work_counter, 0, sizeof(int), st));
    nb_decode_kernel<<<grid, threads, smem, st>>>(p);
    LDPC_CUDA_TRY(cudaGetLastError());
    if (host) {
        LDPC_CUDA_TRY(cudaMemcpyAsync(hard_syms, base + o_out, out_bytes, cudaMemcpyDeviceToHost, st));
        if (o->iters_out)
            LDPC_CUDA_TRY(cudaMemcpyAsync(o->iters_out, base + o_it, (size_t)F * 4, cudaMemcpyDeviceToHost, st));
        if (o->ok_out)
            LDPC_CUDA_TRY(cudaMemcpyAsync(o->ok_out, base + o_ok, (size_t)F * 4, cudaMemcpyDeviceToHost, st));
        LDPC_CUDA_TRY(cudaStreamSynchronize(st));
    }
    return 1;
}
