/*
 * bldpc_oracle.c — CPU oracle for the binary QC-LDPC decode path.  TEST INFRASTRUCTURE
 * ONLY (see bldpc_oracle.h).  B/ = /root/reference/bldpc_实习/ (gsw4869/CUDA_LDPC).
 *
 * Pinned part (checked against the reference's own sources run on the CPU, oracle/_ref):
 *   Get_H, Transform_H (literal), RandomModule, AWGNChannel_CPU, the flooding fp32
 *   VN/CN update rules, the genie batch-wide stop, Statistic and the per-SNR loop.
 * Defined here, not in the reference ("parity unpinned"):
 *   the circulant table (the reference's intended graph), the layered
 *   schedule, alpha, the int8 fixed-point rules below and the fp16 rules further down
 *   (orc_layered_f16; cross-checked against a numpy.float16 restatement, tests/test_oracle_f16.py).
 *
 * int8 layered rules (the CUDA kernel must match bit for bit):
 *   q        = clamp(rintf(y * scale), -127, 127)            APP[n] = q[n]
 *   record(m)= {m1, m2, idx, signs} = 0 before the first iteration
 *   for layer r = 0..J-1, check i = 0..Z-1, edges k = 0..dc-1 (ascending column block):
 *     v_k    = c_k*Z + (i + s_k) mod Z
 *     old_k  = ((signs>>k)&1 ? -1 : +1) * (k == idx ? m2 : m1)
 *     t_k    = APP[v_k] - old_k                               (no clamp, |t| <= 254)
 *     a_k    = min(|t_k|, amax)
 *     min1 <= min2 = two smallest a_k (with multiplicity), idx' = first k with a_k == min1
 *     P      = xor_k (t_k < 0)
 *     m1'    = min1 - ((min1*bnum) >> bshift),  m2' likewise   (bnum = 0: the reference's un-normalised rule)
 *     sign'_k= P ^ (t_k < 0);  new_k = (sign'_k ? -1 : +1) * (k == idx' ? m2' : m1')
 *     APP[v_k] = clamp(t_k + new_k, -127, 127)
 *   hard bit = APP < 0.  ORC_EXIT_SYNDROME tests H x = 0 after each full iteration.
 */
#include "bldpc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI (3.1415926) /* B/define.cuh:58 */

/* ------------------------------------------------------------------ loaders */

int orc_get_h(const char *path, int J, int L, int *H, int *Wc, int *Wv)
{
    /* B/Simulation.cu:310-319: fscanf("%d") x J*L */
    FILE *fp = fopen(path, "r");
    if (!fp) return -1;
    for (int i = 0; i < J * L; i++) {
        int v;
        if (fscanf(fp, "%d", &v) != 1) {
            fclose(fp);
            return -2;
        }
        H[i] = v;
    }
    fclose(fp);
    memset(Wc, 0, (size_t)(J + 1) * sizeof(int));
    memset(Wv, 0, (size_t)(L + 1) * sizeof(int));
    /* B/Simulation.cu:321-340 */
    for (int r = 0; r < J; r++) {
        for (int c = 0; c < L; c++)
            if (H[r * L + c] != -1) Wc[r]++;
        if (Wc[r] > Wc[J]) Wc[J] = Wc[r];
    }
    for (int c = 0; c < L; c++) {
        for (int r = 0; r < J; r++)
            if (H[r * L + c] != -1) Wv[c]++;
        if (Wv[c] > Wv[L]) Wv[L] = Wv[c];
    }
    return 0;
}

void orc_transform_h(const int *H, int J, int L, int Z, const int *Wc, const int *Wv,
                     int *addr, int literal)
{
    /* B/Simulation.cu:363-387 */
    for (int c = 0; c < L; c++) {
        int k = 0;
        for (int r = 0; r < J; r++) {
            int s = H[r * L + c];
            if (s == -1) continue;
            int position = 0;
            for (int c2 = 0; c2 < c; c2++)
                if (H[r * L + c2] != -1) position++;
            int base = (Z - s) % Z;
            for (int i = 0; i < Z; i++) {
                int row;
                if (literal) /* :380 — the non-wrapping branch returns i, not base+i (F3) */
                    row = (base + i >= Z) ? base + i - Z : i;
                else
                    row = (base + i) % Z;
                addr[(c * Z + i) * Wv[L] + k] = (r * Z + row) * Wc[J] + position;
            }
            k++;
        }
    }
}

/* ------------------------------------------------------------------ channel */

float orc_random_module(int *seed)
{
    /* B/LDPC_Encoder.cu:46-56 */
    float temp = 0.0f;
    seed[0] = (seed[0] * 249) % 61967;
    seed[1] = (seed[1] * 251) % 63443;
    seed[2] = (seed[2] * 252) % 63599;
    temp = (((float)seed[0]) / ((float)61967)) + (((float)seed[1]) / ((float)63443)) +
           (((float)seed[2]) / ((float)63599));
    temp -= (int)temp;
    return temp;
}

void orc_awgn(int *seed, float sigma, const int *codeword, float *out, int N, int F)
{
    /* B/LDPC_Encoder.cu:25-41.  The reference is compiled as C++: log/sqrt on a float
     * argument are the float overloads, sin(2*PI*u2) is double, and the sum is formed in
     * double and rounded on store. */
    for (int f = 0; f < F; f++) {
        for (int n = 0; n < N; n++) {
            float u1 = orc_random_module(seed);
            float u2 = orc_random_module(seed);
            float temp = sqrtf((float)(-2) * logf((float)1 - u1));
            int cw = codeword ? codeword[n * F + f] : 0;
            out[n * F + f] =
                (float)((double)sigma * sin(2 * ORC_PI * (double)u2) * (double)temp + 1.0 - 2 * cw);
        }
    }
}

float orc_sigma(int snrtype, float snr_db, float rate)
{
    /* B/main.cu:120-127 */
    if (snrtype == 0) return (float)sqrt(0.5 / (rate * (pow(10.0, (snr_db / 10.0)))));
    return (float)sqrt(0.5 / (pow(10.0, (snr_db / 10.0))));
}

/* ------------------------------------------------------------------ helpers */

typedef struct {
    int dc;
    int col[64];
    int shift[64];
} orc_layer;

static orc_layer *build_layers(int J, int L, const int *H)
{
    orc_layer *ly = (orc_layer *)calloc((size_t)J, sizeof(orc_layer));
    for (int r = 0; r < J; r++)
        for (int c = 0; c < L; c++)
            if (H[r * L + c] != -1) {
                ly[r].col[ly[r].dc] = c;
                ly[r].shift[ly[r].dc] = H[r * L + c];
                ly[r].dc++;
            }
    return ly;
}

void orc_syndrome_ok(int J, int L, int Z, const int *H, const int *D, int F, int *ok)
{
    orc_layer *ly = build_layers(J, L, H);
    for (int f = 0; f < F; f++) ok[f] = 1;
    for (int r = 0; r < J; r++)
        for (int i = 0; i < Z; i++)
            for (int f = 0; f < F; f++) {
                int p = 0;
                for (int k = 0; k < ly[r].dc; k++)
                    p ^= D[(ly[r].col[k] * Z + (i + ly[r].shift[k]) % Z) * F + f] & 1;
                if (p) ok[f] = 0;
            }
    free(ly);
}

/* ------------------------------------------------------------------ flooding fp32 */

static void sortQ(float *MinQ, float *SubMinQ, float *Q, int Weight)
{
    /* B/LDPC_Decoder.cu:374-398: two bubble passes push the two smallest to the end */
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < Weight - 1; j++)
            if (Q[j] < Q[j + 1]) {
                float tmp = Q[j];
                Q[j] = Q[j + 1];
                Q[j + 1] = tmp;
            }
    *MinQ = Q[Weight - 1];
    *SubMinQ = Q[Weight - 2];
}

int orc_flooding_fp32(int J, int L, int Z, const int *H, const int *Wc, const int *Wv,
                      const int *addr, const float *y, int F, int maxit, int exit_mode,
                      int length, int *D, int *iters, float *rq_out)
{
    const int N = L * Z, M = J * Z, Wcm = Wc[J], Wvm = Wv[L];
    float *RQ = (float *)calloc((size_t)M * Wcm * F, sizeof(float)); /* :82 memset 0 */
    int *done = (int *)calloc((size_t)F, sizeof(int));
    int *Dcur = (int *)malloc((size_t)N * F * sizeof(int));
    int *ok = (int *)malloc((size_t)F * sizeof(int));
    if (!RQ || !done || !Dcur || !ok) return -5;
    int it = 0;
    for (int f = 0; f < F; f++) iters[f] = 0;
    memset(D, 0, (size_t)(N + 1) * F * sizeof(int));
    while (it < maxit) { /* :94 */
        it++;
        /* Variablenode(_Shared_)_Kernel :172-261, threads in ascending offset = n*F+f */
        for (int n = 0; n < N; n++) {
            int Weight = Wv[n / Z];
            for (int f = 0; f < F; f++) {
                float R[64];
                int Ad[64];
                float Add_result = 0.0f; /* uninitialised in the reference (F4); defined 0 */
                for (int i = 0; i < Weight; i++) Ad[i] = addr[n * Wvm + i] * F + f;
                for (int i = 0; i < Weight; i++) R[i] = RQ[Ad[i]];
                for (int i = 0; i < Weight; i++) Add_result += R[i];
                Add_result += y[n * F + f];
                Dcur[n * F + f] = (Add_result < 0) ? 1 : 0;
                for (int i = 0; i < Weight; i++) RQ[Ad[i]] = Add_result - R[i];
            }
        }
        /* Checknode(_Shared_)_Kernel :262-372 */
        for (int m = 0; m < M; m++) {
            int Weight = Wc[m / Z];
            for (int f = 0; f < F; f++) {
                float Q[64], Q0[64], MinQ, SubMinQ;
                int Sign[64], SignAll = 1, Index_minQ = 0;
                size_t base = ((size_t)m * Wcm) * F + f;
                for (int i = 0; i < Weight; i++) Q[i] = RQ[base + (size_t)i * F];
                for (int i = 0; i < Weight; i++) {
                    Sign[i] = (Q[i] < 0) ? -1 : 1;
                    Q[i] = (Q[i] < 0) ? -Q[i] : Q[i];
                    Q0[i] = Q[i];
                }
                for (int i = 0; i < Weight; i++) SignAll *= Sign[i];
                sortQ(&MinQ, &SubMinQ, Q, Weight);
                for (int i = 0; i < Weight; i++)
                    if (Q0[i] == MinQ) {
                        Index_minQ = i;
                        break;
                    }
                for (int i = 0; i < Weight; i++)
                    RQ[base + (size_t)i * F] =
                        (i != Index_minQ) ? SignAll * Sign[i] * MinQ : SignAll * Sign[i] * SubMinQ;
            }
        }
        /* host side :134-153 */
        if (exit_mode == ORC_EXIT_SYNDROME) {
            int all = 1;
            orc_syndrome_ok(J, L, Z, H, Dcur, F, ok);
            for (int f = 0; f < F; f++) {
                if (done[f]) continue;
                for (int n = 0; n < N; n++) D[n * F + f] = Dcur[n * F + f];
                iters[f] = it;
                if (ok[f]) {
                    done[f] = 1;
                    D[N * F + f] = 1;
                } else
                    all = 0;
            }
            if (all) break;
        } else {
            int good = 0;
            memcpy(D, Dcur, (size_t)N * F * sizeof(int));
            for (int f = 0; f < F; f++) {
                int s = 0;
                for (int n = 0; n < length; n++) s += D[n * F + f];
                D[N * F + f] = (s == 0) ? 1 : 0;
                good += D[N * F + f];
                iters[f] = it;
            }
            if (exit_mode == ORC_EXIT_GENIE && good == F) break;
        }
    }
    if (exit_mode == ORC_EXIT_NONE) { /* D6: flag row = true syndrome in fixed-iteration mode */
        orc_syndrome_ok(J, L, Z, H, D, F, ok);
        for (int f = 0; f < F; f++) D[N * F + f] = ok[f];
    }
    if (rq_out) memcpy(rq_out, RQ, (size_t)M * Wcm * F * sizeof(float));
    free(RQ);
    free(done);
    free(Dcur);
    free(ok);
    return 0;
}

/* ------------------------------------------------------------------ layered fp32 */

int orc_layered_fp32(int J, int L, int Z, const int *H, const float *y, int F, int maxit,
                     float alpha, int exit_mode, int *D, int *iters, float *app_out)
{
    const int N = L * Z;
    orc_layer *ly = build_layers(J, L, H);
    int dcm = 0;
    for (int r = 0; r < J; r++)
        if (ly[r].dc > dcm) dcm = ly[r].dc;
#pragma omp parallel for schedule(dynamic)
    for (int f = 0; f < F; f++) {
        float *c = (float *)calloc((size_t)J * Z * dcm, sizeof(float)); /* explicit c2v per edge */
        float *a = (float *)malloc((size_t)N * sizeof(float));
        int *h = (int *)malloc((size_t)N * sizeof(int));
        for (int n = 0; n < N; n++) a[n] = y[(size_t)n * F + f];
        int it = 0, okf = 0;
        while (it < maxit) {
            it++;
            for (int r = 0; r < J; r++) {
                const int dc = ly[r].dc;
                for (int i = 0; i < Z; i++) {
                    float t[64], mag[64];
                    int v[64], neg[64], P = 0, idx = 0;
                    float min1 = INFINITY, min2 = INFINITY;
                    float *cm = c + ((size_t)(r * Z + i)) * dcm;
                    for (int k = 0; k < dc; k++) {
                        v[k] = ly[r].col[k] * Z + (i + ly[r].shift[k]) % Z;
                        t[k] = a[v[k]] - cm[k];
                        neg[k] = t[k] < 0;
                        mag[k] = neg[k] ? -t[k] : t[k];
                        P ^= neg[k];
                        if (mag[k] < min1) {
                            min2 = min1;
                            min1 = mag[k];
                            idx = k;
                        } else if (mag[k] < min2)
                            min2 = mag[k];
                    }
                    float m1 = alpha * min1, m2 = alpha * min2;
                    for (int k = 0; k < dc; k++) {
                        float m = (k == idx) ? m2 : m1;
                        float nw = (P ^ neg[k]) ? -m : m;
                        cm[k] = nw;
                        a[v[k]] = t[k] + nw;
                    }
                }
            }
            if (exit_mode == ORC_EXIT_SYNDROME || it == maxit) {
                for (int n = 0; n < N; n++) h[n] = a[n] < 0;
                okf = 1;
                for (int r = 0; r < J && okf; r++)
                    for (int i = 0; i < Z && okf; i++) {
                        int p = 0;
                        for (int k = 0; k < ly[r].dc; k++)
                            p ^= h[ly[r].col[k] * Z + (i + ly[r].shift[k]) % Z];
                        if (p) okf = 0;
                    }
                if (okf && exit_mode == ORC_EXIT_SYNDROME) break;
            }
        }
        for (int n = 0; n < N; n++) {
            D[(size_t)n * F + f] = a[n] < 0;
            if (app_out) app_out[(size_t)n * F + f] = a[n];
        }
        if (exit_mode == ORC_EXIT_GENIE) { /* per-frame genie flag, no early stop */
            okf = 1;
            for (int n = 0; n < N; n++)
                if (a[n] < 0) okf = 0;
        }
        D[(size_t)N * F + f] = okf;
        iters[f] = it;
        free(c);
        free(a);
        free(h);
    }
    free(ly);
    return 0;
}

/* ------------------------------------------------------------------ layered int8 */

static inline int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

int orc_layered_i8(int J, int L, int Z, const int *H, const float *y, int F, int maxit,
                   float scale, int amax, int bnum, int bshift, int exit_mode, int *D,
                   int *iters, int8_t *app_out, uint32_t *rec_out)
{
    const int N = L * Z, M = J * Z;
    orc_layer *ly = build_layers(J, L, H);
    if (amax < 1 || amax > 127) return -3;
#pragma omp parallel for schedule(dynamic)
    for (int f = 0; f < F; f++) {
        int *a = (int *)malloc((size_t)N * sizeof(int));
        uint32_t *rec = (uint32_t *)calloc((size_t)M * 4, sizeof(uint32_t));
        for (int n = 0; n < N; n++)
            /* clamp in float first: (int) of +-inf / huge values / NaN is undefined in C (x86 returns INT_MIN for all of
             * them, which would turn +inf into -127); NaN -> -127, like fmaxf(NaN, -127) in the kernel */
            a[n] = (int)fminf(fmaxf(rintf(y[(size_t)n * F + f] * scale), -127.0f), 127.0f);
        int it = 0, okf = 0;
        while (it < maxit) {
            it++;
            for (int r = 0; r < J; r++) {
                const int dc = ly[r].dc;
                for (int i = 0; i < Z; i++) {
                    uint32_t *rc = rec + ((size_t)(r * Z + i)) * 4;
                    const int m1o = (int)rc[0], m2o = (int)rc[1], idxo = (int)rc[2];
                    const uint32_t sgo = rc[3];
                    int t[64], v[64], neg[64], P = 0, idx = 0;
                    int min1 = 1 << 20, min2 = 1 << 20;
                    for (int k = 0; k < dc; k++) {
                        v[k] = ly[r].col[k] * Z + (i + ly[r].shift[k]) % Z;
                        int old = (k == idxo) ? m2o : m1o;
                        if ((sgo >> k) & 1u) old = -old;
                        t[k] = a[v[k]] - old;
                        neg[k] = t[k] < 0;
                        int ak = neg[k] ? -t[k] : t[k];
                        if (ak > amax) ak = amax;
                        P ^= neg[k];
                        if (ak < min1) {
                            min2 = min1;
                            min1 = ak;
                            idx = k;
                        } else if (ak < min2)
                            min2 = ak;
                    }
                    if (min2 > amax) min2 = amax; /* dc == 1 only */
                    const int m1 = min1 - ((min1 * bnum) >> bshift), m2 = min2 - ((min2 * bnum) >> bshift);
                    uint32_t sg = 0;
                    for (int k = 0; k < dc; k++) {
                        int s = P ^ neg[k];
                        int m = (k == idx) ? m2 : m1;
                        sg |= (uint32_t)s << k;
                        a[v[k]] = clampi(t[k] + (s ? -m : m), -127, 127);
                    }
                    rc[0] = (uint32_t)m1;
                    rc[1] = (uint32_t)m2;
                    rc[2] = (uint32_t)idx;
                    rc[3] = sg;
                }
            }
            if (exit_mode == ORC_EXIT_SYNDROME || it == maxit) {
                okf = 1;
                for (int r = 0; r < J && okf; r++)
                    for (int i = 0; i < Z && okf; i++) {
                        int p = 0;
                        for (int k = 0; k < ly[r].dc; k++)
                            p ^= a[ly[r].col[k] * Z + (i + ly[r].shift[k]) % Z] < 0;
                        if (p) okf = 0;
                    }
                if (okf && exit_mode == ORC_EXIT_SYNDROME) break;
            }
        }
        for (int n = 0; n < N; n++) {
            D[(size_t)n * F + f] = a[n] < 0;
            if (app_out) app_out[(size_t)n * F + f] = (int8_t)a[n];
        }
        if (rec_out)
            for (int m = 0; m < M; m++)
                for (int w = 0; w < 4; w++)
                    rec_out[((size_t)m * 4 + w) * F + f] = rec[(size_t)m * 4 + w];
        D[(size_t)N * F + f] = okf;
        iters[f] = it;
        free(a);
        free(rec);
    }
    free(ly);
    return 0;
}

/* ------------------------------------------------------------------ layered fp16 */

/* IEEE binary16 <-> binary32, round to nearest even, subnormals kept.  Every fp16 operation below is the float
 * operation on the converted operands rounded back: for add/sub/mul of two 11-bit significands a 24-bit
 * intermediate makes the double rounding innocuous (24 >= 2*11 + 2), so this IS the correctly rounded fp16 result
 * — what HADD2 / HMUL2 compute on the GPU. */
static inline float orc_h2f(uint16_t h)
{
    const uint32_t s = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 1023u;
    uint32_t u;
    if (e == 0) {
        if (m == 0) u = s;
        else {
            float f = (float)m * (1.0f / 16777216.0f); /* m * 2^-24, exact */
            memcpy(&u, &f, 4);
            u |= s;
        }
    } else if (e == 31) u = s | 0x7F800000u | (m << 13);
    else u = s | ((e + 112u) << 23) | (m << 13);
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline uint16_t orc_f2h(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint16_t s = (uint16_t)((u >> 16) & 0x8000u);
    u &= 0x7FFFFFFFu;
    if (u >= 0x7F800000u) return (uint16_t)(s | (u > 0x7F800000u ? 0x7E00u : 0x7C00u));
    if (u >= 0x477FF000u) return (uint16_t)(s | 0x7C00u); /* >= 65520 rounds to inf */
    if (u < 0x38800000u) {                                /* below 2^-14: subnormal result, unit 2^-24 */
        float a;
        memcpy(&a, &u, 4);
        const float r = a * 16777216.0f; /* exact scaling */
        return (uint16_t)(s | (uint16_t)lrintf(r)); /* nearest even in the default rounding mode; 1024 = smallest normal */
    }
    const uint32_t mant = u & 0x7FFFFFu, exp = (u >> 23) - 112u;
    uint32_t h = (exp << 10) | (mant >> 13);
    const uint32_t rem = mant & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++; /* carries into the exponent correctly */
    return (uint16_t)(s | h);
}
uint16_t orc_f16_from_f32(float f) { return orc_f2h(f); }
float orc_f16_to_f32(uint16_t h) { return orc_h2f(h); }

/*
 * fp16 layered rules ("parity unpinned": ours, like the int8 rules; the CUDA kernel bldpc_layered_f16.cu must match
 * bit for bit).  All values are binary16 patterns, all operations correctly rounded (nearest even):
 *   APP[n]   = f16(clamp(y * scale, -127, 127))              (float product, float clamp, one rounding; NaN -> -127)
 *   c2v      = +0 before the first iteration
 *   t_k      = APP[v_k] - c2v_old_k ;  a_k = |t_k| (sign bit cleared) ;  neg_k = sign BIT of t_k (-0 counts)
 *   min1 <= min2 = two smallest a_k (with multiplicity);  P = xor_k neg_k
 *   m1'      = min(min1, amax) * c,  m2' = min(min2, amax) * c,   c = 1 - bnum / 2^bshift  (bnum = 0: no product)
 *   new_k    = (P ^ neg_k ? -1 : +1) * (a_k == min1 ? m2' : m1')   (equal to the first-index rule: ties have min2 == min1)
 *   APP[v_k] = t_k + new_k      (no clamp: |APP| <= |channel| + dv * amax + rounding drift, far inside binary16;
 *                                t_k - c2v_old_k is never inf - inf because messages are <= amax)
 *   hard bit = sign BIT of APP.  ORC_EXIT_SYNDROME tests H x = 0 after each full iteration.
 * app_out: [N][F] patterns; msg_out: [M][dc_max][F] patterns (absent edges 0).
 */
int orc_layered_f16(int J, int L, int Z, const int *H, const float *y, int F, int maxit, float scale, int amax,
                    int bnum, int bshift, int exit_mode, int *D, int *iters, uint16_t *app_out, uint16_t *msg_out)
{
    const int N = L * Z, M = J * Z;
    orc_layer *ly = build_layers(J, L, H);
    if (amax < 1 || amax > 127) return -3;
    int dcmax = 0;
    for (int r = 0; r < J; r++) dcmax = ly[r].dc > dcmax ? ly[r].dc : dcmax;
    const uint16_t amaxh = orc_f2h((float)amax);
    const float c = 1.0f - (float)bnum / (float)(1 << bshift);
#pragma omp parallel for schedule(dynamic)
    for (int f = 0; f < F; f++) {
        uint16_t *a = (uint16_t *)malloc((size_t)N * sizeof(uint16_t));
        uint16_t *msg = (uint16_t *)calloc((size_t)M * dcmax, sizeof(uint16_t));
        for (int n = 0; n < N; n++) a[n] = orc_f2h(fminf(fmaxf(y[(size_t)n * F + f] * scale, -127.0f), 127.0f));
        int it = 0, okf = 0;
        while (it < maxit) {
            it++;
            for (int r = 0; r < J; r++) {
                const int dc = ly[r].dc;
                for (int i = 0; i < Z; i++) {
                    uint16_t *mr = msg + (size_t)(r * Z + i) * dcmax;
                    uint16_t t[64];
                    int v[64];
                    unsigned min1 = 0x7C00u, min2 = 0x7C00u, P = 0;
                    for (int k = 0; k < dc; k++) {
                        v[k] = ly[r].col[k] * Z + (i + ly[r].shift[k]) % Z;
                        t[k] = orc_f2h(orc_h2f(a[v[k]]) - orc_h2f(mr[k]));
                        const unsigned ak = t[k] & 0x7FFFu; /* positive patterns order like integers */
                        P ^= t[k] >> 15;
                        if (ak < min1) {
                            min2 = min1;
                            min1 = ak;
                        } else if (ak < min2)
                            min2 = ak;
                    }
                    unsigned m1 = min1 < amaxh ? min1 : amaxh, m2 = min2 < amaxh ? min2 : amaxh;
                    if (bnum != 0) {
                        m1 = orc_f2h(orc_h2f((uint16_t)m1) * c);
                        m2 = orc_f2h(orc_h2f((uint16_t)m2) * c);
                    }
                    for (int k = 0; k < dc; k++) {
                        const unsigned ak = t[k] & 0x7FFFu;
                        const unsigned s = P ^ (t[k] >> 15);
                        const uint16_t nw = (uint16_t)((ak == min1 ? m2 : m1) | (s << 15));
                        a[v[k]] = orc_f2h(orc_h2f(t[k]) + orc_h2f(nw));
                        mr[k] = nw;
                    }
                }
            }
            if (exit_mode == ORC_EXIT_SYNDROME || it == maxit) {
                okf = 1;
                for (int r = 0; r < J && okf; r++)
                    for (int i = 0; i < Z && okf; i++) {
                        unsigned p = 0;
                        for (int k = 0; k < ly[r].dc; k++) p ^= a[ly[r].col[k] * Z + (i + ly[r].shift[k]) % Z] >> 15;
                        if (p) okf = 0;
                    }
                if (okf && exit_mode == ORC_EXIT_SYNDROME) break;
            }
        }
        for (int n = 0; n < N; n++) {
            D[(size_t)n * F + f] = a[n] >> 15;
            if (app_out) app_out[(size_t)n * F + f] = a[n];
        }
        if (msg_out)
            for (int m = 0; m < M; m++)
                for (int k = 0; k < dcmax; k++) msg_out[((size_t)m * dcmax + k) * F + f] = msg[(size_t)m * dcmax + k];
        D[(size_t)N * F + f] = okf;
        iters[f] = it;
        free(a);
        free(msg);
    }
    free(ly);
    return 0;
}

/* ------------------------------------------------------------------ statistics + sim loop */

int orc_statistic(orc_sim_counters *c, const int *codeword, const int *D, const int *iters,
                  int N, int F, int length, long leastErrorFrames, long leastTestFrames)
{
    /* B/Simulation.cu:245-285; num_Frames is advanced by the caller (:113) */
    for (int f = 0; f < F; f++) {
        int err = 0;
        for (int n = 0; n < length; n++) {
            int cw = codeword ? codeword[n * F + f] : 0;
            if (D[n * F + f] != cw) err++;
        }
        int flag = D[N * F + f];
        c->num_Error_Bits += err;
        if (err != 0 || flag == 0) c->num_Error_Frames++;
        if (err == 0 && flag == 0) c->num_Alarm_Frames++;
        if (err != 0 && flag == 1) c->num_False_Frames++;
        c->Total_Iteration += iters[f];
    }
    return (c->num_Error_Frames >= leastErrorFrames && c->num_Frames >= leastTestFrames) ? 1 : 0;
}

int orc_sim_point(int J, int L, int Z, const int *H, const int *Wc, const int *Wv,
                  const int *addr, int snrtype, float snr_db, int F, int maxit, int length,
                  int exit_mode, long leastErrorFrames, long leastTestFrames, long max_frames,
                  int seed0, orc_sim_counters *out)
{
    const int N = L * Z;
    int seed[3] = {seed0, seed0, seed0}; /* B/main.cu:117-119 */
    float rate = (float)length / (float)N; /* B/define.cuh:31 */
    float sigma = orc_sigma(snrtype, snr_db, rate);
    float *y = (float *)malloc((size_t)N * F * sizeof(float));
    int *D = (int *)malloc((size_t)(N + 1) * F * sizeof(int));
    int *iters = (int *)malloc((size_t)F * sizeof(int));
    if (!y || !D || !iters) return -5;
    memset(out, 0, sizeof(*out));
    while (1) { /* B/Simulation.cu:111-156 */
        out->num_Frames += F;
        orc_awgn(seed, sigma, NULL, y, N, F);
        orc_flooding_fp32(J, L, Z, H, Wc, Wv, addr, y, F, maxit, exit_mode, length, D, iters, NULL);
        if (orc_statistic(out, NULL, D, iters, N, F, length, leastErrorFrames, leastTestFrames))
            break;
        if (max_frames > 0 && out->num_Frames >= max_frames) break;
    }
    free(y);
    free(D);
    free(iters);
    return 0;
}
