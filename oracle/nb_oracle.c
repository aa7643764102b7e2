/*
 * nb_oracle.c — CPU oracle for the non-binary GF(q) LDPC decode path.  TEST INFRASTRUCTURE ONLY.
 * Follows NB/src/LDPC_Decoder.cpp (whole file), NB/src/GF.cpp:48-117, NB/src/LDPC_Encoder.cpp,
 * NB/src/Simulation.cpp:313-466, NB/src/main.cu:190-228 of gsw4869/CUDA_LDPC (NB/ = myNBLDPC/).
 * See nb_oracle.h for what is pinned and how.
 */
#include "nb_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NB_PI (3.1415926) /* NB/include/define.h:57 */

struct nb_orc_code {
    int N, M, q, p, dv_max, dc_max, n_qam;
    int *vw, *cw;           /* weights */
    int *v_cn, *v_gf;       /* [N][dv_max] */
    int *c_vn, *c_gf;       /* [M][dc_max] */
    int *v_pos;             /* [N][dv_max]: index_in_CN(col, d)  (NB/src/LDPC_Decoder.cpp:119-130) */
    int *c_pos;             /* [M][dc_max]: index_in_VN(row, d)  (:106-117) */
    int *mul, *inv;         /* [q][q], [q] */
    float *cre, *cim;       /* constellation */
    float rate;
};

static int read_int(FILE *f, int *v) { return fscanf(f, "%d", v) == 1; }

nb_orc_code *nb_orc_load(const char *matrix, const char *gf_table, const char *constellation, int n_qam,
                         int coef_is_exponent)
{
    FILE *f = fopen(matrix, "r");
    if (!f) return NULL;
    nb_orc_code *c = (nb_orc_code *)calloc(1, sizeof(*c));
    int ok = read_int(f, &c->N) && read_int(f, &c->M) && read_int(f, &c->q) && read_int(f, &c->dv_max) &&
             read_int(f, &c->dc_max);
    if (!ok) goto fail;
    c->p = 0;
    while ((1 << c->p) < c->q) c->p++;
    c->rate = (float)(c->N - c->M) / c->N; /* NB/src/Simulation.cpp:363 */
    c->vw = (int *)calloc(c->N, sizeof(int));
    c->cw = (int *)calloc(c->M, sizeof(int));
    c->v_cn = (int *)calloc((size_t)c->N * c->dv_max, sizeof(int));
    c->v_gf = (int *)calloc((size_t)c->N * c->dv_max, sizeof(int));
    c->c_vn = (int *)malloc((size_t)c->M * c->dc_max * sizeof(int));
    c->c_gf = (int *)malloc((size_t)c->M * c->dc_max * sizeof(int));
    c->v_pos = (int *)calloc((size_t)c->N * c->dv_max, sizeof(int));
    c->c_pos = (int *)calloc((size_t)c->M * c->dc_max, sizeof(int));
    for (int i = 0; i < c->M * c->dc_max; i++) c->c_vn[i] = c->c_gf[i] = -1;
    for (int i = 0; i < c->N; i++)
        if (!read_int(f, &c->vw[i]) || c->vw[i] > c->dv_max) goto fail;
    for (int i = 0; i < c->M; i++)
        if (!read_int(f, &c->cw[i]) || c->cw[i] > c->dc_max) goto fail;
    for (int i = 0; i < c->N; i++)
        for (int j = 0; j < c->vw[i]; j++) {
            int a, b;
            if (!read_int(f, &a) || !read_int(f, &b)) goto fail;
            c->v_cn[i * c->dv_max + j] = a - 1;
            c->v_gf[i * c->dv_max + j] = b;
        }
    for (int i = 0; i < c->M; i++)
        for (int j = 0; j < c->cw[i]; j++) {
            int a, b;
            if (!read_int(f, &a) || !read_int(f, &b)) goto fail;
            c->c_vn[i * c->dc_max + j] = a - 1;
            c->c_gf[i * c->dc_max + j] = b;
        }
    fclose(f);
    f = NULL;
    /* GF tables, NB/src/GF.cpp:83-112 */
    f = fopen(gf_table, "r");
    if (!f) goto fail;
    {
        char line[512], w1[64], w2[64];
        const int q = c->q;
        c->mul = (int *)malloc((size_t)q * q * sizeof(int));
        c->inv = (int *)malloc((size_t)q * sizeof(int));
        if (!fgets(line, sizeof line, f)) goto fail;
        if (fscanf(f, "%63s %63s", w1, w2) != 2) goto fail;
        for (int i = 0; i < q * q; i++)
            if (!read_int(f, &c->mul[i])) goto fail;
        if (fscanf(f, "%63s %63s", w1, w2) != 2) goto fail;
        for (int i = 0; i < q * q; i++) {
            int a;
            if (!read_int(f, &a)) goto fail;
            if (a != ((i / q) ^ (i % q))) goto fail; /* GFAdd is XOR, NB/src/GF.cpp:48 */
        }
        if (fscanf(f, "%63s %63s", w1, w2) != 2) goto fail;
        for (int i = 0; i < q; i++)
            if (!read_int(f, &c->inv[i])) goto fail;
    }
    fclose(f);
    f = NULL;
    if (coef_is_exponent) { /* *_exp.txt store the exponent of alpha = 2 */
        for (int pass = 0; pass < 2; pass++) {
            int *g = pass ? c->c_gf : c->v_gf;
            int n = pass ? c->M * c->dc_max : c->N * c->dv_max;
            for (int i = 0; i < n; i++) {
                if (g[i] < 0) continue;
                int a = 1;
                for (int e = 0; e < g[i]; e++) a = c->mul[a * c->q + 2];
                g[i] = a;
            }
        }
    }
    /* index_in_CN / index_in_VN: first match */
    for (int i = 0; i < c->N; i++)
        for (int d = 0; d < c->vw[i]; d++) {
            int cn = c->v_cn[i * c->dv_max + d], pos = -1;
            for (int k = 0; k < c->cw[cn]; k++)
                if (c->c_vn[cn * c->dc_max + k] == i) {
                    pos = k;
                    break;
                }
            if (pos < 0) goto fail;
            c->v_pos[i * c->dv_max + d] = pos;
        }
    for (int i = 0; i < c->M; i++)
        for (int d = 0; d < c->cw[i]; d++) {
            int vn = c->c_vn[i * c->dc_max + d], pos = -1;
            for (int k = 0; k < c->vw[vn]; k++)
                if (c->v_cn[vn * c->dv_max + k] == i) {
                    pos = k;
                    break;
                }
            if (pos < 0) goto fail;
            c->c_pos[i * c->dc_max + d] = pos;
        }
    c->n_qam = n_qam;
    if (constellation) { /* NB/src/Simulation.cpp:313-338: "%s %d %s %f %s %f" per point */
        f = fopen(constellation, "r");
        if (!f) goto fail;
        c->cre = (float *)calloc(c->q > n_qam ? c->q : n_qam, sizeof(float));
        c->cim = (float *)calloc(c->q > n_qam ? c->q : n_qam, sizeof(float));
        for (int k = 0; k < n_qam; k++) {
            char t[100];
            int idx;
            float re, im;
            if (fscanf(f, "%99s %d %99s %f %99s %f", t, &idx, t, &re, t, &im) != 6 || idx < 0 || idx >= n_qam)
                goto fail;
            c->cre[idx] = re;
            c->cim[idx] = im;
        }
        fclose(f);
        f = NULL;
    }
    return c;
fail:
    if (f) fclose(f);
    nb_orc_free(c);
    return NULL;
}

void nb_orc_free(nb_orc_code *c)
{
    if (!c) return;
    free(c->vw); free(c->cw); free(c->v_cn); free(c->v_gf); free(c->c_vn); free(c->c_gf);
    free(c->v_pos); free(c->c_pos); free(c->mul); free(c->inv); free(c->cre); free(c->cim);
    free(c);
}

void nb_orc_info(const nb_orc_code *c, int *out)
{
    out[0] = c->N; out[1] = c->M; out[2] = c->q; out[3] = c->p; out[4] = c->dv_max; out[5] = c->dc_max;
    out[6] = c->n_qam;
}

void nb_orc_tables(const nb_orc_code *c, int *mul, int *inv, int *check_vn, int *check_coef, int *check_w)
{
    if (mul) memcpy(mul, c->mul, (size_t)c->q * c->q * sizeof(int));
    if (inv) memcpy(inv, c->inv, (size_t)c->q * sizeof(int));
    if (check_vn) memcpy(check_vn, c->c_vn, (size_t)c->M * c->dc_max * sizeof(int));
    if (check_coef) memcpy(check_coef, c->c_gf, (size_t)c->M * c->dc_max * sizeof(int));
    if (check_w) memcpy(check_w, c->cw, (size_t)c->M * sizeof(int));
}

void nb_orc_constellation(const nb_orc_code *c, float *re, float *im)
{
    memcpy(re, c->cre, (size_t)c->n_qam * sizeof(float));
    memcpy(im, c->cim, (size_t)c->n_qam * sizeof(float));
}

int nb_orc_syndrome_ok(const nb_orc_code *c, const int *x)
{
    for (int row = 0; row < c->M; row++) {
        int s = 0;
        for (int i = 0; i < c->cw[row]; i++)
            s ^= c->mul[x[c->c_vn[row * c->dc_max + i]] * c->q + c->c_gf[row * c->dc_max + i]];
        if (s) return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------ channel */

static float random_module(int *seed)
{
    /* NB/src/LDPC_Encoder.cpp:70-79 */
    float temp = 0.0f;
    seed[0] = (seed[0] * 249) % 61967;
    seed[1] = (seed[1] * 251) % 63443;
    seed[2] = (seed[2] * 252) % 63599;
    temp = (((float)seed[0]) / ((float)61967)) + (((float)seed[1]) / ((float)63443)) +
           (((float)seed[2]) / ((float)63599));
    temp -= (int)temp;
    return temp;
}

void nb_orc_awgn(int *seed, float sigma, const float *tx, float *rx, int len)
{
    /* NB/src/LDPC_Encoder.cpp:41-68 */
    for (int i = 0; i < len; i++) {
        float u1 = random_module(seed);
        float u2 = random_module(seed);
        float temp = sqrtf((float)(-2) * logf((float)1 - u1));
        rx[2 * i] = (float)((double)sigma * cos(2 * NB_PI * (double)u2) * (double)temp + (double)tx[2 * i]);
        u1 = random_module(seed);
        u2 = random_module(seed);
        temp = sqrtf((float)(-2) * logf((float)1 - u1));
        rx[2 * i + 1] = (float)((double)sigma * cos(2 * NB_PI * (double)u2) * (double)temp + (double)tx[2 * i + 1]);
    }
}

int nb_orc_modulate(const nb_orc_code *c, const int *sym, float *tx)
{
    /* NB/src/main.cu:190-212 + NB/src/LDPC_Encoder.cpp:18-37 */
    if (c->n_qam != 2) {
        for (int s = 0; s < c->N; s++) {
            tx[2 * s] = c->cre[sym[s]];
            tx[2 * s + 1] = c->cim[sym[s]];
        }
        return c->N;
    }
    for (int i = 0; i < c->N; i++)
        for (int j = 0; j < c->p; j++) {
            int bit = (sym[i] & (1 << j)) >> j; /* LSB first */
            tx[2 * (i * c->p + j)] = c->cre[bit];
            tx[2 * (i * c->p + j) + 1] = c->cim[bit];
        }
    return c->N * c->p;
}

float nb_orc_sigma(const nb_orc_code *c, int snrtype, float snr_db)
{
    /* NB/src/main.cu:221-228 */
    if (snrtype == 0)
        return (float)sqrt(0.5 / (log(c->n_qam) / log(2) * c->rate * (pow(10.0, (snr_db / 10.0)))));
    return (float)sqrt(0.5 / (log(c->n_qam) / log(2) * pow(10.0, (snr_db / 10.0))));
}

void nb_orc_demodulate(const nb_orc_code *c, float sigma, const float *rx, float *L_ch)
{
    /* NB/src/LDPC_Decoder.cpp:132-171 */
    const int q = c->q;
    if (c->n_qam == 2) {
        float *llr = (float *)malloc((size_t)c->N * c->p * sizeof(float));
        for (int b = 0; b < c->N * c->p; b++) llr[b] = -2 * rx[2 * b] / (sigma * sigma);
        for (int s = 0; s < c->N; s++)
            for (int a = 1; a < q; a++) {
                float v = 0;
                for (int b = 0; b < c->p; b++)
                    if ((a & (1 << b)) != 0) v += llr[s * c->p + b];
                L_ch[s * (q - 1) + a - 1] = v;
            }
        free(llr);
    } else {
        for (int s = 0; s < c->N; s++)
            for (int a = 1; a < q; a++)
                L_ch[s * (q - 1) + a - 1] =
                    ((2 * rx[2 * s] - c->cre[0] - c->cre[a]) * (c->cre[a] - c->cre[0]) +
                     (2 * rx[2 * s + 1] - c->cim[0] - c->cim[a]) * (c->cim[a] - c->cim[0])) /
                    (2 * sigma * sigma);
    }
}

/* ------------------------------------------------------------------ decoders */

typedef struct {
    const nb_orc_code *c;
    float *v2c;   /* sort_L_v2c [N][dv_max][q] */
    int *ent;     /* sort_Entr_v2c [N][dv_max][q] */
    float *c2v;   /* L_c2v [M][dc_max][q] */
    float *LLR;   /* [N][q] */
    int sum_mode;
} nb_state;

#define V2C(s, col, d) ((s)->v2c + ((size_t)(col) * (s)->c->dv_max + (d)) * (s)->c->q)
#define ENT(s, col, d) ((s)->ent + ((size_t)(col) * (s)->c->dv_max + (d)) * (s)->c->q)
#define C2V(s, row, d) ((s)->c2v + ((size_t)(row) * (s)->c->dc_max + (d)) * (s)->c->q)

static void bubble_desc(float *a, int n, int *index)
{
    /* NB/src/LDPC_Decoder.cpp:17-36: swap only when strictly smaller => stable */
    for (int i = 0; i < n; i++)
        for (int j = 1; j < n - i; j++)
            if (a[j - 1] < a[j]) {
                float x = a[j];
                a[j] = a[j - 1];
                a[j - 1] = x;
                int t = index[j];
                index[j] = index[j - 1];
                index[j - 1] = t;
            }
}

static int decide_max(const float *LLR, int q)
{
    /* DecideLLRVector :71-91 */
    float max = 0;
    int alpha_i = 0;
    for (int a = 0; a < q - 1; a++)
        if (LLR[a] > max) {
            max = LLR[a];
            alpha_i = a + 1;
        }
    return (max <= 0) ? 0 : alpha_i;
}

static int decide_min(const float *LLR, int q)
{
    /* d_DecideLLRVector :92-105 */
    float min = INFINITY; /* float min = DBL_MAX */
    int alpha_i = 0;
    for (int a = 0; a < q; a++)
        if (LLR[a] < min) {
            min = LLR[a];
            alpha_i = a;
        }
    return alpha_i;
}

/* literal ConstructConf, :319-359 */
static void conf_literal(nb_state *s, int Nm, int Nc, int *sumNonele, float *sumNonLLR, int *diff, int begin,
                         int except, int end, int row, float *E)
{
    const nb_orc_code *c = s->c;
    if (begin > end) {
        if (*sumNonLLR > E[*sumNonele]) E[*sumNonele] = *sumNonLLR;
    } else if (begin == except) {
        conf_literal(s, Nm, Nc, sumNonele, sumNonLLR, diff, begin + 1, except, end, row, E);
    } else {
        const int vn = c->c_vn[row * c->dc_max + begin], index = c->c_pos[row * c->dc_max + begin];
        const int h = c->c_gf[row * c->dc_max + begin];
        const float *L = V2C(s, vn, index);
        const int *S = ENT(s, vn, index);
        for (int k = 0; k < Nm; k++) {
            *sumNonele = c->mul[S[k] * c->q + h] ^ *sumNonele;
            *sumNonLLR = *sumNonLLR + L[k];
            *diff += (k != 0) ? 1 : 0;
            if (*diff <= Nc) {
                conf_literal(s, Nm, Nc, sumNonele, sumNonLLR, diff, begin + 1, except, end, row, E);
                *sumNonele = c->mul[S[k] * c->q + h] ^ *sumNonele;
                *sumNonLLR = *sumNonLLR - L[k];
                *diff -= (k != 0) ? 1 : 0;
            } else {
                *sumNonele = c->mul[S[k] * c->q + h] ^ *sumNonele;
                *sumNonLLR = *sumNonLLR - L[k];
                *diff -= (k != 0) ? 1 : 0;
                break;
            }
        }
    }
}

/* `fresh` variant: same configurations, but every leaf sums its inputs from 0.0f in ascending
 * edge position (no add/subtract drift).  ks[] holds the chosen sorted index per input. */
static void conf_fresh(nb_state *s, int Nm, int Nc, int *ks, int diff, int begin, int except, int end, int row,
                       float *E)
{
    const nb_orc_code *c = s->c;
    if (begin > end) {
        int sym = 0;
        float sum = 0.0f;
        for (int b = 0; b <= end; b++) {
            if (b == except) continue;
            const int vn = c->c_vn[row * c->dc_max + b], index = c->c_pos[row * c->dc_max + b];
            sym ^= c->mul[ENT(s, vn, index)[ks[b]] * c->q + c->c_gf[row * c->dc_max + b]];
            sum = sum + V2C(s, vn, index)[ks[b]];
        }
        if (sum > E[sym]) E[sym] = sum;
    } else if (begin == except) {
        conf_fresh(s, Nm, Nc, ks, diff, begin + 1, except, end, row, E);
    } else {
        for (int k = 0; k < Nm; k++) {
            int d2 = diff + ((k != 0) ? 1 : 0);
            if (d2 > Nc) break;
            ks[begin] = k;
            conf_fresh(s, Nm, Nc, ks, d2, begin + 1, except, end, row, E);
        }
    }
}

static int syndrome_first_fail(const nb_orc_code *c, const int *out)
{
    /* :218-231: sum_temp is never reset, but the scan aborts at the first non-zero value */
    int sum_temp = 0;
    for (int row = 0; row < c->M; row++) {
        for (int i = 0; i < c->cw[row]; i++)
            sum_temp ^= c->mul[out[c->c_vn[row * c->dc_max + i]] * c->q + c->c_gf[row * c->dc_max + i]];
        if (sum_temp) return 0;
    }
    return 1;
}

static int decode_ems(nb_state *s, const float *L_ch, int maxit, int Nm_in, int Nc_in, int *out, int *iter_number)
{
    const nb_orc_code *c = s->c;
    const int q = c->q, N = c->N, M = c->M;
    float *E = (float *)malloc((size_t)q * sizeof(float));
    int *index = (int *)malloc((size_t)q * sizeof(int));
    memset(s->c2v, 0, (size_t)M * c->dc_max * q * sizeof(float));
    *iter_number = 0;
    while (*iter_number < maxit) {
        (*iter_number)++;
        for (int col = 0; col < N; col++) {
            float *LLR = s->LLR + (size_t)col * q;
            memcpy(LLR, L_ch + (size_t)col * (q - 1), (size_t)(q - 1) * sizeof(float));
            for (int d = 0; d < c->vw[col]; d++) {
                const float *m = C2V(s, c->v_cn[col * c->dv_max + d], c->v_pos[col * c->dv_max + d]);
                for (int a = 0; a < q - 1; a++) LLR[a] += m[a];
            }
            out[col] = decide_max(LLR, q);
        }
        if (syndrome_first_fail(c, out)) {
            free(E);
            free(index);
            (*iter_number)--;
            return 1;
        }
        for (int col = 0; col < N; col++) {
            const float *LLR = s->LLR + (size_t)col * q;
            for (int dv = 0; dv < c->vw[col]; dv++) {
                const float *m = C2V(s, c->v_cn[col * c->dv_max + dv], c->v_pos[col * c->dv_max + dv]);
                float *v = V2C(s, col, dv);
                int *en = ENT(s, col, dv);
                for (int a = 0; a < q - 1; a++) v[a] = LLR[a] - m[a];
                v[q - 1] = 0;
                for (int i = 0; i < q - 1; i++) index[i] = i + 1;
                index[q - 1] = 0;
                bubble_desc(v, q, index);
                for (int i = 0; i < q; i++) en[i] = index[i];
            }
        }
        for (int row = 0; row < M; row++) {
            const int w = c->cw[row];
            for (int dc = 0; dc < w; dc++) {
                for (int a = 0; a < q; a++) E[a] = -INFINITY; /* EMS_L_c2v[q] = -DBL_MAX as float */
                const int Nc2 = (Nc_in == c->dc_max - 1) ? w - 1 : Nc_in; /* :297-304 */
                if (s->sum_mode == NB_ORC_SUM_LITERAL) {
                    int sumNonele = 0, diff = 0;
                    float sumNonLLR = 0;
                    conf_literal(s, q, 1, &sumNonele, &sumNonLLR, &diff, 0, dc, w - 1, row, E);
                    sumNonele = 0;
                    sumNonLLR = 0;
                    diff = 0;
                    conf_literal(s, Nm_in, Nc2, &sumNonele, &sumNonLLR, &diff, 0, dc, w - 1, row, E);
                } else {
                    int ks[64];
                    memset(ks, 0, sizeof ks);
                    conf_fresh(s, q, 1, ks, 0, 0, dc, w - 1, row, E);
                    memset(ks, 0, sizeof ks);
                    conf_fresh(s, Nm_in, Nc2, ks, 0, 0, dc, w - 1, row, E);
                }
                float *m = C2V(s, row, dc);
                const int h = c->c_gf[row * c->dc_max + dc];
                for (int k = 1; k < q; k++) {
                    int v = c->mul[k * q + h];
                    m[k - 1] = (float)((E[v] - E[0]) / 1.2); /* float difference, double division (:309) */
                }
            }
        }
    }
    free(E);
    free(index);
    return 0;
}

/* TMM check node, :704-817; fills c2v of `row` from the v2c vectors */
static void tmm_check(nb_state *s, int row, int *Zn, float *dU, float *Min1, float *Min2, int *MinCol, float *I,
                      int *Path, float *Ev, float *Lc2p)
{
    const nb_orc_code *c = s->c;
    const int q = c->q, w = c->cw[row];
    int syn = 0;
    for (int dc = 0; dc < w; dc++) { /* d_TMM_Get_Zn */
        const float *v = V2C(s, c->c_vn[row * c->dc_max + dc], c->c_pos[row * c->dc_max + dc]);
        double min = DBL_MAX;
        int min_ele = 0;
        for (int a = 0; a < q; a++)
            if (v[a] < min) {
                min = v[a];
                min_ele = c->mul[a * q + c->c_gf[row * c->dc_max + dc]];
            }
        Zn[dc] = min_ele;
        syn ^= min_ele;
    }
    for (int dc = 0; dc < w; dc++) { /* d_TMM_Get_deltaU */
        const float *v = V2C(s, c->c_vn[row * c->dc_max + dc], c->c_pos[row * c->dc_max + dc]);
        const int hinv = c->inv[c->c_gf[row * c->dc_max + dc]];
        const float min = v[c->mul[hinv * q + Zn[dc]]];
        for (int x = 0; x < q; x++) dU[dc * q + (x ^ Zn[dc])] = v[c->mul[hinv * q + x]] - min;
    }
    for (int a = 0; a < q; a++) { /* TMM_Get_Min */
        Min1[a] = INFINITY;
        Min2[a] = INFINITY;
        MinCol[a] = 0; /* uninitialised in the reference; always overwritten for finite inputs */
        for (int dc = 0; dc < w; dc++) {
            if (dU[dc * q + a] < Min1[a]) {
                Min2[a] = Min1[a];
                Min1[a] = dU[dc * q + a];
                MinCol[a] = dc;
            } else if (dU[dc * q + a] < Min2[a])
                Min2[a] = dU[dc * q + a];
        }
    }
    I[0] = 0; /* TMM_ConstructConf */
    Path[0] = Path[1] = -1;
    Ev[0] = 0;
    for (int i = 1; i < q; i++) {
        I[i] = dU[MinCol[i] * q + i];
        Path[i * 2] = Path[i * 2 + 1] = MinCol[i];
        Ev[i] = Min2[i];
        for (int j = 0; j < q; j++) {
            if (j == i) continue;
            int k = i ^ j;
            if (MinCol[j] != MinCol[k]) {
                double d1 = dU[MinCol[j] * q + j], d2 = dU[MinCol[k] * q + k];
                if (d1 > d2 && d1 < I[i]) {
                    I[i] = (float)d1;
                    Path[i * 2] = MinCol[j];
                    Path[i * 2 + 1] = MinCol[k];
                    Ev[i] = Min1[i];
                } else if (d1 < d2 && d2 < I[i]) {
                    I[i] = (float)d2;
                    Path[i * 2] = MinCol[j];
                    Path[i * 2 + 1] = MinCol[k];
                    Ev[i] = Min1[i];
                }
            }
        }
    }
    for (int dc = 0; dc < w; dc++) { /* :496-521 */
        Lc2p[0] = 0;
        for (int eta = 1; eta < q; eta++)
            Lc2p[eta] = (dc != Path[eta * 2] && dc != Path[eta * 2 + 1]) ? I[eta] : Ev[eta];
        const int hinv = c->inv[c->c_gf[row * c->dc_max + dc]];
        const int beta_syn = syn ^ Zn[dc];
        float *m = C2V(s, row, dc);
        for (int eta = 0; eta < q; eta++) m[c->mul[hinv * q + (eta ^ beta_syn)]] = (float)(Lc2p[eta] * 0.8);
    }
}

static int decode_tmm(nb_state *s, const float *L_ch, int maxit, int layered, int *out, int *iter_number)
{
    const nb_orc_code *c = s->c;
    const int q = c->q, N = c->N, M = c->M;
    for (int col = 0; col < N; col++) { /* :363-390 */
        float max = -INFINITY;
        for (int a = 0; a < q - 1; a++)
            if (L_ch[(size_t)col * (q - 1) + a] > max) max = L_ch[(size_t)col * (q - 1) + a];
        float *LLR = s->LLR + (size_t)col * q;
        LLR[0] = max;
        for (int a = 1; a < q; a++) LLR[a] = max - L_ch[(size_t)col * (q - 1) + a - 1];
        for (int d = 0; d < c->vw[col]; d++) memcpy(V2C(s, col, d), LLR, (size_t)q * sizeof(float));
    }
    memset(s->c2v, 0, (size_t)M * c->dc_max * q * sizeof(float));
    int *Zn = (int *)malloc((size_t)c->dc_max * sizeof(int));
    float *dU = (float *)malloc((size_t)c->dc_max * q * sizeof(float));
    float *Min1 = (float *)malloc((size_t)q * sizeof(float)), *Min2 = (float *)malloc((size_t)q * sizeof(float));
    int *MinCol = (int *)malloc((size_t)q * sizeof(int)), *Path = (int *)malloc((size_t)q * 2 * sizeof(int));
    float *I = (float *)malloc((size_t)q * sizeof(float)), *Ev = (float *)malloc((size_t)q * sizeof(float));
    float *Lc2p = (float *)malloc((size_t)q * sizeof(float));
    int ret = 0;
    *iter_number = 0;
    while (*iter_number < maxit) {
        (*iter_number)++;
        for (int col = 0; col < N; col++) {
            float *LLR = s->LLR + (size_t)col * q;
            if (!layered) /* :423-434: LLR is never reset, c2v accumulates over the iterations */
                for (int d = 0; d < c->vw[col]; d++) {
                    const float *m = C2V(s, c->v_cn[col * c->dv_max + d], c->v_pos[col * c->dv_max + d]);
                    for (int a = 0; a < q; a++) LLR[a] += m[a];
                }
            out[col] = decide_min(LLR, q);
        }
        if (syndrome_first_fail(c, out)) {
            (*iter_number)--;
            ret = 1;
            break;
        }
        if (!layered) {
            for (int col = 0; col < N; col++) {
                const float *LLR = s->LLR + (size_t)col * q;
                for (int dv = 0; dv < c->vw[col]; dv++) {
                    const float *m = C2V(s, c->v_cn[col * c->dv_max + dv], c->v_pos[col * c->dv_max + dv]);
                    float *v = V2C(s, col, dv);
                    for (int a = 0; a < q; a++) v[a] = LLR[a] - m[a];
                }
            }
            for (int row = 0; row < M; row++) tmm_check(s, row, Zn, dU, Min1, Min2, MinCol, I, Path, Ev, Lc2p);
        } else {
            for (int row = 0; row < M; row++) { /* :630-690 */
                for (int d = 0; d < c->cw[row]; d++) {
                    const int vn = c->c_vn[row * c->dc_max + d];
                    float *v = V2C(s, vn, c->c_pos[row * c->dc_max + d]);
                    const float *LLR = s->LLR + (size_t)vn * q, *m = C2V(s, row, d);
                    for (int a = 0; a < q; a++) v[a] = LLR[a] - m[a];
                }
                tmm_check(s, row, Zn, dU, Min1, Min2, MinCol, I, Path, Ev, Lc2p);
                for (int d = 0; d < c->cw[row]; d++) {
                    const int vn = c->c_vn[row * c->dc_max + d];
                    const float *v = V2C(s, vn, c->c_pos[row * c->dc_max + d]), *m = C2V(s, row, d);
                    float *LLR = s->LLR + (size_t)vn * q;
                    for (int a = 0; a < q; a++) LLR[a] = v[a] + m[a];
                }
            }
        }
    }
    free(Zn); free(dU); free(Min1); free(Min2); free(MinCol); free(Path); free(I); free(Ev); free(Lc2p);
    return ret;
}

/* ------------------------------------------------------------------ FFT-BP (parity unpinned)
 * Probability-domain sum-product over GF(2^p); the reference has no such decoder, so these
 * rules ARE the specification the CUDA kernel is held to (decisions / iteration counts; messages agree
 * to float rounding because expf differs between libm and CUDA):
 *   p_ch[0] = exp(0 - mx), p_ch[a] = exp(L_ch[a-1] - mx), mx = max(0, max L_ch); normalised by the sum
 *   taken in ascending a.  c2v = 1 before the first iteration.
 *   iteration: post[a] = (p_ch[a] * c2v_0[a]) * c2v_1[a] ...; decide = first argmax; syndrome stop with the
 *   reference's counting (iterations-1 on success).
 *   v2c_d[a] = p_ch[a] * prod_{d' != d, ascending} c2v_d'[a], normalised by its ascending sum.
 *   check: P_d[h_d*x] = v2c_d[x]; F_d = WHT(P_d) (in-place butterflies, len = 1, 2, 4, ...: (u, v) ->
 *   (u+v, u-v)); G = prod_{d' != d, ascending} F_d'; g = WHT(G); g = max(g, 1e-30f);
 *   c2v_d[x] = g[h_d*x], normalised by its ascending sum. */
static void wht(float *v, int q)
{
    for (int len = 1; len < q; len <<= 1)
        for (int i = 0; i < q; i += 2 * len)
            for (int j = i; j < i + len; j++) {
                float a = v[j], b = v[j + len];
                v[j] = a + b;
                v[j + len] = a - b;
            }
}

/* Sums of the FFT-BP decoder are pairwise trees, t[x] += t[x + len] for len = q/2 ... 1 (q is a power of
 * two): the order a GPU reduction follows, so that the kernel can be compared bit for bit. */
static float tree_sum(const float *v, int q)
{
    float t[512];
    for (int a = 0; a < q; a++) t[a] = v[a];
    for (int len = q / 2; len >= 1; len >>= 1)
        for (int x = 0; x < len; x++) t[x] = t[x] + t[x + len];
    return t[0];
}

static void normalise(float *v, int q)
{
    const float s = tree_sum(v, q);
    for (int a = 0; a < q; a++) v[a] = v[a] / s;
}

static int decode_fftbp(nb_state *s, const float *L_ch, int maxit, int *out, int *iter_number)
{
    const nb_orc_code *c = s->c;
    const int q = c->q, N = c->N, M = c->M;
    float *pch = (float *)malloc((size_t)N * q * sizeof(float));
    float *F = (float *)malloc((size_t)c->dc_max * q * sizeof(float)), *G = (float *)malloc((size_t)q * sizeof(float));
    for (int col = 0; col < N; col++) {
        float mx = 0.0f;
        for (int a = 0; a < q - 1; a++)
            if (L_ch[(size_t)col * (q - 1) + a] > mx) mx = L_ch[(size_t)col * (q - 1) + a];
        float *p = pch + (size_t)col * q;
        p[0] = expf(0.0f - mx);
        for (int a = 1; a < q; a++) p[a] = expf(L_ch[(size_t)col * (q - 1) + a - 1] - mx);
        normalise(p, q);
    }
    for (size_t i = 0; i < (size_t)M * c->dc_max * q; i++) s->c2v[i] = 1.0f;
    int ret = 0;
    *iter_number = 0;
    while (*iter_number < maxit) {
        (*iter_number)++;
        for (int col = 0; col < N; col++) {
            const float *p = pch + (size_t)col * q;
            float best = -1.0f;
            int bi = 0;
            for (int a = 0; a < q; a++) {
                float v = p[a];
                for (int d = 0; d < c->vw[col]; d++)
                    v = v * C2V(s, c->v_cn[col * c->dv_max + d], c->v_pos[col * c->dv_max + d])[a];
                if (v > best) {
                    best = v;
                    bi = a;
                }
            }
            out[col] = bi;
        }
        if (syndrome_first_fail(c, out)) {
            (*iter_number)--;
            ret = 1;
            break;
        }
        for (int col = 0; col < N; col++)
            for (int d = 0; d < c->vw[col]; d++) {
                float *v = V2C(s, col, d);
                for (int a = 0; a < q; a++) {
                    float x = pch[(size_t)col * q + a];
                    for (int d2 = 0; d2 < c->vw[col]; d2++)
                        if (d2 != d) x = x * C2V(s, c->v_cn[col * c->dv_max + d2], c->v_pos[col * c->dv_max + d2])[a];
                    v[a] = x;
                }
                normalise(v, q);
            }
        for (int row = 0; row < M; row++) {
            const int w = c->cw[row];
            for (int d = 0; d < w; d++) {
                const float *v = V2C(s, c->c_vn[row * c->dc_max + d], c->c_pos[row * c->dc_max + d]);
                const int h = c->c_gf[row * c->dc_max + d];
                for (int x = 0; x < q; x++) F[d * q + c->mul[x * q + h]] = v[x];
                wht(F + d * q, q);
            }
            for (int d = 0; d < w; d++) {
                for (int y = 0; y < q; y++) {
                    float g = 1.0f;
                    int first = 1;
                    for (int d2 = 0; d2 < w; d2++) {
                        if (d2 == d) continue;
                        g = first ? F[d2 * q + y] : g * F[d2 * q + y];
                        first = 0;
                    }
                    G[y] = g;
                }
                wht(G, q);
                float *m = C2V(s, row, d);
                const int h = c->c_gf[row * c->dc_max + d];
                for (int x = 0; x < q; x++) {
                    float g = G[c->mul[x * q + h]];
                    m[x] = g > 1e-30f ? g : 1e-30f;
                }
                normalise(m, q);
            }
        }
    }
    free(pch);
    free(F);
    free(G);
    return ret;
}

int nb_orc_decode(const nb_orc_code *c, int algo, int sum_mode, const float *L_ch, int maxit, int ems_nm,
                  int ems_nc, int *out, int *iter_number)
{
    nb_state s;
    s.c = c;
    s.sum_mode = sum_mode;
    s.v2c = (float *)calloc((size_t)c->N * c->dv_max * c->q, sizeof(float));
    s.ent = (int *)calloc((size_t)c->N * c->dv_max * c->q, sizeof(int));
    s.c2v = (float *)calloc((size_t)c->M * c->dc_max * c->q, sizeof(float));
    s.LLR = (float *)calloc((size_t)c->N * c->q, sizeof(float));
    int r;
    if (algo == NB_ORC_EMS)
        r = decode_ems(&s, L_ch, maxit, ems_nm, ems_nc, out, iter_number);
    else if (algo == NB_ORC_FFT_BP)
        r = decode_fftbp(&s, L_ch, maxit, out, iter_number);
    else
        r = decode_tmm(&s, L_ch, maxit, algo == NB_ORC_LAYERED_TMM, out, iter_number);
    free(s.v2c); free(s.ent); free(s.c2v); free(s.LLR);
    return r;
}

void nb_orc_decode_batch(const nb_orc_code *c, int algo, int sum_mode, const float *L_ch, int F, int maxit,
                         int ems_nm, int ems_nc, int *out, int *iters, int *ok)
{
#pragma omp parallel for schedule(dynamic)
    for (int f = 0; f < F; f++)
        ok[f] = nb_orc_decode(c, algo, sum_mode, L_ch + (size_t)f * c->N * (c->q - 1), maxit, ems_nm, ems_nc,
                              out + (size_t)f * c->N, &iters[f]);
}
