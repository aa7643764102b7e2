// Host-side loader of the non-binary path: replaces Get_H (NB/src/Simulation.cpp:347-466),
// GFInitial (NB/src/GF.cpp:68-117), Get_CONSTELLATION (NB/src/Simulation.cpp:313-338) and the
// link-table flattening of NB/src/main.cu:101-188.  File formats are the reference's, verbatim.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.h"
#include "nb_common.h"

namespace {

bool read_int(FILE *f, int *v) { return fscanf(f, "%d", v) == 1; }

int primitive_poly(int q)
{
    switch (q) {  // the polynomials of the reference's GF/Arith.Table.GF.<q>.txt headers
        case 4: return 7;
        case 8: return 11;
        case 16: return 19;
        case 32: return 37;
        case 64: return 67;
        case 128: return 137;
        case 256: return 285;
        case 512: return 529;
        default: return 0;
    }
}

// polynomial-basis tables, alpha = 2
bool generate_gf(int q, std::vector<uint16_t> &mul, std::vector<uint16_t> &inv)
{
    const int poly = primitive_poly(q);
    if (!poly) return false;
    mul.assign((size_t)q * q, 0);
    inv.assign(q, 0);
    for (int a = 0; a < q; a++)
        for (int b = 0; b < q; b++) {
            int r = 0, x = a, y = b;
            while (y) {
                if (y & 1) r ^= x;
                y >>= 1;
                x <<= 1;
                if (x & q) x ^= poly;
            }
            mul[(size_t)a * q + b] = (uint16_t)r;
        }
    for (int a = 1; a < q; a++)
        for (int b = 1; b < q; b++)
            if (mul[(size_t)a * q + b] == 1) {
                inv[a] = (uint16_t)b;
                break;
            }
    return true;
}

// "GF(q) with Primitive Polynomial: P." / "Multiply Table:" q*q / "Add Table:" q*q / "Inverse Table:" q
int parse_gf_file(const char *path, int q, std::vector<uint16_t> &mul, std::vector<uint16_t> &inv)
{
    FILE *f = fopen(path, "r");
    if (!f) return LDPC_ERR_IO;
    char line[512], w1[64], w2[64];
    int rc = LDPC_ERR_FORMAT;
    mul.assign((size_t)q * q, 0);
    inv.assign(q, 0);
    do {
        if (!fgets(line, sizeof line, f)) break;
        if (fscanf(f, "%63s %63s", w1, w2) != 2) break;
        bool ok = true;
        for (int i = 0; i < q * q && ok; i++) {
            int v;
            ok = read_int(f, &v) && v >= 0 && v < q;
            if (ok) mul[i] = (uint16_t)v;
        }
        if (!ok || fscanf(f, "%63s %63s", w1, w2) != 2) break;
        for (int i = 0; i < q * q && ok; i++) {
            int v;
            ok = read_int(f, &v) && v == ((i / q) ^ (i % q));  // GFAdd is XOR (NB/src/GF.cpp:48)
        }
        if (!ok || fscanf(f, "%63s %63s", w1, w2) != 2) break;
        for (int i = 0; i < q && ok; i++) {
            int v;
            ok = read_int(f, &v) && v >= 0 && v < q;
            if (ok) inv[i] = (uint16_t)v;
        }
        if (!ok) break;
        for (int a = 1; a < q && ok; a++) ok = mul[(size_t)a * q + inv[a]] == 1;
        if (ok) rc = LDPC_OK;
    } while (0);
    fclose(f);
    return rc;
}

}  // namespace

extern "C" int nb_ldpc_load_code(const char *matrix, const char *gf_table, const char *constellation,
                                 int coef_is_exponent, nb_ldpc_code_t **out)
{
    if (!matrix || !out) return LDPC_ERR_ARG;
    *out = nullptr;
    FILE *f = fopen(matrix, "r");
    if (!f) return LDPC_ERR_IO;
    nb_ldpc_code *c = new (std::nothrow) nb_ldpc_code();
    if (!c) {
        fclose(f);
        return LDPC_ERR_NOMEM;
    }
    c->device = -1;
    c->d_mul = nullptr;
    c->scratch = nullptr;
    c->scratch_bytes = 0;
    c->last_use = nullptr;
    c->last_use_valid = false;
    int rc = LDPC_ERR_FORMAT;
    do {
        if (!(read_int(f, &c->N) && read_int(f, &c->M) && read_int(f, &c->q) && read_int(f, &c->dv_max) &&
              read_int(f, &c->dc_max)))
            break;
        if (c->N <= 0 || c->M <= 0 || c->M >= c->N || c->dv_max <= 0 || c->dc_max <= 0) break;
        c->p = 0;
        while ((1 << c->p) < c->q) c->p++;
        if ((1 << c->p) != c->q || c->q < 4 || c->q > 512 || c->dc_max > 32 || c->dv_max > 8) {
            rc = LDPC_ERR_UNSUPPORTED;
            break;
        }
        c->rate = (float)(c->N - c->M) / c->N;  // NB/src/Simulation.cpp:363
        c->vw.assign(c->N, 0);
        c->cw.assign(c->M, 0);
        c->v_cn.assign((size_t)c->N * c->dv_max, -1);
        c->v_gf.assign((size_t)c->N * c->dv_max, -1);
        c->v_pos.assign((size_t)c->N * c->dv_max, -1);
        c->c_vn.assign((size_t)c->M * c->dc_max, -1);
        c->c_gf.assign((size_t)c->M * c->dc_max, -1);
        c->c_pos.assign((size_t)c->M * c->dc_max, -1);
        bool ok = true;
        for (int i = 0; i < c->N && ok; i++) ok = read_int(f, &c->vw[i]) && c->vw[i] >= 1 && c->vw[i] <= c->dv_max;
        for (int i = 0; i < c->M && ok; i++) ok = read_int(f, &c->cw[i]) && c->cw[i] >= 1 && c->cw[i] <= c->dc_max;
        c->E = 0;
        for (int i = 0; i < c->N && ok; i++)
            for (int j = 0; j < c->vw[i] && ok; j++) {
                int a, b;
                ok = read_int(f, &a) && read_int(f, &b) && a >= 1 && a <= c->M && b >= 0;  // 1-based check index
                c->v_cn[(size_t)i * c->dv_max + j] = a - 1;
                c->v_gf[(size_t)i * c->dv_max + j] = b;
                c->E++;
            }
        for (int i = 0; i < c->M && ok; i++)
            for (int j = 0; j < c->cw[i] && ok; j++) {
                int a, b;
                ok = read_int(f, &a) && read_int(f, &b) && a >= 1 && a <= c->N && b >= 0;
                c->c_vn[(size_t)i * c->dc_max + j] = a - 1;
                c->c_gf[(size_t)i * c->dc_max + j] = b;
            }
        if (!ok) break;
        // GF arithmetic
        if (gf_table) {
            rc = parse_gf_file(gf_table, c->q, c->mul, c->inv);
            if (rc != LDPC_OK) break;
            rc = LDPC_ERR_FORMAT;
        } else if (!generate_gf(c->q, c->mul, c->inv)) {
            rc = LDPC_ERR_UNSUPPORTED;
            break;
        }
        // coefficients: element form, or exponent of alpha = 2 (the *_exp.txt files)
        for (int pass = 0; pass < 2 && ok; pass++) {
            std::vector<int> &g = pass ? c->c_gf : c->v_gf;
            for (size_t i = 0; i < g.size() && ok; i++) {
                if (g[i] < 0) continue;
                if (coef_is_exponent) {
                    int a = 1;
                    for (int e = 0; e < g[i] % (c->q - 1); e++) a = c->mul[(size_t)a * c->q + 2];
                    g[i] = a;
                }
                // element 0 is tolerated (the reference reads *_exp.txt raw and meets 0 there);
                // the trellis min-max decoders need h^-1 and refuse such a code at decode time
                ok = g[i] >= 0 && g[i] < c->q;
            }
        }
        if (!ok) break;
        // index_in_CN / index_in_VN (NB/src/LDPC_Decoder.cpp:106-130): first match; the two edge
        // lists of the file must describe the same graph with the same coefficients
        for (int i = 0; i < c->N && ok; i++)
            for (int d = 0; d < c->vw[i] && ok; d++) {
                const int cn = c->v_cn[(size_t)i * c->dv_max + d];
                int pos = -1;
                for (int k = 0; k < c->cw[cn]; k++)
                    if (c->c_vn[(size_t)cn * c->dc_max + k] == i) {
                        pos = k;
                        break;
                    }
                ok = pos >= 0 && c->c_gf[(size_t)cn * c->dc_max + pos] == c->v_gf[(size_t)i * c->dv_max + d];
                c->v_pos[(size_t)i * c->dv_max + d] = pos;
            }
        for (int i = 0; i < c->M && ok; i++)
            for (int d = 0; d < c->cw[i] && ok; d++) {
                const int vn = c->c_vn[(size_t)i * c->dc_max + d];
                int pos = -1;
                for (int k = 0; k < c->vw[vn]; k++)
                    if (c->v_cn[(size_t)vn * c->dv_max + k] == i) {
                        pos = k;
                        break;
                    }
                ok = pos >= 0;
                c->c_pos[(size_t)i * c->dc_max + d] = pos;
            }
        if (!ok) break;
        rc = LDPC_OK;
    } while (0);
    fclose(f);
    c->n_const = 0;
    if (rc == LDPC_OK && constellation) {
        // "Point: i Real: x Imag(e): y" — the labels are skipped like the reference's %s (:326-334)
        FILE *g = fopen(constellation, "r");
        if (!g) {
            rc = LDPC_ERR_IO;
        } else {
            char t[100];
            int idx;
            float re, im;
            std::vector<float> cre, cim;
            while (fscanf(g, "%99s %d %99s %f %99s %f", t, &idx, t, &re, t, &im) == 6) {
                if (idx < 0 || idx >= 4096) {
                    rc = LDPC_ERR_FORMAT;
                    break;
                }
                if ((int)cre.size() <= idx) {
                    cre.resize(idx + 1, 0.f);
                    cim.resize(idx + 1, 0.f);
                }
                cre[idx] = re;
                cim[idx] = im;
            }
            fclose(g);
            c->n_const = (int)cre.size();
            c->cre = cre;
            c->cim = cim;
            if (c->n_const != 2 && c->n_const != c->q) rc = LDPC_ERR_FORMAT;  // BPSK or one point per symbol
        }
    }
    if (rc != LDPC_OK) {
        delete c;
        return rc;
    }
    *out = c;
    return LDPC_OK;
}

extern "C" void nb_ldpc_free_code(nb_ldpc_code_t *c)
{
    if (!c) return;
    if (c->last_use) cudaEventDestroy(c->last_use);
    if (c->d_mul) {
        cudaFree(c->d_mul);
        cudaFree(c->d_inv);
        cudaFree(c->d_vw);
        cudaFree(c->d_cw);
        cudaFree(c->d_v_cn);
        cudaFree(c->d_v_pos);
        cudaFree(c->d_c_vn);
        cudaFree(c->d_c_gf);
        cudaFree(c->d_c_pos);
        cudaFree(c->d_cre);
        cudaFree(c->d_cim);
    }
    if (c->scratch) cudaFree(c->scratch);
    delete c;
}

extern "C" int nb_ldpc_code_info(const nb_ldpc_code_t *c, nb_ldpc_code_info_t *info)
{
    if (!c || !info) return LDPC_ERR_ARG;
    info->N = c->N;
    info->M = c->M;
    info->q = c->q;
    info->p = c->p;
    info->dv_max = c->dv_max;
    info->dc_max = c->dc_max;
    info->n_const = c->n_const;
    return LDPC_OK;
}

extern "C" int nb_ldpc_code_tables(const nb_ldpc_code_t *c, uint16_t *mul, uint16_t *inv, int *check_vn,
                                   int *check_coef, int *check_weight)
{
    if (!c) return LDPC_ERR_ARG;
    if (mul) memcpy(mul, c->mul.data(), c->mul.size() * sizeof(uint16_t));
    if (inv) memcpy(inv, c->inv.data(), c->inv.size() * sizeof(uint16_t));
    if (check_vn) memcpy(check_vn, c->c_vn.data(), c->c_vn.size() * sizeof(int));
    if (check_coef) memcpy(check_coef, c->c_gf.data(), c->c_gf.size() * sizeof(int));
    if (check_weight) memcpy(check_weight, c->cw.data(), c->cw.size() * sizeof(int));
    return LDPC_OK;
}

extern "C" void nb_decode_opts_default(nb_decode_opts_t *o)
{
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->struct_size = (int)sizeof(*o);
    o->algo = NB_ALGO_EMS;          // NB/include/define.h:37 decoder_method 0
    o->in_kind = NB_IN_SYMBOL_LLR;
    o->mem_space = LDPC_MEM_HOST;
    o->ems_nm = 2;                  // NB/include/define.h:31-32
    o->ems_nc = 2;
    o->sigma = 1.0f;
}
