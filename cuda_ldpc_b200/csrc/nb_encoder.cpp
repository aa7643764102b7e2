// Systematic encoder over GF(q) for generating test vectors (host).  The reference has no encoder: its only
// non-zero codeword is the hard-coded CodeWord_sym_test[96] of the BDS GF(64) code (NB/include/codeword_test.h:1,
// used at NB/src/main.cu:190-212), every other matrix runs on the all-zero word — for a QAM constellation that is
// one corner point, which biases the FER.
//
// H (M x N over GF(q), from the loaded check lists) is brought to reduced row-echelon form by Gauss-Jordan elimination
// with the handle's multiply / inverse tables; pivots are searched from the LAST column backwards so that the
// information symbols are the first positions whenever the right-hand M x M part of H is invertible.  Non-pivot
// columns = information positions (K = N - rank), and every pivot symbol is a GF(q) combination of them:
//   c[piv_i] = sum_j P[i][j] * c[info_j]            (characteristic 2: minus = plus)
// The result is cached in the handle.
#include <string.h>

#include <mutex>

#include "common.h"
#include "nb_common.h"

namespace {
std::mutex g_nb_enc_mu;

void build(nb_ldpc_code *c)
{
    const int N = c->N, M = c->M, q = c->q;
    std::vector<uint16_t> A((size_t)M * N, 0);
    for (int m = 0; m < M; m++)
        for (int k = 0; k < c->cw[m]; k++) {
            const int v = c->c_vn[m * c->dc_max + k], h = c->c_gf[m * c->dc_max + k];
            A[(size_t)m * N + v] ^= (uint16_t)h;  // a repeated (check, variable) pair adds up in GF(2^p)
        }
    const uint16_t *mul = c->mul.data();
    std::vector<int> piv_col;  // pivot column of row i (rows 0..rank-1 after the elimination)
    std::vector<char> is_piv(N, 0);
    int row = 0;
    for (int col = N - 1; col >= 0 && row < M; col--) {
        int p = -1;
        for (int r = row; r < M; r++)
            if (A[(size_t)r * N + col]) {
                p = r;
                break;
            }
        if (p < 0) continue;
        if (p != row)
            for (int x = 0; x < N; x++) std::swap(A[(size_t)p * N + x], A[(size_t)row * N + x]);
        uint16_t *pr = &A[(size_t)row * N];
        const uint16_t *minv = mul + (size_t)c->inv[pr[col]] * q;  // scale the pivot row to 1
        for (int x = 0; x < N; x++) pr[x] = minv[pr[x]];
#pragma omp parallel for schedule(static)
        for (int r = 0; r < M; r++) {
            const uint16_t f = A[(size_t)r * N + col];
            if (r == row || !f) continue;
            uint16_t *ar = &A[(size_t)r * N];
            const uint16_t *mf = mul + (size_t)f * q;
            for (int x = 0; x < N; x++) ar[x] ^= mf[pr[x]];
        }
        piv_col.push_back(col);
        is_piv[col] = 1;
        row++;
    }
    const int rank = row;
    c->enc_info_pos.clear();
    for (int x = 0; x < N; x++)
        if (!is_piv[x]) c->enc_info_pos.push_back(x);
    const int K = (int)c->enc_info_pos.size();
    c->enc_piv_col = piv_col;
    c->enc_P.assign((size_t)rank * K, 0);
    for (int i = 0; i < rank; i++)
        for (int j = 0; j < K; j++) c->enc_P[(size_t)i * K + j] = A[(size_t)i * N + c->enc_info_pos[j]];
    c->enc_state = 1;
}
}  // namespace

extern "C" int nb_ldpc_encode_info(nb_ldpc_code_t *c, int *K, int *info_positions)
{
    if (!c) return LDPC_ERR_ARG;
    {
        std::lock_guard<std::mutex> lk(g_nb_enc_mu);
        if (c->enc_state == 0) build(c);
    }
    if (K) *K = (int)c->enc_info_pos.size();
    if (info_positions) memcpy(info_positions, c->enc_info_pos.data(), c->enc_info_pos.size() * sizeof(int));
    return LDPC_OK;
}

extern "C" int nb_ldpc_encode(nb_ldpc_code_t *c, const uint16_t *info_syms, uint16_t *codeword_syms)
{
    if (!c || !info_syms || !codeword_syms) return LDPC_ERR_ARG;
    int rc = nb_ldpc_encode_info(c, nullptr, nullptr);
    if (rc != LDPC_OK) return rc;
    const int K = (int)c->enc_info_pos.size(), rank = (int)c->enc_piv_col.size(), q = c->q;
    for (int j = 0; j < K; j++)
        if (info_syms[j] >= q) return LDPC_ERR_ARG;
    for (int j = 0; j < K; j++) codeword_syms[c->enc_info_pos[j]] = info_syms[j];
    const uint16_t *mul = c->mul.data();
    for (int i = 0; i < rank; i++) {
        unsigned acc = 0;
        const uint16_t *Pi = &c->enc_P[(size_t)i * K];
        for (int j = 0; j < K; j++) acc ^= mul[(size_t)Pi[j] * q + info_syms[j]];
        codeword_syms[c->enc_piv_col[i]] = (uint16_t)acc;
    }
    return LDPC_OK;
}
