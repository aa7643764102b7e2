// Host-side loader for the J*_L*_Z*_BlockH.txt / PON_LDPC.txt quasi-cyclic shift matrices.
// Replaces Get_H and Transform_H (B/Simulation.cu:292-387) and the table upload in
// B/main.cu:88-104.  No CUDA kernels here; the tables travel to the kernels by value.
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.h"

namespace ldpcb {

static thread_local std::string g_last_cuda_error;

void set_cuda_error(cudaError_t e, const char *where)
{
    g_last_cuda_error = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " at " + where;
    (void)cudaGetLastError();  // clear the sticky-less error state
}

static std::mutex g_scratch_mu;

int ensure_scratch(const ldpc_code *code_c, size_t bytes, void **out)
{
    ldpc_code *code = const_cast<ldpc_code *>(code_c);
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    if (code->scratch_bytes < bytes) {
        if (code->scratch) {
            LDPC_CUDA_TRY(cudaDeviceSynchronize());
            LDPC_CUDA_TRY(cudaFree(code->scratch));
            code->scratch = nullptr;
            code->scratch_bytes = 0;
        }
        size_t want = bytes + (bytes >> 3);
        LDPC_CUDA_TRY(cudaMalloc(&code->scratch, want));
        code->scratch_bytes = want;
    }
    *out = code->scratch;
    return LDPC_OK;
}

// "…/J4_L24_Z96_BlockH.txt" -> 4, 24, 96
static bool geometry_from_name(const char *path, int *J, int *L, int *Z)
{
    const char *base = strrchr(path, '/');
    base = base ? base + 1 : path;
    int j, l, z;
    if (sscanf(base, "J%d_L%d_Z%d", &j, &l, &z) == 3 && j > 0 && l > 0 && z > 0) {
        *J = j;
        *L = l;
        *Z = z;
        return true;
    }
    return false;
}

}  // namespace ldpcb

using namespace ldpcb;

extern "C" const char *ldpc_last_cuda_error(void) { return g_last_cuda_error.c_str(); }

extern "C" const char *ldpc_version(void) { return "ldpc_b200 0.1 (sm_100a)"; }

extern "C" const char *ldpc_strerror(int code)
{
    switch (code) {
        case LDPC_OK: return "ok";
        case LDPC_ERR_IO: return "cannot open file";
        case LDPC_ERR_FORMAT: return "file does not parse or geometry is inconsistent";
        case LDPC_ERR_ARG: return "bad argument";
        case LDPC_ERR_CUDA: return "CUDA error (see ldpc_last_cuda_error)";
        case LDPC_ERR_NOMEM: return "out of memory";
        case LDPC_ERR_UNSUPPORTED: return "unsupported combination";
        case LDPC_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
        default: return code > 0 ? "ok (positive value)" : "unknown error";
    }
}

extern "C" int ldpc_load_code(const char *path, int J, int L, int Z, ldpc_code_t **out)
{
    if (!path || !out) return LDPC_ERR_ARG;
    *out = nullptr;
    if (J <= 0 || L <= 0 || Z <= 0) {
        if (!geometry_from_name(path, &J, &L, &Z)) return LDPC_ERR_ARG;
    }
    if (J > kMaxLayers || L > 128 || J >= L + 1 || Z > 65535) return LDPC_ERR_UNSUPPORTED;
    FILE *fp = fopen(path, "r");
    if (!fp) return LDPC_ERR_IO;
    std::vector<int> H((size_t)J * L);
    // whitespace/CRLF-insensitive integer scan, exactly J*L entries (B/Simulation.cu:315-319);
    // trailing non-space garbage or extra integers are a format error (the reference ignores them)
    for (size_t i = 0; i < H.size(); i++) {
        int v;
        if (fscanf(fp, "%d", &v) != 1) {
            fclose(fp);
            return LDPC_ERR_FORMAT;
        }
        if (v < -1 || v >= Z) {
            fclose(fp);
            return LDPC_ERR_FORMAT;
        }
        H[i] = v;
    }
    int extra;
    if (fscanf(fp, "%d", &extra) == 1) {
        fclose(fp);
        return LDPC_ERR_FORMAT;  // more than J*L integers: wrong geometry
    }
    fclose(fp);

    ldpc_code *c = new (std::nothrow) ldpc_code();
    if (!c) return LDPC_ERR_NOMEM;
    c->J = J;
    c->L = L;
    c->Z = Z;
    c->N = L * Z;
    c->M = J * Z;
    c->K = (L - J) * Z;
    c->H = H;
    c->Wc.assign(J + 1, 0);
    c->Wv.assign(L + 1, 0);
    c->scratch = nullptr;
    c->scratch_bytes = 0;
    c->pipe_stream[0] = c->pipe_stream[1] = nullptr;
    c->pack_host[0] = c->pack_host[1] = nullptr;
    c->pack_host_bytes = 0;
    c->pack_ev[0] = c->pack_ev[1] = nullptr;
    c->last_h2d_bytes = 0;
    c->last_use = nullptr;
    c->last_use_valid = false;
    c->enc_state = 0;
    memset(&c->lt, 0, sizeof(c->lt));
    memset(&c->ct, 0, sizeof(c->ct));
    int nb = 0;
    for (int r = 0; r < J; r++) {
        c->lt.off[r] = (unsigned short)nb;
        for (int col = 0; col < L; col++) {
            int s = H[(size_t)r * L + col];
            if (s < 0) continue;
            if (nb >= kMaxBlocks) {
                delete c;
                return LDPC_ERR_UNSUPPORTED;
            }
            c->lt.col[nb] = (unsigned char)col;
            c->lt.shift[nb] = (unsigned short)s;
            nb++;
            c->Wc[r]++;
            c->Wv[col]++;
        }
        c->lt.dc[r] = (unsigned char)c->Wc[r];
        if (c->Wc[r] > c->Wc[J]) c->Wc[J] = c->Wc[r];
    }
    c->E = nb * Z;
    int nv = 0;
    c->dv_min = 1 << 30;
    c->dc_min = 1 << 30;
    for (int r = 0; r < J; r++)
        if (c->Wc[r] < c->dc_min) c->dc_min = c->Wc[r];
    for (int col = 0; col < L; col++) {
        c->ct.voff[col] = (unsigned short)nv;
        c->ct.dv[col] = (unsigned char)c->Wv[col];
        for (int r = 0; r < J; r++) {
            int s = H[(size_t)r * L + col];
            if (s < 0) continue;
            int pos = 0;
            for (int c2 = 0; c2 < col; c2++)
                if (H[(size_t)r * L + c2] >= 0) pos++;
            c->ct.row[nv] = (unsigned char)r;
            c->ct.pos[nv] = (unsigned char)pos;
            c->ct.shift[nv] = (unsigned short)s;
            nv++;
        }
        if (c->Wv[col] > c->Wv[L]) c->Wv[L] = c->Wv[col];
        if (c->Wv[col] < c->dv_min) c->dv_min = c->Wv[col];
    }
    c->dc_max = c->Wc[J];
    c->dv_max = c->Wv[L];
    if (c->dc_max > kMaxDc || c->dv_max > kMaxDv || c->dc_max < 1) {
        delete c;
        return LDPC_ERR_UNSUPPORTED;
    }
    // device properties are looked up lazily by the first decode call: loading (and
    // ldpc_code_tables) works on a machine without a GPU, decoding does not.
    c->device = -1;
    c->num_sms = 0;
    *out = c;
    return LDPC_OK;
}

extern "C" void ldpc_free_code(ldpc_code_t *code)
{
    if (!code) return;
    if (code->last_use) cudaEventDestroy(code->last_use);
    if (code->scratch) cudaFree(code->scratch);
    for (int i = 0; i < 2; i++) {
        if (code->pipe_stream[i]) cudaStreamDestroy(code->pipe_stream[i]);
        if (code->pack_host[i]) cudaFreeHost(code->pack_host[i]);
        if (code->pack_ev[i]) cudaEventDestroy(code->pack_ev[i]);
    }
    delete code;
}

extern "C" int ldpc_code_info(const ldpc_code_t *c, ldpc_code_info_t *info)
{
    if (!c || !info) return LDPC_ERR_ARG;
    info->J = c->J;
    info->L = c->L;
    info->Z = c->Z;
    info->N = c->N;
    info->K = c->K;
    info->M = c->M;
    info->E = c->E;
    info->dc_max = c->dc_max;
    info->dv_max = c->dv_max;
    info->dc_min = c->dc_min;
    info->dv_min = c->dv_min;
    return LDPC_OK;
}

extern "C" int ldpc_code_tables(const ldpc_code_t *c, int *H, int *Wc, int *Wv, int *addr)
{
    if (!c) return LDPC_ERR_ARG;
    if (H) memcpy(H, c->H.data(), c->H.size() * sizeof(int));
    if (Wc) memcpy(Wc, c->Wc.data(), c->Wc.size() * sizeof(int));
    if (Wv) memcpy(Wv, c->Wv.data(), c->Wv.size() * sizeof(int));
    if (addr) {
        // Address_Variablenode as a fixed Transform_H would fill it: variable n = c*Z + j, k-th
        // connected row block r -> slot (r*Z + (j - s) mod Z) * Wc_max + position
        const int Wvm = c->Wv[c->L], Wcm = c->Wc[c->J];
        for (size_t i = 0; i < (size_t)c->N * Wvm; i++) addr[i] = -1;
        for (int col = 0; col < c->L; col++)
            for (int k = 0; k < c->ct.dv[col]; k++) {
                int e = c->ct.voff[col] + k;
                int r = c->ct.row[e], pos = c->ct.pos[e], s = c->ct.shift[e];
                for (int j = 0; j < c->Z; j++) {
                    int row = (j - s) % c->Z;
                    if (row < 0) row += c->Z;
                    addr[(size_t)(col * c->Z + j) * Wvm + k] = (r * c->Z + row) * Wcm + pos;
                }
            }
    }
    return LDPC_OK;
}

extern "C" float ldpc_sigma(int snrtype, float snr_db, float rate)
{
    // B/main.cu:120-127
    if (snrtype == 0) return (float)sqrt(0.5 / (rate * (pow(10.0, (snr_db / 10.0)))));
    return (float)sqrt(0.5 / (pow(10.0, (snr_db / 10.0))));
}
