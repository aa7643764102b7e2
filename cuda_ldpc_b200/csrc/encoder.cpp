// Systematic encoder for generating test vectors (host).  The reference has no encoder at all:
// its binary simulator only ever sends the all-zero codeword (B/Simulation.cu:96-109, the
// PN_Message == 1 branch is empty).  Information bits are the first (L-J)*Z positions
// (B/LDPC_Decoder.cu:36, msgLen); H = [A | B] with B the M x M parity part.  B^-1 is computed once
// per handle by Gauss-Jordan elimination over GF(2) on 64-bit words and cached:
//   parity = B^-1 * (A * info).
#include <string.h>

#include <mutex>

#include "common.h"

namespace {

std::mutex g_enc_mu;

// returns false if B is singular
bool invert_parity_part(const ldpc_code *c, std::vector<uint64_t> &inv)
{
    const int M = c->M, K = c->K, Z = c->Z, W = (M + 63) / 64;
    std::vector<uint64_t> B((size_t)M * W, 0);
    inv.assign((size_t)M * W, 0);
    for (int r = 0; r < c->J; r++)
        for (int k = 0; k < c->lt.dc[r]; k++) {
            const int e = c->lt.off[r] + k, col = c->lt.col[e], s = c->lt.shift[e];
            if (col * Z < K) continue;
            for (int i = 0; i < Z; i++) {
                const int m = r * Z + i, pc = col * Z + (i + s) % Z - K;  // check row i of the block hits column (i+s) mod Z
                B[(size_t)m * W + (pc >> 6)] ^= 1ull << (pc & 63);
            }
        }
    for (int m = 0; m < M; m++) inv[(size_t)m * W + (m >> 6)] = 1ull << (m & 63);
    for (int col = 0; col < M; col++) {
        const int w = col >> 6;
        const uint64_t bit = 1ull << (col & 63);
        int piv = -1;
        for (int r = col; r < M; r++)
            if (B[(size_t)r * W + w] & bit) {
                piv = r;
                break;
            }
        if (piv < 0) return false;
        if (piv != col)
            for (int x = 0; x < W; x++) {
                std::swap(B[(size_t)piv * W + x], B[(size_t)col * W + x]);
                std::swap(inv[(size_t)piv * W + x], inv[(size_t)col * W + x]);
            }
        const uint64_t *bp = &B[(size_t)col * W], *ip = &inv[(size_t)col * W];
#pragma omp parallel for schedule(static)
        for (int r = 0; r < M; r++) {
            if (r == col || !(B[(size_t)r * W + w] & bit)) continue;
            uint64_t *br = &B[(size_t)r * W], *ir = &inv[(size_t)r * W];
            for (int x = w; x < W; x++) br[x] ^= bp[x];  // columns left of the pivot are already clear
            for (int x = 0; x < W; x++) ir[x] ^= ip[x];
        }
    }
    return true;
}

}  // namespace

extern "C" int ldpc_encode(ldpc_code_t *c, const uint8_t *info, uint8_t *cw)
{
    if (!c || !info || !cw) return LDPC_ERR_ARG;
    const int M = c->M, K = c->K, Z = c->Z, W = (M + 63) / 64;
    {
        std::lock_guard<std::mutex> lk(g_enc_mu);
        if (c->enc_state == 0) {
            std::vector<uint64_t> inv;
            const bool ok = invert_parity_part(c, inv);
            c->enc_cache.clear();
            if (ok) {
                c->enc_cache.resize(inv.size() * 2);
                memcpy(c->enc_cache.data(), inv.data(), inv.size() * sizeof(uint64_t));
            }
            c->enc_state = ok ? 1 : -1;
        }
    }
    if (c->enc_state < 0) return LDPC_ERR_UNSUPPORTED;  // parity part of H is singular
    const uint64_t *inv = reinterpret_cast<const uint64_t *>(c->enc_cache.data());
    // s = A * info
    std::vector<uint64_t> s(W, 0);
    for (int r = 0; r < c->J; r++)
        for (int k = 0; k < c->lt.dc[r]; k++) {
            const int e = c->lt.off[r] + k, col = c->lt.col[e], sh = c->lt.shift[e];
            if (col * Z >= K) continue;
            for (int i = 0; i < Z; i++)
                if (info[col * Z + (i + sh) % Z] & 1) s[(r * Z + i) >> 6] ^= 1ull << ((r * Z + i) & 63);
        }
    memcpy(cw, info, K);
    for (int m = 0; m < M; m++) {
        uint64_t acc = 0;
        const uint64_t *row = inv + (size_t)m * W;
        for (int x = 0; x < W; x++) acc ^= row[x] & s[x];
        cw[K + m] = (uint8_t)(__builtin_popcountll(acc) & 1);
    }
    return LDPC_OK;
}
