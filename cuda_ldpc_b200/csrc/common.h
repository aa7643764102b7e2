// Internal definitions shared by the translation units of libldpc_b200.so.
// Reference paths: B/ = bldpc_实习/, NB/ = myNBLDPC/ (gsw4869/CUDA_LDPC).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/ldpc_b200.h"

namespace ldpcb {

void set_cuda_error(cudaError_t e, const char *where);
// returns LDPC_ERR_CUDA after recording the message
#define LDPC_CUDA_TRY(expr)                                   \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) {                              \
            ::ldpcb::set_cuda_error(_e, #expr);               \
            return LDPC_ERR_CUDA;                             \
        }                                                     \
    } while (0)

constexpr int kMaxLayers = 64;   // J <= 64 (largest shipped file: J48_L60_Z160)
constexpr int kMaxBlocks = 512;  // non-empty circulants (largest shipped: PON 275)
constexpr int kMaxDc = 32;
constexpr int kMaxDv = 16;       // B/LDPC_Decoder.cu:175 sizes R[15]

// Per-layer tables handed to the layered kernels by value (constant bank): layer r owns
// entries [off[r], off[r]+dc[r]) of col/shift, ascending column block (the reference's edge
// order inside a check, B/Simulation.cu:373-377 `position`).
struct LayerTables {
    unsigned short off[kMaxLayers];
    unsigned char dc[kMaxLayers];
    unsigned char col[kMaxBlocks];
    unsigned short shift[kMaxBlocks];
};

// Column-major twin for the flooding VN pass: column block c owns entries
// [voff[c], voff[c]+dv[c]) of (row block, position in the check, shift), ascending row block
// (B/Simulation.cu:367-384, index2).
struct ColumnTables {
    unsigned short voff[128];
    unsigned char dv[128];
    unsigned char row[kMaxBlocks];
    unsigned char pos[kMaxBlocks];
    unsigned short shift[kMaxBlocks];
};

}  // namespace ldpcb

struct ldpc_code {
    int J, L, Z, N, K, M, E;
    int dc_max, dv_max, dc_min, dv_min;
    std::vector<int> H, Wc, Wv;
    ldpcb::LayerTables lt;
    ldpcb::ColumnTables ct;
    int device;
    int num_sms;
    // One scratch arena per handle: calls on one handle SERIALISE.  `call_mu` is held for the whole of every decode
    // call (host threads), and `last_use` (recorded on the call's stream when it returns) is waited for by the next
    // call's stream, so work enqueued on different streams never shares the arena in time.  Concurrency = one handle
    // per stream / host thread (a handle is a few KB of tables plus its arena).
    std::mutex call_mu;
    cudaEvent_t last_use;
    bool last_use_valid;
    // scratch arena (device), grown on demand
    void *scratch;
    size_t scratch_bytes;
    // two internal streams for the chunked host path (H2D of chunk k+1 under the decode of chunk k)
    cudaStream_t pipe_stream[2];
    // pinned int8 staging slots + "slot free again" events of the host-pack path (ldpc_decode_opts_t::host_pack_threads)
    void *pack_host[2];
    size_t pack_host_bytes;
    cudaEvent_t pack_ev[2];
    size_t last_h2d_bytes;  // channel-value bytes the last host-buffer call uploaded (ldpc_last_h2d_bytes)
    // encoder cache (host): parity-part inverse, built lazily
    std::vector<uint32_t> enc_cache;
    int enc_state;  // 0 = not built, 1 = ok, -1 = singular
    std::vector<int> enc_pivot_cols;
};

namespace ldpcb {
// grows code->scratch to at least `bytes` (not thread safe across concurrent first calls with
// different sizes; decode calls serialise on it)
int ensure_scratch(const ldpc_code *code, size_t bytes, void **out);

int launch_flooding_fp32(const ldpc_code *code, const float *y_nf, int F, int iters, int exit_mode,
                         unsigned char *hard_nf, int *iters_dev, int *ok_dev, float *msgs, int *flag_scratch,
                         cudaStream_t st, int *launches, bool may_block);

struct LayeredArgs {
    const void *llr;
    int llr_dtype, layout, F, iters, exit_mode, out_format;
    float scale;
    int msg_max, beta_num, beta_shift;
    float alpha;
    void *out;
    int *iters_out, *ok_out;
    void *dbg_app, *dbg_rec;
    void *scratch;         // device scratch for the check records
    size_t scratch_bytes;
    // fused channel (llr_dtype == LDPC_DTYPE_CHANNEL)
    float ch_sigma;
    unsigned long long ch_seed, ch_first;
    const unsigned char *ch_cw;
};
// frames one full wave of resident CTAs decodes (grid x codewords per group)
int layered_i8_wave_frames(const ldpc_code *code, int *frames);
int layered_f16_wave_frames(const ldpc_code *code, int *frames);
// bytes of record scratch the layered kernels want for a batch of F frames
int layered_i8_scratch_bytes(const ldpc_code *code, int F, int beta_num, size_t *bytes);
int launch_layered_i8(const ldpc_code *code, const LayeredArgs &a, cudaStream_t st, int *launches);
// fp16 message mode (bldpc_layered_f16.cu): same arguments; dbg_app / dbg_rec are binary16 [N][F] / [M][dc_max][F]
int layered_f16_scratch_bytes(const ldpc_code *code, int F, size_t *bytes);
int launch_layered_f16(const ldpc_code *code, const LayeredArgs &a, cudaStream_t st, int *launches);
int launch_layered_f32_nf(const ldpc_code *code, const float *y_nf, int F, int iters, int exit_mode, float alpha,
                          float *app, float *msgs, unsigned char *hard_nf, int *iters_dev, int *ok_dev,
                          int *flag_scratch, cudaStream_t st, int *launches, bool may_block);
}  // namespace ldpcb
