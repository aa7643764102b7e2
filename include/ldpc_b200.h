/*
 * ldpc_b200.h — C-ABI of libldpc_b200.so, the B200-native (sm_100a) batched LDPC decode
 * engine that drops in for the decode path of gsw4869/CUDA_LDPC.
 *
 * Plain C: opaque handles, plain pointers and sizes, integer return codes (0 = ok,
 * negative = error; the library never prints and never exits — the reference printf()s and
 * exit(0)s on every failure, e.g. bldpc_实习/LDPC_Decoder.cu:39-44).
 *
 * Reference paths: B/ = bldpc_实习/, NB/ = myNBLDPC/.
 */
#ifndef LDPC_B200_H
#define LDPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ errors */
#define LDPC_OK 0
#define LDPC_ERR_IO (-1)          /* file cannot be opened                                  */
#define LDPC_ERR_FORMAT (-2)      /* file does not parse / inconsistent geometry            */
#define LDPC_ERR_ARG (-3)         /* bad argument                                           */
#define LDPC_ERR_CUDA (-4)        /* CUDA runtime error; see ldpc_last_cuda_error()         */
#define LDPC_ERR_NOMEM (-5)
#define LDPC_ERR_UNSUPPORTED (-6) /* combination not implemented (e.g. dc > 32)             */
#define LDPC_ERR_NO_DEVICE (-7)   /* no sm_100 device: there is NO CPU fallback             */

const char *ldpc_strerror(int code);
/* last CUDA error string seen by the calling thread ("" if none) */
const char *ldpc_last_cuda_error(void);
/* library version string, e.g. "ldpc_b200 0.1 (sm_100a)" */
const char *ldpc_version(void);

/* ------------------------------------------------------------------ binary QC-LDPC */

typedef struct ldpc_code ldpc_code_t;

/* Replaces Get_H + Transform_H (B/Simulation.cu:292-387) and the table upload in
 * B/main.cu:88-104.  Reads a J x L block matrix of circulant shifts (-1 = zero block) in the
 * reference's text format (whitespace/CRLF tolerant).  J, L, Z <= 0: parsed from
 * "J<j>_L<l>_Z<z>" in the file name; PON_LDPC.txt needs 12/69/256 explicitly.
 * The graph is the true circulant row = (col - s) mod Z (the reference's Transform_H mis-wires
 * it).  The handle owns its tables and ONE scratch arena, bound to the device that is current at its
 * first decode call (later calls switch to that device and back).  Calls on one handle are safe from any host
 * thread and any stream but SERIALISE (a per-handle lock for the call, an event chain between the streams); for
 * concurrent decoding load one handle per stream or host thread — a handle is a few KB of tables plus its arena.  */
int ldpc_load_code(const char *blockh_path, int J, int L, int Z, ldpc_code_t **out);
void ldpc_free_code(ldpc_code_t *code);

typedef struct {
    int J, L, Z;
    int N, K, M;      /* code bits, info bits (first (L-J)*Z positions), checks            */
    int E;            /* edges                                                              */
    int dc_max, dv_max, dc_min, dv_min;
} ldpc_code_info_t;
int ldpc_code_info(const ldpc_code_t *code, ldpc_code_info_t *info);

/* Host copies of the tables the reference builds (for parity tests and for callers that keep
 * the reference's own kernels): H[J*L], Wc[J+1], Wv[L+1] as Get_H fills them, and
 * Address_Variablenode[N*Wv[L]] as a FIXED Transform_H would.  Any pointer may be NULL.    */
int ldpc_code_tables(const ldpc_code_t *code, int *H, int *Wc, int *Wv, int *address_variablenode);

/* enums for ldpc_decode_opts_t */
#define LDPC_LAYOUT_NF 0 /* reference layout: value[n*F + f], frame index fastest (B/LDPC_Encoder.cu:38) */
#define LDPC_LAYOUT_FN 1 /* value[f*N + n]                                                              */

#define LDPC_DTYPE_FP32 0
#define LDPC_DTYPE_FP16 1
#define LDPC_DTYPE_INT8 2
#define LDPC_DTYPE_CHANNEL 3 /* no input buffer: the channel values are generated inside the layered int8 kernel
                                (same Philox stream as ldpc_awgn_bpsk, fused into the load of the first iteration);
                                `llr` is ignored, see ldpc_decode_opts_t.channel_*                              */

#define LDPC_MEM_HOST 0
#define LDPC_MEM_DEVICE 1

#define LDPC_SCHED_FLOODING 0 /* reference schedule: VN pass then CN pass per iteration */
#define LDPC_SCHED_LAYERED 1  /* row-block layered (throughput mode)                    */

#define LDPC_EXIT_NONE 0     /* exactly `iters` iterations                                          */
#define LDPC_EXIT_GENIE 1    /* reference rule: stop when ALL frames are all-zero on the info bits  */
#define LDPC_EXIT_SYNDROME 2 /* per frame: latch at the first iteration with H x = 0                */

#define LDPC_OUT_INT32_REF 0 /* reference: int32 D[(N+1)*F], row N = per-frame ok flag (B/LDPC_Decoder.cu:45,134-149) */
#define LDPC_OUT_U8 1        /* uint8 bits[N*F] in the input layout                                 */
#define LDPC_OUT_BITPACK 2   /* uint32 words[F][ceil(N/32)], bit n%32 of word n/32                  */

typedef struct {
    int struct_size;   /* = sizeof(ldpc_decode_opts_t), for ABI growth                              */
    int batch;         /* F, number of codewords in this call                                       */
    int layout;        /* LDPC_LAYOUT_*  (input LLRs and U8 output)                                 */
    int llr_dtype;     /* LDPC_DTYPE_*   (input)                                                    */
    int mem_space;     /* LDPC_MEM_*     (llr, hard_bits, iters_out, ok_out all live there)         */
    int schedule;      /* LDPC_SCHED_*                                                              */
    int msg_dtype;     /* arithmetic of the messages: LDPC_DTYPE_FP32 (flooding, layered), LDPC_DTYPE_INT8 or
                          LDPC_DTYPE_FP16 (layered: the packed throughput modes, 4 / 2 codewords per word)  */
    int early_exit;    /* LDPC_EXIT_*                                                               */
    int out_format;    /* LDPC_OUT_*                                                                */
    float alpha;       /* fp32 layered: multiplies min1/min2 (1.0 = the reference's un-normalised rule) */
    float llr_scale;   /* int8: q = sat127(rint(y*llr_scale)) when llr_dtype is fp32/fp16; fp16 messages: APP =
                          f16(clamp(y*llr_scale, +-127)), no integer rounding                             */
    int msg_max;       /* int8 / fp16: message magnitude clip (1..127)                              */
    int beta_num;      /* int8: m' = m - ((m*beta_num) >> beta_shift); 0 = the reference's un-normalised   */
    int beta_shift;    /*       rule, (1,2) = x0.75, (1,3) = x0.875, (3,4) = x0.8125; beta_num <= 8.
                          fp16 messages: m' = m * (1 - beta_num / 2^beta_shift), one fp16 product         */
    int *iters_out;    /* [F] iterations executed per frame, or NULL                                */
    int *ok_out;       /* [F] 1 = syndrome satisfied (genie mode: all-zero info bits), or NULL      */
    void *stream;      /* cudaStream_t, NULL = default stream                                       */
    void *debug_app;   /* optional device/host buffer receiving the final APP [N][F] (layered: fp32 / int8 /
                          binary16 by msg_dtype) — tests                                                  */
    void *debug_msgs;  /* optional buffer receiving the final check-to-variable messages [M][dc_max][F] (fp32 modes:
                          float, the reference's Memory_RQ slots; layered int8: int8, layered fp16: binary16; 0 for
                          absent edges) — tests */
    /* fused channel (llr_dtype == LDPC_DTYPE_CHANNEL): y = 1 - 2c + sigma*N(0,1), Philox keyed by
     * (channel_seed, channel_first_frame + f, bit/4) exactly like ldpc_awgn_bpsk                    */
    float channel_sigma;
    uint64_t channel_seed;
    uint64_t channel_first_frame;
    const uint8_t *channel_codeword; /* device uint8 [N] broadcast to all frames, NULL = all-zero      */
    /* Host path, layered int8 (ignored by the fp16 message mode), fp32 [N][F] input: the library can quantise each chunk to int8 on host threads (the
     * kernel's own rule q = sat127(rint(y*llr_scale)), bit-identical results) into pinned staging buffers and copy a
     * quarter of the bytes over PCIe; the quantisation of chunk k+1 runs under the copy and decode of chunk k.
     * > 0: that many threads.  0 (default): automatic — 12 threads when the calling process' affinity mask holds at
     * least 12 CPUs, else a plain fp32 copy.  < 0: never (a busy host, or several ranks sharing
     * few cores).                                                                                            */
    int host_pack_threads;
} ldpc_decode_opts_t;

void ldpc_decode_opts_default(ldpc_decode_opts_t *opts);

/* Replaces LDPC_Decoder_GPU (B/LDPC_Decoder.cuh:5, called from B/Simulation.cu:143).
 * llr: channel values (the reference feeds the raw sample y = 1-2c+sigma*n, B/LDPC_Encoder.cu:38).
 * With mem_space = HOST the call copies in/out and synchronises; with DEVICE it only enqueues work on
 * opts->stream and never waits for the device — early-exit modes terminate on the device (layered int8: inside the
 * kernel; flooding / layered fp32: the remaining iterations are enqueued and return at their first instruction once
 * the stop condition holds), except when the scratch arena has to grow (first call, or a larger batch).
 * Returns the number of kernel launches enqueued (>= 0) or an error.                          */
int ldpc_decode_batch(const ldpc_code_t *code, const void *llr, void *hard_bits, int iters,
                      const ldpc_decode_opts_t *opts);

/* bytes of hard_bits for a given format */
size_t ldpc_out_bytes(const ldpc_code_t *code, int batch, int out_format);

/* Frames one full wave of the layered kernel decodes on this device (resident CTAs x codewords per CTA group;
 * msg_dtype = LDPC_DTYPE_INT8 or LDPC_DTYPE_FP16), or a negative error.  In fixed-iteration mode every group costs the
 * same and groups are dealt to the CTAs round-robin, so a batch that is a multiple of this figure gives every CTA the
 * same number of groups.  Measured on B200 the effect is small (PON +0.5 %, J4_L24_Z96 none: the CTAs sharing an SM
 * absorb an uneven last round); it matters for batches of only a few waves.  No reference counterpart: the
 * reference's batch is the compile-time `message_length` (B/define.cuh:25).                                           */
int ldpc_wave_frames(const ldpc_code_t *code, int msg_dtype);

/* Host-to-device bytes of channel values the last HOST-buffer ldpc_decode_batch call on this handle copied (the chunked
 * feed mixes int8 chunks quantised on the host with fp32 chunks in a share that adapts to the host; benchmarks report
 * this figure).  No reference counterpart: the reference copies N*F*4 bytes per batch (B/Simulation.cu:138).          */
size_t ldpc_last_h2d_bytes(const ldpc_code_t *code);

/* ------------------------------------------------------------------ channel + simulation loop */

/* On-device AWGN for BPSK: y[n,f] = 1 - 2*c[n,f] + sigma * N(0,1), Philox4x32-10 keyed by
 * (seed, global frame index, n) so that results do not depend on how frames are sharded
 * (replaces AWGNChannel_CPU, B/LDPC_Encoder.cu:25-41).  codeword_bits may be NULL (all-zero,
 * the reference's only mode, B/Simulation.cu:96-109); otherwise uint8 bits [N] broadcast to all
 * frames.  Output is a DEVICE buffer of fp32 in `layout`.                                   */
int ldpc_awgn_bpsk(const ldpc_code_t *code, float *y_dev, int batch, int layout, float sigma,
                   uint64_t seed, uint64_t first_frame, const uint8_t *codeword_bits_dev, void *stream);

/* B/struct.cuh:16-33 counters */
typedef struct {
    int64_t num_Frames, num_Error_Frames, num_Error_Bits, Total_Iteration, num_False_Frames,
        num_Alarm_Frames;
} ldpc_sim_counters_t;

/* Replaces one call of Statistic (B/Simulation.cu:245-285) on device-resident decoder output:
 * accumulates the six counters into `counters_dev` (device, ldpc_sim_counters_t layout).
 * D_dev is in `out_format` (LDPC_OUT_INT32_REF or LDPC_OUT_U8, layout NF, or LDPC_OUT_BITPACK);
 * ok_dev [F] = per-frame flag (NULL with INT32_REF: the flag row N is used).
 * codeword_bits_dev NULL = all-zero codeword.  length = leading bits compared (msgLen).      */
int ldpc_statistic(const ldpc_code_t *code, const void *D_dev, int out_format, const int *ok_dev,
                   const int *iters_dev, int batch, int length, const uint8_t *codeword_bits_dev,
                   int64_t *counters_dev, void *stream);

/* The counter-based generator behind ldpc_awgn_bpsk (Philox4x32-10, Salmon et al. SC'11), exposed
 * for known-answer tests: out[4] = philox(counter[4], key[2]).                                */
void ldpc_philox4x32(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);

/* B/main.cu:120-127 */
float ldpc_sigma(int snrtype, float snr_db, float rate);

/* Systematic encoder for test vectors (the reference has none): info uint8 [K] ->
 * codeword uint8 [N] on the host, by GF(2) elimination cached in the handle.  Returns
 * LDPC_ERR_UNSUPPORTED if the parity part of H is singular.                                 */
int ldpc_encode(ldpc_code_t *code, const uint8_t *info_bits, uint8_t *codeword_bits);

/* ------------------------------------------------------------------ non-binary GF(q) LDPC */

typedef struct nb_ldpc_code nb_ldpc_code_t;

/* Replaces NB Get_H (NB/src/Simulation.cpp:347-466), GFInitial (NB/src/GF.cpp:68-117),
 * Get_CONSTELLATION (NB/src/Simulation.cpp:313-338) and the link-table flattening in
 * NB/src/main.cu:101-188.  gf_table may be NULL: tables are then generated from the primitive
 * polynomial (7,11,19,37,67,137,285,529 for q = 4..512).  constellation may be NULL (BPSK).
 * coef_is_exponent != 0: edge coefficients are exponents of alpha (the *_exp.txt files).   */
int nb_ldpc_load_code(const char *matrix, const char *gf_table, const char *constellation,
                      int coef_is_exponent, nb_ldpc_code_t **out);
void nb_ldpc_free_code(nb_ldpc_code_t *code);

typedef struct {
    int N, M, q, p;       /* symbols, checks, field size, bits per symbol */
    int dv_max, dc_max;
    int n_const;          /* constellation points (0 = none loaded)       */
} nb_ldpc_code_info_t;
int nb_ldpc_code_info(const nb_ldpc_code_t *code, nb_ldpc_code_info_t *info);
/* host copies: mul[q*q], inv[q], and the check lists vn[M*dc_max], coef[M*dc_max] (-1 padded) */
int nb_ldpc_code_tables(const nb_ldpc_code_t *code, uint16_t *mul, uint16_t *inv, int *check_vn,
                        int *check_coef, int *check_weight);

#define NB_ALGO_EMS 0         /* NB/src/LDPC_Decoder.cpp:172-359  (scale 1/1.2)            */
#define NB_ALGO_TMM 1         /* NB/src/LDPC_Decoder.cpp:361-542  (scale 0.8)              */
#define NB_ALGO_LAYERED_TMM 3 /* NB/src/LDPC_Decoder.cpp:544-702                           */
#define NB_ALGO_FFT_BP 4      /* NOT in the reference ("parity unpinned"): probability-domain BP,
                                 check node = product in the Walsh-Hadamard domain; specified by the oracle */

#define NB_IN_SYMBOL_LLR 0 /* fp32 [F][N][q-1]: L_ch as Demodulate produces it (log P(a)/P(0)) */
#define NB_IN_BPSK 1       /* fp32 [F][N*p] real BPSK samples + sigma: fused Demodulate BPSK branch (NB/src/LDPC_Decoder.cpp:137-158) */
#define NB_IN_QAM 2        /* fp32 [F][N][2] complex samples + sigma: fused QAM branch (:160-169) */

typedef struct {
    int struct_size;
    int batch;       /* F */
    int algo;        /* NB_ALGO_* */
    int in_kind;     /* NB_IN_*   */
    int mem_space;   /* LDPC_MEM_* */
    int ems_nm;      /* EMS_NM (NB/include/define.h:31): 1..4 sorted entries per input, or q with ems_nc = dc_max-1 =
                        the reference's log-QSPA setting, decoder_method 2 (NB/src/Simulation.cpp:64-67) */
    int ems_nc;      /* EMS_NC (:32) */
    float sigma;     /* for the fused demappers */
    int *iters_out;  /* [F] iter_number as the reference returns it (iterations-1 on success) */
    int *ok_out;     /* [F] 1 = syndrome satisfied */
    void *stream;
} nb_decode_opts_t;

void nb_decode_opts_default(nb_decode_opts_t *opts);

/* Replaces Demodulate + Decoding_EMS(_GPU) / Decoding_TMM(_GPU) / Decoding_layered_TMM
 * (NB/include/Decode_GPU.cuh:17,19; called from NB/src/Simulation.cpp:130-138), batched over F
 * frames.  hard_syms: uint16 [F][N].  iters = maxIT (NB/include/define.h:35).  Threading as for the binary
 * handle: calls on one handle are safe from any host thread / stream and serialise on its scratch arena.     */
int nb_ldpc_decode_batch(const nb_ldpc_code_t *code, const void *in, uint16_t *hard_syms, int iters,
                         const nb_decode_opts_t *opts);

/* NB/src/main.cu:221-228: sigma from SNR; bits per channel use = log2(n_qam), rate = (N-M)/N.
 * snrtype 0 = Eb/N0, 1 = Es/N0.  n_qam <= 0: the loaded constellation's size.                 */
float nb_ldpc_sigma(const nb_ldpc_code_t *code, int snrtype, float snr_db, int n_qam);

/* On-device transmitter + channel for the non-binary simulator: replaces BitToSym / Modulate
 * (NB/src/LDPC_Encoder.cpp:6-37, NB/src/main.cu:190-212) and the complex AWGNChannel_CPU
 * (NB/src/LDPC_Encoder.cpp:41-68).  codeword_syms_dev: uint16 [N] broadcast to all frames (NULL =
 * all-zero).  BPSK constellation: out = float [F][N*p] (real parts; the reference also draws and then
 * ignores an imaginary part); otherwise out = float [F][N][2].  Noise: Philox4x32-10 keyed by
 * (seed, global frame index, sample index).  Output feeds nb_ldpc_decode_batch (NB_IN_BPSK / NB_IN_QAM). */
int nb_ldpc_modulate_awgn(const nb_ldpc_code_t *code, float *out_dev, int batch, float sigma, uint64_t seed,
                          uint64_t first_frame, const uint16_t *codeword_syms_dev, void *stream);

/* Replaces NB Statistic (NB/src/Simulation.cpp:256-311) on device-resident decoder output
 * hard_syms_dev uint16 [F][N]: counters_dev (ldpc_sim_counters_t layout) accumulates frames, error
 * frames (any symbol wrong), symbol errors over ALL N symbols (what the reference prints as "BER"),
 * iterations, false frames (wrong but syndrome ok) and alarm frames (right but flagged).      */
int nb_ldpc_statistic(const nb_ldpc_code_t *code, const uint16_t *hard_syms_dev, const int *iters_dev,
                      const int *ok_dev, int batch, const uint16_t *codeword_syms_dev, int64_t *counters_dev,
                      void *stream);

/* Systematic encoder over GF(q) for test vectors (the reference has none: its only non-zero codeword is the
 * hard-coded CodeWord_sym_test of the BDS code, NB/include/codeword_test.h:1, NB/src/main.cu:190-212).  Gauss-Jordan
 * elimination of H with the loaded tables, cached in the handle.  nb_ldpc_encode_info: *K = N - rank(H) information
 * symbols and their positions in the codeword (ascending; the first K positions whenever the last M columns of H are
 * invertible).  nb_ldpc_encode: info_syms uint16 [K] -> codeword_syms uint16 [N] with H c = 0 (host buffers).        */
int nb_ldpc_encode_info(nb_ldpc_code_t *code, int *K, int *info_positions);
int nb_ldpc_encode(nb_ldpc_code_t *code, const uint16_t *info_syms, uint16_t *codeword_syms);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_H */
