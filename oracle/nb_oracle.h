/*
 * nb_oracle.h — CPU oracle for the non-binary GF(q) LDPC decode path.  TEST INFRASTRUCTURE ONLY
 * (see bldpc_oracle.h for the rules).  NB/ = /root/reference/myNBLDPC/ (gsw4869/CUDA_LDPC).
 *
 * Pinned against the reference's own CPU decoder (NB/src/LDPC_Decoder.cpp compiled unchanged into
 * oracle/_ref/libnbldpc_ref_*.so by oracle/build_ref_nb.sh): Demodulate, Decoding_EMS (literal
 * running-sum recursion), Decoding_TMM, Decoding_layered_TMM, the complex AWGN channel and RNG —
 * frame-exact on the golden vectors under tests/golden/nb_*.npz.
 * Defined here ("parity unpinned"): the `fresh` EMS summation order the CUDA kernel is held to
 * (SURVEY C.3), and the exponent -> alpha^e coefficient mapping for *_exp.txt files.
 */
#ifndef NB_ORACLE_H
#define NB_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nb_orc_code nb_orc_code;

/* NB/src/Simulation.cpp:347-466 (Get_H) + NB/src/GF.cpp:68-117 (GFInitial) +
 * NB/src/Simulation.cpp:313-338 (Get_CONSTELLATION).  constellation may be NULL.
 * coef_is_exponent: map coefficient e -> alpha^e (alpha = 2).  Returns NULL on error.       */
nb_orc_code *nb_orc_load(const char *matrix, const char *gf_table, const char *constellation, int n_qam,
                         int coef_is_exponent);
void nb_orc_free(nb_orc_code *c);
/* out[0..6] = N, M, q, p (bits/symbol), dv_max, dc_max, n_qam */
void nb_orc_info(const nb_orc_code *c, int *out);
/* mul[q*q], inv[q]; check lists vn[M*dc_max], coef[M*dc_max] (-1 padded), weight[M]; any may be NULL */
void nb_orc_tables(const nb_orc_code *c, int *mul, int *inv, int *check_vn, int *check_coef, int *check_w);
/* constellation points re[n_qam], im[n_qam] */
void nb_orc_constellation(const nb_orc_code *c, float *re, float *im);

/* syndrome of symbols x[N]: returns 1 if all checks are satisfied (NB/src/LDPC_Decoder.cpp:218-231) */
int nb_orc_syndrome_ok(const nb_orc_code *c, const int *x);

/* NB/src/LDPC_Encoder.cpp:41-79: complex AWGN, cos branch, 4 uniforms per sample.
 * tx/rx: interleaved (re, im) pairs, len samples.                                          */
void nb_orc_awgn(int *seed, float sigma, const float *tx, float *rx, int len);
/* NB/src/LDPC_Encoder.cpp:18-37 + NB/src/main.cu:190-212: symbols -> channel input.
 * BPSK (n_qam == 2): N*p samples from the LSB-first bits; QAM: N samples.  Returns len.     */
int nb_orc_modulate(const nb_orc_code *c, const int *sym, float *tx);
/* NB/src/main.cu:221-228 */
float nb_orc_sigma(const nb_orc_code *c, int snrtype, float snr_db);

/* NB/src/LDPC_Decoder.cpp:132-171 (Demodulate): rx interleaved complex -> L_ch[N*(q-1)]       */
void nb_orc_demodulate(const nb_orc_code *c, float sigma, const float *rx, float *L_ch);

#define NB_ORC_EMS 0
#define NB_ORC_TMM 1
#define NB_ORC_LAYERED_TMM 3
#define NB_ORC_FFT_BP 4      /* NOT in the reference ("parity unpinned"): probability-domain BP with
                                the check node as a Walsh-Hadamard-domain product; rules in nb_oracle.c */
#define NB_ORC_SUM_LITERAL 0 /* the reference's running sum (inc before / dec after recursion) */
#define NB_ORC_SUM_FRESH 1   /* sum of the inputs in ascending edge position, from 0.0f          */

/* NB/src/LDPC_Decoder.cpp:172-317 / :361-542 / :544-702.  L_ch[N*(q-1)] -> out[N] symbols.
 * Returns 1 if the syndrome was satisfied; *iter_number as the reference sets it
 * (iterations-1 on success, maxit on failure).                                              */
int nb_orc_decode(const nb_orc_code *c, int algo, int sum_mode, const float *L_ch, int maxit, int ems_nm,
                  int ems_nc, int *out, int *iter_number);

/* batch helper for baselines/tests: F frames, OpenMP over frames. L_ch[F][N*(q-1)], out[F][N]  */
void nb_orc_decode_batch(const nb_orc_code *c, int algo, int sum_mode, const float *L_ch, int F, int maxit,
                         int ems_nm, int ems_nc, int *out, int *iters, int *ok);

#ifdef __cplusplus
}
#endif
#endif
