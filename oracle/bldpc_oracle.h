/*
 * bldpc_oracle.h — CPU oracle for the binary QC-LDPC decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under cuda_ldpc_b200/ may include, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * It is a plain-C restatement of the reference's algorithm (gsw4869/CUDA_LDPC,
 * tree "bldpc_实习", abbreviated B/ below); every function cites the reference
 * file:line it follows.  Parity pin: the flooding fp32 path in `literal` table
 * mode, driven by the reference RNG, is checked frame-count-exact against the
 * reference's own sources executed on the CPU (oracle/_ref, built by
 * oracle/build_ref.sh) — see tests/test_oracle_binary.py and tests/golden/.
 * The layered schedule and the int8 quantisation are NOT in the reference
 * ("parity unpinned" for those two modes: the update rules a4/a5 are pinned, the
 * schedule and the fixed-point rules are defined here and the CUDA kernels are
 * held bit-exact to them).
 */
#ifndef BLDPC_ORACLE_H
#define BLDPC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* early-exit modes */
#define ORC_EXIT_NONE 0      /* run exactly maxit iterations                          */
#define ORC_EXIT_GENIE 1     /* reference: stop when ALL frames decode to all-zero    */
#define ORC_EXIT_SYNDROME 2  /* per frame: latch at the first zero syndrome           */

/* B/Simulation.cu:292-354 (Get_H).  H[J*L], Wc[J+1], Wv[L+1] (last slot = max). */
int orc_get_h(const char *path, int J, int L, int *H, int *Wc, int *Wv);

/* B/Simulation.cu:363-387 (Transform_H).  addr[N*Wv[L]], pre-filled with -1.
 * literal != 0: bit-for-bit the reference ternary (wrong rows);
 * literal == 0: the intended circulant, row = (col - s) mod Z.                    */
void orc_transform_h(const int *H, int J, int L, int Z, const int *Wc, const int *Wv,
                     int *addr, int literal);

/* B/LDPC_Encoder.cu:46-56 (RandomModule). */
float orc_random_module(int *seed);

/* B/LDPC_Encoder.cu:25-41 (AWGNChannel_CPU): out[n*F+f], frame-major draw order. */
void orc_awgn(int *seed, float sigma, const int *codeword, float *out, int N, int F);

/* B/main.cu:120-127: sigma from SNR (snrtype 0 = Eb/N0 with `rate`, 1 = Es/N0). */
float orc_sigma(int snrtype, float snr_db, float rate);

/* Flooding un-normalised min-sum, fp32 — B/LDPC_Decoder.cu:94-153 host loop,
 * :172-261 VN kernels, :262-372 CN kernels, :374-398 sortQ.
 * y[N*F] (n-major, f fastest); D[(N+1)*F] int: rows 0..N-1 hard bits, row N flag.
 * iters[F] per-frame iterations (GENIE / NONE: same value in every slot).
 * rq (optional, M*Wc[J]*F floats) receives the final Memory_RQ contents.
 * length = number of leading bits the genie test looks at (msgLen).               */
int orc_flooding_fp32(int J, int L, int Z, const int *H, const int *Wc, const int *Wv,
                      const int *addr, const float *y, int F, int maxit, int exit_mode,
                      int length, int *D, int *iters, float *rq);

/* Layered (row-block) min-sum, fp32; alpha multiplies min1/min2 (1.0f = reference
 * rule).  app_out (optional) N*F final APP values; D/iters as above, flag row =
 * true syndrome check.  Schedule not in the reference.               */
int orc_layered_fp32(int J, int L, int Z, const int *H, const float *y, int F, int maxit,
                     float alpha, int exit_mode, int *D, int *iters, float *app_out);

/* Layered min-sum, int8 state (rules in bldpc_oracle.c header).
 * y fp32 [N*F]; q = sat127(rint(y*scale)).  app_out (optional) int8 [N*F];
 * rec_out (optional) uint32 [M*4*F]: {m1, m2, idx, signs} per check per frame.     */
int orc_layered_i8(int J, int L, int Z, const int *H, const float *y, int F, int maxit,
                   float scale, int amax, int bnum, int bshift, int exit_mode, int *D,
                   int *iters, int8_t *app_out, uint32_t *rec_out);

/* Layered min-sum, fp16 state and messages (rules in bldpc_oracle.c "fp16 layered rules"; ours, parity unpinned).
 * app_out (optional) binary16 patterns [N*F]; msg_out (optional) patterns [M*dc_max*F], absent edges 0.             */
int orc_layered_f16(int J, int L, int Z, const int *H, const float *y, int F, int maxit, float scale, int amax,
                    int bnum, int bshift, int exit_mode, int *D, int *iters, uint16_t *app_out, uint16_t *msg_out);
uint16_t orc_f16_from_f32(float f);
float orc_f16_to_f32(uint16_t h);

/* true syndrome of hard bits x[N*F] (bit per int) -> ok[F] (1 = all checks satisfied) */
void orc_syndrome_ok(int J, int L, int Z, const int *H, const int *D, int F, int *ok);

/* B/Simulation.cu:245-285 (Statistic) counters, B/struct.cuh:16-33. */
typedef struct {
    long num_Frames, num_Error_Frames, num_Error_Bits, Total_Iteration, num_False_Frames,
        num_Alarm_Frames;
} orc_sim_counters;

/* One batch of Statistic(): D as produced above, codeword[N*F], iteraTime per frame. */
int orc_statistic(orc_sim_counters *c, const int *codeword, const int *D, const int *iters,
                  int N, int F, int length, long leastErrorFrames, long leastTestFrames);

/* B/Simulation.cu:12-171 + B/main.cu:114-157: one SNR point of the reference loop
 * (all-zero codeword, reference RNG reseeded to `seed0`, flooding decoder with the
 * reference's genie exit) until the stop rule fires or max_frames is reached.      */
int orc_sim_point(int J, int L, int Z, const int *H, const int *Wc, const int *Wv,
                  const int *addr, int snrtype, float snr_db, int F, int maxit, int length,
                  int exit_mode, long leastErrorFrames, long leastTestFrames, long max_frames,
                  int seed0, orc_sim_counters *out);

#ifdef __cplusplus
}
#endif
#endif
