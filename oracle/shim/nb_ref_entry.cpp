/*
 * C entry points onto the reference's non-binary CPU decoder (myNBLDPC/src/LDPC_Decoder.cpp etc.)
 * built by oracle/build_ref_nb.sh into oracle/_ref/libnbldpc_ref_<cfg>.so.  TEST INFRASTRUCTURE
 * ONLY.  Setup restates NB/src/main.cu:44-72 (allocate VN/CN, Get_H, GFInitial,
 * Get_CONSTELLATION); everything else is the reference's own code.  Must be used with the
 * working directory set to the reference's myNBLDPC/ (relative ./GF and ./Constellation paths).
 */
#include "define.h"
#include "LDPC_Decoder.h"
#include "LDPC_Encoder.h"
#include "GF.h"
#include "Decode_GPU.cuh"

shim_dim3 threadIdx, blockIdx, blockDim, gridDim;

/* the GPU twins are referenced by Simulation.cpp's worker but never called here */
int Decoding_EMS_GPU(const LDPCCode *, VN *, CN *, int, int, int *, const unsigned *, const unsigned *, const int *,
                     const int *, const int *, const int *, const int *, int &) { return -1; }
int Decoding_TMM_GPU(const LDPCCode *, VN *, CN *, int, int, int *, const unsigned *, const unsigned *,
                     const unsigned *, const int *, const int *, const int *, const int *, const int *, int &) { return -1; }

static LDPCCode *g_H = 0;
static VN *g_vn = 0;
static CN *g_cn = 0;
static CComplex *g_const = 0;

extern "C" int nbref_init()
{
    if (g_H) return 0;
    g_H = (LDPCCode *)malloc(sizeof(LDPCCode));
    FILE *fp = fopen(Matrixfile, "r");
    if (!fp) return -1;
    if (fscanf(fp, "%d", &g_H->Variablenode_num) != 1 || fscanf(fp, "%d", &g_H->Checknode_num) != 1) return -2;
    fclose(fp);
    g_vn = (VN *)malloc(g_H->Variablenode_num * sizeof(VN));
    g_cn = (CN *)malloc(g_H->Checknode_num * sizeof(CN));
    Get_H(g_H, g_vn, g_cn);
    GFInitial(GFQ);
    g_const = Get_CONSTELLATION(g_H);
    return 0;
}

extern "C" void nbref_info(int *out) /* N M q p dv dc n_qam maxIT */
{
    int v[8] = {g_H->Variablenode_num, g_H->Checknode_num, GFQ, g_H->q_bit, maxdv, maxdc, n_QAM, maxIT};
    memcpy(out, v, sizeof v);
}

extern "C" float nbref_sigma(float snr) /* NB/src/main.cu:221-224 (snrtype 0) */
{
    return (float)sqrt(0.5 / (log(n_QAM) / log(2) * g_H->rate * (pow(10.0, (snr / 10.0)))));
}

/* modulate `sym`, add noise with the reference RNG (seed updated in place), demodulate */
extern "C" void nbref_channel(int *seed, float sigma, const int *sym, float *rx, float *L_ch)
{
    const int N = g_H->Variablenode_num, len = (n_QAM != 2) ? N : g_H->bit_length;
    int *bits = (int *)calloc(g_H->bit_length, sizeof(int));
    int *syms = (int *)malloc(N * sizeof(int));
    CComplex *tx = (CComplex *)malloc(len * sizeof(CComplex)), *out = (CComplex *)malloc(len * sizeof(CComplex));
    memcpy(syms, sym, N * sizeof(int));
    if (n_QAM != 2) {
        Modulate(g_H, g_const, tx, syms);
    } else { /* NB/src/main.cu:200-211 */
        for (int i = 0; i < N; i++)
            for (int j = 0; j < g_H->q_bit; j++) bits[i * g_H->q_bit + j] = (sym[i] & (1 << j)) >> j;
        Modulate(g_H, g_const, tx, bits);
    }
    AWGNChannel a;
    a.seed[0] = seed[0]; a.seed[1] = seed[1]; a.seed[2] = seed[2];
    a.sigma = sigma;
    AWGNChannel_CPU(g_H, &a, out, tx);
    seed[0] = a.seed[0]; seed[1] = a.seed[1]; seed[2] = a.seed[2];
    for (int i = 0; i < len; i++) {
        rx[2 * i] = out[i].Real;
        rx[2 * i + 1] = out[i].Image;
    }
    Demodulate(g_H, &a, g_const, g_vn, out);
    for (int s = 0; s < N; s++)
        for (int q = 0; q < GFQ - 1; q++) L_ch[s * (GFQ - 1) + q] = g_vn[s].L_ch[q];
    free(bits); free(syms); free(tx); free(out);
}

/* algo: 0 EMS(EMS_NM, EMS_NC), 1 TMM, 3 layered TMM.  Returns the decoder's return value. */
extern "C" int nbref_decode(int algo, const float *L_ch, int *out, int *iter_number)
{
    const int N = g_H->Variablenode_num;
    for (int s = 0; s < N; s++)
        for (int q = 0; q < GFQ - 1; q++) g_vn[s].L_ch[q] = L_ch[s * (GFQ - 1) + q];
    int it = 0, r;
    if (algo == 0)
        r = Decoding_EMS(g_H, g_vn, g_cn, EMS_NM, EMS_NC, out, it);
    else if (algo == 1)
        r = Decoding_TMM(g_H, g_vn, g_cn, EMS_NM, EMS_NC, out, it);
    else
        r = Decoding_layered_TMM(g_H, g_vn, g_cn, EMS_NM, EMS_NC, out, it);
    *iter_number = it;
    return r;
}
