/*
 * C entry points onto the reference's binary decoder built under the CPU shim
 * (oracle/_ref/libbldpc_ref_<cfg>.so).  TEST INFRASTRUCTURE ONLY.  The setup and per-SNR
 * preamble restate B/main.cu:88-135 (memset -1, Get_H, Transform_H, reseed, sigma, zero
 * counters); everything else is the reference's own code.
 */
#include "define.cuh"
#include "LDPC_Decoder.cuh"
#include "LDPC_Encoder.cuh"

shim_dim3 threadIdx, blockIdx, blockDim, gridDim;

static int g_H[J * L], g_Wc[J + 1], g_Wv[L + 1];
static int *g_addr = 0;

extern "C" int ref_init()
{
    if (g_addr) return 0;
    g_addr = (int *)malloc(sizeof(int) * J * L * Z);
    memset(g_Wc, 0, sizeof(g_Wc));
    memset(g_Wv, 0, sizeof(g_Wv));
    memset(g_H, 0, sizeof(g_H));
    memset(g_addr, -1, sizeof(int) * J * L * Z);
    Get_H(g_H, g_Wc, g_Wv);
    Transform_H(g_H, g_Wc, g_Wv, g_addr);
    return 0;
}

extern "C" void ref_geometry(int *out) /* J L Z N K M F maxIT Wc_max Wv_max */
{
    ref_init();
    int v[10] = {J, L, Z, CW_Len, msgLen, parLen, Num_Frames_OneTime, maxIT, g_Wc[J], g_Wv[L]};
    memcpy(out, v, sizeof(v));
}

extern "C" void ref_tables(int *H, int *Wc, int *Wv, int *addr)
{
    ref_init();
    memcpy(H, g_H, sizeof(g_H));
    memcpy(Wc, g_Wc, sizeof(g_Wc));
    memcpy(Wv, g_Wv, sizeof(g_Wv));
    memcpy(addr, g_addr, sizeof(int) * CW_Len * g_Wv[L]);
}

extern "C" void ref_awgn(int *seed, float sigma, float *out) /* all-zero codeword */
{
    AWGNChannel a;
    a.seed[0] = seed[0];
    a.seed[1] = seed[1];
    a.seed[2] = seed[2];
    a.sigma = sigma;
    int *cw = (int *)calloc((size_t)Num_Frames_OneTime * CW_Len, sizeof(int));
    AWGNChannel_CPU(&a, out, cw);
    free(cw);
    seed[0] = a.seed[0];
    seed[1] = a.seed[1];
    seed[2] = a.seed[2];
}

extern "C" int ref_decode(const float *y, int *D)
{
    ref_init();
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    LDPCCode ldpc;
    ldpc.iteraTime = 0;
    LDPC_Decoder_GPU(D, (float *)y, prop, g_addr, g_Wc, g_Wv, &ldpc);
    return ldpc.iteraTime;
}

extern "C" void ref_sim_point(float snr, long *counters6)
{
    ref_init();
    AWGNChannel awgn;
    Simulation sim;
    memset(&sim, 0, sizeof(sim));
    sim.SNR = snr;
    awgn.seed[0] = ix_define;
    awgn.seed[1] = iy_define;
    awgn.seed[2] = iz_define;
    if (snrtype == 0)
        awgn.sigma = (float)sqrt(0.5 / (rate * (pow(10.0, (sim.SNR / 10.0)))));
    else
        awgn.sigma = (float)sqrt(0.5 / (pow(10.0, (sim.SNR / 10.0))));
    float sigma_dev = awgn.sigma;
    Simulation_GPU(&awgn, &sigma_dev, &sim, g_addr, g_Wc, g_Wv);
    counters6[0] = sim.num_Frames;
    counters6[1] = sim.num_Error_Frames;
    counters6[2] = sim.num_Error_Bits;
    counters6[3] = sim.Total_Iteration;
    counters6[4] = sim.num_False_Frames;
    counters6[5] = sim.num_Alarm_Frames;
}
