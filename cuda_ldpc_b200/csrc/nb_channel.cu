// Device-side transmitter, channel and statistics of the non-binary simulator.
//  * nb_ldpc_modulate_awgn replaces BitToSym / Modulate / AWGNChannel_CPU (NB/src/LDPC_Encoder.cpp:6-68,
//    NB/src/main.cu:190-212): the reference draws 4 uniforms per complex sample from three 16-bit
//    LCGs under a global mutex (NB/src/Simulation.cpp:44-48); here each sample is a pure function of
//    (seed, global frame, sample index) through Philox4x32-10 (one call = one complex sample).
//  * nb_ldpc_statistic replaces Statistic (NB/src/Simulation.cpp:256-311).
#include <math.h>

#include <mutex>

#include "common.h"
#include "nb_common.h"

namespace ldpcb {

__device__ __forceinline__ void nb_philox(uint32_t c[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1;
        c[3] = (uint32_t)p0;
        c[0] = n0;
        c[2] = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ float nb_u01(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }
// radius uniform with 32-bit resolution near 0 (the normal's tail reaches 6.76 sigma; philox.cuh)
__device__ __forceinline__ float nb_u01_tail(uint32_t x) { return __fmaf_rn((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }

// one thread = one channel use (a BPSK bit or a QAM symbol) of one frame
__global__ void __launch_bounds__(256)
nb_modulate_awgn_kernel(float *__restrict__ out, int N, int p, int bpsk, int F, float sigma, uint32_t k0, uint32_t k1,
                        unsigned long long first_frame, const uint16_t *__restrict__ cw, const float *__restrict__ cre,
                        const float *__restrict__ cim)
{
    const int L = bpsk ? N * p : N;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)L * F) return;
    const int f = (int)(tid / L), i = (int)(tid % L);
    const unsigned long long gf = first_frame + (unsigned long long)f;
    uint32_t c[4] = {(uint32_t)gf, (uint32_t)(gf >> 32), (uint32_t)i, 0x4E424C44u /* "NBLD" */};
    nb_philox(c, k0, k1);
    const float r0 = sqrtf(-2.0f * __logf(nb_u01_tail(c[0]))), r1 = sqrtf(-2.0f * __logf(nb_u01_tail(c[2])));
    const float g0 = r0 * __cosf(6.283185307179586f * nb_u01(c[1]));  // the reference's cos branch
    const float g1 = r1 * __cosf(6.283185307179586f * nb_u01(c[3]));
    if (bpsk) {
        const int s = i / p, b = i - s * p;
        const int bit = cw ? ((cw[s] >> b) & 1) : 0;  // LSB first (NB/src/main.cu:206)
        out[(size_t)f * L + i] = cre[bit] + sigma * g0;
    } else {
        const int sym = cw ? cw[i] : 0;
        out[((size_t)f * L + i) * 2] = cre[sym] + sigma * g0;
        out[((size_t)f * L + i) * 2 + 1] = cim[sym] + sigma * g1;
    }
}

__global__ void __launch_bounds__(256)
nb_statistic_kernel(const uint16_t *__restrict__ sym, const int *__restrict__ iters, const int *__restrict__ ok, int N,
                    int F, const uint16_t *__restrict__ cw, unsigned long long *__restrict__ cnt)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v[6] = {0, 0, 0, 0, 0, 0};
    if (f < F) {
        int err = 0;
        for (int n = 0; n < N; n++) err += sym[(size_t)f * N + n] != (cw ? cw[n] : (uint16_t)0);
        v[0] = 1;
        v[1] = err != 0;                 // NB/src/Simulation.cpp:270
        v[2] = (unsigned long long)err;  // :268 (symbol errors; printed in the "BER" column)
        v[3] = (unsigned long long)iters[f];
        v[4] = (err != 0 && ok[f] == 1);
        v[5] = (err == 0 && ok[f] == 0);
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
        unsigned long long x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(cnt + k, x);
    }
}

int nb_upload_tables(nb_ldpc_code *c);  // nb_decode.cu

}  // namespace ldpcb

using namespace ldpcb;

extern "C" float nb_ldpc_sigma(const nb_ldpc_code_t *c, int snrtype, float snr_db, int n_qam)
{
    if (!c) return 0.0f;
    if (n_qam <= 0) n_qam = c->n_const > 0 ? c->n_const : 2;
    // NB/src/main.cu:221-228
    if (snrtype == 0)
        return (float)sqrt(0.5 / (log((double)n_qam) / log(2.0) * c->rate * (pow(10.0, (snr_db / 10.0)))));
    return (float)sqrt(0.5 / (log((double)n_qam) / log(2.0) * pow(10.0, (snr_db / 10.0))));
}

extern "C" int nb_ldpc_modulate_awgn(const nb_ldpc_code_t *cc, float *out, int F, float sigma, uint64_t seed,
                                     uint64_t first_frame, const uint16_t *cw, void *stream)
{
    if (!cc || !out || F <= 0) return LDPC_ERR_ARG;
    nb_ldpc_code *c = const_cast<nb_ldpc_code *>(cc);
    if (c->n_const != 2 && c->n_const != c->q) return LDPC_ERR_ARG;  // needs a constellation
    int rc = nb_upload_tables(c);
    if (rc != LDPC_OK) return rc;
    const int bpsk = c->n_const == 2;
    const long long n = (long long)(bpsk ? c->N * c->p : c->N) * F;
    nb_modulate_awgn_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        out, c->N, c->p, bpsk, F, sigma, (uint32_t)seed, (uint32_t)(seed >> 32), first_frame, cw, c->d_cre, c->d_cim);
    LDPC_CUDA_TRY(cudaGetLastError());
    return 1;
}

extern "C" int nb_ldpc_statistic(const nb_ldpc_code_t *c, const uint16_t *sym, const int *iters, const int *ok, int F,
                                 const uint16_t *cw, int64_t *counters, void *stream)
{
    if (!c || !sym || !iters || !ok || !counters || F <= 0) return LDPC_ERR_ARG;
    nb_statistic_kernel<<<(F + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        sym, iters, ok, c->N, F, cw, reinterpret_cast<unsigned long long *>(counters));
    LDPC_CUDA_TRY(cudaGetLastError());
    return 1;
}
