// On-device channel and statistics for the simulation loop.
//  * ldpc_awgn_bpsk replaces AWGNChannel_CPU + RandomModule (B/LDPC_Encoder.cu:25-56): the reference
//    draws 2 uniforms per sample from three 16-bit LCGs on ONE host thread and copies F*N floats over
//    PCIe per batch (B/Simulation.cu:137-138); here every sample is a pure function of
//    (seed, global frame index, bit index) through Philox4x32-10, so any sharding of the frame range
//    over GPUs produces the same noise, and nothing crosses PCIe.
//  * ldpc_statistic replaces Statistic (B/Simulation.cu:245-285) with the same counter semantics,
//    reduced on the device.
#include "common.h"
#include "philox.cuh"

namespace ldpcb {

// one thread = 4 consecutive bits n..n+3 of one frame (one Philox call, two Box-Muller pairs)
__global__ void __launch_bounds__(256)
awgn_bpsk_kernel(float *__restrict__ y, int N, int F, int layout, float sigma, uint32_t k0, uint32_t k1,
                 unsigned long long first_frame, const uint8_t *__restrict__ cw)
{
    const int NB = (N + 3) / 4;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)NB * F) return;
    // frame index fastest for the reference layout [N][F] (coalesced stores), bit block fastest otherwise
    int f, nb;
    if (layout == LDPC_LAYOUT_NF) {
        f = (int)(tid % F);
        nb = (int)(tid / F);
    } else {
        nb = (int)(tid % NB);
        f = (int)(tid / NB);
    }
    const unsigned long long gf = first_frame + (unsigned long long)f;
    float g[4];
    awgn_normals4(gf, nb, k0, k1, g);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int n = nb * 4 + j;
        if (n >= N) break;
        const size_t o = (layout == LDPC_LAYOUT_NF) ? (size_t)n * F + f : (size_t)f * N + n;
        y[o] = awgn_bpsk_sample(cw ? (cw[n] & 1) : 0, sigma, g[j]);  // BPSK, B/LDPC_Encoder.cu:10-17
    }
}

// one thread per frame; counters: frames, error frames, error bits, iterations, false, alarm
__global__ void __launch_bounds__(256)
statistic_kernel(const void *__restrict__ D, int fmt, const int *__restrict__ ok, const int *__restrict__ iters,
                 int N, int F, int length, const uint8_t *__restrict__ cw, unsigned long long *__restrict__ cnt)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v[6] = {0, 0, 0, 0, 0, 0};
    if (f < F) {
        int err = 0;
        if (fmt == LDPC_OUT_BITPACK) {
            const unsigned *w = reinterpret_cast<const unsigned *>(D) + (size_t)f * ((N + 31) / 32);
            for (int n0 = 0; n0 < length; n0 += 32) {
                unsigned x = w[n0 >> 5], c = 0;
                if (cw)
                    for (int b = 0; b < 32 && n0 + b < N; b++) c |= (unsigned)(cw[n0 + b] & 1) << b;
                x ^= c;
                if (length - n0 < 32) x &= (1u << (length - n0)) - 1u;
                err += __popc(x);
            }
        } else {
            for (int n = 0; n < length; n++) {
                const int d = (fmt == LDPC_OUT_INT32_REF) ? reinterpret_cast<const int *>(D)[(size_t)n * F + f]
                                                          : (int)reinterpret_cast<const unsigned char *>(D)[(size_t)n * F + f];
                err += (d != (cw ? (int)cw[n] : 0));
            }
        }
        const int flag = ok ? ok[f] : reinterpret_cast<const int *>(D)[(size_t)N * F + f];
        v[0] = 1;
        v[1] = (err != 0 || flag == 0);  // B/Simulation.cu:257
        v[2] = (unsigned long long)err;  // :256
        v[3] = (unsigned long long)iters[f];
        v[4] = (err != 0 && flag == 1);  // false frame :259
        v[5] = (err == 0 && flag == 0);  // alarm frame :258
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
        unsigned long long x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(cnt + k, x);
    }
}

}  // namespace ldpcb

using namespace ldpcb;

extern "C" void ldpc_philox4x32(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = {counter[0], counter[1], counter[2], counter[3]};
    philox4x32_10(c, key[0], key[1]);
    for (int i = 0; i < 4; i++) out[i] = c[i];
}

extern "C" int ldpc_awgn_bpsk(const ldpc_code_t *c, float *y, int F, int layout, float sigma, uint64_t seed,
                              uint64_t first_frame, const uint8_t *cw, void *stream)
{
    if (!c || !y || F <= 0 || (layout != LDPC_LAYOUT_NF && layout != LDPC_LAYOUT_FN)) return LDPC_ERR_ARG;
    const long long n = (long long)((c->N + 3) / 4) * F;
    awgn_bpsk_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        y, c->N, F, layout, sigma, (uint32_t)seed, (uint32_t)(seed >> 32), first_frame, cw);
    LDPC_CUDA_TRY(cudaGetLastError());
    return 1;
}

extern "C" int ldpc_statistic(const ldpc_code_t *c, const void *D, int fmt, const int *ok, const int *iters, int F,
                              int length, const uint8_t *cw, int64_t *counters, void *stream)
{
    if (!c || !D || !iters || !counters || F <= 0 || length <= 0 || length > c->N) return LDPC_ERR_ARG;
    if (fmt < 0 || fmt > 2 || (fmt != LDPC_OUT_INT32_REF && !ok)) return LDPC_ERR_ARG;
    statistic_kernel<<<(F + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        D, fmt, ok, iters, c->N, F, length, cw, reinterpret_cast<unsigned long long *>(counters));
    LDPC_CUDA_TRY(cudaGetLastError());
    return 1;
}
