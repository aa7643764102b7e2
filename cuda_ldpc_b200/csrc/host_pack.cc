// Host-side packing for the host-buffer path of ldpc_decode_batch (ldpc_decode_opts_t::host_pack_threads):
// fp32 channel values [N][F] (the reference's Channel_Out layout, B/Simulation.cu:138) -> int8 [N][fc] of one
// frame chunk, with the layered kernel's own quantisation rule q = sat127(rint(y * scale)) (round to nearest
// even, like __float2int_rn), so the decode is bit-identical to copying the fp32 values — at a quarter of the
// PCIe bytes.  Plain C++ compiled by g++ (AVX2 clone selected at load time); measured 93 GB/s of fp32 read with
// 16 threads on the B200 box's host (tools/ubench/hostquant.c), against ~55 GB/s of PCIe.
#include <math.h>
#include <omp.h>
#include <stddef.h>

namespace {
__attribute__((target_clones("avx2", "default"))) void quant_row(const float *__restrict__ y, signed char *__restrict__ q,
                                                                  int n, float scale)
{
    for (int i = 0; i < n; i++) {
        float v = nearbyintf(y[i] * scale);
        v = v > 127.0f ? 127.0f : (v < -127.0f ? -127.0f : v);
        q[i] = (signed char)(int)v;
    }
}
}  // namespace

// y: host fp32 [N][ldF]; out: int8 [N][fc] = columns [f0, f0 + fc) of y
extern "C" void ldpcb_host_pack_nf(const float *y, size_t ldF, int N, int f0, int fc, float scale, signed char *out,
                                   int threads)
{
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int n = 0; n < N; n++) quant_row(y + (size_t)n * ldF + f0, out + (size_t)n * fc, fc, scale);
}
