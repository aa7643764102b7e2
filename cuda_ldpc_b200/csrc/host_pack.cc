// Host-side packing for the host-buffer path of ldpc_decode_batch (ldpc_decode_opts_t::host_pack_threads):
// fp32 channel values [N][F] (the reference's Channel_Out layout, B/Simulation.cu:138) -> int8 [N][fc] of one
// frame chunk, with the layered kernel's own quantisation rule q = sat127(rint(y * scale)) (round to nearest
// even; NaN -> -127 like the kernel's clamp and the oracle), so the decode is bit-identical to copying the fp32 values — at a quarter
// of the PCIe bytes.  The loop is bound by the host's DRAM: it reads the fp32 values once and writes the bytes with
// non-temporal stores (no read-for-ownership of the pinned staging buffer).  AVX-512 path chosen at run time
// (__builtin_cpu_supports), portable scalar path otherwise; both follow the same rule bit for bit
// (tests/test_binary_gpu.py::test_host_pack_extreme_values).
#include <immintrin.h>
#include <math.h>
#include <omp.h>
#include <sched.h>
#include <stddef.h>
#include <stdint.h>

namespace {

inline signed char quant1(float y, float scale)
{
    float v = nearbyintf(y * scale);
    v = v > 127.0f ? 127.0f : (v < -127.0f ? -127.0f : v);
    return (v != v) ? (signed char)-127 : (signed char)(int)v;
}

void quant_row_scalar(const float *__restrict__ y, signed char *__restrict__ q, int n, float scale)
{
    for (int i = 0; i < n; i++) q[i] = quant1(y[i], scale);
}

// 16 floats -> 16 bytes (low 128 bits): clamp in float (the bounds are integers, so clamping commutes with the
// rounding), convert with the current rounding mode (nearest even).  NaN lanes -> -127: max_ps returns its SECOND
// operand when either is NaN, which is the kernel's fmaxf(NaN, -127) = -127
__attribute__((target("avx512f,avx512bw"))) inline __m128i quant16(const float *y, __m512 vs)
{
    const __m512 v = _mm512_mul_ps(_mm512_loadu_ps(y), vs);
    const __m512 c = _mm512_min_ps(_mm512_max_ps(v, _mm512_set1_ps(-127.0f)), _mm512_set1_ps(127.0f));
    return _mm512_cvtsepi32_epi8(_mm512_cvtps_epi32(c));
}

// `ynext`: the row this thread converts next.  A row segment of a chunk is ~1.2 pages long and the next one lies
// a whole row of the batch further on, so the hardware prefetcher restarts on every row; prefetching the next
// segment line by line while this one is converted keeps the loop bandwidth-bound instead of latency-bound.
__attribute__((target("avx512f,avx512bw"))) void quant_row_avx512(const float *__restrict__ y, signed char *__restrict__ q,
                                                                 int n, float scale, const float *ynext)
{
    const __m512 vs = _mm512_set1_ps(scale);
    int i = 0;
    // head: up to the first 64-byte boundary of the output
    const int head = (int)((64 - ((uintptr_t)q & 63)) & 63);
    for (; i < head && i < n; i++) q[i] = quant1(y[i], scale);
    for (; i + 64 <= n; i += 64) {  // 64 floats in, one 64-byte non-temporal store out
        if (ynext) {
            _mm_prefetch(reinterpret_cast<const char *>(ynext + i), _MM_HINT_NTA);
            _mm_prefetch(reinterpret_cast<const char *>(ynext + i + 16), _MM_HINT_NTA);
            _mm_prefetch(reinterpret_cast<const char *>(ynext + i + 32), _MM_HINT_NTA);
            _mm_prefetch(reinterpret_cast<const char *>(ynext + i + 48), _MM_HINT_NTA);
        }
        const __m128i a = quant16(y + i, vs), b = quant16(y + i + 16, vs), c = quant16(y + i + 32, vs),
                      d = quant16(y + i + 48, vs);
        __m512i o = _mm512_castsi128_si512(a);
        o = _mm512_inserti32x4(o, b, 1);
        o = _mm512_inserti32x4(o, c, 2);
        o = _mm512_inserti32x4(o, d, 3);
        _mm512_stream_si512(reinterpret_cast<__m512i *>(q + i), o);
    }
    for (; i < n; i++) q[i] = quant1(y[i], scale);
}

}  // namespace

// CPUs this process may run on (the pool of the auto mode of host_pack_threads)
extern "C" int ldpcb_host_cpus(void)
{
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) != 0) return 1;
    const int n = CPU_COUNT(&set);
    return n < 1 ? 1 : n;
}

// y: host fp32 [N][ldF]; out: int8 [N][fc] = columns [f0, f0 + fc) of y
extern "C" void ldpcb_host_pack_nf(const float *y, size_t ldF, int N, int f0, int fc, float scale, signed char *out,
                                   int threads)
{
    if (threads < 1) threads = 1;
    static const bool avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
#pragma omp parallel num_threads(threads)
    {
#pragma omp for schedule(static) nowait
        for (int n = 0; n < N; n++) {
            if (avx512)
                quant_row_avx512(y + (size_t)n * ldF + f0, out + (size_t)n * fc, fc, scale,
                                 n + 1 < N ? y + (size_t)(n + 1) * ldF + f0 : nullptr);
            else
                quant_row_scalar(y + (size_t)n * ldF + f0, out + (size_t)n * fc, fc, scale);
        }
        _mm_sfence();  // this thread's non-temporal stores are visible before the staging buffer goes to cudaMemcpyAsync
    }
}
