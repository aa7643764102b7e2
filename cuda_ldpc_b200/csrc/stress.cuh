#pragma once
// Race-hunting builds (tests/test_stress_gpu.py; compute-sanitizer is closed on the GPU pool, so racecheck cannot be
// run there): -DLDPC_STRESS=1 makes a pseudo-random quarter of the warps sleep up to ~4 us after every CTA barrier
// and before every read of a shared flag, which turns "a fast warp is a few hundred cycles ahead" into microseconds
// of skew: a missing barrier then corrupts results within a handful of groups instead of once per million.
// -DLDPC_R1_RACES=1 re-opens the two races round 1 shipped (one s_fail word; no barrier between the output of
// latched frames and the next sweep) — the canary that proves the stress test can see them.
#ifndef LDPC_STRESS
#define LDPC_STRESS 0
#endif
#ifndef LDPC_R1_RACES
#define LDPC_R1_RACES 0
#endif
#if LDPC_STRESS
__device__ __forceinline__ void stress_point(unsigned salt)
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    unsigned h = (unsigned)(t >> 5) * 2654435761u + (threadIdx.x >> 5) * 40503u + blockIdx.x * 9176u + salt * 7919u;
    h ^= h >> 15;
    h = __shfl_sync(0xffffffffu, h, 0);
    if ((h & 3u) == 0u) __nanosleep((h >> 8) & 4095u);
}
#define LDPC_STRESS_POINT(salt) stress_point(salt)
#else
#define LDPC_STRESS_POINT(salt) ((void)0)
#endif

