// Developer microbenchmark: host-side fp32 -> int8 quantisation bandwidth (OpenMP), to decide whether quantising on
// the host before the H2D copy could beat copying fp32 (PCIe ~55 GB/s).  gcc -O3 -march=native -fopenmp hostquant.c -lm
#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int main(void)
{
    const size_t n = (size_t)38400 * 9472;
    float *y = (float *)malloc(n * 4);
    signed char *q = (signed char *)malloc(n);
#pragma omp parallel for
    for (size_t i = 0; i < n; i++) { y[i] = (float)(i % 977) * 0.01f - 4.0f; q[i] = 0; }
    for (int rep = 0; rep < 4; rep++) {
        double t0 = omp_get_wtime();
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++) {
            float v = nearbyintf(y[i] * 8.0f);
            v = v > 127.0f ? 127.0f : (v < -127.0f ? -127.0f : v);
            q[i] = (signed char)v;
        }
        double dt = omp_get_wtime() - t0;
        printf("threads %d  %.2f ms  read %.1f GB/s\n", omp_get_max_threads(), dt * 1e3, n * 4 / dt / 1e9);
    }
    // the library's pattern: 8 column chunks of an [N][F] array, rows of fc floats at stride F
    {
        const int N = 38400, F = 9472, fc = 1184;
        for (int rep = 0; rep < 3; rep++) {
            double t0 = omp_get_wtime();
            for (int f0 = 0; f0 < F; f0 += fc) {
#pragma omp parallel for schedule(static)
                for (int r = 0; r < N; r++) {
                    const float *src = y + (size_t)r * F + f0;
                    signed char *dst = q + (size_t)r * fc;
                    for (int i = 0; i < fc; i++) {
                        float v = nearbyintf(src[i] * 8.0f);
                        v = v > 127.0f ? 127.0f : (v < -127.0f ? -127.0f : v);
                        dst[i] = (signed char)v;
                    }
                }
            }
            double dt = omp_get_wtime() - t0;
            printf("chunked pattern: %.2f ms  read %.1f GB/s\n", dt * 1e3, n * 4 / dt / 1e9);
        }
    }
    printf("%d\n", q[12345]);
    return 0;
}
