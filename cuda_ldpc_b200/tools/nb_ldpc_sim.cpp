// nb_ldpc_sim — the non-binary simulator's driver over the C-ABI of libldpc_b200.so.
// Same loop as the reference (myNBLDPC/src/main.cu:215-249 SNR sweep, reseed + sigma per point;
// src/Simulation.cpp:115-158 frame loop; Statistic src/Simulation.cpp:256-311; stop rule
// errFrames >= 50 && frames >= 1000, include/define.h:52-53; result row " SNR frames errFrames FER
// SER avgIter sec/frame" of src/Simulation.cpp:198), but frames are decoded in batches on the device
// (the reference: one frame per decoder call), the channel runs on the device, and the frame range is
// sharded over --gpus devices.  Configuration is run-time (the reference: macros in include/define.h).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ldpc_b200.h"

struct Args {
    std::string matrix, gf, constellation, codeword, results;
    int encode = 0;
    int exp = 0, algo = NB_ALGO_EMS, nm = 2, nc = 2, maxit = 20, batch = 2048, gpus = 1, snrtype = 0;
    double snr_start = 0, snr_stop = 5, snr_step = 0.5;  // define.h:47-49
    long least_errors = 50, least_frames = 1000, max_frames = 0;
    unsigned long long seed = 173;
};

static void usage()
{
    printf("usage: nb_ldpc_sim --matrix FILE --constellation FILE [--gf FILE] [--exp] [--algo ems|tmm|ltmm|fftbp|logqspa]\n"
           "       [--nm n --nc n] [--snr a b step] [--snrtype 0|1] [--maxit n] [--batch F] [--least-errors n]\n"
           "       [--least-frames n] [--max-frames n] [--gpus g] [--seed s] [--codeword FILE | --encode]\n"
           "       [--results FILE]\n"
           "  --encode        transmit a random codeword (nb_ldpc_encode of random information symbols drawn from --seed)\n"
           "                  instead of the all-zero word; the reference can only send zeros or its one hard-coded word\n"
           "  --results FILE  append the reference's result row per finished SNR point (it always appends to results.txt,\n"
           "                  myNBLDPC/src/Simulation.cpp:199-206)\n");
}

struct Gpu {
    int dev, rc = 0;
    std::string err;  // ldpc_last_cuda_error() is per thread: captured inside the worker
    nb_ldpc_code_t *code = nullptr;
    cudaStream_t st = nullptr;
    float *x = nullptr;
    uint16_t *out = nullptr, *cw = nullptr;
    int *iters = nullptr, *ok = nullptr;
    int64_t *cnt = nullptr;
};

int main(int argc, char **argv)
{
    Args a;
    for (int i = 1; i < argc; i++) {
        std::string s = argv[i];
        auto next = [&](int k = 1) { if (i + k >= argc) { usage(); exit(2); } return argv[i + k]; };
        if (s == "--matrix") a.matrix = next(), i++;
        else if (s == "--gf") a.gf = next(), i++;
        else if (s == "--constellation") a.constellation = next(), i++;
        else if (s == "--codeword") a.codeword = next(), i++;
        else if (s == "--exp") a.exp = 1;
        else if (s == "--encode") a.encode = 1;
        else if (s == "--results") a.results = next(), i++;
        else if (s == "--algo") {
            std::string m = next(); i++;
            a.algo = m == "tmm" ? NB_ALGO_TMM : m == "ltmm" ? NB_ALGO_LAYERED_TMM : m == "fftbp" ? NB_ALGO_FFT_BP : NB_ALGO_EMS;
            if (m == "logqspa") a.nm = a.nc = 1 << 20;  // decoder_method 2: Nm = q, Nc = dc-1 (clamped by the library's test >=)
        }
        else if (s == "--nm") a.nm = atoi(next()), i++;
        else if (s == "--nc") a.nc = atoi(next()), i++;
        else if (s == "--snr") a.snr_start = atof(next(1)), a.snr_stop = atof(next(2)), a.snr_step = atof(next(3)), i += 3;
        else if (s == "--snrtype") a.snrtype = atoi(next()), i++;
        else if (s == "--maxit") a.maxit = atoi(next()), i++;
        else if (s == "--batch") a.batch = atoi(next()), i++;
        else if (s == "--least-errors") a.least_errors = atol(next()), i++;
        else if (s == "--least-frames") a.least_frames = atol(next()), i++;
        else if (s == "--max-frames") a.max_frames = atol(next()), i++;
        else if (s == "--gpus") a.gpus = atoi(next()), i++;
        else if (s == "--seed") a.seed = strtoull(next(), nullptr, 10), i++;
        else { usage(); return 2; }
    }
    if (a.matrix.empty() || a.constellation.empty()) { usage(); return 2; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { fprintf(stderr, "nb_ldpc_sim: %s\n", ldpc_strerror(LDPC_ERR_NO_DEVICE)); return 1; }
    if (a.gpus > ndev) a.gpus = ndev;
    std::vector<Gpu> g(a.gpus);
    nb_ldpc_code_info_t info;
    std::vector<uint16_t> cw;
    for (int d = 0; d < a.gpus; d++) {
        g[d].dev = d;
        cudaSetDevice(d);
        int rc = nb_ldpc_load_code(a.matrix.c_str(), a.gf.empty() ? nullptr : a.gf.c_str(), a.constellation.c_str(), a.exp, &g[d].code);
        if (rc != LDPC_OK) { fprintf(stderr, "nb_ldpc_sim: %s\n", ldpc_strerror(rc)); return 1; }
        nb_ldpc_code_info(g[d].code, &info);
        if (d == 0 && !a.codeword.empty()) {  // whitespace / comma separated symbols, e.g. CodeWord_sym_test
            FILE *f = fopen(a.codeword.c_str(), "r");
            if (!f) { fprintf(stderr, "nb_ldpc_sim: cannot open %s\n", a.codeword.c_str()); return 1; }
            int v;
            while ((int)cw.size() < info.N && fscanf(f, " %d%*[, \t\r\n]", &v) == 1) cw.push_back((uint16_t)v);
            fclose(f);
            if ((int)cw.size() != info.N) { fprintf(stderr, "nb_ldpc_sim: codeword file needs %d symbols\n", info.N); return 1; }
        }
        if (d == 0 && a.encode && cw.empty()) {  // random information symbols -> codeword (H c = 0 by construction)
            int K = 0;
            rc = nb_ldpc_encode_info(g[d].code, &K, nullptr);
            std::vector<uint16_t> u(K > 0 ? K : 1);
            unsigned long long s = a.seed * 6364136223846793005ull + 1442695040888963407ull;
            for (int j = 0; j < K; j++) {
                s = s * 6364136223846793005ull + 1442695040888963407ull;
                u[j] = (uint16_t)((s >> 33) % (unsigned)info.q);
            }
            cw.assign(info.N, 0);
            if (rc == LDPC_OK) rc = nb_ldpc_encode(g[d].code, u.data(), cw.data());
            if (rc != LDPC_OK) { fprintf(stderr, "nb_ldpc_sim: encoder: %s\n", ldpc_strerror(rc)); return 1; }
        }
        const bool bpsk = info.n_const == 2;
        cudaStreamCreate(&g[d].st);
        cudaMalloc(&g[d].x, (size_t)a.batch * (bpsk ? info.N * info.p : info.N * 2) * sizeof(float));
        cudaMalloc(&g[d].out, (size_t)a.batch * info.N * sizeof(uint16_t));
        cudaMalloc(&g[d].iters, a.batch * sizeof(int));
        cudaMalloc(&g[d].ok, a.batch * sizeof(int));
        cudaMalloc(&g[d].cnt, 6 * sizeof(int64_t));
        if (!cw.empty()) {
            cudaMalloc(&g[d].cw, info.N * sizeof(uint16_t));
            cudaMemcpy(g[d].cw, cw.data(), info.N * sizeof(uint16_t), cudaMemcpyHostToDevice);
        }
    }
    const int kbits = (info.N - info.M) * info.p;
    printf("* %s  non-binary LDPC simulation\n* N=%d M=%d GF(%d) dv<=%d dc<=%d, %d-point constellation, decoder %s, max %d iterations\n",
           ldpc_version(), info.N, info.M, info.q, info.dv_max, info.dc_max, info.n_const,
           a.algo == NB_ALGO_EMS ? "EMS" : a.algo == NB_ALGO_TMM ? "TMM" : a.algo == NB_ALGO_FFT_BP ? "FFT-BP" : "layered TMM", a.maxit);
    printf(" SNR   frames errFrames      FER         SER     avgIter  sec/frame   info Mbit/s\n");
    const int npts = (int)floor((a.snr_stop - a.snr_start) / a.snr_step + 1e-9) + 1;
    for (int pt = 0; pt < npts; pt++) {
        const float snr = (float)(a.snr_start + pt * a.snr_step);
        const float sigma = nb_ldpc_sigma(g[0].code, a.snrtype, snr, 0);
        for (auto &x : g) { cudaSetDevice(x.dev); cudaMemsetAsync(x.cnt, 0, 6 * sizeof(int64_t), x.st); }
        unsigned long long next_frame = 0;
        int64_t tot[6] = {0, 0, 0, 0, 0, 0};
        auto t0 = std::chrono::steady_clock::now();
        bool stop = false;
        while (!stop) {
            std::vector<std::thread> th;
            for (int d = 0; d < a.gpus; d++) {
                const unsigned long long first = next_frame + (unsigned long long)d * a.batch;
                th.emplace_back([&, d, first]() {
                    Gpu &x = g[d];
                    cudaSetDevice(x.dev);
                    nb_decode_opts_t o;
                    nb_decode_opts_default(&o);
                    o.batch = a.batch; o.algo = a.algo; o.mem_space = LDPC_MEM_DEVICE; o.ems_nm = a.nm; o.ems_nc = a.nc;
                    o.in_kind = info.n_const == 2 ? NB_IN_BPSK : NB_IN_QAM; o.sigma = sigma;
                    o.iters_out = x.iters; o.ok_out = x.ok; o.stream = x.st;
                    int rc = nb_ldpc_modulate_awgn(x.code, x.x, a.batch, sigma, a.seed, first, x.cw, x.st);
                    if (rc >= 0) rc = nb_ldpc_decode_batch(x.code, x.x, x.out, a.maxit, &o);
                    if (rc >= 0) rc = nb_ldpc_statistic(x.code, x.out, x.iters, x.ok, a.batch, x.cw, x.cnt, x.st);
                    x.rc = rc;
                    if (rc < 0) x.err = ldpc_last_cuda_error();
                });
            }
            for (auto &t : th) t.join();
            next_frame += (unsigned long long)a.gpus * a.batch;
            memset(tot, 0, sizeof tot);
            for (auto &x : g) {
                if (x.rc < 0) { fprintf(stderr, "nb_ldpc_sim: gpu %d: %s %s\n", x.dev, ldpc_strerror(x.rc), x.err.c_str()); return 1; }
                cudaSetDevice(x.dev);
                int64_t c[6];
                cudaMemcpyAsync(c, x.cnt, sizeof c, cudaMemcpyDeviceToHost, x.st);
                cudaStreamSynchronize(x.st);
                for (int k = 0; k < 6; k++) tot[k] += c[k];
            }
            stop = (tot[1] >= a.least_errors && tot[0] >= a.least_frames) || (a.max_frames > 0 && tot[0] >= a.max_frames);
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const double nf = (double)tot[0];
        printf(" %.1f %8ld  %4ld  %6.4e  %6.4e  %.2f  %6.4esec  %9.3f\n", snr, (long)tot[0], (long)tot[1], tot[1] / nf,
               tot[2] / nf / info.N, tot[3] / nf, sec / nf, nf * kbits / sec / 1e6);
        fflush(stdout);
        if (!a.results.empty()) {  // the reference's row, byte for byte its format (Simulation.cpp:198,205)
            if (FILE *fp = fopen(a.results.c_str(), "a")) {
                fprintf(fp, " %.1f %8ld  %4ld  %6.4e  %6.4e  %.2f  %6.4esec\n", snr, (long)tot[0], (long)tot[1], tot[1] / nf,
                        tot[2] / nf / info.N, tot[3] / nf, sec / nf);
                fclose(fp);
            } else {
                fprintf(stderr, "nb_ldpc_sim: can not open file: %s\n", a.results.c_str());
            }
        }
    }
    for (auto &x : g) {
        cudaSetDevice(x.dev);
        cudaFree(x.x); cudaFree(x.out); cudaFree(x.iters); cudaFree(x.ok); cudaFree(x.cnt); cudaFree(x.cw);
        cudaStreamDestroy(x.st);
        nb_ldpc_free_code(x.code);
    }
    return 0;
}
