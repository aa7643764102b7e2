// ldpc_sim — the binary simulator's driver, re-stated over the C-ABI of libldpc_b200.so.
// Same loop as the reference (bldpc_实习/main.cu:106-157 SNR sweep, reseed + sigma per point;
// Simulation.cu:111-156 batch loop; Simulation.cu:245-285 Statistic and stop rule; the result row
// " SNR NTF NEF FER BER AverIT FER_F FER_A" of Simulation.cu:239,271,281), but configuration is
// run-time (the reference recompiles for every code / SNR range, define.cuh), the channel is
// generated on the device (Philox keyed by the global frame index), nothing but six counters
// crosses PCIe, and the frame range is sharded over --gpus devices (one host thread + stream each,
// counters summed on the host: codewords are independent, no collective is needed).
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
    std::string code;
    int J = 0, L = 0, Z = 0;
    double snr_start = 0, snr_stop = 3, snr_step = 0.5;
    int snrtype = 1;  // define.cuh:45 as committed: Es/N0
    int maxit = 50;   // define.cuh:35
    int batch = 4096; // define.cuh:60 Num_Frames_OneTime
    int schedule = LDPC_SCHED_FLOODING, msg = LDPC_DTYPE_FP32, exit_mode = LDPC_EXIT_SYNDROME;
    long least_errors = 50, least_frames = 10000, max_frames = 0;  // define.cuh:52-53
    int gpus = 1, msg_max = 31, beta_num = 0, beta_shift = 0, encode = 0;
    float llr_scale = 8.0f;
    unsigned long long seed = 173;  // define.cuh:40-42
};

static void usage()
{
    printf("usage: ldpc_sim --code FILE [--J j --L l --Z z] [--snr a b step] [--snrtype 0|1] [--maxit n] [--batch F]\n"
           "       [--layered | --layered-fp16] [--exit none|genie|syndrome] [--least-errors n] [--least-frames n] [--max-frames n]\n"
           "       [--gpus g] [--seed s] [--msg-max m] [--beta num shift] [--llr-scale s] [--encode]\n");
}

struct Gpu {
    std::string err;  // ldpc_last_cuda_error() is per thread: captured inside the worker
    int dev;
    ldpc_code_t *code = nullptr;
    cudaStream_t st = nullptr;
    float *y = nullptr;
    void *out = nullptr;
    int *iters = nullptr, *ok = nullptr;
    int64_t *cnt = nullptr;
    uint8_t *cw = nullptr;
    int rc = 0;
};

int main(int argc, char **argv)
{
    Args a;
    for (int i = 1; i < argc; i++) {
        std::string s = argv[i];
        auto next = [&](int k = 1) { if (i + k >= argc) { usage(); exit(2); } return argv[i + k]; };
        if (s == "--code") a.code = next(), i++;
        else if (s == "--J") a.J = atoi(next()), i++;
        else if (s == "--L") a.L = atoi(next()), i++;
        else if (s == "--Z") a.Z = atoi(next()), i++;
        else if (s == "--snr") a.snr_start = atof(next(1)), a.snr_stop = atof(next(2)), a.snr_step = atof(next(3)), i += 3;
        else if (s == "--snrtype") a.snrtype = atoi(next()), i++;
        else if (s == "--maxit") a.maxit = atoi(next()), i++;
        else if (s == "--batch") a.batch = atoi(next()), i++;
        else if (s == "--layered") a.schedule = LDPC_SCHED_LAYERED, a.msg = LDPC_DTYPE_INT8;
        else if (s == "--layered-fp16") a.schedule = LDPC_SCHED_LAYERED, a.msg = LDPC_DTYPE_FP16;  // fp16 message mode
        else if (s == "--exit") { std::string m = next(); i++; a.exit_mode = m == "none" ? LDPC_EXIT_NONE : m == "genie" ? LDPC_EXIT_GENIE : LDPC_EXIT_SYNDROME; }
        else if (s == "--least-errors") a.least_errors = atol(next()), i++;
        else if (s == "--least-frames") a.least_frames = atol(next()), i++;
        else if (s == "--max-frames") a.max_frames = atol(next()), i++;
        else if (s == "--gpus") a.gpus = atoi(next()), i++;
        else if (s == "--seed") a.seed = strtoull(next(), nullptr, 10), i++;
        else if (s == "--msg-max") a.msg_max = atoi(next()), i++;
        else if (s == "--beta") a.beta_num = atoi(next(1)), a.beta_shift = atoi(next(2)), i += 2;
        else if (s == "--llr-scale") a.llr_scale = (float)atof(next()), i++;
        else if (s == "--encode") a.encode = 1;
        else { usage(); return 2; }
    }
    if (a.code.empty()) { usage(); return 2; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { fprintf(stderr, "ldpc_sim: %s\n", ldpc_strerror(LDPC_ERR_NO_DEVICE)); return 1; }
    if (a.gpus > ndev) a.gpus = ndev;
    std::vector<Gpu> g(a.gpus);
    ldpc_code_info_t info;
    std::vector<uint8_t> cw_host;
    for (int d = 0; d < a.gpus; d++) {
        g[d].dev = d;
        cudaSetDevice(d);
        int rc = ldpc_load_code(a.code.c_str(), a.J, a.L, a.Z, &g[d].code);
        if (rc != LDPC_OK) { fprintf(stderr, "ldpc_sim: %s: %s\n", a.code.c_str(), ldpc_strerror(rc)); return 1; }
        ldpc_code_info(g[d].code, &info);
        if (d == 0 && a.encode) {
            cw_host.resize(info.N);
            std::vector<uint8_t> u(info.K);
            unsigned x = (unsigned)a.seed * 2654435761u + 1u;
            for (auto &b : u) { x = x * 1664525u + 1013904223u; b = (x >> 16) & 1; }
            rc = ldpc_encode(g[0].code, u.data(), cw_host.data());
            if (rc != LDPC_OK) { fprintf(stderr, "ldpc_sim: encode: %s\n", ldpc_strerror(rc)); return 1; }
        }
        cudaStreamCreate(&g[d].st);
        cudaMalloc(&g[d].y, (size_t)info.N * a.batch * sizeof(float));
        cudaMalloc(&g[d].out, ldpc_out_bytes(g[d].code, a.batch, LDPC_OUT_INT32_REF));
        cudaMalloc(&g[d].iters, a.batch * sizeof(int));
        cudaMalloc(&g[d].ok, a.batch * sizeof(int));
        cudaMalloc(&g[d].cnt, 6 * sizeof(int64_t));
        if (a.encode) {
            cudaMalloc(&g[d].cw, info.N);
            cudaMemcpy(g[d].cw, cw_host.data(), info.N, cudaMemcpyHostToDevice);
        }
    }
    const float rate = (float)info.K / (float)info.N;
    // banner, after WriteLogo (Simulation.cu:176-240)
    printf("*******************Binary LDPC Simulation (ldpc_b200)*******************\n");
    printf("* %s\n* Message bits' length of LDPC is %d\n* Parity bits' length of LDPC is %d\n", ldpc_version(), info.K, info.M);
    printf("* CodeWord length of LDPC is %d\n* H's row is divided into %d blocks, and column divided into %d blocks. Dimension Z is %d\n",
           info.N, info.J, info.L, info.Z);
    printf("* The encoding rate for current LDPC is %f\n* Maximum iterations for LDPC_decoder is %d\n", rate, a.maxit);
    printf("* schedule %s, messages %s, exit %s, %d frames per batch, %d GPU(s), codeword %s\n",
           a.schedule == LDPC_SCHED_LAYERED ? "layered" : "flooding", a.msg == LDPC_DTYPE_INT8 ? "int8" : (a.msg == LDPC_DTYPE_FP16 ? "fp16" : "fp32"),
           a.exit_mode == LDPC_EXIT_NONE ? "none" : a.exit_mode == LDPC_EXIT_GENIE ? "genie" : "syndrome", a.batch,
           a.gpus, a.encode ? "random (encoded)" : "all-zero");
    printf("* The type of SNR is %s\n", a.snrtype == 0 ? "Eb/No" : "Es/No");
    printf(" SNR   %5s   %5s   %7s    %7s     %7s  %7s   %7s   %9s\n", "NTF", "NEF", "FER", "BER", "AverIT", "FER_F", "FER_A", "infoGbit/s");
    const int npts = (int)floor((a.snr_stop - a.snr_start) / a.snr_step + 1e-9) + 1;  // start + i*step
    for (int pt = 0; pt < npts; pt++) {
        const float snr = (float)(a.snr_start + pt * a.snr_step);
        const float sigma = ldpc_sigma(a.snrtype, snr, rate);
        ldpc_sim_counters_t tot;
        memset(&tot, 0, sizeof(tot));
        for (auto &x : g) { cudaSetDevice(x.dev); cudaMemsetAsync(x.cnt, 0, 6 * sizeof(int64_t), x.st); }
        unsigned long long next_frame = 0;  // per-SNR reset = the reference's reseed (main.cu:117-119)
        auto t0 = std::chrono::steady_clock::now();
        bool stop = false;
        while (!stop) {
            std::vector<std::thread> th;
            for (int d = 0; d < a.gpus; d++) {
                const unsigned long long first = next_frame + (unsigned long long)d * a.batch;
                th.emplace_back([&, d, first]() {
                    Gpu &x = g[d];
                    cudaSetDevice(x.dev);
                    ldpc_decode_opts_t o;
                    ldpc_decode_opts_default(&o);
                    o.batch = a.batch; o.mem_space = LDPC_MEM_DEVICE; o.schedule = a.schedule; o.msg_dtype = a.msg;
                    o.early_exit = a.exit_mode; o.llr_scale = a.llr_scale;
                    o.msg_max = a.msg_max; o.beta_num = a.beta_num; o.beta_shift = a.beta_shift;
                    o.iters_out = x.iters; o.ok_out = x.ok; o.stream = x.st;
                    int rc;
                    if (a.schedule == LDPC_SCHED_LAYERED) {
                        // channel fused into the decoder's load phase, bit-packed decisions: no F*N buffer anywhere
                        o.llr_dtype = LDPC_DTYPE_CHANNEL; o.out_format = LDPC_OUT_BITPACK;
                        o.channel_sigma = sigma; o.channel_seed = a.seed; o.channel_first_frame = first; o.channel_codeword = x.cw;
                        rc = ldpc_decode_batch(x.code, nullptr, x.out, a.maxit, &o);
                        if (rc >= 0) rc = ldpc_statistic(x.code, x.out, LDPC_OUT_BITPACK, x.ok, x.iters, a.batch, info.K, x.cw, x.cnt, x.st);
                    } else {
                        o.out_format = LDPC_OUT_INT32_REF;
                        rc = ldpc_awgn_bpsk(x.code, x.y, a.batch, LDPC_LAYOUT_NF, sigma, a.seed, first, x.cw, x.st);
                        if (rc >= 0) rc = ldpc_decode_batch(x.code, x.y, x.out, a.maxit, &o);
                        if (rc >= 0) rc = ldpc_statistic(x.code, x.out, LDPC_OUT_INT32_REF, nullptr, x.iters, a.batch, info.K, x.cw, x.cnt, x.st);
                    }
                    x.rc = rc;
                    if (rc < 0) x.err = ldpc_last_cuda_error();
                });
            }
            for (auto &t : th) t.join();
            next_frame += (unsigned long long)a.gpus * a.batch;
            memset(&tot, 0, sizeof(tot));
            for (auto &x : g) {
                if (x.rc < 0) { fprintf(stderr, "ldpc_sim: gpu %d: %s %s\n", x.dev, ldpc_strerror(x.rc), x.err.c_str()); return 1; }
                cudaSetDevice(x.dev);
                int64_t c[6];
                cudaMemcpyAsync(c, x.cnt, sizeof c, cudaMemcpyDeviceToHost, x.st);
                cudaStreamSynchronize(x.st);
                tot.num_Frames += c[0]; tot.num_Error_Frames += c[1]; tot.num_Error_Bits += c[2];
                tot.Total_Iteration += c[3]; tot.num_False_Frames += c[4]; tot.num_Alarm_Frames += c[5];
            }
            // stop rule, Simulation.cu:273
            stop = (tot.num_Error_Frames >= a.least_errors && tot.num_Frames >= a.least_frames) ||
                   (a.max_frames > 0 && tot.num_Frames >= a.max_frames);
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const double nf = (double)tot.num_Frames;
        printf(" %.1f %8ld  %4ld  %6.4e  %6.4e  %.2f  %6.4e %6.4e  %9.3f\n", snr, (long)tot.num_Frames, (long)tot.num_Error_Frames,
               tot.num_Error_Frames / nf, tot.num_Error_Bits / nf / info.K, tot.Total_Iteration / nf, tot.num_False_Frames / nf,
               tot.num_Alarm_Frames / nf, nf * info.K / sec / 1e9);
        fflush(stdout);
    }
    for (auto &x : g) {
        cudaSetDevice(x.dev);
        cudaFree(x.y); cudaFree(x.out); cudaFree(x.iters); cudaFree(x.ok); cudaFree(x.cnt); cudaFree(x.cw);
        cudaStreamDestroy(x.st);
        ldpc_free_code(x.code);
    }
    printf("\ntask finish\n");
    return 0;
}
