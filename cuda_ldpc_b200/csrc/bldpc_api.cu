// C-ABI entry points of the binary QC-LDPC decode path (include/ldpc_b200.h).
// ldpc_decode_batch replaces LDPC_Decoder_GPU (B/LDPC_Decoder.cuh:5, B/LDPC_Decoder.cu:23-164):
// same inputs (channel values, frames interleaved [N][F]) and the same output array
// D[(N+1)][F], but scratch is owned by the code handle instead of cudaMalloc/cudaFree per call
// (:38-88, :160-163), nothing is copied to the host per iteration (:135) and errors are
// returned, not printed.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.h"

namespace ldpcb {

static std::mutex g_dev_mu;

static int ensure_device(const ldpc_code *cc)
{
    ldpc_code *c = const_cast<ldpc_code *>(cc);
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (c->device >= 0) return LDPC_OK;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        return LDPC_ERR_NO_DEVICE;
    }
    int dev = 0, major = 0, sms = 0;
    LDPC_CUDA_TRY(cudaGetDevice(&dev));
    LDPC_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    LDPC_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (major != 10) return LDPC_ERR_NO_DEVICE;  // the library carries sm_100a code only
    LDPC_CUDA_TRY(cudaStreamCreateWithFlags(&c->pipe_stream[0], cudaStreamNonBlocking));
    LDPC_CUDA_TRY(cudaStreamCreateWithFlags(&c->pipe_stream[1], cudaStreamNonBlocking));
    LDPC_CUDA_TRY(cudaEventCreateWithFlags(&c->last_use, cudaEventDisableTiming));
    c->device = dev;
    c->num_sms = sms;
    return LDPC_OK;
}

// any dtype/layout -> fp32 [N][F]
__global__ void __launch_bounds__(256)
to_f32_nf_kernel(const void *__restrict__ in, float *__restrict__ out, int dtype, int layout, int N, int F)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)N * F) return;
    const int n = (int)(tid / F), f = (int)(tid % F);
    const size_t o = (layout == LDPC_LAYOUT_NF) ? (size_t)tid : (size_t)f * N + n;
    float v;
    if (dtype == LDPC_DTYPE_FP32)
        v = reinterpret_cast<const float *>(in)[o];
    else if (dtype == LDPC_DTYPE_FP16)
        v = __half2float(reinterpret_cast<const __half *>(in)[o]);
    else
        v = (float)reinterpret_cast<const signed char *>(in)[o];
    out[tid] = v;
}

// hard[N][F] (+ ok[F]) -> requested output format
__global__ void __launch_bounds__(256)
emit_hard_kernel(const unsigned char *__restrict__ hard, const int *__restrict__ ok, void *__restrict__ out,
                 int out_format, int layout, int N, int F)
{
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (out_format == LDPC_OUT_BITPACK) {
        const int W = (N + 31) / 32;
        if (tid >= (long long)W * F) return;
        const int f = (int)(tid % F), w = (int)(tid / F);
        unsigned bits = 0;
        for (int b = 0; b < 32; b++) {
            const int n = w * 32 + b;
            if (n < N && hard[(size_t)n * F + f]) bits |= 1u << b;
        }
        reinterpret_cast<unsigned *>(out)[(size_t)f * W + w] = bits;
        return;
    }
    if (tid >= (long long)(N + 1) * F) return;
    const int n = (int)(tid / F), f = (int)(tid % F);
    if (out_format == LDPC_OUT_INT32_REF) {
        reinterpret_cast<int *>(out)[tid] = (n < N) ? (int)hard[tid] : ok[f];
    } else if (n < N) {
        const size_t o = (layout == LDPC_LAYOUT_NF) ? (size_t)tid : (size_t)f * N + n;
        reinterpret_cast<unsigned char *>(out)[o] = hard[tid];
    }
}

static size_t dtype_bytes(int dt) { return dt == LDPC_DTYPE_FP32 ? 4 : (dt == LDPC_DTYPE_FP16 ? 2 : 1); }

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// Host buffers, layered int8: the batch is cut into chunks of 2 groups per SM and the chunks alternate
// between two internal streams, so the H2D copy of chunk k+1 and the D2H copy of chunk k-1 run under
// the decode of chunk k (the reference copies the whole batch, decodes, and copies N*F ints back
// EVERY iteration, B/Simulation.cu:138, B/LDPC_Decoder.cu:135).  A [N][F] chunk is a strided column
// block (cudaMemcpy2DAsync), a [F][N] chunk is contiguous.  Pinned host memory (inputs AND result
// buffers) is needed for real overlap; pageable memory still works (the copies then serialise).
extern "C" void ldpcb_host_pack_nf(const float *y, size_t ldF, int N, int f0, int fc, float scale, signed char *out,
                                   int threads);  // host_pack.cc
extern "C" int ldpcb_host_cpus(void);

// ldpc_decode_opts_t::host_pack_threads -> threads of the host quantiser (0 = copy the fp32 values)
static int pack_threads_of(const ldpc_decode_opts_t *o)
{
    if (o->llr_dtype != LDPC_DTYPE_FP32 || o->layout != LDPC_LAYOUT_NF) return 0;
    if (o->host_pack_threads > 0) return o->host_pack_threads;
    if (o->host_pack_threads < 0) return 0;
    // auto: the quantiser is worth running once it reads host DRAM about as fast as PCIe moves it (~55 GB/s), which
    // takes about a dozen cores; more threads than that only contend for the same DRAM (24-CPU box: 12 threads 10.4,
    // 24 threads 9.9 Gbit/s, profiles/r02_e2e.txt); fewer CPUs in this process' affinity mask -> plain copy
    const int cpus = ldpcb_host_cpus();
    return cpus >= 12 ? 12 : 0;
}

// pinned staging slots of the host-pack path, grown on demand
static int ensure_pack_slots(const ldpc_code *cc, size_t bytes)
{
    ldpc_code *c = const_cast<ldpc_code *>(cc);
    std::lock_guard<std::mutex> lk(g_dev_mu);
    for (int i = 0; i < 2; i++)
        if (!c->pack_ev[i]) LDPC_CUDA_TRY(cudaEventCreateWithFlags(&c->pack_ev[i], cudaEventDisableTiming));
    if (c->pack_host_bytes >= bytes) return LDPC_OK;
    for (int i = 0; i < 2; i++) {
        if (c->pack_host[i]) LDPC_CUDA_TRY(cudaFreeHost(c->pack_host[i]));
        c->pack_host[i] = nullptr;
        LDPC_CUDA_TRY(cudaHostAlloc(&c->pack_host[i], bytes, cudaHostAllocDefault));
    }
    c->pack_host_bytes = bytes;
    return LDPC_OK;
}

static int decode_host_pipelined(const ldpc_code *c, const void *llr, void *hard_bits, int iters,
                                 const ldpc_decode_opts_t *o, int Fc)
{
    const int F = o->batch, N = c->N;
    // Hybrid feed (fp32 [N][F] input with host threads available): chunks alternate between two ways into the GPU
    // that run CONCURRENTLY — "pack" chunks are quantised to int8 by host threads (the kernel's own rule, a quarter
    // of the PCIe bytes, same decode bit for bit) while the DMA engine copies the previous "copy" chunk as fp32.
    // Host cores reading DRAM and PCIe each sustain 50-60 GB/s on the B200 box; neither alone beats the other
    // (profiles/r02_e2e.txt), together they add up to 70-80 GB/s.  copy chunk = 1/2 of a pack chunk by default (swept 0-150 %): the
    // DMA engine also moves the pack chunks' int8 bytes (tuning: LDPC_B200_HYBRID_COPY_PCT, 0 = pack every chunk).
    const bool f16 = (o->msg_dtype == LDPC_DTYPE_FP16);  // fp16 message mode: fp32 copy chunks only (an int8 pack
                                                          // would round what this mode exists not to round)
    const int pack_threads = f16 ? 0 : pack_threads_of(o);
    const bool pack = pack_threads >= 1;
    int copy_pct = 50;
    if (const char *e = getenv("LDPC_B200_HYBRID_COPY_PCT")) copy_pct = atoi(e) < 0 ? 0 : (atoi(e) > 400 ? 400 : atoi(e));
    // A share that adapts to the host (copy chunk grown while the DMA engine is found idle after a pack chunk, shrunk
    // while it is backed up) was tried: +5-14 % on a box whose host memory was loaded, -1-4 % on a quiet one, no gain on
    // a heavily loaded one where even the fixed share falls to the plain copy's 6.3-6.8 Gbit/s (profiles/r02_e2e.txt).
    // Kept fixed.
    ldpc_code *cm = const_cast<ldpc_code *>(c);
    int Fcopy = pack ? ((int)((long long)Fc * copy_pct / 100) & ~3) : Fc;  // frames of a copy chunk (whole groups)
    if (pack) {
        int rcp = ensure_pack_slots(c, (size_t)N * Fc);
        if (rcp != LDPC_OK) return rcp;
    }
    const size_t esz = dtype_bytes(o->llr_dtype);
    const int W = (N + 31) / 32;
    const int Fmax = Fc > Fcopy ? Fc : Fcopy;
    size_t rec_bytes = 0;
    int rc = f16 ? layered_f16_scratch_bytes(c, Fmax, &rec_bytes) : layered_i8_scratch_bytes(c, Fmax, o->beta_num, &rec_bytes);
    if (rc != LDPC_OK) return rc;
    const size_t in_b = align_up((size_t)N * Fmax * esz), out_b = align_up(ldpc_out_bytes(c, Fmax, o->out_format));
    const size_t fl_b = align_up((size_t)Fmax * 4), slot = in_b + out_b + 2 * fl_b + align_up(rec_bytes);
    unsigned char *base = nullptr;
    rc = ensure_scratch(c, 2 * slot, reinterpret_cast<void **>(&base));
    if (rc != LDPC_OK) return rc;
    cudaStream_t user = reinterpret_cast<cudaStream_t>(o->stream);
    LDPC_CUDA_TRY(cudaStreamSynchronize(user));  // work queued before this call is complete
    int launches = 0, npack = 0;
    size_t h2d_bytes = 0;
    rc = LDPC_OK;
    for (int f0 = 0, k = 0; f0 < F && rc == LDPC_OK; k++) {
        // even k: pack chunk (stream 0, slot 0), odd k: copy chunk (stream 1, slot 1); without host threads every chunk
        // is a copy chunk and the two streams / slots simply alternate
        const bool packed = pack && ((k & 1) == 0 || Fcopy == 0);
        const int want = packed ? Fc : (pack ? Fcopy : Fc);
        const int fc = (F - f0 < want) ? F - f0 : want;
        cudaStream_t st = c->pipe_stream[k & 1];
        unsigned char *sl = base + (size_t)(k & 1) * slot;
        unsigned char *d_in = sl, *d_out = sl + in_b;
        int *d_it = reinterpret_cast<int *>(sl + in_b + out_b), *d_ok = reinterpret_cast<int *>(sl + in_b + out_b + fl_b);
        const unsigned char *h_in = reinterpret_cast<const unsigned char *>(llr);
        cudaError_t ce = cudaSuccess;
        if (packed) {
            // the staging buffer was last read by the copy of the pack chunk before last: wait for it, quantise this
            // chunk into it (the GPU and the DMA engine are busy with the previous chunks meanwhile)
            const int sb = npack & 1;
            if (npack >= 2) ce = cudaEventSynchronize(c->pack_ev[sb]);
            signed char *stage = reinterpret_cast<signed char *>(c->pack_host[sb]);
            if (ce == cudaSuccess) {
                ldpcb_host_pack_nf(reinterpret_cast<const float *>(llr), (size_t)F, N, f0, fc, o->llr_scale, stage,
                                   pack_threads);
                ce = cudaMemcpyAsync(d_in, stage, (size_t)N * fc, cudaMemcpyHostToDevice, st);
                h2d_bytes += (size_t)N * fc;
            }
            if (ce == cudaSuccess) ce = cudaEventRecord(c->pack_ev[sb], st);
            npack++;
        } else if (o->layout == LDPC_LAYOUT_NF)
            ce = cudaMemcpy2DAsync(d_in, (size_t)fc * esz, h_in + (size_t)f0 * esz, (size_t)F * esz, (size_t)fc * esz, N,
                                   cudaMemcpyHostToDevice, st);
        else
            ce = cudaMemcpyAsync(d_in, h_in + (size_t)f0 * N * esz, (size_t)fc * N * esz, cudaMemcpyHostToDevice, st);
        if (!packed) h2d_bytes += (size_t)N * fc * esz;
        if (ce != cudaSuccess) {
            set_cuda_error(ce, "host chunk upload");
            rc = LDPC_ERR_CUDA;
            break;
        }
        LayeredArgs a;
        a.llr = d_in;
        a.llr_dtype = packed ? LDPC_DTYPE_INT8 : o->llr_dtype;
        a.layout = o->layout;
        a.F = fc;
        a.iters = iters;
        a.exit_mode = o->early_exit;
        a.out_format = o->out_format;
        a.scale = o->llr_scale;
        a.msg_max = o->msg_max;
        a.beta_num = o->beta_num;
        a.beta_shift = o->beta_shift;
        a.alpha = o->alpha;
        a.out = d_out;
        a.iters_out = d_it;
        a.ok_out = d_ok;
        a.dbg_app = nullptr;
        a.dbg_rec = nullptr;
        a.scratch = sl + in_b + out_b + 2 * fl_b;
        a.scratch_bytes = align_up(rec_bytes);
        a.ch_sigma = 0.0f;
        a.ch_seed = a.ch_first = 0;
        a.ch_cw = nullptr;
        rc = f16 ? launch_layered_f16(c, a, st, &launches) : launch_layered_i8(c, a, st, &launches);
        if (rc != LDPC_OK) break;
        unsigned char *h_out = reinterpret_cast<unsigned char *>(hard_bits);
        if (o->out_format == LDPC_OUT_BITPACK) {
            ce = cudaMemcpyAsync(h_out + (size_t)f0 * W * 4, d_out, (size_t)fc * W * 4, cudaMemcpyDeviceToHost, st);
        } else if (o->out_format == LDPC_OUT_U8 && o->layout == LDPC_LAYOUT_FN) {
            ce = cudaMemcpyAsync(h_out + (size_t)f0 * N, d_out, (size_t)fc * N, cudaMemcpyDeviceToHost, st);
        } else {
            const size_t es = (o->out_format == LDPC_OUT_INT32_REF) ? 4 : 1;
            const int rows = (o->out_format == LDPC_OUT_INT32_REF) ? N + 1 : N;
            ce = cudaMemcpy2DAsync(h_out + (size_t)f0 * es, (size_t)F * es, d_out, (size_t)fc * es, (size_t)fc * es, rows,
                                   cudaMemcpyDeviceToHost, st);
        }
        if (ce == cudaSuccess && o->iters_out)
            ce = cudaMemcpyAsync(o->iters_out + f0, d_it, (size_t)fc * 4, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess && o->ok_out)
            ce = cudaMemcpyAsync(o->ok_out + f0, d_ok, (size_t)fc * 4, cudaMemcpyDeviceToHost, st);
        if (ce != cudaSuccess) {
            set_cuda_error(ce, "host chunk download");
            rc = LDPC_ERR_CUDA;
            break;
        }
        f0 += fc;
    }
    cm->last_h2d_bytes = h2d_bytes;
    // every exit path waits for both internal streams: async copies still target the caller's buffers and the
    // pinned staging slots
    const cudaError_t s0 = cudaStreamSynchronize(c->pipe_stream[0]), s1 = cudaStreamSynchronize(c->pipe_stream[1]);
    if (rc != LDPC_OK) return rc;
    if (s0 != cudaSuccess || s1 != cudaSuccess) {
        set_cuda_error(s0 != cudaSuccess ? s0 : s1, "cudaStreamSynchronize(pipe_stream)");
        return LDPC_ERR_CUDA;
    }
    return launches;
}

}  // namespace ldpcb

using namespace ldpcb;

extern "C" void ldpc_decode_opts_default(ldpc_decode_opts_t *o)
{
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->struct_size = (int)sizeof(*o);
    o->batch = 0;
    o->layout = LDPC_LAYOUT_NF;
    o->llr_dtype = LDPC_DTYPE_FP32;
    o->mem_space = LDPC_MEM_HOST;
    o->schedule = LDPC_SCHED_FLOODING;  // the reference's schedule and arithmetic
    o->msg_dtype = LDPC_DTYPE_FP32;
    o->early_exit = LDPC_EXIT_NONE;
    o->out_format = LDPC_OUT_INT32_REF;
    o->alpha = 1.0f;
    o->llr_scale = 8.0f;
    o->msg_max = 127;
    o->beta_num = 0;
    o->beta_shift = 0;
}

extern "C" size_t ldpc_last_h2d_bytes(const ldpc_code_t *c) { return c ? c->last_h2d_bytes : 0; }

extern "C" size_t ldpc_out_bytes(const ldpc_code_t *c, int F, int fmt)
{
    if (!c || F <= 0) return 0;
    switch (fmt) {
        case LDPC_OUT_INT32_REF: return (size_t)(c->N + 1) * F * sizeof(int);
        case LDPC_OUT_U8: return (size_t)c->N * F;
        case LDPC_OUT_BITPACK: return (size_t)((c->N + 31) / 32) * F * sizeof(unsigned);
        default: return 0;
    }
}

static int decode_batch_locked(const ldpc_code_t *c, const void *llr, void *hard_bits, int iters,
                               const ldpc_decode_opts_t *o);

// restores the caller's current device on every exit path
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    int enter(int want)
    {
        LDPC_CUDA_TRY(cudaGetDevice(&prev));
        if (prev != want) {
            LDPC_CUDA_TRY(cudaSetDevice(want));
            switched = true;
        }
        return LDPC_OK;
    }
    ~DeviceGuard()
    {
        if (switched) (void)cudaSetDevice(prev);
    }
};

extern "C" int ldpc_wave_frames(const ldpc_code_t *cc, int msg_dtype)
{
    if (!cc) return LDPC_ERR_ARG;
    if (msg_dtype != LDPC_DTYPE_INT8 && msg_dtype != LDPC_DTYPE_FP16) return LDPC_ERR_UNSUPPORTED;
    ldpc_code *c = const_cast<ldpc_code *>(cc);
    std::lock_guard<std::mutex> lk(c->call_mu);
    int rc = ensure_device(c);
    if (rc != LDPC_OK) return rc;
    DeviceGuard dg;
    rc = dg.enter(c->device);
    if (rc != LDPC_OK) return rc;
    int frames = 0;
    rc = msg_dtype == LDPC_DTYPE_FP16 ? layered_f16_wave_frames(c, &frames) : layered_i8_wave_frames(c, &frames);
    return rc != LDPC_OK ? rc : frames;
}

extern "C" int ldpc_decode_batch(const ldpc_code_t *cc, const void *llr, void *hard_bits, int iters,
                                 const ldpc_decode_opts_t *o)
{
    if (!cc || !hard_bits || !o || iters <= 0) return LDPC_ERR_ARG;
    ldpc_code *c = const_cast<ldpc_code *>(cc);
    // calls on one handle serialise (one scratch arena, common.h): host threads on call_mu, streams on last_use
    std::lock_guard<std::mutex> lk(c->call_mu);
    int rc = ensure_device(c);  // binds the handle to the device current at its first decode
    if (rc != LDPC_OK) return rc;
    DeviceGuard dg;             // later calls run on that device whatever the calling thread's current device is
    rc = dg.enter(c->device);
    if (rc != LDPC_OK) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(o->stream);
    if (c->last_use_valid) LDPC_CUDA_TRY(cudaStreamWaitEvent(st, c->last_use, 0));
    rc = decode_batch_locked(c, llr, hard_bits, iters, o);
    if (cudaEventRecord(c->last_use, st) == cudaSuccess) c->last_use_valid = true;
    return rc;
}

static int decode_batch_locked(const ldpc_code_t *c, const void *llr, void *hard_bits, int iters,
                               const ldpc_decode_opts_t *o)
{
    if (!c || !hard_bits || !o || iters <= 0) return LDPC_ERR_ARG;
    const bool fused_channel = (o->llr_dtype == LDPC_DTYPE_CHANNEL);
    if (!llr && !fused_channel) return LDPC_ERR_ARG;
    if (o->struct_size != (int)sizeof(ldpc_decode_opts_t)) return LDPC_ERR_ARG;
    const int F = o->batch;
    if (F <= 0) return LDPC_ERR_ARG;
    if (o->layout != LDPC_LAYOUT_NF && o->layout != LDPC_LAYOUT_FN) return LDPC_ERR_ARG;
    if (o->llr_dtype < 0 || o->llr_dtype > 3 || o->out_format < 0 || o->out_format > 2) return LDPC_ERR_ARG;
    if (o->early_exit < 0 || o->early_exit > 2) return LDPC_ERR_ARG;
    if (o->schedule != LDPC_SCHED_FLOODING && o->schedule != LDPC_SCHED_LAYERED) return LDPC_ERR_ARG;
    // "flooding" below = the streaming fp32 paths (messages in HBM, hard bits emitted by a conversion
    // kernel): flooding fp32 and layered fp32.  The layered int8 path has its own I/O inside the kernel.
    const bool layered_f32 = (o->schedule == LDPC_SCHED_LAYERED && o->msg_dtype == LDPC_DTYPE_FP32);
    const bool flooding = (o->schedule == LDPC_SCHED_FLOODING) || layered_f32;
    if (flooding && o->msg_dtype != LDPC_DTYPE_FP32) return LDPC_ERR_UNSUPPORTED;
    const bool layered_f16 = (!flooding && o->msg_dtype == LDPC_DTYPE_FP16);
    if (!flooding && o->msg_dtype != LDPC_DTYPE_INT8 && !layered_f16) return LDPC_ERR_UNSUPPORTED;
    if (o->schedule == LDPC_SCHED_LAYERED && o->early_exit == LDPC_EXIT_GENIE) return LDPC_ERR_UNSUPPORTED;
    if (fused_channel && (flooding || !(o->channel_sigma >= 0.0f))) return LDPC_ERR_UNSUPPORTED;  // layered int8 / fp16 only
    int rc = ensure_device(c);
    if (rc != LDPC_OK) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(o->stream);
    const bool host = (o->mem_space == LDPC_MEM_HOST);
    // large host batches: chunked copy/compute overlap (2 groups per SM and chunk).  Measured on B200 for
    // the bench workload (1.45 GB of fp32 per call, pinned): 37.9 ms unchunked -> 28.5 ms with 8 chunks,
    // i.e. PCIe-bound (55 GB/s) instead of copy + decode in series (profiles/r01_h2d_probe.txt).
    if (host && !flooding && !fused_channel && !o->debug_app && !o->debug_msgs) {  // layered int8 / fp16
        int Fc = 4 * c->num_sms * 2;
        if (const char *e = getenv("LDPC_B200_CHUNK_GROUPS")) Fc = 4 * c->num_sms * (atoi(e) > 0 ? atoi(e) : 2);  // tuning
        if (F >= 2 * Fc) return decode_host_pipelined(c, llr, hard_bits, iters, o, Fc);
    }
    const size_t in_bytes = fused_channel ? 0 : (size_t)c->N * F * dtype_bytes(o->llr_dtype);
    const size_t out_bytes = ldpc_out_bytes(c, F, o->out_format);
    const size_t dbg_app_bytes =
        o->debug_app ? (size_t)c->N * F * (layered_f32 ? 4 : (flooding ? 0 : (layered_f16 ? 2 : 1))) : 0;
    const size_t dbg_msg_bytes =
        o->debug_msgs ? (size_t)c->M * c->dc_max * F * (flooding ? 4 : (layered_f16 ? 2 : 1)) : 0;  // fp32 / fp16 / int8 c2v messages

    // ---- scratch carve-up (one arena per handle; decode calls on one handle serialise on it)
    size_t need = 0;
    const size_t o_in = need;      need += host ? align_up(in_bytes) : 0;
    const size_t o_out = need;     need += host ? align_up(out_bytes) : 0;
    const size_t o_it = need;      need += align_up((size_t)F * 4);
    const size_t o_ok = need;      need += align_up((size_t)F * 4);
    const size_t o_dapp = need;    need += host ? align_up(dbg_app_bytes) : 0;
    const size_t o_dmsg = need;    need += (host && !flooding) ? align_up(dbg_msg_bytes) : 0;
    size_t o_y = 0, o_msgs = 0, o_hard = 0, o_flags = 0, o_app = 0;
    const bool need_conv = flooding && (o->llr_dtype != LDPC_DTYPE_FP32 || o->layout != LDPC_LAYOUT_NF);
    if (flooding) {
        o_y = need;     need += need_conv ? align_up((size_t)c->N * F * 4) : 0;
        o_msgs = need;  need += align_up((size_t)c->M * c->dc_max * F * 4);
        o_hard = need;  need += align_up((size_t)c->N * F);
        o_flags = need; need += align_up((size_t)(2 * F + 2) * 4);
        o_app = need;   need += layered_f32 ? align_up((size_t)c->N * F * 4) : 0;
    }
    const size_t o_layer = need;  // the layered kernel sizes its own slice behind this offset
    unsigned char *base = nullptr;
    size_t layered_extra = 0;
    if (!flooding) {
        rc = layered_f16 ? layered_f16_scratch_bytes(c, F, &layered_extra)
                         : layered_i8_scratch_bytes(c, F, o->beta_num, &layered_extra);
        if (rc != LDPC_OK) return rc;
    }
    rc = ensure_scratch(c, need + layered_extra, reinterpret_cast<void **>(&base));
    if (rc != LDPC_OK) return rc;

    const void *d_in = llr;
    void *d_out = hard_bits;
    int *d_it = o->iters_out, *d_ok = o->ok_out;
    void *d_dapp = o->debug_app, *d_dmsg = o->debug_msgs;
    if (host) {
        d_in = base + o_in;
        d_out = base + o_out;
        if (in_bytes) LDPC_CUDA_TRY(cudaMemcpyAsync(base + o_in, llr, in_bytes, cudaMemcpyHostToDevice, st));
        const_cast<ldpc_code *>(c)->last_h2d_bytes = in_bytes;
        d_it = reinterpret_cast<int *>(base + o_it);
        d_ok = reinterpret_cast<int *>(base + o_ok);
        if (o->debug_app) d_dapp = base + o_dapp;
        if (o->debug_msgs) d_dmsg = base + o_dmsg;
    } else {
        if (!d_it) d_it = reinterpret_cast<int *>(base + o_it);
        if (!d_ok) d_ok = reinterpret_cast<int *>(base + o_ok);
    }
    int launches = 0;
    if (flooding) {
        const float *y = reinterpret_cast<const float *>(d_in);
        if (need_conv) {
            float *yc = reinterpret_cast<float *>(base + o_y);
            const long long n = (long long)c->N * F;
            to_f32_nf_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_in, yc, o->llr_dtype, o->layout, c->N, F);
            launches++;
            y = yc;
        }
        float *msgs = reinterpret_cast<float *>(base + o_msgs);
        unsigned char *hard = base + o_hard;
        if (layered_f32)
            rc = launch_layered_f32_nf(c, y, F, iters, o->early_exit, o->alpha, reinterpret_cast<float *>(base + o_app),
                                       msgs, hard, d_it, d_ok, reinterpret_cast<int *>(base + o_flags), st, &launches, host);
        else
            rc = launch_flooding_fp32(c, y, F, iters, o->early_exit, hard, d_it, d_ok, msgs,
                                      reinterpret_cast<int *>(base + o_flags), st, &launches, host);
        if (rc != LDPC_OK) return rc;
        if (layered_f32 && o->debug_app)
            LDPC_CUDA_TRY(cudaMemcpyAsync(o->debug_app, base + o_app, dbg_app_bytes,
                                          host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
        const long long n = (o->out_format == LDPC_OUT_BITPACK) ? (long long)((c->N + 31) / 32) * F
                                                                : (long long)(c->N + 1) * F;
        emit_hard_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hard, d_ok, d_out, o->out_format, o->layout,
                                                                    c->N, F);
        launches++;
        LDPC_CUDA_TRY(cudaGetLastError());
        if (o->debug_msgs)
            LDPC_CUDA_TRY(cudaMemcpyAsync(o->debug_msgs, msgs, dbg_msg_bytes,
                                          host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
    } else {
        LayeredArgs a;
        a.llr = d_in;
        a.llr_dtype = o->llr_dtype;
        a.layout = o->layout;
        a.F = F;
        a.iters = iters;
        a.exit_mode = o->early_exit;
        a.out_format = o->out_format;
        a.scale = o->llr_scale;
        a.msg_max = o->msg_max;
        a.beta_num = o->beta_num;
        a.beta_shift = o->beta_shift;
        a.alpha = o->alpha;
        a.out = d_out;
        a.iters_out = d_it;
        a.ok_out = d_ok;
        a.dbg_app = d_dapp;
        a.dbg_rec = d_dmsg;
        a.scratch = base + o_layer;
        a.scratch_bytes = layered_extra;
        a.ch_sigma = o->channel_sigma;
        a.ch_seed = o->channel_seed;
        a.ch_first = o->channel_first_frame;
        a.ch_cw = o->channel_codeword;
        rc = layered_f16 ? launch_layered_f16(c, a, st, &launches) : launch_layered_i8(c, a, st, &launches);
        if (rc != LDPC_OK) return rc;
        if (host && o->debug_msgs)
            LDPC_CUDA_TRY(cudaMemcpyAsync(o->debug_msgs, d_dmsg, dbg_msg_bytes, cudaMemcpyDeviceToHost, st));
    }
    if (host) {
        LDPC_CUDA_TRY(cudaMemcpyAsync(hard_bits, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
        if (o->iters_out)
            LDPC_CUDA_TRY(cudaMemcpyAsync(o->iters_out, d_it, (size_t)F * 4, cudaMemcpyDeviceToHost, st));
        if (o->ok_out) LDPC_CUDA_TRY(cudaMemcpyAsync(o->ok_out, d_ok, (size_t)F * 4, cudaMemcpyDeviceToHost, st));
        if (o->debug_app && dbg_app_bytes && !flooding)
            LDPC_CUDA_TRY(cudaMemcpyAsync(o->debug_app, d_dapp, dbg_app_bytes, cudaMemcpyDeviceToHost, st));
        LDPC_CUDA_TRY(cudaStreamSynchronize(st));
    }
    return launches;
}
