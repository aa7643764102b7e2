#!/usr/bin/env python
"""bench.py — decoded info Gbit/s at fixed iterations (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): binary QC-LDPC J15_L30_Z1280 (N=38400, K=19200, rate 1/2),
BPSK-AWGN synthetic frames (all-zero codeword, the reference's only mode), layered normalised
min-sum with int8 state, 10 iterations fixed, no early exit.  One step = one pass of the hot path
(ldpc_decode_batch) over one batch of F frames per GPU.  Under torchrun each rank decodes its own
shard of the frame range (weak scaling, no data-path collective); the time is the max over ranks.

Printed JSON line: value (device-resident I/O), e2e (same call with HOST buffers, H2D/D2H inside the
timed region), roofline (SURVEY §8d byte model over the kernel's CUDA-event time), cpu_baseline
(the C oracle on the host cores, bounded sample), clocks, gpu_launches.
--impl reference times the reference's own decoder sources compiled for the CPU (oracle/_ref).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CODE_FILE = "J15_L30_Z1280_BlockH.txt"
J, L, Z = 15, 30, 1280
N, K, M, E = L * Z, (L - J) * Z, J * Z, 147200
ITERS = 10
EBN0_DB = 2.0
# B_cw(it) = it*(2*E*w + 2*M*rec) + N*4 + N/8, int8: w = 1, rec = 4  -> 4 638 400 B at 10 it
B_CW = ITERS * (2 * E * 1 + 2 * M * 4) + N * 4 + N // 8
# dram__bytes_read.sum + dram__bytes_write.sum of one ldpc_layered_i8_kernel launch over 2368 frames, from the
# `ncu --set full` capture summarised in profiles/r01_ncu_layered_i8_summary.txt (1.49 GB + 3.29 GB): per frame
NCU_DRAM_BYTES_PER_FRAME = (1.486490e9 + 3.289583e9) / 2368
METRIC = "decoded info Gbit/s at fixed iters"
WORKLOAD = ("binary QC-LDPC J15_L30_Z1280 (N=38400, K=19200), BPSK-AWGN Eb/N0 2.0 dB, layered normalised "
            "min-sum (x0.875), int8 state, 10 iterations fixed, no early exit")


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.stop, self.index = [], False, index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([x.strip() for x in out.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------ CPU legs (oracle = checker / baseline only)

def cpu_baseline_port(frames=4096):
    """The C oracle (oracle/liboracle.so, OpenMP over frames) on a bounded sample of the workload."""
    import numpy as np
    lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    H = np.zeros(J * L, np.int32); Wc = np.zeros(J + 1, np.int32); Wv = np.zeros(L + 1, np.int32)
    path = os.path.join(ROOT, "cuda_ldpc_b200", "data", "bldpc", CODE_FILE)
    assert lib.orc_get_h(path.encode(), J, L, ip(H), ip(Wc), ip(Wv)) == 0
    lib.orc_layered_i8.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int), C.c_void_p, C.c_int, C.c_int, C.c_float,
                                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_void_p]
    cores = os.cpu_count() or 1
    frames = max(frames, 2 * cores)
    rng = np.random.default_rng(1)
    sigma = (0.5 / (0.5 * 10 ** (EBN0_DB / 10))) ** 0.5
    y = (1.0 + sigma * rng.standard_normal((N, frames))).astype(np.float32)
    D = np.zeros((N + 1) * frames, np.int32); it = np.zeros(frames, np.int32)
    t0 = time.perf_counter()
    rc = lib.orc_layered_i8(J, L, Z, ip(H), y.ctypes.data, frames, ITERS, 8.0, 31, 1, 3, 0, D.ctypes.data,
                            it.ctypes.data, None, None)
    dt = time.perf_counter() - t0
    assert rc == 0
    return {"value": frames * K / dt / 1e9, "unit": "Gbit/s", "cores": cores, "kind": "port",
            "sample": f"{frames} frames of the workload, oracle layered int8, {ITERS} it, OpenMP over frames, {dt:.2f} s"}


def _ref_worker(args):
    so, reps, seed = args
    import numpy as np
    lib = C.CDLL(so)
    lib.ref_init()
    g = np.zeros(10, np.int32)
    lib.ref_geometry(g.ctypes.data_as(C.POINTER(C.c_int)))
    n, f = int(g[3]), int(g[6])
    rng = np.random.default_rng(seed)
    sigma = (0.5 / (0.5 * 10 ** (EBN0_DB / 10))) ** 0.5
    y = (1.0 + sigma * rng.standard_normal((n, f))).astype(np.float32)
    D = np.zeros((n + 1) * f, np.int32)
    t0 = time.perf_counter()
    for _ in range(reps):
        lib.ref_decode(y.ctypes.data_as(C.POINTER(C.c_float)), D.ctypes.data_as(C.POINTER(C.c_int)))
    return f * reps, time.perf_counter() - t0


def run_reference(args):
    """The reference's own LDPC_Decoder_GPU + kernels (B/LDPC_Decoder.cu), compiled for the CPU by
    oracle/build_ref.sh, one process per host core, each step = one batch of 16 frames per core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    so = os.path.join(ROOT, "oracle", "_ref", "libbldpc_ref_C2.so")
    cores = os.cpu_count() or 1
    base = {"metric": METRIC, "unit": "Gbit/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference", "config": {"workload": WORKLOAD, "note": "reference decoder is flooding fp32 "
                                            "(its only mode), same code, same iterations"}}
    if not os.path.exists(so):
        cb = cpu_baseline_port()
        base.update({"value": cb["value"], "ms_per_step": None, "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base))
        return 0
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        if args.warmup:
            pool.map(_ref_worker, [(so, 1, 100 + i) for i in range(cores)])
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(so, args.steps, i) for i in range(cores)])
        dt = time.perf_counter() - t0
    frames = sum(r[0] for r in res)
    val = frames * K / dt / 1e9
    cb = {"value": val, "unit": "Gbit/s", "cores": cores, "kind": "reference",
          "sample": f"{frames} frames: reference LDPC_Decoder_GPU sources run on the CPU (shim), {cores} processes x "
                    f"{args.steps} batches of 16 frames, {ITERS} it flooding fp32"}
    base.update({"value": val, "ms_per_step": dt / max(args.steps, 1) * 1e3, "cpu_baseline": cb,
                 "e2e": {"value": val, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))
    return 0


# ------------------------------------------------------------------ our arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    import cuda_ldpc_b200 as m

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = "unset"
    try:  # run on (and allocate pinned host buffers from) the CPUs next to this GPU: the e2e copies stay NUMA-local
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        numa = f"{len(os.sched_getaffinity(0))} cpus"
    except Exception as e:  # affinity is an optimisation only
        numa = f"unavailable ({type(e).__name__})"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", CODE_FILE))
    F = args.frames  # per GPU; input 38400*F*4 B (1.45 GB at F=9472) is far larger than the 126 MB L2
    sigma = m.sigma_from_snr(0, EBN0_DB, code.rate)
    g = torch.Generator(device=dev).manual_seed(173 + rank)  # each rank its own shard of the frame range
    y = (1.0 + sigma * torch.randn(code.N, F, device=dev, generator=g)).contiguous()
    out = torch.empty(code.out_bytes(F, m.OUT_BITPACK), dtype=torch.uint8, device=dev)
    it_d = torch.empty(F, dtype=torch.int32, device=dev)
    ok_d = torch.empty(F, dtype=torch.int32, device=dev)
    kw = dict(schedule=m.SCHED_LAYERED, out_format=m.OUT_BITPACK, msg_max=args.msg_max, llr_scale=args.llr_scale,
              beta_num=args.beta_num, beta_shift=args.beta_shift)

    def step():
        return code.decode(y, ITERS, out=out, iters_out=it_d, ok_out=ok_d, **kw).launches

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks / throttle reasons are sampled from the warm-up to the end of the e2e region (nvidia-smi answers in
    # ~50 ms, the device-timed region alone lasts ~0.2 s), all of it under decode load
    clk = ClockSampler(local)
    clk.__enter__()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches = 0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        launches += step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = sum(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)) / args.steps  # one launch per step
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * F * args.steps * K / (total_ms * 1e-3) / 1e9
    ok_frac = float(ok_d.float().mean().item())

    # ---- e2e: the same C-ABI call with HOST buffers (pinned), H2D + D2H inside the timed region
    Fe = args.e2e_frames
    yh = y[:, :Fe].contiguous().cpu().pin_memory().numpy()
    e2e_steps = max(2, min(args.steps, 4))
    # pinned result buffers (hard bits, iterations, flags): the step's result is read back into them
    ho = torch.empty(code.out_bytes(Fe, m.OUT_BITPACK), dtype=torch.uint8).pin_memory().numpy()
    hi = torch.empty(Fe, dtype=torch.int32).pin_memory().numpy()
    hk = torch.empty(Fe, dtype=torch.int32).pin_memory().numpy()
    hkw = dict(kw, out=ho, iters_out=hi, ok_out=hk)
    code.decode(yh, ITERS, **hkw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r = code.decode(yh, ITERS, **hkw)
        launches_e2e = r.launches
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * Fe * e2e_steps * K / float(te.item()) / 1e9
    h2d = Fe * code.N * 4
    d2h = code.out_bytes(Fe, m.OUT_BITPACK) + 8 * Fe
    # same call with the host buffer in the [F][N] layout: frame chunks are contiguous, so the library
    # overlaps the H2D copy of chunk k+1 with the decode of chunk k (informational, not the headline)
    yh_fn = y[:, :Fe].t().contiguous().cpu().pin_memory().numpy()
    code.decode(yh_fn, ITERS, layout=m.LAYOUT_FN, **hkw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        code.decode(yh_fn, ITERS, layout=m.LAYOUT_FN, **hkw)
    torch.cuda.synchronize()
    e2e_fn_val = Fe * e2e_steps * K / (time.perf_counter() - t0) / 1e9
    del yh_fn
    # context for "PCIe-bound": the same call fed fp16 channel values (half the bytes; NOT the reference's input type)
    yh16 = torch.from_numpy(yh).to(torch.float16).pin_memory().numpy()
    code.decode(yh16, ITERS, **hkw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        code.decode(yh16, ITERS, **hkw)
    torch.cuda.synchronize()
    e2e_fp16_val = Fe * e2e_steps * K / (time.perf_counter() - t0) / 1e9
    del yh16
    # and with host-side int8 packing of the chunks (ldpc_decode_opts_t::host_pack_threads): a quarter of the PCIe
    # bytes, same bits, but it occupies the host cores — informational; only tried when this rank has >= 12 of them
    pack_threads = (os.cpu_count() or 1) // world
    e2e_pack_val = None
    if pack_threads >= 12:
        code.decode(yh, ITERS, host_pack_threads=pack_threads, **hkw)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            code.decode(yh, ITERS, host_pack_threads=pack_threads, **hkw)
        torch.cuda.synchronize()
        e2e_pack_val = Fe * e2e_steps * K / (time.perf_counter() - t0) / 1e9
    clk.__exit__()

    if rank == 0:
        peak, which = measured_peaks()
        achieved = F * B_CW / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": F, "msg_max": args.msg_max,
                       "llr_scale": args.llr_scale, "beta": [args.beta_num, args.beta_shift],
                       "input": "fp32 [N][F] channel values resident in HBM (1.45 GB per step at F=9472, larger than L2)",
                       "output": "bit-packed hard decisions + per-frame syndrome flag",
                       "converged_fraction": ok_frac},
            "e2e": {"value": e2e_val, "unit": "Gbit/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "frames_per_step": Fe, "steps": e2e_steps, "launches_per_step": launches_e2e,
                    "cpu_affinity": numa,
                    "layout": "[N][F] fp32 pinned host buffer (the reference's Channel_Out layout); the library cuts "
                              "the batch into chunks of 2 groups per SM on two streams (H2D / decode / D2H overlap)",
                    "rank0_value_with_FN_layout_chunked_overlap": e2e_fn_val,
                    "rank0_value_with_host_int8_packing": e2e_pack_val, "host_pack_threads": pack_threads,
                    "rank0_value_with_fp16_host_input": e2e_fp16_val},
            "gpu_launches": launches,
            "clocks": clk.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": args.traffic if args.traffic is not None else NCU_DRAM_BYTES_PER_FRAME * F,
                         "traffic_source": "ncu --set full at F=2368 (profiles/r01_ncu_layered_i8_summary.txt), scaled "
                                           "to this launch's frame count; mostly check records spilling from L2",
                         "peak_source": f"{which} (MEASURED_PEAKS.json, burst copy)",
                         "kernel": "ldpc_layered_i8_kernel", "kernel_ms": kern_ms,
                         "bytes_model": f"SURVEY 8(d): B_cw(10)={B_CW} B x {F} frames per launch (algorithmic bytes of a "
                                        "streaming layered decoder).  This design keeps the APP state in shared "
                                        "memory, so its real DRAM traffic is lower (traffic) and the pipe that binds is "
                                        "the SM issue rate, not HBM: see issue_bound",
                         "issue_bound": {"inst_per_edge_4frames": 57.5, "ipc_per_smsp": 0.71,
                                         "practical_ipc_peak_per_smsp": 0.80,
                                         "source": "profiles/r01_ncu_layered_i8_summary.txt, profiles/r01_pipe_ubench.txt"}},
        }
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))  # the CPU baseline uses every host core again
        except Exception:
            pass
        try:
            line["cpu_baseline"] = None if args.no_cpu_baseline else cpu_baseline_port()
        except Exception as e:  # the bench value never depends on the oracle
            line["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=148 * 4 * 16)
    ap.add_argument("--e2e-frames", type=int, default=148 * 4 * 16)
    ap.add_argument("--msg-max", type=int, default=31)
    ap.add_argument("--llr-scale", type=float, default=8.0)
    ap.add_argument("--beta-num", type=int, default=1)
    ap.add_argument("--beta-shift", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the oracle timing (profiling runs)")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per launch from the ncu capture")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
