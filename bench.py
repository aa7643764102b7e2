#!/usr/bin/env python
"""bench.py — decoded info Gbit/s at fixed iterations (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): binary QC-LDPC J15_L30_Z1280 (N=38400, K=19200, rate 1/2),
BPSK-AWGN synthetic frames (all-zero codeword, the reference's only mode), layered normalised
min-sum with int8 state, 10 iterations fixed, no early exit.  One step = one pass of the hot path
(ldpc_decode_batch) over one batch of F frames per GPU.  Under torchrun each rank decodes its own
shard of the frame range (weak scaling, no data-path collective); the time is the max over ranks.

Printed JSON line: value (device-resident I/O), e2e (same call with HOST buffers, H2D/D2H inside the
timed region, results compared byte for byte with the device-resident ones: e2e.parity_ok), sustained
(>= 3 s back to back), roofline (SURVEY §8d byte model over the kernel's CUDA-event time; traffic from the
tracked ncu summary), clocks, gpu_launches, and at N = 1 also configs (C1 / C3 / flooding fp32),
reference_gpu (the reference's own GPU decoder on this box) and cpu_baseline (the C oracle on the host
cores, bounded sample).
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

def cpu_baseline_port(frames=16384):
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

# The three binary configs of BASELINE.json: geometry, edges, largest check degree, frames per launch,
# operating point (snrtype, dB).  B_cw(it) = it*(2*E*w + 2*M*rec) + N*4 + N/8 with w = 1, rec = 2*w + 1 + ceil(dc_max/8).
CONFIGS = {
    "C1": dict(file="J4_L24_Z96_BlockH.txt", geo=(4, 24, 96), E=7680, dc_max=20, F=65536, snr=(1, 3.0)),
    "C2": dict(file=CODE_FILE, geo=(J, L, Z), E=E, dc_max=8, F=148 * 4 * 16, snr=(0, EBN0_DB)),
    "C3": dict(file="PON_LDPC.txt", geo=(12, 69, 256), E=70400, dc_max=23, F=16384, snr=(1, 4.5)),
}


def b_cw_int8(cfg, iters):
    j, l, z = cfg["geo"]
    n, m_ = l * z, j * z
    rec = 2 + 1 + (cfg["dc_max"] + 7) // 8
    return iters * (2 * cfg["E"] + 2 * m_ * rec) + n * 4 + n // 8


def b_cw_fp16(cfg, iters):
    """SURVEY 8d's fp16 rows: w = 2, rec = 2*w + 1 + ceil(dc_max/8)"""
    j, l, z = cfg["geo"]
    n, m_ = l * z, j * z
    rec = 4 + 1 + (cfg["dc_max"] + 7) // 8
    return iters * (2 * cfg["E"] * 2 + 2 * m_ * rec) + n * 4 + n // 8


def b_cw_flooding_fp32(cfg, iters):
    """the reference's own layout: (4E + 2N) * 4 bytes per codeword-iteration"""
    j, l, z = cfg["geo"]
    return iters * (4 * cfg["E"] + 2 * l * z) * 4


def ncu_traffic_per_frame():
    """dram bytes per frame of the dominant kernel, from the tracked ncu summary written by tools/ncu_summary.py --json"""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_layered_i8.json")) as f:
            d = json.load(f)["C2"]
        return (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / d["frames"], "profiles/r02_ncu_layered_i8.json"
    except Exception:
        return None, None


def ncu_issue_pipes():
    """an on-chip-resident design is bound by the SM issue pipes, not HBM; the pipe utilisation of the
    dominant kernel comes from the same tracked ncu capture as roofline.traffic (not measured live: ncu is a profiler)"""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_layered_i8.json")) as f:
            d = json.load(f)["C2"]
        edges_iters = E * 10  # edge updates of one codeword in 10 iterations; a thread serves one check row of 4 codewords
        return {"source": "profiles/r02_ncu_layered_i8.json (ncu --set full, C2, %d frames)" % d["frames"],
                "alu_pipe_pct": d["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"],
                "fma_pipe_pct": d["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"],
                "fp16_pipe_pct": d["sm__inst_executed_pipe_fma_type_fp16.avg.pct_of_peak_sustained_active"],
                "lsu_pipe_pct": d["sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"],
                "issue_slots_busy": d["smsp__issue_active.avg.per_cycle_active"],
                "thread_inst_per_edge_per_4_codewords": 32 * d["smsp__inst_executed.sum"] / (edges_iters * d["frames"] / 4)}
    except Exception:
        return None


def timed_launches(fn, steps, torch):
    """CUDA events on the launching stream around every step; returns (total ms, mean ms per step)"""
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    n = 0
    for i in range(steps):
        n += fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[-1]), sum(ev[i].elapsed_time(ev[i + 1]) for i in range(steps)) / steps, n


def other_configs(m, torch, dev, peak, kw):
    """C1 / C3 layered int8 at 10 fixed iterations, C3 throughput mode, and the same-arithmetic figure (flooding fp32,
    the reference's schedule, layout and update rule) — N = 1 only, informational beside the headline."""
    res = []
    for name in ("C1", "C3"):
        cfg = CONFIGS[name]
        code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", cfg["file"]), *cfg["geo"])
        wave = code.wave_frames()  # whole waves of resident CTAs (ldpc_wave_frames): every CTA gets the same number of groups
        F = max(1, round(cfg["F"] / wave)) * wave
        sigma = m.sigma_from_snr(cfg["snr"][0], cfg["snr"][1], code.rate)
        y = (1.0 + sigma * torch.randn(code.N, F, device=dev)).contiguous()
        out = torch.empty(code.out_bytes(F, m.OUT_BITPACK), dtype=torch.uint8, device=dev)
        it_d = torch.empty(F, dtype=torch.int32, device=dev)
        ok_d = torch.empty(F, dtype=torch.int32, device=dev)
        step = lambda **x: code.decode(y, x.pop("iters", ITERS), out=out, iters_out=it_d, ok_out=ok_d, **dict(kw, **x)).launches
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        _, ms, _ = timed_launches(step, 10, torch)
        ach = F * b_cw_int8(cfg, ITERS) / (ms * 1e-3) / 1e9
        res.append({"config": name, "mode": "layered int8, 10 iterations fixed", "frames_per_launch": F, "frames_per_wave": wave,
                    "value": F * code.K / (ms * 1e-3) / 1e9, "unit": "Gbit/s", "kernel_ms": ms,
                    "roofline_frac": ach / peak, "b_cw_bytes": b_cw_int8(cfg, ITERS)})
        if name == "C3":  # BASELINE.json configs[2]: throughput mode = syndrome early termination, max 50 iterations
            step2 = lambda: step(iters=50, early_exit=m.EXIT_SYNDROME)
            for _ in range(3):
                step2()
            torch.cuda.synchronize()
            _, ms2, _ = timed_launches(step2, 10, torch)
            res.append({"config": "C3", "mode": "layered int8, syndrome early termination, max 50 iterations, Es/N0 4.5 dB",
                        "frames_per_launch": F, "value": F * code.K / (ms2 * 1e-3) / 1e9, "unit": "Gbit/s",
                        "kernel_ms": ms2, "avg_iterations": float(it_d.float().mean().item()),
                        "converged_fraction": float(ok_d.float().mean().item())})
        del y, out
    for name in ("C2", "C1", "C3"):  # fp16 message mode (★g2): same workload, half2 state, 2 codewords per CTA
        cfg = CONFIGS[name]
        code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", cfg["file"]), *cfg["geo"])
        wave = code.wave_frames(m.DTYPE_FP16)
        F = max(1, round(cfg["F"] / 2 / wave)) * wave
        sigma = m.sigma_from_snr(cfg["snr"][0], cfg["snr"][1], code.rate)
        y = (1.0 + sigma * torch.randn(code.N, F, device=dev)).contiguous()
        out = torch.empty(code.out_bytes(F, m.OUT_BITPACK), dtype=torch.uint8, device=dev)
        it_d = torch.empty(F, dtype=torch.int32, device=dev)
        ok_d = torch.empty(F, dtype=torch.int32, device=dev)
        step = lambda: code.decode(y, ITERS, out=out, iters_out=it_d, ok_out=ok_d, **dict(kw, msg_dtype=m.DTYPE_FP16)).launches
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        _, ms, _ = timed_launches(step, 10, torch)
        ach = F * b_cw_fp16(cfg, ITERS) / (ms * 1e-3) / 1e9
        res.append({"config": name, "mode": "layered fp16 messages, 10 iterations fixed", "frames_per_launch": F, "frames_per_wave": wave,
                    "value": F * code.K / (ms * 1e-3) / 1e9, "unit": "Gbit/s", "kernel_ms": ms,
                    "roofline_frac": ach / peak, "b_cw_bytes": b_cw_fp16(cfg, ITERS),
                    "note": "above 1: SURVEY 8d's fp16 model streams the APP values through HBM; they live in shared memory here"})
        del y, out
    for name, F in (("C1", 4096), ("C2", 1024)):  # flooding fp32: the reference's arithmetic, F = 4096 is its batch
        cfg = CONFIGS[name]
        code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", cfg["file"]), *cfg["geo"])
        sigma = m.sigma_from_snr(cfg["snr"][0], cfg["snr"][1], code.rate)
        y = (1.0 + sigma * torch.randn(code.N, F, device=dev)).contiguous()
        out = torch.empty(code.out_bytes(F, m.OUT_INT32_REF), dtype=torch.uint8, device=dev)
        it_d = torch.empty(F, dtype=torch.int32, device=dev)
        ok_d = torch.empty(F, dtype=torch.int32, device=dev)
        step = lambda: code.decode(y, ITERS, out=out, iters_out=it_d, ok_out=ok_d).launches
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        _, ms, nl = timed_launches(step, 5, torch)
        ach = F * b_cw_flooding_fp32(cfg, ITERS) / (ms * 1e-3) / 1e9
        res.append({"config": name, "mode": "flooding fp32 un-normalised min-sum, 10 iterations (the reference's rules and "
                    "Memory_RQ layout: same arithmetic as the reference arm)", "frames_per_launch": F,
                    "value": F * code.K / (ms * 1e-3) / 1e9, "unit": "Gbit/s", "ms_per_step": ms, "launches_per_step": nl // 5,
                    "roofline_frac": ach / peak, "bytes_model": "(4E + 2N) * 4 B per codeword-iteration"})
        del y, out
    return res


def reference_gpu():
    """The reference's own GPU decoder on this box (baseline/_ref/bldpc_gpu_C2_time, built from /root/reference by
    baseline/build_all.sh): live if the binary travelled here, else the tracked measurement of the same command."""
    exe = os.path.join(ROOT, "baseline", "_ref", "bldpc_gpu_C2_time")
    try:
        if os.path.exists(exe):
            p = subprocess.run([exe, "time", str(EBN0_DB), "2"], capture_output=True, text=True, timeout=240)
            for ln in p.stdout.splitlines():
                if ln.startswith("{"):
                    d = json.loads(ln)
                    d["source"] = "live: baseline/_ref/bldpc_gpu_C2_time"
                    return d
        with open(os.path.join(ROOT, "profiles", "r02_reference_gpu.jsonl")) as f:
            for ln in f:
                d = json.loads(ln)
                if d.get("binary") == "C2_time":
                    d["source"] = "profiles/r02_reference_gpu.jsonl (recorded on a B200 of this pool)"
                    return d
    except Exception as e:
        return {"error": repr(e)}
    return None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import cuda_ldpc_b200 as m
    from cuda_ldpc_b200 import sim as simlib

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

    def allmax(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # clocks / throttle reasons are sampled from the warm-up to the end of the e2e region (nvidia-smi answers in
    # ~50 ms, the K-step device-timed region alone lasts ~0.15 s), all of it under decode load
    clk = ClockSampler(local)
    clk.__enter__()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    total_ms, kern_ms, launches = timed_launches(step, args.steps, torch)
    barrier()
    total_ms = allmax(total_ms)
    value = world * F * args.steps * K / (total_ms * 1e-3) / 1e9
    ok_frac = float(ok_d.float().mean().item())

    # ---- sustained: the same step back to back for >= 3 s (does the kernel hold its clock?)
    sus_clk = ClockSampler(local)
    sus_clk.__enter__()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_sus = max(int(args.sustain_s * 1e3 / kern_ms), 1)
    e0.record()
    for _ in range(n_sus):
        step()
    e1.record()
    torch.cuda.synchronize()
    sus_ms = allmax(e0.elapsed_time(e1))
    sus_clk.__exit__()
    sustained = {"value": world * F * n_sus * K / (sus_ms * 1e-3) / 1e9, "unit": "Gbit/s", "seconds": sus_ms * 1e-3,
                 "steps": n_sus, "clocks": sus_clk.summary()}
    dev_res = (out.clone(), it_d.clone(), ok_d.clone())  # device-resident result of this input: the e2e parity reference

    # ---- e2e: the same C-ABI call with HOST buffers (pinned), H2D + D2H inside the timed region
    Fe = F
    yh = y.cpu().pin_memory().numpy()
    e2e_steps = max(2, min(args.steps, 6))
    ho = torch.empty(code.out_bytes(Fe, m.OUT_BITPACK), dtype=torch.uint8).pin_memory().numpy()
    hi = torch.empty(Fe, dtype=torch.int32).pin_memory().numpy()
    hk = torch.empty(Fe, dtype=torch.int32).pin_memory().numpy()
    hkw = dict(kw, out=ho, iters_out=hi, ok_out=hk)
    cpus = len(os.sched_getaffinity(0))
    # host quantiser: the library's automatic choice when this process has the host to itself.  With several ranks on
    # one box the host's DRAM is the shared resource and a plain fp32 DMA copy costs it one read per byte where the
    # quantiser costs a read, a write and the DMA's read (measured at N = 2: 10.3 vs 12.0 Gbit/s): off.
    pack = 0 if world == 1 else -1

    def host_parity():
        return bool((torch.from_numpy(ho).to(dev) == dev_res[0]).all().item() and
                    (torch.from_numpy(hi).to(dev) == dev_res[1]).all().item() and
                    (torch.from_numpy(hk).to(dev) == dev_res[2]).all().item())

    def time_host(n, **extra):
        for _ in range(2):  # warm-up: staging buffers, host thread team, scratch arena
            code.decode(yh, ITERS, **dict(hkw, **extra))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            r = code.decode(yh, ITERS, **dict(hkw, **extra))
        torch.cuda.synchronize()
        return time.perf_counter() - t0, r.launches

    packed = pack > 0 or (pack == 0 and cpus >= 12)  # the library's own rule (csrc/bldpc_api.cu pack_threads_of)

    e2e_s, launches_e2e = time_host(e2e_steps, host_pack_threads=pack)
    h2d_e2e = int(m.lib.ldpc_last_h2d_bytes(code.handle))  # what the library uploaded in the last step (int8 + fp32 chunks)
    parity_ok = host_parity()
    e2e_val = world * Fe * e2e_steps * K / allmax(e2e_s) / 1e9
    h2d = Fe * code.N * 4
    d2h = code.out_bytes(Fe, m.OUT_BITPACK) + 8 * Fe
    ho[:] = 0
    copy_s, _ = time_host(e2e_steps, host_pack_threads=-1)  # the same call with the quantiser off: fp32 over PCIe
    parity_ok = parity_ok and host_parity()
    e2e_copy_val = world * Fe * e2e_steps * K / allmax(copy_s) / 1e9

    # ---- the other two end-to-end shapes of the reference's loop
    # (b) the reference's actual interface (B/LDPC_Decoder.cuh:5): Channel_Out is a DEVICE pointer, D comes back to the host
    def step_dev_in_host_out():
        n = code.decode(y, ITERS, out=out, iters_out=it_d, ok_out=ok_d, **kw).launches
        torch.from_numpy(ho).copy_(out, non_blocking=True)
        torch.from_numpy(hi).copy_(it_d, non_blocking=True)
        torch.from_numpy(hk).copy_(ok_d, non_blocking=True)
        return n
    step_dev_in_host_out()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_dev_in_host_out()
    torch.cuda.synchronize()
    e2e_b = world * Fe * e2e_steps * K / allmax(time.perf_counter() - t0) / 1e9
    parity_ok = parity_ok and host_parity()
    # (d) a caller that already holds 8-bit channel values (llr_dtype = LDPC_DTYPE_INT8, the kernel's own quantiser rule
    # applied upstream): a quarter of the PCIe bytes, no host pass over the data inside the call
    yq = torch.clamp(torch.round(y * args.llr_scale), -127, 127).to(torch.int8)
    yq_h = torch.empty(yq.shape, dtype=torch.int8).pin_memory()
    yq_h.copy_(yq)
    torch.cuda.synchronize()
    del yq
    yq_n = yq_h.numpy()
    ho[:] = 0
    for _ in range(2):
        code.decode(yq_n, ITERS, **hkw)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        code.decode(yq_n, ITERS, **hkw)
    torch.cuda.synchronize()
    e2e_d = world * Fe * e2e_steps * K / allmax(time.perf_counter() - t0) / 1e9
    parity_ok = parity_ok and host_parity()
    # (c) the simulation-loop shape (B/Simulation.cu:111-156): no input at all — the channel is generated inside the
    # kernel (LDPC_DTYPE_CHANNEL), Statistic runs on the device, 48 bytes of counters come back per batch
    runner = simlib.CudaBatchRunner(code, F, maxit=ITERS, early_exit=m.EXIT_NONE, msg_max=args.msg_max,
                                    llr_scale=args.llr_scale, beta_num=args.beta_num, beta_shift=args.beta_shift)
    runner.run(sigma, rank * F)
    runner.counters()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        runner.run(sigma, (i * world + rank) * F)
        cnt = runner.counters()  # D2H of the six int64 counters + sync
    e2e_c = world * Fe * e2e_steps * K / allmax(time.perf_counter() - t0) / 1e9
    clk.__exit__()

    if rank == 0:
        peak, which = measured_peaks()
        achieved = F * B_CW / (kern_ms * 1e-3) / 1e9
        tpf, tsrc = ncu_traffic_per_frame()
        line = {
            "metric": METRIC, "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": F, "msg_max": args.msg_max,
                       "llr_scale": args.llr_scale, "beta": [args.beta_num, args.beta_shift],
                       "input": "fp32 [N][F] channel values resident in HBM (1.45 GB per step at F=9472, larger than L2)",
                       "output": "bit-packed hard decisions + per-frame syndrome flag",
                       "converged_fraction": ok_frac},
            "e2e": {"value": e2e_val, "unit": "Gbit/s", "h2d_bytes_per_step": h2d_e2e,
                    "d2h_bytes_per_step": d2h, "frames_per_step": Fe, "steps": e2e_steps,
                    "launches_per_step": launches_e2e, "cpu_affinity": numa, "parity_ok": parity_ok,
                    "parity_check": "hard bits, iteration counts and flags of every host-buffer shape below == the "
                                    "device-resident result of the same frames, byte for byte",
                    "layout": "[N][F] fp32 pinned host buffer (the reference's Channel_Out layout) through "
                              "ldpc_decode_batch(mem_space = HOST): chunks of 2 groups per SM on two streams; each chunk is "
                              "quantised to int8 by the library's host threads (the kernel's own rule, bit-identical) while "
                              "the previous chunk is copied and decoded" if packed else
                              "[N][F] fp32 pinned host buffer copied as fp32 (too few host cores per rank for the quantiser)",
                    "host_pack_threads": ("auto (12)" if packed else "auto (off: < 12 cpus)") if pack == 0 else pack, "host_cpus": cpus,
                    "host_bytes_read_per_step": Fe * code.N * 4,
                    "value_with_fp32_copy_over_pcie": e2e_copy_val,
                    "shape_device_in_host_out": {"value": e2e_b, "unit": "Gbit/s", "what": "Channel_Out on the device (the "
                                                 "reference's own interface, B/LDPC_Decoder.cuh:5), hard decisions + "
                                                 "iterations + flags read back to pinned host memory every step",
                                                 "d2h_bytes_per_step": d2h},
                    "shape_host_int8_in": {"value": e2e_d, "unit": "Gbit/s", "what": "pinned host int8 [N][F] input (caller "
                                           "quantised with the kernel's rule; results equal the fp32 path's byte for "
                                           "byte), same outputs to the host", "h2d_bytes_per_step": Fe * code.N,
                                           "d2h_bytes_per_step": d2h},
                    "shape_sim_loop": {"value": e2e_c, "unit": "Gbit/s", "what": "fused device channel (LDPC_DTYPE_CHANNEL) -> "
                                       "decode -> ldpc_statistic -> 48 bytes of counters to the host per batch "
                                       "(B/Simulation.cu:111-156 without its F*N buffers)", "d2h_bytes_per_step": 48,
                                       "counters_last_batch": [int(v) for v in cnt]}},
            "gpu_launches": launches,
            "clocks": clk.summary(),
            "sustained": sustained,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": args.traffic if args.traffic is not None else (tpf * F if tpf else None),
                         "traffic_source": tsrc and f"{tsrc}: ncu --set full of this kernel, dram bytes per frame x {F} frames "
                                                    "(mostly c2v message blocks spilling from L2)",
                         "peak_source": f"{which} (MEASURED_PEAKS.json, burst copy)",
                         "kernel": "ldpc_layered_i8_kernel<8>", "kernel_ms": kern_ms,
                         "issue_pipes": ncu_issue_pipes(),
                         "bytes_model": f"SURVEY 8(d): B_cw(10)={B_CW} B x {F} frames per launch (algorithmic bytes of a "
                                        "streaming layered decoder).  This design keeps the APP state in shared "
                                        "memory, so its real DRAM traffic is lower (traffic) and the pipes that bind are "
                                        "the SM's ALU / FMA issue pipes: profiles/r02_ncu_layered_i8_summary.txt"},
        }
        if world == 1:
            try:
                line["configs"] = [{"config": "C2", "mode": "layered int8, 10 iterations fixed (the headline)",
                                    "frames_per_launch": F, "value": value, "unit": "Gbit/s", "kernel_ms": kern_ms,
                                    "roofline_frac": achieved / peak, "b_cw_bytes": B_CW}] + other_configs(m, torch, dev, peak, kw)
            except Exception as e:
                line["configs"] = {"error": repr(e)}
            line["reference_gpu"] = reference_gpu()
            try:
                os.sched_setaffinity(0, range(os.cpu_count() or 1))  # the CPU baseline uses every host core again
            except Exception:
                pass
            try:
                line["cpu_baseline"] = None if args.no_cpu_baseline else cpu_baseline_port()
            except Exception as e:  # the bench value never depends on the oracle
                line["cpu_baseline"] = {"error": repr(e)}
        else:
            # rank 0's OpenMP team would fight the other ranks' spin-waits (round 1: 30x too slow): N = 1 only
            line["cpu_baseline"] = None
        print(json.dumps(line))
        if not parity_ok:
            print("e2e parity check FAILED: host-buffer results differ from the device-resident results", file=sys.stderr)
    if world > 1:
        dist.destroy_process_group()
    return 0 if parity_ok else 3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=148 * 4 * 16)
    ap.add_argument("--sustain-s", type=float, default=3.0, help="length of the back-to-back sustained region")
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
