#!/usr/bin/env python
"""Developer timing loop (not the contract bench): times ldpc_decode_batch on device-resident
inputs for a few codes / modes with CUDA events."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuda_ldpc_b200 as m

BL = os.path.join(m.DATA_DIR, "bldpc")


def run(name, path, geo, F, iters, snr, reps=5, **kw):
    code = m.LdpcCode(os.path.join(BL, path), *geo)
    sigma = m.sigma_from_snr(0, snr, code.rate)
    y = 1.0 + sigma * torch.randn(code.N, F, device="cuda")
    out = torch.empty(code.out_bytes(F, kw.get("out_format", m.OUT_BITPACK)), dtype=torch.uint8, device="cuda")
    it = torch.empty(F, dtype=torch.int32, device="cuda"); ok = torch.empty(F, dtype=torch.int32, device="cuda")
    kw.setdefault("out_format", m.OUT_BITPACK)
    for _ in range(2):
        r = code.decode(y, iters, out=out, iters_out=it, ok_out=ok, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = code.decode(y, iters, out=out, iters_out=it, ok_out=ok, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbps = F * code.K / (ms * 1e-3) / 1e9
    print(f"{name:34s} F={F:7d} it={iters:2d} {ms:9.3f} ms  {gbps:8.2f} info Gbit/s  ok={ok.float().mean().item():.3f} "
          f"avg_it={it.float().mean().item():.2f} launches={r.launches}", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    L = dict(schedule=m.SCHED_LAYERED)
    run("C2 layered i8 fixed10", "J15_L30_Z1280_BlockH.txt", (0, 0, 0), 148 * 4 * 8, 10, 1.5, **L)
    run("C2 layered i8 fixed10 beta.75", "J15_L30_Z1280_BlockH.txt", (0, 0, 0), 148 * 4 * 8, 10, 1.5, beta_num=1, beta_shift=2, **L)
    run("C2 layered i8 synd", "J15_L30_Z1280_BlockH.txt", (0, 0, 0), 148 * 4 * 8, 10, 2.0, early_exit=2, **L)
    run("C1 layered i8 fixed10", "J4_L24_Z96_BlockH.txt", (0, 0, 0), 65536, 10, 4.0, **L)
    run("C3 layered i8 fixed10", "PON_LDPC.txt", (12, 69, 256), 16384, 10, 4.5, **L)
    run("C3 layered i8 synd50", "PON_LDPC.txt", (12, 69, 256), 16384, 50, 4.5, early_exit=2, **L)
    run("C1 flooding fp32 fixed10", "J4_L24_Z96_BlockH.txt", (0, 0, 0), 4096, 10, 4.0)
    run("C2 flooding fp32 fixed10", "J15_L30_Z1280_BlockH.txt", (0, 0, 0), 1024, 10, 1.5)
