#!/usr/bin/env python
"""Developer timing of the fp32 modes (flooding = the reference's rules, layered fp32 = the float twin of the int8 kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cuda_ldpc_b200 as m
BL = os.path.join(m.DATA_DIR, "bldpc")
for name, path, geo, F, snr in [("C1", "J4_L24_Z96_BlockH.txt", (0, 0, 0), 16384, 3.0), ("C2", "J15_L30_Z1280_BlockH.txt", (0, 0, 0), 1024, 2.0),
                                ("C3", "PON_LDPC.txt", (12, 69, 256), 2048, 4.0)]:
    code = m.LdpcCode(os.path.join(BL, path), *geo)
    y = 1.0 + m.sigma_from_snr(0, snr, code.rate) * torch.randn(code.N, F, device="cuda")
    for label, kw in [("flooding fp32", dict()), ("layered fp32", dict(schedule=m.SCHED_LAYERED, msg_dtype=m.DTYPE_FP32))]:
        for _ in range(2): r = code.decode(y, 10, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): r = code.decode(y, 10, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"{name} {label:14s} F={F:6d} {ms:8.3f} ms {F * code.K / ms / 1e6:7.2f} info Gbit/s  launches {r.launches}  ok {float(r.ok.float().mean()):.3f}", flush=True)
