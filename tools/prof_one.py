#!/usr/bin/env python
"""Runs one decode configuration a few times (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuda_ldpc_b200 as m
name, F, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dt = m.DTYPE_FP16 if len(sys.argv) > 5 and sys.argv[5] == "fp16" else m.DTYPE_INT8
geo = {"C1": ("J4_L24_Z96_BlockH.txt", (0, 0, 0), 4.0), "C2": ("J15_L30_Z1280_BlockH.txt", (0, 0, 0), 2.0),
       "C3": ("PON_LDPC.txt", (12, 69, 256), 4.5)}[name]
code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", geo[0]), *geo[1])
y = 1.0 + m.sigma_from_snr(0, geo[2], code.rate) * torch.randn(code.N, F, device="cuda")
for _ in range(reps):
    r = code.decode(y, iters, schedule=m.SCHED_LAYERED, msg_dtype=dt, out_format=m.OUT_BITPACK, msg_max=31, beta_num=1, beta_shift=3)
torch.cuda.synchronize()
print("ok", float(r.ok.float().mean()))
