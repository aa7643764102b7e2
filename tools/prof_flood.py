#!/usr/bin/env python
"""Runs the flooding fp32 mode a few times (for ncu captures): prof_flood.py C1|C2|C3 F iters"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuda_ldpc_b200 as m
name, F, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
geo = {"C1": ("J4_L24_Z96_BlockH.txt", (0, 0, 0), 4.0), "C2": ("J15_L30_Z1280_BlockH.txt", (0, 0, 0), 2.0),
       "C3": ("PON_LDPC.txt", (12, 69, 256), 4.5)}[name]
code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", geo[0]), *geo[1])
y = 1.0 + m.sigma_from_snr(0, geo[2], code.rate) * torch.randn(code.N, F, device="cuda")
for _ in range(2):
    r = code.decode(y, iters)
torch.cuda.synchronize()
print("ok", float(r.ok.float().mean()))
