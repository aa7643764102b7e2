#!/usr/bin/env python
"""Runs one NB decode configuration a few times (for ncu captures): prof_nb.py ems|tmm|ltmm|fftbp C4|C5 F"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cuda_ldpc_b200 as m
algo = {"ems": m.ALGO_EMS, "tmm": m.ALGO_TMM, "ltmm": m.ALGO_LAYERED_TMM, "fftbp": m.ALGO_FFT_BP}[sys.argv[1]]
cfg = sys.argv[2]; F = int(sys.argv[3])
NB = os.path.join(m.DATA_DIR, "nbldpc")
mat, cst, nqam, ebn0 = {"C4": ("LDPC_N576_K288_GF64_d1_exp.txt", "GRAY_64QAM.txt", 64, 10.0),
                        "C5": ("LDPC_N576_K480_GF256_exp.txt", "BPSK.txt", 2, 4.5)}[cfg]
code = m.NbLdpcCode(os.path.join(NB, mat), None, os.path.join(NB, "Constellation", cst), coef_is_exponent=True)
sigma = m.lib.nb_ldpc_sigma(code._h, 0, ebn0, 0)
per = code.in_elems(m.IN_BPSK if nqam == 2 else m.IN_QAM)
x = torch.empty(F * per, dtype=torch.float32, device="cuda")
m.lib.nb_ldpc_modulate_awgn(code._h, x.data_ptr(), F, sigma, 1, 0, None, torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    out, it, ok = code.decode(x.view(F, per), 20, algo=algo, in_kind=m.IN_BPSK if nqam == 2 else m.IN_QAM, sigma=sigma)
torch.cuda.synchronize()
print("FER", 1 - ok.float().mean().item(), "avg_it", it.float().mean().item())
