#!/usr/bin/env python
"""Developer timing of the NB matrices no BASELINE config names (Tanner GF(16) N=9472, N96 GF(256)), BPSK."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cuda_ldpc_b200 as m
NB = os.path.join(m.DATA_DIR, "nbldpc")
for mat, exp, ebn0, F in [("Tanner_74_9_Z128_GF16.txt", False, 5.5, 512), ("LDPC_N96_K48_GF256_d1_exp.txt", True, 4.0, 8192)]:
    code = m.NbLdpcCode(os.path.join(NB, mat), None, os.path.join(NB, "Constellation", "BPSK.txt"), coef_is_exponent=exp)
    sigma = m.lib.nb_ldpc_sigma(code._h, 0, ebn0, 0)
    per = code.in_elems(m.IN_BPSK)
    x = torch.empty(F * per, dtype=torch.float32, device="cuda")
    m.lib.nb_ldpc_modulate_awgn(code._h, x.data_ptr(), F, sigma, 1, 0, None, torch.cuda.current_stream().cuda_stream)
    x = x.view(F, per)
    for name, algo in [("EMS(2,2)", m.ALGO_EMS), ("TMM", m.ALGO_TMM), ("layered TMM", m.ALGO_LAYERED_TMM), ("FFT-BP", m.ALGO_FFT_BP)]:
        for _ in range(2): out, it, ok = code.decode(x, 20, algo=algo, in_kind=m.IN_BPSK, sigma=sigma)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2): out, it, ok = code.decode(x, 20, algo=algo, in_kind=m.IN_BPSK, sigma=sigma)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        print(f"{mat:32s} {name:12s} F={F:5d} {ms:9.2f} ms {F * code.K_bits / ms / 1e3:9.2f} info Mbit/s  FER={1 - ok.float().mean().item():.3f} avg_it={it.float().mean().item():.2f}", flush=True)
