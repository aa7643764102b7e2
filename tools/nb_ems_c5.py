#!/usr/bin/env python
"""C5 (GF(256), N = 576, K = 480 symbols, BPSK, Eb/N0 4.5 dB) with the EMS(2,2) decoder: frames/s and info Mbit/s
(tools/nb_bench.py times the TMM / FFT-BP decoders on C5; EMS over GF(256) is the slow one)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuda_ldpc_b200 as m
F = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
NB = os.path.join(m.DATA_DIR, "nbldpc")
code = m.NbLdpcCode(os.path.join(NB, "LDPC_N576_K480_GF256_exp.txt"), None, os.path.join(NB, "Constellation", "BPSK.txt"),
                    coef_is_exponent=True)
sigma = m.lib.nb_ldpc_sigma(code._h, 0, 4.5, 0)
per = code.in_elems(m.IN_BPSK)
x = torch.empty(F * per, dtype=torch.float32, device="cuda")
m.lib.nb_ldpc_modulate_awgn(code._h, x.data_ptr(), F, sigma, 1, 0, None, torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    out, it, ok = code.decode(x.view(F, per), 20, algo=m.ALGO_EMS, in_kind=m.IN_BPSK, sigma=sigma)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    out, it, ok = code.decode(x.view(F, per), 20, algo=m.ALGO_EMS, in_kind=m.IN_BPSK, sigma=sigma)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"C5 GF256 BPSK EMS(2,2) F={F} {ms:.2f} ms {F / ms * 1e3:.0f} frames/s {F * code.K_bits / ms / 1e3:.2f} info Mbit/s "
      f"FER={1 - ok.float().mean().item():.3f} avg_it={it.float().mean().item():.2f}")
