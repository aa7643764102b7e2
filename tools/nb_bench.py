#!/usr/bin/env python
"""Developer timing of the non-binary path (not the contract bench): frames/s and info Mbit/s of
nb_ldpc_decode_batch on device-resident channel samples."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cuda_ldpc_b200 as m

NB = os.path.join(m.DATA_DIR, "nbldpc")
CASES = [("C4 GF64 64QAM EMS(2,2)", "LDPC_N576_K288_GF64_d1_exp.txt", "GRAY_64QAM.txt", 64, 10.0, m.ALGO_EMS, 8192),
         ("C4 GF64 64QAM TMM", "LDPC_N576_K288_GF64_d1_exp.txt", "GRAY_64QAM.txt", 64, 10.0, m.ALGO_TMM, 8192),
         ("C4 GF64 64QAM layered TMM", "LDPC_N576_K288_GF64_d1_exp.txt", "GRAY_64QAM.txt", 64, 10.0, m.ALGO_LAYERED_TMM, 8192),
         ("BDS GF64 BPSK layered TMM", "BDS.576.288.GF.64.txt", "BPSK.txt", 2, 2.0, m.ALGO_LAYERED_TMM, 8192),
         ("C4 GF64 64QAM FFT-BP", "LDPC_N576_K288_GF64_d1_exp.txt", "GRAY_64QAM.txt", 64, 10.0, m.ALGO_FFT_BP, 8192),
         ("C5 GF256 BPSK FFT-BP", "LDPC_N576_K480_GF256_exp.txt", "BPSK.txt", 2, 4.5, m.ALGO_FFT_BP, 4096),
         ("C5 GF256 BPSK TMM", "LDPC_N576_K480_GF256_exp.txt", "BPSK.txt", 2, 4.5, m.ALGO_TMM, 4096),
         ("C5 GF256 BPSK layered TMM", "LDPC_N576_K480_GF256_exp.txt", "BPSK.txt", 2, 4.5, m.ALGO_LAYERED_TMM, 4096)]
print(torch.cuda.get_device_name(0))
FSCALE = int(os.environ.get("NB_BENCH_FSCALE", "1"))  # SURVEY 8(d) asks for F >= 16384: NB_BENCH_FSCALE=2 / 4
for name, mat, cst, nqam, ebn0, algo, F in CASES:
    F *= FSCALE
    code = m.NbLdpcCode(os.path.join(NB, mat), None, os.path.join(NB, "Constellation", cst), coef_is_exponent=mat.endswith("_exp.txt"))
    rate = (code.N - code.M) / code.N
    sigma = float(np.sqrt(0.5 / (np.log2(nqam) * rate * 10 ** (ebn0 / 10))))
    g = torch.Generator(device="cuda").manual_seed(1)
    if nqam == 2:
        x = 1.0 + sigma * torch.randn(F, code.N * code.p, device="cuda", generator=g)
        kind = m.IN_BPSK
    else:  # all-zero codeword = constellation point 0 (the reference's only non-BDS test vector)
        _, _, _, _, _ = code.tables()
        pts = np.loadtxt(os.path.join(NB, "Constellation", cst), usecols=(3, 5))
        x = torch.tensor(pts[0], dtype=torch.float32, device="cuda").repeat(F, code.N, 1) + sigma * torch.randn(F, code.N, 2, device="cuda", generator=g)
        x = x.reshape(F, code.N * 2).contiguous()
        kind = m.IN_QAM
    for _ in range(2):
        out, it, ok = code.decode(x, 20, algo=algo, in_kind=kind, sigma=sigma)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out, it, ok = code.decode(x, 20, algo=algo, in_kind=kind, sigma=sigma)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{name:28s} F={F:5d} {ms:9.2f} ms  {F / ms * 1e3:10.0f} frames/s  {F * code.K_bits / ms / 1e3:9.2f} info Mbit/s  "
          f"FER={1 - ok.float().mean().item():.3f} avg_it={it.float().mean().item():.2f}", flush=True)
