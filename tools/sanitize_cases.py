#!/usr/bin/env python
"""Small decode calls for compute-sanitizer (tools/sanitize.sh): every kernel family of libldpc_b200.so on
batches small enough for racecheck (~100x slowdown).  Operating points are chosen so that frames of one
4-codeword group converge at DIFFERENT iterations (the latch / syndrome-flag paths that raced in round 1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cuda_ldpc_b200 as m

BL = os.path.join(m.DATA_DIR, "bldpc")
NB = os.path.join(m.DATA_DIR, "nbldpc")
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def binary(name, fname, geo, F, snr_es, iters):
    code = m.LdpcCode(os.path.join(BL, fname), *geo)
    sigma = m.sigma_from_snr(1, snr_es, 1.0)
    g = torch.Generator(device="cuda").manual_seed(3)
    y = 1.0 + sigma * torch.randn(code.N, F, device="cuda", generator=g)
    for mode in (m.EXIT_SYNDROME, m.EXIT_NONE):
        for fmt in (m.OUT_INT32_REF, m.OUT_BITPACK):
            r = code.decode(y, iters, schedule=m.SCHED_LAYERED, early_exit=mode, out_format=fmt, msg_max=31,
                            beta_num=1, beta_shift=3, debug=(fmt == m.OUT_INT32_REF))
            torch.cuda.synchronize()
            print(f"{name} layered i8 mode={mode} fmt={fmt}: ok {int(r.ok.sum())}/{F} it {r.iters.min().item()}..{r.iters.max().item()}",
                  flush=True)
    r = code.decode_channel(F, iters, sigma, seed=5, msg_max=31, beta_num=1, beta_shift=3)
    torch.cuda.synchronize()
    print(f"{name} fused channel: ok {int(r.ok.sum())}/{F}", flush=True)
    for mode in (m.EXIT_SYNDROME, m.EXIT_NONE):  # fp16 message mode (2-codeword groups, records, dynamic scheduling)
        r = code.decode(y, iters, schedule=m.SCHED_LAYERED, msg_dtype=m.DTYPE_FP16, early_exit=mode, msg_max=31,
                        beta_num=1, beta_shift=3, debug=True)
        torch.cuda.synchronize()
        print(f"{name} layered fp16 mode={mode}: ok {int(r.ok.sum())}/{F} it {r.iters.min().item()}..{r.iters.max().item()}",
              flush=True)
    yh = y.cpu().numpy()
    r = code.decode(yh, iters, schedule=m.SCHED_LAYERED, early_exit=m.EXIT_SYNDROME, out_format=m.OUT_U8, msg_max=31)
    print(f"{name} host path: ok {int(r.ok.sum())}/{F}", flush=True)
    if code.N < 20000:
        r = code.decode(y, 3, early_exit=m.EXIT_SYNDROME)   # flooding fp32
        r = code.decode(y, 3, schedule=m.SCHED_LAYERED, msg_dtype=m.DTYPE_FP32, early_exit=m.EXIT_SYNDROME)
        torch.cuda.synchronize()
        print(f"{name} flooding / layered fp32 done", flush=True)


def nonbinary(name, matrix, q, const, exp, snr, F, algos, iters=6):
    from cuda_ldpc_b200.gf import write_table_file
    import tempfile
    d = tempfile.mkdtemp()
    gf = os.path.join(d, f"Arith.Table.GF.{q}.txt")
    write_table_file(q, gf)
    code = m.NbLdpcCode(os.path.join(NB, matrix), gf, os.path.join(NB, "Constellation", const), exp)
    sigma = code.sigma(0, snr)
    rx = code.modulate_awgn(F, sigma, seed=7)
    kind = m.IN_BPSK if const.startswith("BPSK") else m.IN_QAM
    for algo in algos:
        r = code.decode(rx, iters, algo=algo, in_kind=kind, sigma=sigma)
        torch.cuda.synchronize()
        print(f"{name} algo={algo}: ok {int(r.ok.sum())}/{F}", flush=True)


if which in ("all", "c2"):
    binary("C2<8>", "J15_L30_Z1280_BlockH.txt", (15, 30, 1280), 12, -0.6, 6)
if which in ("all", "c1"):
    binary("C1<20>", "J4_L24_Z96_BlockH.txt", (4, 24, 96), 42, 2.6, 8)
if which in ("all", "c3"):
    binary("C3<24>", "PON_LDPC.txt", (12, 69, 256), 14, 3.1, 8)
if which in ("all", "nb"):
    nonbinary("C4", "LDPC_N576_K288_GF64_d1_exp.txt", 64, "GRAY_64QAM.txt", 1, 9.5, 6,
              (m.ALGO_EMS, m.ALGO_TMM, m.ALGO_LAYERED_TMM, m.ALGO_FFT_BP))
    nonbinary("C5", "LDPC_N576_K480_GF256_exp.txt", 256, "BPSK.txt", 1, 4.5, 3,
              (m.ALGO_TMM, m.ALGO_LAYERED_TMM, m.ALGO_FFT_BP), iters=4)
print("sanitize cases done")
