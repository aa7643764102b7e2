import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, cuda_ldpc_b200 as m
for name, path, geo, F, snr, iters in [("C3 exit 4.5dB", "PON_LDPC.txt", (12, 69, 256), 16384, 4.5, 50), ("C3 exit 3.0dB", "PON_LDPC.txt", (12, 69, 256), 16384, 3.0, 50),
                                       ("C2 exit 2.0dB", "J15_L30_Z1280_BlockH.txt", (0, 0, 0), 4736, 2.0, 10), ("C1 exit 3.0dB", "J4_L24_Z96_BlockH.txt", (0, 0, 0), 65536, 3.0, 10)]:
    code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", path), *geo)
    st = 1 if not name.startswith("C2") else 0
    y = 1.0 + m.sigma_from_snr(st, snr, code.rate) * torch.randn(code.N, F, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    out = torch.empty(code.out_bytes(F, m.OUT_BITPACK), dtype=torch.uint8, device="cuda")
    it = torch.empty(F, dtype=torch.int32, device="cuda"); ok = torch.empty(F, dtype=torch.int32, device="cuda")
    kw = dict(schedule=m.SCHED_LAYERED, early_exit=m.EXIT_SYNDROME, out_format=m.OUT_BITPACK, msg_max=31, beta_num=1, beta_shift=3, out=out, iters_out=it, ok_out=ok)
    for _ in range(3): code.decode(y, iters, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): code.decode(y, iters, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms:.3f} ms {F * code.K / ms / 1e6:.2f} Gbit/s avg_it {it.float().mean().item():.2f} ok {ok.float().mean().item():.3f}")
