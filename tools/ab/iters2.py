import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, cuda_ldpc_b200 as m
for name, path, geo, F, snr in [("C1", "J4_L24_Z96_BlockH.txt", (0, 0, 0), 65536, 4.0), ("C3", "PON_LDPC.txt", (12, 69, 256), 16384, 4.5)]:
    code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", path), *geo)
    y = 1.0 + m.sigma_from_snr(0, snr, code.rate) * torch.randn(code.N, F, device="cuda")
    out = torch.empty(code.out_bytes(F, m.OUT_BITPACK), dtype=torch.uint8, device="cuda")
    it = torch.empty(F, dtype=torch.int32, device="cuda"); ok = torch.empty(F, dtype=torch.int32, device="cuda")
    kw = dict(schedule=m.SCHED_LAYERED, out_format=m.OUT_BITPACK, msg_max=31, beta_num=1, beta_shift=3, out=out, iters_out=it, ok_out=ok)
    res = []
    for iters in range(1, 8):
        for _ in range(3): code.decode(y, iters, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): code.decode(y, iters, **kw)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 10)
    print(name, "t(it):", " ".join(f"{r:.3f}" for r in res), "| increments:", " ".join(f"{b - a:.3f}" for a, b in zip(res, res[1:])))
