import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, cuda_ldpc_b200 as m
CASES = [("C2", "J15_L30_Z1280_BlockH.txt", (0, 0, 0), 148 * 4 * 16, 2.0), ("C1", "J4_L24_Z96_BlockH.txt", (0, 0, 0), 65536, 4.0),
         ("C3", "PON_LDPC.txt", (12, 69, 256), 16384, 4.5), ("J10", "J10_L60_Z160_BlockH.txt", (0, 0, 0), 16384, 4.0)]
line = os.environ.get("LDPC_B200_LIB", "base") + ": "
for name, path, geo, F, snr in CASES:
    code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", path), *geo)
    y = 1.0 + m.sigma_from_snr(0, snr, code.rate) * torch.randn(code.N, F, device="cuda")
    out = torch.empty(code.out_bytes(F, m.OUT_BITPACK), dtype=torch.uint8, device="cuda")
    it = torch.empty(F, dtype=torch.int32, device="cuda"); ok = torch.empty(F, dtype=torch.int32, device="cuda")
    kw = dict(schedule=m.SCHED_LAYERED, msg_dtype=m.DTYPE_FP16, out_format=m.OUT_BITPACK, msg_max=31, beta_num=1, beta_shift=3, out=out, iters_out=it, ok_out=ok)
    for _ in range(3): code.decode(y, 10, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): code.decode(y, 10, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    line += f"{name} {ms:.3f} ms {code.K * F / ms / 1e6:.1f}; "
print(line, flush=True)
