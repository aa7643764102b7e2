import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cuda_ldpc_b200 as m
code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", "J15_L30_Z1280_BlockH.txt"))
F = 148 * 4 * 8
y = torch.empty(code.N, F, dtype=torch.float32).pin_memory()
y.copy_(1.0 + m.sigma_from_snr(0, 2.0, code.rate) * torch.randn(code.N, F))
yn = y.numpy()
out = torch.empty(code.out_bytes(F, m.OUT_BITPACK), dtype=torch.uint8).pin_memory().numpy()
it = torch.empty(F, dtype=torch.int32).pin_memory().numpy(); ok = torch.empty(F, dtype=torch.int32).pin_memory().numpy()
for label, dt in (("int8", m.DTYPE_INT8), ("fp16", m.DTYPE_FP16)):
    kw = dict(schedule=m.SCHED_LAYERED, msg_dtype=dt, out_format=m.OUT_BITPACK, msg_max=31, beta_num=1, beta_shift=3, out=out, iters_out=it, ok_out=ok)
    for _ in range(2): r = code.decode(yn, 10, **kw)
    t0 = time.perf_counter()
    for _ in range(5): r = code.decode(yn, 10, **kw)
    dt_ = (time.perf_counter() - t0) / 5
    print(f"{label}: host-buffer call {dt_*1e3:.2f} ms = {code.K*F/dt_/1e9:.2f} Gbit/s, launches {r.launches}")
