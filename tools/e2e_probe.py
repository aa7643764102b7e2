import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cuda_ldpc_b200 as m
code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", "J15_L30_Z1280_BlockH.txt"))
F = 9472
y = (1.0 + 0.8 * torch.randn(code.N, F)).pin_memory().numpy()
ho = torch.empty(code.out_bytes(F, m.OUT_BITPACK), dtype=torch.uint8).pin_memory().numpy()
hi = torch.empty(F, dtype=torch.int32).pin_memory().numpy(); hk = torch.empty(F, dtype=torch.int32).pin_memory().numpy()
kw = dict(schedule=m.SCHED_LAYERED, out_format=m.OUT_BITPACK, msg_max=31, beta_num=1, beta_shift=3, out=ho, iters_out=hi, ok_out=hk)
for hp in [int(a) for a in sys.argv[1:]] or [0]:
    for _ in range(2): r = code.decode(y, 10, host_pack_threads=hp, **kw)
    t0 = time.perf_counter()
    for _ in range(4): r = code.decode(y, 10, host_pack_threads=hp, **kw)
    dt = (time.perf_counter() - t0) / 4
    print("host_pack_threads %2d: %.2f ms  %.2f Gbit/s  launches %d" % (hp, dt * 1e3, F * code.K / dt / 1e9, r.launches), flush=True)
