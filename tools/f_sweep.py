#!/usr/bin/env python
"""Developer tool: C2 layered int8, 10 iterations, throughput against the batch size F (frames per launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cuda_ldpc_b200 as m
code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", "J15_L30_Z1280_BlockH.txt"))
for F in [int(a) for a in sys.argv[1:]] or [2368, 4736, 9472, 18944]:
    y = 1.0 + m.sigma_from_snr(0, 2.0, code.rate) * torch.randn(code.N, F, device="cuda")
    out = torch.empty(code.out_bytes(F, m.OUT_BITPACK), dtype=torch.uint8, device="cuda")
    it = torch.empty(F, dtype=torch.int32, device="cuda"); ok = torch.empty(F, dtype=torch.int32, device="cuda")
    kw = dict(schedule=m.SCHED_LAYERED, out_format=m.OUT_BITPACK, msg_max=31, beta_num=1, beta_shift=3, out=out, iters_out=it, ok_out=ok)
    for _ in range(3): code.decode(y, 10, **kw)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): code.decode(y, 10, **kw)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 5)
    print(f"F={F:6d} groups/CTA={F/4/148:5.1f}  ms {min(ts):8.3f}  Gbit/s {F * code.K / min(ts) / 1e6:6.2f}", flush=True)
    del y
