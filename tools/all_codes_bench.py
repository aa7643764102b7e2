#!/usr/bin/env python
"""Developer tool: layered int8 throughput (10 iterations fixed) of every shipped binary H file, with the edge-update rate."""
import glob, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cuda_ldpc_b200 as m
for path in sorted(glob.glob(os.path.join(m.DATA_DIR, "bldpc", "*.txt"))):
    b = os.path.basename(path)
    geo = (12, 69, 256) if b.startswith("PON") else (0, 0, 0)
    code = m.LdpcCode(path, *geo)
    F = max(4096, min(262144, (1 << 28) // code.N // 4 * 4))
    y = 1.0 + 0.6 * torch.randn(code.N, F, device="cuda")
    out = torch.empty(code.out_bytes(F, m.OUT_BITPACK), dtype=torch.uint8, device="cuda")
    it = torch.empty(F, dtype=torch.int32, device="cuda"); ok = torch.empty(F, dtype=torch.int32, device="cuda")
    kw = dict(schedule=m.SCHED_LAYERED, out_format=m.OUT_BITPACK, msg_max=31, beta_num=1, beta_shift=3, out=out, iters_out=it, ok_out=ok)
    for _ in range(2): code.decode(y, 10, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): code.decode(y, 10, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    E = int((code.tables()[0] >= 0).sum()) * code.Z
    print(f"{b:28s} N={code.N:6d} F={F:7d} {ms:8.3f} ms {F * code.K / ms / 1e6:7.2f} info Gbit/s  {F / 4 * E * 10 / ms / 1e6:7.1f} G edge-groups/s", flush=True)
    del y, out
