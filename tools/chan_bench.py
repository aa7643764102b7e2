#!/usr/bin/env python
"""Developer timing of the channel / statistic kernels against their HBM byte counts (C2, F = 9472)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cuda_ldpc_b200 as m
code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", "J15_L30_Z1280_BlockH.txt"))
F = 9472
y = torch.empty(code.N, F, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = timeit(lambda: m.lib.ldpc_awgn_bpsk(code.handle, y.data_ptr(), F, m.LAYOUT_NF, 0.8, 173, 0, None, st))
print(f"ldpc_awgn_bpsk  [N][F] fp32: {ms:.3f} ms  {code.N * F * 4 / ms / 1e6:.0f} GB/s written, {code.N * F / ms / 1e6:.1f} G samples/s")
r = code.decode(y, 10, schedule=m.SCHED_LAYERED, out_format=m.OUT_BITPACK, msg_max=31, early_exit=m.EXIT_SYNDROME)
cnt = torch.zeros(6, dtype=torch.int64, device="cuda")
ms = timeit(lambda: m.lib.ldpc_statistic(code.handle, r.D.data_ptr(), m.OUT_BITPACK, r.ok.data_ptr(), r.iters.data_ptr(), F, code.K, None, cnt.data_ptr(), st))
print(f"ldpc_statistic  bit-packed:  {ms:.3f} ms  ({F * ((code.N + 31) // 32) * 4 / ms / 1e6:.0f} GB/s read)")
