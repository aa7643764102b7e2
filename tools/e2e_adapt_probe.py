"""Host-buffer path, fixed vs adaptive copy share of the hybrid feed (set LDPC_B200_HYBRID_COPY_PCT to pin the share).
One process per setting, alternating (tools/e2e_adapt_probe.sh)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cuda_ldpc_b200 as m
code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", "J15_L30_Z1280_BlockH.txt"))
F = 148 * 4 * 16
y = torch.empty(code.N, F, dtype=torch.float32).pin_memory()
y.copy_(1.0 + m.sigma_from_snr(0, 2.0, code.rate) * torch.randn(code.N, F))
yn = y.numpy()
out = torch.empty(code.out_bytes(F, m.OUT_BITPACK), dtype=torch.uint8).pin_memory().numpy()
it = torch.empty(F, dtype=torch.int32).pin_memory().numpy(); ok = torch.empty(F, dtype=torch.int32).pin_memory().numpy()
kw = dict(schedule=m.SCHED_LAYERED, out_format=m.OUT_BITPACK, msg_max=31, beta_num=1, beta_shift=3, out=out, iters_out=it, ok_out=ok)
for _ in range(3): code.decode(yn, 10, **kw)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); r = code.decode(yn, 10, **kw); ts.append(time.perf_counter() - t0)
ts = np.array(ts)
print(f"{os.environ.get('LDPC_B200_HYBRID_COPY_PCT', 'adaptive'):>8s}: median {np.median(ts)*1e3:6.2f} ms = {code.K*F/np.median(ts)/1e9:5.2f} Gbit/s  (min {ts.min()*1e3:.2f}, max {ts.max()*1e3:.2f}); "
      f"h2d {m.lib.ldpc_last_h2d_bytes(code.handle)/1e6:.0f} MB of {code.N*F*4/1e6:.0f}; launches {r.launches}", flush=True)
