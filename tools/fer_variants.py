#!/usr/bin/env python
"""int8 quantiser / normalisation variants at the two points where the 50-iteration layered int8 curve sits right of
the flooding fp32 one (tools/fer_sweep2.py): which fixed-point setting closes the gap?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_ldpc_b200 as m
from cuda_ldpc_b200 import sim
from fer_sweep2 import cp
BL = os.path.join(m.DATA_DIR, "bldpc")
V = [("m31 s8 x0.875", dict(msg_max=31, llr_scale=8.0, beta_num=1, beta_shift=3)),
     ("m31 s8 x1", dict(msg_max=31, llr_scale=8.0)),
     ("m127 s8 x0.875", dict(msg_max=127, llr_scale=8.0, beta_num=1, beta_shift=3)),
     ("m63 s16 x0.875", dict(msg_max=63, llr_scale=16.0, beta_num=1, beta_shift=3)),
     ("m127 s16 x0.875", dict(msg_max=127, llr_scale=16.0, beta_num=1, beta_shift=3)),
     ("m127 s16 x0.8125", dict(msg_max=127, llr_scale=16.0, beta_num=3, beta_shift=4)),
     ("m127 s24 x0.875", dict(msg_max=127, llr_scale=24.0, beta_num=1, beta_shift=3)),
     ("m127 s16 x0.75", dict(msg_max=127, llr_scale=16.0, beta_num=1, beta_shift=2))]
for name, f, geo, st, snr, F, maxf in [("C3", "PON_LDPC.txt", (12, 69, 256), 1, 2.6, 16384, 1 << 19),
                                       ("C2", "J15_L30_Z1280_BlockH.txt", (0, 0, 0), 0, 1.8, 4736, 1 << 19)]:
    code = m.LdpcCode(os.path.join(BL, f), *geo)
    for vn, kw in V:
        run = sim.CudaBatchRunner(code, F, maxit=50, early_exit=m.EXIT_SYNDROME, **kw)
        r = sim.run_snr_point(run, snr, m.sigma_from_snr(st, snr, code.rate), least_errors=100, least_frames=F, max_frames=maxf, length=code.K)
        lo, hi = cp(r.num_Error_Frames, r.num_Frames)
        print(f"{name} 50 it {snr} dB  layered int8 {vn:18s} {r.num_Frames:8d} {r.num_Error_Frames:5d}  {r.FER:.3e} [{lo:.3e}, {hi:.3e}]  it {r.AverageIT:.2f}", flush=True)
