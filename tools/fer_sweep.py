#!/usr/bin/env python
"""FER of the engine's modes on the device channel (developer tool; results go to profiles/)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuda_ldpc_b200 as m
from cuda_ldpc_b200 import sim

BL = os.path.join(m.DATA_DIR, "bldpc")
CODES = {"C1": ("J4_L24_Z96_BlockH.txt", (0, 0, 0), 1, [2.5, 3.0, 3.5]),
         "C2": ("J15_L30_Z1280_BlockH.txt", (0, 0, 0), 0, [1.6, 1.8, 2.0, 2.2]),
         "C3": ("PON_LDPC.txt", (12, 69, 256), 1, [3.6, 3.9, 4.2])}
VARIANTS = [("flooding fp32", dict(schedule=m.SCHED_FLOODING)),
            ("layered i8 m31 s8", dict(msg_max=31, llr_scale=8.0)),
            ("layered i8 m31 s8 b.875", dict(msg_max=31, llr_scale=8.0, beta_num=1, beta_shift=3)),
            ("layered i8 m31 s8 b.75", dict(msg_max=31, llr_scale=8.0, beta_num=1, beta_shift=2)),
            ("layered i8 m15 s4 b.75", dict(msg_max=15, llr_scale=4.0, beta_num=1, beta_shift=2)),
            ("layered i8 m63 s16 b.75", dict(msg_max=63, llr_scale=16.0, beta_num=1, beta_shift=2)),
            ("layered i8 m127 s8", dict(msg_max=127, llr_scale=8.0))]
names = sys.argv[1:] or list(CODES)
for name in names:
    f, geo, snrtype, snrs = CODES[name]
    code = m.LdpcCode(os.path.join(BL, f), *geo)
    for vname, kw in VARIANTS:
        row = []
        for snr in snrs:
            kw2 = dict(kw)
            sched = kw2.pop("schedule", m.SCHED_LAYERED)
            F = 2048 if sched == m.SCHED_FLOODING and name == "C2" else 8192
            run = sim.CudaBatchRunner(code, F, maxit=10, schedule=sched, early_exit=m.EXIT_SYNDROME, **kw2)
            res = sim.run_snr_point(run, snr, m.sigma_from_snr(snrtype, snr, code.rate), least_errors=200,
                                    least_frames=F, max_frames=8 * F, length=code.K)
            row.append(f"{snr}dB FER {res.FER:.3e} it {res.AverageIT:.2f} ({res.num_Frames})")
        print(f"{name} {vname:26s} | " + " | ".join(row), flush=True)
