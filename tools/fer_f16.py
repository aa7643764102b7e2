#!/usr/bin/env python
"""FER of the fp16 message mode beside the reference rule (flooding fp32) and the int8 mode, same device Philox channel,
Clopper-Pearson 95 % intervals (results -> profiles/r02_fer_fp16.txt).  Points: the non-saturated ones of
tools/fer_sweep2.py where the int8 mode (msg_max 31, scale 8, x0.875) was WORSE than the reference rule at 50
iterations (PON), plus J15_L30_Z1280 at 50 iterations and J4_L24_Z96 at the reference's 10."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fer_sweep2 import cp, BL
import cuda_ldpc_b200 as m
from cuda_ldpc_b200 import sim

PLAN = [
    ("C3", "PON_LDPC.txt", (12, 69, 256), 1, 50, [(2.4, 1 << 16), (2.5, 1 << 18), (2.6, 1 << 20), (2.7, 1 << 21)], 16384, 4096),
    ("C2", "J15_L30_Z1280_BlockH.txt", (0, 0, 0), 0, 50, [(1.6, 1 << 18), (1.8, 1 << 19)], 4736, 1024),
    ("C1", "J4_L24_Z96_BlockH.txt", (0, 0, 0), 1, 10, [(3.0, 1 << 18), (3.5, 1 << 21)], 65536, 8192),
]
Q = dict(msg_max=31, llr_scale=8.0, beta_num=1, beta_shift=3)
VARIANTS = [("flooding fp32 (reference rule)", dict(schedule=m.SCHED_FLOODING)),
            ("layered int8 m31 s8 x0.875", dict(Q)),
            ("layered fp16 m31 s8 x0.875", dict(Q, msg_dtype=m.DTYPE_FP16))]
EXTRA = {"C2": [("layered fp16 m31 s8 x0.9375", dict(Q, msg_dtype=m.DTYPE_FP16, beta_num=1, beta_shift=4)),
                ("layered fp16 m31 s8 x1", dict(Q, msg_dtype=m.DTYPE_FP16, beta_num=0, beta_shift=0)),
                ("layered int8 m31 s8 x1", dict(Q, beta_num=0, beta_shift=0))]}
only = [a for a in sys.argv[1:] if not a.startswith("--")]
extra_only = "--extra-only" in sys.argv
print("# cfg maxit  Es|Eb/N0   decoder                          frames   errors   FER        95 % Clopper-Pearson        avg it   vs reference rule")
for name, f, geo, snrtype, maxit, pts, Fl, Ff in PLAN:
    if only and name not in only:
        continue
    code = m.LdpcCode(os.path.join(BL, f), *geo)
    for snr, maxf in pts:
        rows = []
        for vname, kw in ((VARIANTS[:1] if extra_only else VARIANTS) + EXTRA.get(name, [])):
            kw2 = dict(kw)
            sched = kw2.pop("schedule", m.SCHED_LAYERED)
            F = Ff if sched == m.SCHED_FLOODING else Fl
            mf = maxf if sched != m.SCHED_FLOODING else min(maxf, 1 << 18)
            run = sim.CudaBatchRunner(code, F, maxit=maxit, schedule=sched, early_exit=m.EXIT_SYNDROME, **kw2)
            res = sim.run_snr_point(run, snr, m.sigma_from_snr(snrtype, snr, code.rate), least_errors=100,
                                    least_frames=F, max_frames=mf, length=code.K)
            lo, hi = cp(res.num_Error_Frames, res.num_Frames)
            rows.append((vname, res, lo, hi))
        _, rf, flo, fhi = rows[0]
        for vname, res, lo, hi in rows:
            verdict = "" if vname.startswith("flooding") else ("intervals overlap" if not (hi < flo or fhi < lo) else ("BETTER" if hi < flo else "WORSE"))
            print(f"{name:3s} {maxit:3d}   {snr:4.1f} dB ({'Es' if snrtype else 'Eb'})  {vname:32s} {res.num_Frames:8d} {res.num_Error_Frames:7d}   "
                  f"{res.FER:.3e}  [{lo:.3e}, {hi:.3e}]   {res.AverageIT:5.2f}   {verdict}", flush=True)
