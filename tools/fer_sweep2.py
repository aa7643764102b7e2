#!/usr/bin/env python
"""Round-2 FER evidence (results -> profiles/r02_fer_sweep.txt): layered int8 (the throughput mode) beside flooding
fp32 (the reference's update rule, schedule and arithmetic, B/LDPC_Decoder.cu:172-372 on the fixed graph) on the
device channel, with Clopper-Pearson 95 % intervals, at operating points where the curves are NOT saturated:
  C3 (PON): Es/N0 swept downwards until FER ~ 1e-1 and ~ 1e-3, plus the 4.5 dB "throughput mode" point with an
            upper bound from the frames run; max 50 iterations, syndrome exit (B/define.cuh:35, BASELINE configs[2])
  C2      : 10 iterations where flooding fp32 leaves FER = 1 (2.6 - 3.4 dB Eb/N0) and 50 iterations at 1.8 - 2.2 dB
  C1      : 10 iterations, 2.5 - 3.5 dB Es/N0 (the points of the reference-GPU run, profiles/r02_reference_gpu.txt)
Stop rule per point: >= 100 frame errors and >= one batch, or max_frames (B/Simulation.cu:245-285 uses 50 / 10000)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipy.stats import beta
import cuda_ldpc_b200 as m
from cuda_ldpc_b200 import sim

BL = os.path.join(m.DATA_DIR, "bldpc")


def cp(k, n, conf=0.95):
    a = 1 - conf
    lo = 0.0 if k == 0 else beta.ppf(a / 2, k, n - k + 1)
    hi = 1.0 if k == n else beta.ppf(1 - a / 2, k + 1, n - k)
    return lo, hi


PLAN = [
    # name, file, geo, snrtype, maxit, [(snr, max_frames)], F_layered, F_flooding
    ("C3", "PON_LDPC.txt", (12, 69, 256), 1, 50, [(2.1, 1 << 16), (2.3, 1 << 17), (2.4, 1 << 18), (2.5, 1 << 19), (2.6, 1 << 20), (2.7, 1 << 21), (4.5, 1 << 23)], 16384, 4096),
    ("C2", "J15_L30_Z1280_BlockH.txt", (0, 0, 0), 0, 10, [(2.2, 1 << 18), (2.6, 1 << 20), (3.0, 1 << 21), (3.4, 1 << 21)], 4736, 1024),
    ("C2", "J15_L30_Z1280_BlockH.txt", (0, 0, 0), 0, 50, [(1.4, 1 << 15), (1.6, 1 << 18), (1.8, 1 << 19), (2.0, 1 << 20)], 4736, 1024),
    ("C1", "J4_L24_Z96_BlockH.txt", (0, 0, 0), 1, 10, [(2.5, 1 << 16), (3.0, 1 << 18), (3.5, 1 << 21)], 65536, 8192),
]
VARIANTS = [("flooding fp32 (reference rule)", dict(schedule=m.SCHED_FLOODING)),
            ("layered int8 m31 s8 x0.875", dict(msg_max=31, llr_scale=8.0, beta_num=1, beta_shift=3))]
only = sys.argv[1:] if __name__ == "__main__" else ["none"]
print("# cfg maxit  Es|Eb/N0   decoder                          frames   errors   FER        95 % Clopper-Pearson        avg it   verdict")
for name, f, geo, snrtype, maxit, pts, Fl, Ff in PLAN:
    if only and name not in only:
        continue
    code = m.LdpcCode(os.path.join(BL, f), *geo)
    for snr, maxf in pts:
        rows = []
        for vname, kw in VARIANTS:
            kw2 = dict(kw)
            sched = kw2.pop("schedule", m.SCHED_LAYERED)
            F = Ff if sched == m.SCHED_FLOODING else Fl
            mf = maxf if sched != m.SCHED_FLOODING else min(maxf, 1 << 18)   # the fp32 streaming mode is 6-10x slower
            run = sim.CudaBatchRunner(code, F, maxit=maxit, schedule=sched, early_exit=m.EXIT_SYNDROME, **kw2)
            res = sim.run_snr_point(run, snr, m.sigma_from_snr(snrtype, snr, code.rate), least_errors=100,
                                    least_frames=F, max_frames=mf, length=code.K)
            lo, hi = cp(res.num_Error_Frames, res.num_Frames)
            rows.append((vname, res, lo, hi))
        (_, rf, flo, fhi), (_, rl, llo, lhi) = rows
        verdict = ("intervals overlap" if not (lhi < flo or fhi < llo) else
                   ("layered int8 BETTER than the reference rule" if lhi < flo else "layered int8 WORSE"))
        for vname, res, lo, hi in rows:
            print(f"{name:3s} {maxit:3d}   {snr:4.1f} dB ({'Es' if snrtype else 'Eb'})  {vname:32s} {res.num_Frames:8d} {res.num_Error_Frames:7d}   "
                  f"{res.FER:.3e}  [{lo:.3e}, {hi:.3e}]   {res.AverageIT:5.2f}   {verdict if vname.startswith('layered') else ''}", flush=True)
