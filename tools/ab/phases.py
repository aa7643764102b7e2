"""LDPC_B200_LIB=libldpc_b200_phase.so python tools/ab/phases.py — per-phase clock64 cycles of thread 0 of every CTA."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, cuda_ldpc_b200 as m
L = C.CDLL(m.lib_path)
for name, path, geo, F, snr in [("C2", "J15_L30_Z1280_BlockH.txt", (0, 0, 0), 148 * 4 * 16, 2.0), ("C1", "J4_L24_Z96_BlockH.txt", (0, 0, 0), 65536, 4.0), ("C3", "PON_LDPC.txt", (12, 69, 256), 16384, 4.5)]:
    code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", path), *geo)
    y = 1.0 + m.sigma_from_snr(0, snr, code.rate) * torch.randn(code.N, F, device="cuda")
    out = torch.empty(code.out_bytes(F, m.OUT_BITPACK), dtype=torch.uint8, device="cuda")
    it = torch.empty(F, dtype=torch.int32, device="cuda"); ok = torch.empty(F, dtype=torch.int32, device="cuda")
    kw = dict(schedule=m.SCHED_LAYERED, out_format=m.OUT_BITPACK, msg_max=31, beta_num=1, beta_shift=3, out=out, iters_out=it, ok_out=ok)
    for _ in range(2): code.decode(y, 10, **kw)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 8)()
    L.ldpcb_debug_phase_cycles(buf, 1)
    code.decode(y, 10, **kw); torch.cuda.synchronize()
    L.ldpcb_debug_phase_cycles(buf, 1)
    v = np.array(list(buf)[:5], float); tot = v.sum()
    print(name, "load %.1f%%  first sweep %.1f%%  later sweeps %.1f%% (per sweep %.1f%%)  syndrome %.1f%%  outputs %.1f%%" %
          (v[0] / tot * 100, v[1] / tot * 100, v[2] / tot * 100, v[2] / tot * 100 / 9, v[3] / tot * 100, v[4] / tot * 100))
