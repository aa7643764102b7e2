"""fp16 message mode beside the int8 mode: decode time at 10 fixed iterations, early-exit time, and frame errors
on the same Philox channel (tools/f16_probe.py; device-resident inputs, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cuda_ldpc_b200 as m

CASES = [("C2", "J15_L30_Z1280_BlockH.txt", (0, 0, 0), 148 * 4 * 16, 2.0, 1.6),
         ("C1", "J4_L24_Z96_BlockH.txt", (0, 0, 0), 65536, 4.0, 3.2),
         ("C3", "PON_LDPC.txt", (12, 69, 256), 16384, 4.5, 3.4)]
for name, path, geo, F, snr, snr_fer in CASES:
    code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", path), *geo)
    y = 1.0 + m.sigma_from_snr(0, snr, code.rate) * torch.randn(code.N, F, device="cuda")
    out = torch.empty(code.out_bytes(F, m.OUT_BITPACK), dtype=torch.uint8, device="cuda")
    it = torch.empty(F, dtype=torch.int32, device="cuda"); ok = torch.empty(F, dtype=torch.int32, device="cuda")
    kw = dict(schedule=m.SCHED_LAYERED, out_format=m.OUT_BITPACK, msg_max=31, beta_num=1, beta_shift=3, out=out, iters_out=it, ok_out=ok)
    line = f"{name}: "
    for label, dt in (("int8", m.DTYPE_INT8), ("fp16", m.DTYPE_FP16)):
        for mode, iters in ((m.EXIT_NONE, 10), (m.EXIT_SYNDROME, 50)):
            for _ in range(3): code.decode(y, iters, msg_dtype=dt, early_exit=mode, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): code.decode(y, iters, msg_dtype=dt, early_exit=mode, **kw)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            line += f"{label} {'fixed10' if mode == m.EXIT_NONE else 'exit50'} {ms:.3f} ms = {code.K * F / ms / 1e6:.1f} Gbit/s; "
    print(line, flush=True)
    # frame errors at a waterfall point, same channel stream for both modes
    sigma = m.sigma_from_snr(0, snr_fer, code.rate)
    Ff = 148 * 64
    tot = {"int8": [0, 0.0], "fp16": [0, 0.0]}
    nb = 8
    for b in range(nb):
        for label, dt in (("int8", m.DTYPE_INT8), ("fp16", m.DTYPE_FP16)):
            r = code.decode_channel(Ff, 20, sigma, seed=7, first_frame=b * Ff, msg_dtype=dt, msg_max=31, beta_num=1, beta_shift=3)
            torch.cuda.synchronize()
            W = r.D.view(torch.int32)
            K32 = code.K // 32
            err = (W[:, :K32] != 0).any(1) | (r.ok == 0)
            tot[label][0] += int(err.sum()); tot[label][1] += float(r.iters.float().sum())
    n = nb * Ff
    print(f"   FER at Eb/N0 {snr_fer} dB, 20 it, {n} frames: " + "; ".join(f"{k} {v[0]}/{n} = {v[0]/n:.2e}, avg it {v[1]/n:.2f}" for k, v in tot.items()), flush=True)
