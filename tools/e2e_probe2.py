#!/usr/bin/env python
"""Developer probe: host-buffer ldpc_decode_batch (C2, F = 9472, 10 iterations) with the host quantiser on / off and
different chunk sizes (LDPC_B200_CHUNK_GROUPS = groups per SM and chunk)."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, cuda_ldpc_b200 as m
    code = m.LdpcCode(os.path.join(m.DATA_DIR, "bldpc", "J15_L30_Z1280_BlockH.txt"))
    F = 148 * 4 * 16
    y = (1.0 + 0.79 * torch.randn(code.N, F)).pin_memory().numpy()
    ho = torch.empty(code.out_bytes(F, m.OUT_BITPACK), dtype=torch.uint8).pin_memory().numpy()
    hi = torch.empty(F, dtype=torch.int32).pin_memory().numpy(); hk = torch.empty(F, dtype=torch.int32).pin_memory().numpy()
    kw = dict(schedule=m.SCHED_LAYERED, out_format=m.OUT_BITPACK, msg_max=31, beta_num=1, beta_shift=3, out=ho, iters_out=hi, ok_out=hk)
    for pack in [int(x) for x in sys.argv[2:]]:
        code.decode(y, 10, host_pack_threads=pack, **kw)
        t0 = time.perf_counter()
        for _ in range(5): code.decode(y, 10, host_pack_threads=pack, **kw)
        dt = (time.perf_counter() - t0) / 5
        print(f"  chunk_groups={os.environ.get('LDPC_B200_CHUNK_GROUPS','2')} pack={pack:3d}: {dt*1e3:7.2f} ms  {F*code.K/dt/1e9:6.2f} Gbit/s  host read {y.nbytes/dt/1e9:5.1f} GB/s", flush=True)
else:
    combos = (("2", "0"), ("2", "50"), ("2", "75"), ("2", "100"), ("2", "150"), ("4", "75"), ("4", "100"), ("1", "75"), ("1", "100"))
    if len(sys.argv) > 1 and sys.argv[1] == "pct":
        combos = tuple(("2", x) for x in sys.argv[2:])
    print("host cpus", len(os.sched_getaffinity(0)), flush=True)
    for cg, pct in combos:
        print("chunk groups", cg, "copy pct", pct, flush=True)
        subprocess.run([sys.executable, __file__, "child", "-1", "0", "12", "0", "12"],
                       env=dict(os.environ, LDPC_B200_CHUNK_GROUPS=cg, LDPC_B200_HYBRID_COPY_PCT=pct))
