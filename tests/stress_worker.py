"""Worker of tests/test_stress_gpu.py: runs in its own process with LDPC_B200_LIB pointing at a race-hunting build
of the library (stress.cuh).  Decodes batches whose groups mix early and late frames, several times, through the
host and device entries, and compares every result with the CPU oracle (binary) / with run 0 (non-binary, whose
oracle comparison lives in test_nb_gpu.py).  Exit code 0 = every comparison held, 3 = a mismatch (printed)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C

import numpy as np
import torch

import cuda_ldpc_b200 as m
from conftest import DATA, OracleCode, fp, ip

assert os.path.basename(m.lib_path) == os.environ.get("LDPC_B200_LIB", "libldpc_b200.so")
orc = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
orc.orc_sigma.restype = C.c_float
orc.orc_sigma.argtypes = [C.c_int, C.c_float, C.c_float]
orc.orc_awgn.argtypes = [C.POINTER(C.c_int), C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int, C.c_int]
orc.orc_layered_i8.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_float,
                                                C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                                C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
orc.orc_layered_f16.argtypes = orc.orc_layered_i8.argtypes
bad = 0


def check(cond, what):
    global bad
    if not cond:
        bad += 1
        print("MISMATCH:", what, flush=True)


def binary(fname, geo, F, snr, iters, reps):
    path = os.path.join(DATA, "bldpc", fname)
    code, oc = m.LdpcCode(path, *geo), OracleCode(orc, path, *geo)
    N = code.N
    s = np.array([173, 173, 173], np.int32)
    y = np.zeros(N * F, np.float32)
    orc.orc_awgn(ip(s), orc.orc_sigma(1, snr, 1.0), None, fp(y), N, F)
    y = y.reshape(N, F)
    D = np.zeros((N + 1) * F, np.int32)
    its = np.zeros(F, np.int32)
    app = np.zeros(N * F, np.int8)
    rec = np.zeros(oc.M * 4 * F, np.uint32)
    assert orc.orc_layered_i8(oc.J, oc.L, oc.Z, ip(oc.H), fp(y), F, iters, 8.0, 31, 1, 3, 2, ip(D), ip(its),
                              app.ctypes.data, rec.ctypes.data) == 0
    D = D.reshape(N + 1, F)
    from test_binary_gpu import expand_messages
    msgs = expand_messages(rec, oc, F)
    assert its.min() < its.max()
    yd = torch.as_tensor(y, device="cuda")
    kw = dict(schedule=m.SCHED_LAYERED, early_exit=m.EXIT_SYNDROME, msg_max=31, beta_num=1, beta_shift=3)
    for rep in range(reps):
        r = code.decode(yd, iters, debug=True, **kw)
        torch.cuda.synchronize()
        check((r.D.cpu().numpy() == D).all(), f"{fname} device rep {rep}: hard bits / flags")
        check((r.iters.cpu().numpy() == its).all(), f"{fname} device rep {rep}: iterations")
        check((r.app.cpu().numpy() == app.reshape(N, F)).all(), f"{fname} device rep {rep}: APP")
        check((r.msgs.cpu().numpy() == msgs).all(), f"{fname} device rep {rep}: c2v messages")
        rh = code.decode(y, iters, out_format=m.OUT_U8, **kw)
        check((rh.D == D[:N]).all() and (rh.iters == its).all() and (rh.ok == D[N]).all(), f"{fname} host rep {rep}")
    print(f"{fname}: {reps} repeats x {F} frames, iterations {its.min()}..{its.max()}", flush=True)


def binary_f16(fname, geo, F, snr, iters, reps):
    """fp16 message mode (bldpc_layered_f16.cu): mixed early / late frames per 2-frame group, dynamic group scheduling."""
    path = os.path.join(DATA, "bldpc", fname)
    code, oc = m.LdpcCode(path, *geo), OracleCode(orc, path, *geo)
    N = code.N
    s = np.array([173, 173, 173], np.int32)
    y = np.zeros(N * F, np.float32)
    orc.orc_awgn(ip(s), orc.orc_sigma(1, snr, 1.0), None, fp(y), N, F)
    y = y.reshape(N, F)
    D = np.zeros((N + 1) * F, np.int32)
    its = np.zeros(F, np.int32)
    app = np.zeros(N * F, np.uint16)
    msg = np.zeros(oc.M * code.dc_max * F, np.uint16)
    assert orc.orc_layered_f16(oc.J, oc.L, oc.Z, ip(oc.H), fp(y), F, iters, 8.0, 31, 1, 3, 2, ip(D), ip(its),
                               app.ctypes.data, msg.ctypes.data) == 0
    D = D.reshape(N + 1, F)
    assert its.min() < its.max()
    yd = torch.as_tensor(y, device="cuda")
    kw = dict(schedule=m.SCHED_LAYERED, msg_dtype=m.DTYPE_FP16, early_exit=m.EXIT_SYNDROME, msg_max=31, beta_num=1,
              beta_shift=3)
    for rep in range(reps):
        r = code.decode(yd, iters, debug=True, **kw)
        torch.cuda.synchronize()
        check((r.D.cpu().numpy() == D).all(), f"{fname} fp16 device rep {rep}: hard bits / flags")
        check((r.iters.cpu().numpy() == its).all(), f"{fname} fp16 device rep {rep}: iterations")
        check((r.app.cpu().numpy().view(np.uint16) == app.reshape(N, F)).all(), f"{fname} fp16 device rep {rep}: APP")
        check((r.msgs.cpu().numpy().view(np.uint16) == msg).all(), f"{fname} fp16 device rep {rep}: c2v messages")
        rh = code.decode(y, iters, out_format=m.OUT_U8, **kw)
        check((rh.D == D[:N]).all() and (rh.iters == its).all() and (rh.ok == D[N]).all(), f"{fname} fp16 host rep {rep}")
    print(f"{fname} fp16: {reps} repeats x {F} frames, iterations {its.min()}..{its.max()}", flush=True)


def nonbinary(matrix, q, const, snr, F, algos, iters, reps):
    import tempfile
    from cuda_ldpc_b200.gf import write_table_file
    gf = os.path.join(tempfile.mkdtemp(), f"Arith.Table.GF.{q}.txt")
    write_table_file(q, gf)
    nbd = os.path.join(DATA, "nbldpc")
    code = m.NbLdpcCode(os.path.join(nbd, matrix), gf, os.path.join(nbd, "Constellation", const), True)
    sigma = code.sigma(0, snr)
    rx = code.modulate_awgn(F, sigma, seed=7)
    kind = m.IN_BPSK if const.startswith("BPSK") else m.IN_QAM
    for algo in algos:
        ref = None
        for rep in range(reps):
            out, it, ok = code.decode(rx, iters, algo=algo, in_kind=kind, sigma=sigma)
            torch.cuda.synchronize()
            got = (out.cpu().numpy().copy(), it.cpu().numpy().copy(), ok.cpu().numpy().copy())
            if ref is None:
                ref = got
            check(all((a == b).all() for a, b in zip(got, ref)), f"{matrix} algo {algo} rep {rep}")
        print(f"{matrix} algo {algo}: ok {int(ref[2].sum())}/{F}", flush=True)


what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "binary"):
    binary("J4_L24_Z96_BlockH.txt", (4, 24, 96), 1188, 2.7, 10, 6)
    binary("PON_LDPC.txt", (12, 69, 256), 600, 3.2, 10, 3)
    binary("J15_L30_Z1280_BlockH.txt", (15, 30, 1280), 600, -0.4, 10, 2)
    binary_f16("J4_L24_Z96_BlockH.txt", (4, 24, 96), 1187, 2.7, 10, 4)
    binary_f16("PON_LDPC.txt", (12, 69, 256), 599, 3.2, 10, 2)
if what in ("all", "nb"):
    nonbinary("LDPC_N576_K288_GF64_d1_exp.txt", 64, "GRAY_64QAM.txt", 9.5, 600,
              (m.ALGO_EMS, m.ALGO_TMM, m.ALGO_LAYERED_TMM, m.ALGO_FFT_BP), 8, 3)
    nonbinary("LDPC_N576_K480_GF256_exp.txt", 256, "BPSK.txt", 4.5, 300, (m.ALGO_TMM, m.ALGO_LAYERED_TMM, m.ALGO_FFT_BP), 6, 2)
print(f"stress worker done: {bad} mismatches", flush=True)
sys.exit(3 if bad else 0)
