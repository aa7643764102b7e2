"""Channel, statistics, encoder and simulation loop: CPU known-answer tests plus GPU tests through
the C-ABI."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import DATA, GOLDEN, ROOT, OracleCode, ip, fp

import cuda_ldpc_b200 as m
from cuda_ldpc_b200 import sim

BL = os.path.join(DATA, "bldpc")


def test_philox_known_answers():
    """Philox4x32-10 KATs from Random123 (Salmon et al., SC'11)."""
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        c = np.array(ctr, np.uint32); k = np.array(key, np.uint32); o = np.zeros(4, np.uint32)
        m.lib.ldpc_philox4x32(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        assert tuple(int(x) for x in o) == want


@pytest.mark.parametrize("f,geo", [("J4_L24_Z96_BlockH.txt", (0, 0, 0)), ("PON_LDPC.txt", (12, 69, 256)),
                                   ("J15_L30_Z1280_BlockH.txt", (0, 0, 0)), ("J10_L60_Z160_BlockH.txt", (0, 0, 0))])
def test_encoder_produces_codewords(oracle, f, geo):
    """ldpc_encode (the reference has no encoder): systematic, H c = 0, linear."""
    code = m.LdpcCode(os.path.join(BL, f), *geo)
    oc = OracleCode(oracle, os.path.join(BL, f), code.J, code.L, code.Z)
    rng = np.random.default_rng(5)
    us = rng.integers(0, 2, (3, code.K)).astype(np.uint8)
    us[2] = us[0] ^ us[1]
    cws = np.zeros((3, code.N), np.uint8)
    for i in range(3):
        assert m.lib.ldpc_encode(code.handle, us[i].ctypes.data, cws[i].ctypes.data) == 0
        assert (cws[i, : code.K] == us[i]).all()
    D = np.concatenate([cws.T.astype(np.int32), np.zeros((1, 3), np.int32)], 0)
    ok = np.zeros(3, np.int32)
    oracle.orc_syndrome_ok(oc.J, oc.L, oc.Z, ip(oc.H), ip(np.ascontiguousarray(D)), 3, ip(ok))
    assert ok.tolist() == [1, 1, 1]
    assert (cws[2] == (cws[0] ^ cws[1])).all()


# ------------------------------------------------------------------ GPU

@pytest.mark.gpu
def test_awgn_statistics_and_shard_independence():
    import torch
    code = m.LdpcCode(os.path.join(BL, "J4_L24_Z96_BlockH.txt"))
    N, F, sigma = code.N, 512, 0.7

    def gen(F, first, layout, cw=None):
        y = torch.empty(N * F, dtype=torch.float32, device="cuda")
        cwp = cw.data_ptr() if cw is not None else None
        rc = m.lib.ldpc_awgn_bpsk(code.handle, y.data_ptr(), F, layout, sigma, 173, first, cwp,
                                  torch.cuda.current_stream().cuda_stream)
        assert rc >= 0
        torch.cuda.synchronize()
        return y.cpu().numpy()
    y = gen(F, 0, m.LAYOUT_NF).reshape(N, F)
    n = (y - 1.0) / sigma
    assert abs(n.mean()) < 0.01 and abs(n.var() - 1.0) < 0.01
    assert abs((n ** 4).mean() - 3.0) < 0.08 and abs((n ** 3).mean()) < 0.03
    assert abs(np.corrcoef(n[:-1].ravel(), n[1:].ravel())[0, 1]) < 0.01
    # any sharding of the frame range gives the same samples (keyed by the global frame index)
    a = gen(200, 0, m.LAYOUT_NF).reshape(N, 200)
    b = gen(312, 200, m.LAYOUT_NF).reshape(N, 312)
    assert (np.concatenate([a, b], 1).view(np.uint32) == y.view(np.uint32)).all()
    # [F][N] layout is the transpose; a different seed decorrelates; a codeword flips the signal
    yt = gen(F, 0, m.LAYOUT_FN).reshape(F, N)
    assert (yt.T.view(np.uint32) == y.view(np.uint32)).all()
    cw = torch.as_tensor(np.random.default_rng(1).integers(0, 2, N).astype(np.uint8), device="cuda")
    yc = gen(F, 0, m.LAYOUT_NF, cw).reshape(N, F)
    assert np.allclose(yc - (1.0 - 2.0 * cw.cpu().numpy()[:, None]), y - 1.0, atol=1e-6)


@pytest.mark.gpu
def test_fused_channel_equals_awgn_then_decode():
    """LDPC_DTYPE_CHANNEL (channel generated inside the layered kernel's load phase) gives bit for bit the
    result of ldpc_awgn_bpsk followed by ldpc_decode_batch, for any sharding offset and a non-zero codeword."""
    import torch
    code = m.LdpcCode(os.path.join(BL, "PON_LDPC.txt"), 12, 69, 256)
    u = np.random.default_rng(3).integers(0, 2, code.K).astype(np.uint8)
    cw = np.zeros(code.N, np.uint8)
    assert m.lib.ldpc_encode(code.handle, u.ctypes.data, cw.ctypes.data) == 0
    cwd = torch.as_tensor(cw, device="cuda")
    F, sigma, first = 203, 0.42, 12345
    y = torch.empty(code.N * F, dtype=torch.float32, device="cuda")
    assert m.lib.ldpc_awgn_bpsk(code.handle, y.data_ptr(), F, m.LAYOUT_NF, sigma, 99, first, cwd.data_ptr(),
                                torch.cuda.current_stream().cuda_stream) >= 0
    kw = dict(early_exit=m.EXIT_SYNDROME, out_format=m.OUT_BITPACK, msg_max=31, beta_num=1, beta_shift=3)
    a = code.decode(y.view(code.N, F), 12, schedule=m.SCHED_LAYERED, **kw)
    b = code.decode_channel(F, 12, sigma, seed=99, first_frame=first, codeword=cwd, **kw)
    torch.cuda.synchronize()
    assert (a.D == b.D).all() and (a.iters == b.iters).all() and (a.ok == b.ok).all()
    assert b.ok.float().mean() > 0.3 and a.launches == 1 and b.launches == 1


@pytest.mark.gpu
def test_statistic_kernel_all_formats():
    import torch
    code = m.LdpcCode(os.path.join(BL, "J4_L24_Z96_BlockH.txt"))
    N, K, F = code.N, code.K, 300
    rng = np.random.default_rng(2)
    cw = rng.integers(0, 2, N).astype(np.uint8)
    bits = (cw[:, None] ^ (rng.random((N, F)) < 0.001)).astype(np.uint8)
    bits[:, ::3] = cw[:, None]
    flag = (rng.random(F) < 0.8).astype(np.int32)
    it = rng.integers(1, 11, F).astype(np.int32)
    err = (bits[:K] != cw[:K, None]).sum(0)
    want = [F, ((err != 0) | (flag == 0)).sum(), err.sum(), it.sum(), ((err != 0) & (flag == 1)).sum(),
            ((err == 0) & (flag == 0)).sum()]
    W = (N + 31) // 32
    padded = np.zeros((W * 32, F), np.uint8); padded[:N] = bits
    packed = (padded.T.reshape(F, W, 32).astype(np.uint32) << np.arange(32, dtype=np.uint32)).sum(2).astype(np.uint32)
    ref32 = np.concatenate([bits.astype(np.int32), flag[None]], 0)
    for fmt, arr, okp in ((m.OUT_INT32_REF, ref32, False), (m.OUT_U8, bits, True), (m.OUT_BITPACK, packed, True)):
        D = torch.as_tensor(np.ascontiguousarray(arr), device="cuda")
        okd = torch.as_tensor(flag, device="cuda"); itd = torch.as_tensor(it, device="cuda")
        cwd = torch.as_tensor(cw, device="cuda"); cnt = torch.zeros(6, dtype=torch.int64, device="cuda")
        rc = m.lib.ldpc_statistic(code.handle, D.data_ptr(), fmt, okd.data_ptr() if okp else None, itd.data_ptr(), F, K,
                                  cwd.data_ptr(), cnt.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc >= 0
        assert cnt.cpu().numpy().tolist() == [int(x) for x in want], fmt


@pytest.mark.gpu
def test_reference_fer_counters_reproduced_through_the_cabi(oracle):
    """FER parity, seed-exact: the reference's own noise stream (RandomModule, seeds 173) decoded by
    the flooding fp32 kernels with the reference's genie stop rule gives the reference's own
    Simulation_GPU counters (one-line-fixed Transform_H; BASELINE.md §2: 55/1024 at 3.0 dB)."""
    import torch
    with open(os.path.join(GOLDEN, "binary_ref.json")) as f:
        pts = [p for p in json.load(f)["sim_points"] if p["variant"] == "fixed"]
    code = m.LdpcCode(os.path.join(BL, "J4_L24_Z96_BlockH.txt"))
    for p in pts:
        seed = np.array([173, 173, 173], np.int32)
        sigma = oracle.orc_sigma(p["snrtype"], p["snr_db"], code.rate)
        cnt = torch.zeros(6, dtype=torch.int64, device="cuda")
        frames = 0
        while True:
            y = np.zeros(code.N * p["F"], np.float32)
            oracle.orc_awgn(ip(seed), sigma, None, fp(y), code.N, p["F"])
            r = code.decode(torch.as_tensor(y.reshape(code.N, p["F"]), device="cuda"), p["maxit"], early_exit=m.EXIT_GENIE)
            m.lib.ldpc_statistic(code.handle, r.D.data_ptr(), m.OUT_INT32_REF, None, r.iters.data_ptr(), p["F"], code.K,
                                 None, cnt.data_ptr(), torch.cuda.current_stream().cuda_stream)
            frames += p["F"]
            c = cnt.cpu().numpy()
            if c[1] >= 50 and frames >= p["leastTestFrames"]:
                break
        assert c.tolist() == p["counters"], p


@pytest.mark.gpu
def test_nonzero_codeword_roundtrip_and_sim_loop(oracle):
    """encode -> AWGN -> decode recovers the codeword; the on-device simulation loop counts it as
    error-free; a noisy point reaches the stop rule with an FER consistent with the oracle's."""
    import torch
    code = m.LdpcCode(os.path.join(BL, "J4_L24_Z96_BlockH.txt"))
    oc = OracleCode(oracle, os.path.join(BL, "J4_L24_Z96_BlockH.txt"), 4, 24, 96)
    u = np.random.default_rng(9).integers(0, 2, code.K).astype(np.uint8)
    cw = np.zeros(code.N, np.uint8)
    assert m.lib.ldpc_encode(code.handle, u.ctypes.data, cw.ctypes.data) == 0
    run = sim.CudaBatchRunner(code, 512, maxit=10, codeword=cw)
    res = sim.run_snr_point(run, 5.0, m.sigma_from_snr(1, 5.0, code.rate), least_errors=10 ** 9, max_frames=1024,
                            length=code.K)
    assert res.num_Frames == 1024 and res.num_Error_Frames == 0 and res.num_Error_Bits == 0
    hard = ((run.out.view(torch.int32).view(512, -1).cpu().numpy().astype(np.uint32)[:, :, None]
             >> np.arange(32, dtype=np.uint32)) & 1).reshape(512, -1)[:, : code.N]
    assert (hard == cw[None]).all()
    # noisy point, all-zero codeword: FER of the engine vs FER of the oracle on an independent sample
    run0 = sim.CudaBatchRunner(code, 1024, maxit=10)
    sg = m.sigma_from_snr(1, 2.6, code.rate)
    res = sim.run_snr_point(run0, 2.6, sg, least_errors=100, least_frames=4096, length=code.K)
    F = 384
    y = (1.0 + sg * np.random.default_rng(4).standard_normal((code.N, F))).astype(np.float32)
    D = np.zeros((code.N + 1) * F, np.int32); it = np.zeros(F, np.int32)
    oracle.orc_layered_i8(oc.J, oc.L, oc.Z, ip(oc.H), fp(y), F, 10, 8.0, 31, 0, 0, 2, ip(D), ip(it), None, None)
    fer_o = 1.0 - D.reshape(code.N + 1, F)[code.N].mean()
    se = np.sqrt(max(fer_o * (1 - fer_o), 1e-4) / F + max(res.FER * (1 - res.FER), 1e-4) / res.num_Frames)
    assert abs(res.FER - fer_o) < 4 * se + 0.01, (res.FER, fer_o)
    assert 1.0 <= res.AverageIT <= 10.0


@pytest.mark.gpu
def test_ldpc_sim_cli_runs():
    exe = os.path.join(ROOT, "cuda_ldpc_b200", "ldpc_sim")
    out = subprocess.run([exe, "--code", os.path.join(BL, "J4_L24_Z96_BlockH.txt"), "--snr", "2.5", "3.5", "0.5",
                          "--layered", "--maxit", "10", "--batch", "2048", "--least-frames", "4096", "--max-frames",
                          "65536", "--encode"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    rows = [l.split() for l in out.stdout.splitlines() if l.startswith(" ") and l.split()[0] in ("2.5", "3.0", "3.5")]
    assert len(rows) == 3
    fer = [float(r[3]) for r in rows]
    assert fer[0] > fer[1] > fer[2] and "task finish" in out.stdout
