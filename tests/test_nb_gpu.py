"""GPU parity tests of the non-binary decode path through the C-ABI: golden vectors recorded from
the reference's own CPU decoder, and the CPU oracle on fresh seeded frames."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, NB_DATA

import cuda_ldpc_b200 as m

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def meta():
    with open(os.path.join(GOLDEN, "nb_ref.json")) as f:
        return json.load(f)


def paths(cfg, gf_dir):
    return (os.path.join(NB_DATA, cfg["matrix"]), os.path.join(gf_dir, f"Arith.Table.GF.{cfg['q']}.txt"),
            os.path.join(NB_DATA, cfg["constellation"].replace("./", "")))


def orc_load(nb_oracle, cfg, gf_dir, exp):
    mt, gf, cs = paths(cfg, gf_dir)
    h = nb_oracle.nb_orc_load(mt.encode(), gf.encode(), cs.encode(), cfg["n_qam"], exp)
    assert h
    return C.c_void_p(h)


def channel_input(cfg, rx, N, p):
    """golden rx is interleaved complex; BPSK input of the C-ABI is the real part per bit"""
    if cfg["n_qam"] == 2:
        return np.ascontiguousarray(rx[:, 0::2]), m.IN_BPSK
    return np.ascontiguousarray(rx), m.IN_QAM


@pytest.mark.parametrize("name", ["BDS", "C4", "C5"])
def test_matches_reference_golden_vectors(gf_dir, meta, name):
    """Fused demapper + decoder on the reference's recorded channel samples: decoded symbols, return
    value and iter_number of the reference's CPU decoder.  TMM / layered TMM are exact; EMS is exact
    on every converged frame (non-converged frames amplify the reference's running-sum rounding,
    SURVEY C.3 — the documented tie class)."""
    cfg = meta["configs"][name]
    g = np.load(os.path.join(GOLDEN, f"nb_{name}.npz"))
    mt, gf, cs = paths(cfg, gf_dir)
    code = m.NbLdpcCode(mt, gf, cs, coef_is_exponent=False)  # raw coefficients, as the reference reads them
    inp, kind = channel_input(cfg, g["rx"], code.N, code.p)
    for a in cfg["algos"]:
        out, it, ok = code.decode(inp, cfg["maxit"], algo=a, in_kind=kind, sigma=float(g["sigma"]))
        want_ok, want_it, want_out = g[f"ret_{a}"], g[f"iter_{a}"], g[f"out_{a}"]
        sel = np.ones(len(ok), bool) if a != m.ALGO_EMS else (want_ok == 1)
        assert sel.sum() >= 3
        assert (ok[sel] == want_ok[sel]).all() and (it[sel] == want_it[sel]).all(), (name, a)
        assert (out[sel].astype(np.int32) == want_out[sel]).all(), (name, a)
        if a == m.ALGO_EMS:
            assert (ok == want_ok).sum() >= len(ok) - 2


@pytest.mark.parametrize("name,exp,snr,F", [("BDS", 0, 2.5, 48), ("C4", 1, 9.5, 40), ("C5", 1, 4.0, 24)])
def test_bit_exact_vs_oracle_on_seeded_frames(nb_oracle, gf_dir, meta, name, exp, snr, F):
    """Symbol-LLR input (Demodulate output of the oracle) -> identical decisions, iteration counts and
    flags as the oracle for EMS (`fresh` sums: the kernel's specification), TMM and layered TMM,
    with the exponent -> alpha^e coefficient mapping where the file needs it."""
    cfg = meta["configs"][name]
    mt, gf, cs = paths(cfg, gf_dir)
    h = orc_load(nb_oracle, cfg, gf_dir, exp)
    code = m.NbLdpcCode(mt, None, cs, coef_is_exponent=bool(exp))  # tables generated from the polynomial
    N, q, p = code.N, code.q, code.p
    sym = np.zeros(N, np.int32)
    L = N * p if cfg["n_qam"] == 2 else N
    tx = np.zeros(2 * L, np.float32)
    nb_oracle.nb_orc_modulate(h, sym.ctypes.data, tx.ctypes.data)
    sigma = nb_oracle.nb_orc_sigma(h, 0, snr)
    seed = np.array([173, 173, 173], np.int32)
    lch = np.zeros((F, N * (q - 1)), np.float32)
    rxs = np.zeros((F, 2 * L), np.float32)
    for f in range(F):
        nb_oracle.nb_orc_awgn(seed.ctypes.data, sigma, tx.ctypes.data, rxs[f].ctypes.data, L)
        nb_oracle.nb_orc_demodulate(h, sigma, rxs[f].ctypes.data, lch[f].ctypes.data)
    algos = [(m.ALGO_EMS, 1), (m.ALGO_TMM, 0), (m.ALGO_LAYERED_TMM, 0)]
    if name == "C5":
        algos = algos[1:] + [(m.ALGO_EMS, 1)]
    converged = 0
    for a, summode in algos:
        Fa = F if not (name == "C5" and a == m.ALGO_EMS) else 6
        o_out = np.zeros((Fa, N), np.int32); o_it = np.zeros(Fa, np.int32); o_ok = np.zeros(Fa, np.int32)
        nb_oracle.nb_orc_decode_batch(h, a, summode, lch.ctypes.data, Fa, 20, 2, 2, o_out.ctypes.data,
                                      o_it.ctypes.data, o_ok.ctypes.data)
        out, it, ok = code.decode(lch[:Fa], 20, algo=a)
        assert (ok == o_ok).all() and (it == o_it).all(), (name, a)
        assert (out.astype(np.int32) == o_out).all(), (name, a)
        # fused demapper gives the same result as the oracle's Demodulate
        inp, kind = channel_input(cfg, rxs[:Fa], N, p)
        out2, it2, ok2 = code.decode(inp, 20, algo=a, in_kind=kind, sigma=sigma)
        assert (out2 == out).all() and (it2 == it).all() and (ok2 == ok).all()
        converged += int(o_ok.sum())
        assert (o_out[o_ok == 1] == 0).sum() >= 0.9 * o_out[o_ok == 1].size  # all-zero word was sent
    assert converged > 0


def test_device_path_and_errors(gf_dir, meta):
    import torch
    cfg = meta["configs"]["BDS"]
    mt, gf, cs = paths(cfg, gf_dir)
    code = m.NbLdpcCode(mt, gf, cs)
    g = np.load(os.path.join(GOLDEN, "nb_BDS.npz"))
    inp, kind = channel_input(cfg, g["rx"], code.N, code.p)
    big = torch.from_numpy(np.tile(inp, (40, 1))).cuda()  # 480 frames
    out, it, ok = code.decode(big, 20, algo=m.ALGO_LAYERED_TMM, in_kind=kind, sigma=float(g["sigma"]))
    torch.cuda.synchronize()
    out = out.cpu().numpy().astype(np.int32).reshape(40, len(inp), -1)
    assert (out == g["out_3"][None]).all() and (it.cpu().numpy().reshape(40, -1) == g["iter_3"][None]).all()
    # a code read with raw exponent coefficients containing 0 has no h^-1: TMM refuses it
    c4 = meta["configs"]["C4"]
    mt4, gf4, cs4 = paths(c4, gf_dir)
    raw = m.NbLdpcCode(mt4, gf4, cs4, coef_is_exponent=False)
    with pytest.raises(m.LdpcError) as e:
        raw.decode(np.zeros((1, raw.N * 2), np.float32), 5, algo=m.ALGO_TMM, in_kind=m.IN_QAM, sigma=0.2)
    assert e.value.code == -6


def test_nb_channel_statistic_and_cli(nb_oracle, gf_dir, meta):
    """Device transmitter + channel (BPSK and 64-QAM), statistic kernel and the nb_ldpc_sim driver:
    noise statistics, shard independence, counter formulas of NB Statistic, and an SER / FER at 3 dB
    (BDS, layered TMM) consistent with the oracle decoding an independent sample."""
    import subprocess
    import torch
    from conftest import ROOT
    cfg = meta["configs"]["BDS"]
    mt, gf, cs = paths(cfg, gf_dir)
    code = m.NbLdpcCode(mt, gf, cs)
    cw = np.array(meta["CodeWord_sym_test"], np.uint16)
    cwd = torch.as_tensor(cw.astype(np.int16), device="cuda")
    F, L = 1024, code.N * code.p
    st = torch.cuda.current_stream().cuda_stream
    sigma = m.lib.nb_ldpc_sigma(code._h, 0, 3.0, 0)
    h = orc_load(nb_oracle, cfg, gf_dir, 0)
    assert sigma == pytest.approx(nb_oracle.nb_orc_sigma(h, 0, 3.0), rel=1e-7)

    def gen(F, first):
        x = torch.empty(F * L, dtype=torch.float32, device="cuda")
        assert m.lib.nb_ldpc_modulate_awgn(code._h, x.data_ptr(), F, sigma, 173, first, cwd.data_ptr(), st) >= 0
        torch.cuda.synchronize()
        return x.view(F, L)
    x = gen(F, 0)
    bits = ((cw[:, None] >> np.arange(code.p)) & 1).reshape(-1)
    n = ((x.cpu().numpy() - (1.0 - 2.0 * bits)[None]) / sigma)
    assert abs(n.mean()) < 0.01 and abs(n.var() - 1) < 0.01 and abs((n ** 4).mean() - 3) < 0.1
    assert (torch.cat([gen(300, 0), gen(724, 300)]) == x).all()
    out, it, ok = code.decode(x, 20, algo=m.ALGO_LAYERED_TMM, in_kind=m.IN_BPSK, sigma=sigma)
    cnt = torch.zeros(6, dtype=torch.int64, device="cuda")
    assert m.lib.nb_ldpc_statistic(code._h, out.data_ptr(), it.data_ptr(), ok.data_ptr(), F, cwd.data_ptr(),
                                   cnt.data_ptr(), st) >= 0
    o, i_, k_ = out.cpu().numpy().astype(np.int32), it.cpu().numpy(), ok.cpu().numpy()
    err = (o != cw[None].astype(np.int32)).sum(1)
    want = [F, (err != 0).sum(), err.sum(), i_.sum(), ((err != 0) & (k_ == 1)).sum(), ((err == 0) & (k_ == 0)).sum()]
    assert cnt.cpu().numpy().tolist() == [int(v) for v in want]
    # oracle FER on an independent sample of the same channel
    Fo = 256
    lch = np.zeros((Fo, code.N * (code.q - 1)), np.float32)
    tx = np.zeros(2 * L, np.float32)
    sym = cw.astype(np.int32)
    nb_oracle.nb_orc_modulate(h, sym.ctypes.data, tx.ctypes.data)
    seed = np.array([7, 11, 13], np.int32)
    rx = np.zeros(2 * L, np.float32)
    for f in range(Fo):
        nb_oracle.nb_orc_awgn(seed.ctypes.data, sigma, tx.ctypes.data, rx.ctypes.data, L)
        nb_oracle.nb_orc_demodulate(h, sigma, rx.ctypes.data, lch[f].ctypes.data)
    oo = np.zeros((Fo, code.N), np.int32); oi = np.zeros(Fo, np.int32); ok_o = np.zeros(Fo, np.int32)
    nb_oracle.nb_orc_decode_batch(h, 3, 0, lch.ctypes.data, Fo, 20, 2, 2, oo.ctypes.data, oi.ctypes.data, ok_o.ctypes.data)
    fer_o = ((oo != sym[None]).sum(1) != 0).mean()
    fer_e = want[1] / F
    se = np.sqrt(max(fer_o * (1 - fer_o), 1e-3) / Fo + max(fer_e * (1 - fer_e), 1e-3) / F)
    assert abs(fer_e - fer_o) < 4 * se + 0.01, (fer_e, fer_o)
    # 64-QAM transmitter: mean of the received points = the constellation point of the sent symbol
    c4 = meta["configs"]["C4"]
    mt4, gf4, cs4 = paths(c4, gf_dir)
    qam = m.NbLdpcCode(mt4, gf4, cs4, coef_is_exponent=True)
    y = torch.empty(2048 * qam.N * 2, dtype=torch.float32, device="cuda")
    assert m.lib.nb_ldpc_modulate_awgn(qam._h, y.data_ptr(), 2048, 0.1, 5, 0, None, st) >= 0
    pts = np.loadtxt(cs4, usecols=(3, 5))
    mean = y.view(2048, qam.N, 2).mean((0, 1)).cpu().numpy()
    assert np.allclose(mean, pts[0], atol=2e-3)
    # driver
    exe = os.path.join(ROOT, "cuda_ldpc_b200", "nb_ldpc_sim")
    r = subprocess.run([exe, "--matrix", mt, "--gf", gf, "--constellation", cs, "--algo", "ltmm", "--snr", "2.0", "3.0", "1.0",
                        "--batch", "1024", "--max-frames", "8192", "--codeword", os.path.join(GOLDEN, "nb_BDS_codeword.txt")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    rows = [l.split() for l in r.stdout.splitlines() if l.startswith(" 2.0") or l.startswith(" 3.0")]
    assert len(rows) == 2 and float(rows[0][3]) > float(rows[1][3]) > 0
    assert abs(float(rows[1][3]) - fer_e) < 0.02  # same point (3 dB), same decoder


@pytest.mark.parametrize("name,exp,snr,F", [("BDS", 0, 2.0, 48), ("C5", 1, 3.8, 32), ("C4", 1, 8.5, 32)])
def test_fft_bp_vs_oracle(nb_oracle, gf_dir, meta, name, exp, snr, F):
    """FFT-BP (Walsh-Hadamard check node) is NOT in the reference ("parity unpinned"): the kernel is held
    to the oracle's specification — same decisions / iteration counts (expf differs in the last ulp
    between libm and CUDA, so a marginal frame may differ), and it decodes at least as many frames as
    the reference's TMM on the same noise."""
    cfg = meta["configs"][name]
    mt, gf, cs = paths(cfg, gf_dir)
    h = orc_load(nb_oracle, cfg, gf_dir, exp)
    code = m.NbLdpcCode(mt, gf, cs, coef_is_exponent=bool(exp))
    N, q, p = code.N, code.q, code.p
    L = N * p if cfg["n_qam"] == 2 else N
    sym = np.zeros(N, np.int32)
    tx = np.zeros(2 * L, np.float32)
    nb_oracle.nb_orc_modulate(h, sym.ctypes.data, tx.ctypes.data)
    sigma = nb_oracle.nb_orc_sigma(h, 0, snr)
    seed = np.array([173, 173, 173], np.int32)
    lch = np.zeros((F, N * (q - 1)), np.float32)
    rx = np.zeros(2 * L, np.float32)
    for f in range(F):
        nb_oracle.nb_orc_awgn(seed.ctypes.data, sigma, tx.ctypes.data, rx.ctypes.data, L)
        nb_oracle.nb_orc_demodulate(h, sigma, rx.ctypes.data, lch[f].ctypes.data)
    o_out = np.zeros((F, N), np.int32); o_it = np.zeros(F, np.int32); o_ok = np.zeros(F, np.int32)
    nb_oracle.nb_orc_decode_batch(h, 4, 0, lch.ctypes.data, F, 20, 2, 2, o_out.ctypes.data, o_it.ctypes.data, o_ok.ctypes.data)
    out, it, ok = code.decode(lch, 20, algo=m.ALGO_FFT_BP)
    same = (out.astype(np.int32) == o_out).all(1) & (it == o_it) & (ok == o_ok)
    assert same.sum() >= F - 1, (name, int(same.sum()))
    assert same[o_ok == 1].all()
    out_t, it_t, ok_t = code.decode(lch, 20, algo=m.ALGO_TMM)
    assert ok.sum() >= ok_t.sum() and ok.sum() >= F // 2


@pytest.mark.parametrize("matrix,q,exp,snr,F", [("Tanner_74_9_Z128_GF16.txt", 16, 0, 5.5, 3),
                                                 ("LDPC_N96_K48_GF256_d1_exp.txt", 256, 1, 4.0, 12)])
def test_remaining_shipped_matrices(nb_oracle, gf_dir, matrix, q, exp, snr, F):
    """The two NB files of the reference no BASELINE config names: the only large code
    (N = 9472 symbols over GF(16), dc 20-21) and the short GF(256) one, BPSK, all three decoders."""
    cfg = {"matrix": matrix, "q": q, "constellation": "Constellation/BPSK.txt", "n_qam": 2}
    mt, gf, cs = paths(cfg, gf_dir)
    h = orc_load(nb_oracle, cfg, gf_dir, exp)
    code = m.NbLdpcCode(mt, None, cs, coef_is_exponent=bool(exp))
    N, p = code.N, code.p
    L = N * p
    sym = np.zeros(N, np.int32)
    tx = np.zeros(2 * L, np.float32)
    nb_oracle.nb_orc_modulate(h, sym.ctypes.data, tx.ctypes.data)
    sigma = nb_oracle.nb_orc_sigma(h, 0, snr)
    seed = np.array([173, 173, 173], np.int32)
    lch = np.zeros((F, N * (q - 1)), np.float32)
    rx = np.zeros(2 * L, np.float32)
    for f in range(F):
        nb_oracle.nb_orc_awgn(seed.ctypes.data, sigma, tx.ctypes.data, rx.ctypes.data, L)
        nb_oracle.nb_orc_demodulate(h, sigma, rx.ctypes.data, lch[f].ctypes.data)
    for a, summode in [(m.ALGO_EMS, 1), (m.ALGO_TMM, 0), (m.ALGO_LAYERED_TMM, 0)]:
        o_out = np.zeros((F, N), np.int32); o_it = np.zeros(F, np.int32); o_ok = np.zeros(F, np.int32)
        nb_oracle.nb_orc_decode_batch(h, a, summode, lch.ctypes.data, F, 8, 2, 2, o_out.ctypes.data,
                                      o_it.ctypes.data, o_ok.ctypes.data)
        out, it, ok = code.decode(lch, 8, algo=a)
        assert (ok == o_ok).all() and (it == o_it).all(), (matrix, a)
        assert (out.astype(np.int32) == o_out).all(), (matrix, a)


@pytest.mark.parametrize("nm,nc", [(2, 1), (3, 2), (3, 3), (4, 2), (2, 3)])
def test_ems_configuration_set_variants(nb_oracle, gf_dir, meta, nm, nc):
    """conf(Nm,Nc) for other budgets than the reference's (2,2) default (NB/include/define.h:31-32):
    pairs-only enumeration (Nc <= 2) and the general odometer (Nc = dc-1 = 3 here) against the oracle."""
    cfg = meta["configs"]["BDS"]
    mt, gf, cs = paths(cfg, gf_dir)
    h = orc_load(nb_oracle, cfg, gf_dir, 0)
    code = m.NbLdpcCode(mt, None, cs, coef_is_exponent=False)
    N, q, p = code.N, code.q, code.p
    F = 24
    sym = np.zeros(N, np.int32)
    L = N * p
    tx = np.zeros(2 * L, np.float32)
    nb_oracle.nb_orc_modulate(h, sym.ctypes.data, tx.ctypes.data)
    sigma = nb_oracle.nb_orc_sigma(h, 0, 2.5)
    seed = np.array([173, 173, 173], np.int32)
    lch = np.zeros((F, N * (q - 1)), np.float32)
    rx = np.zeros(2 * L, np.float32)
    for f in range(F):
        nb_oracle.nb_orc_awgn(seed.ctypes.data, sigma, tx.ctypes.data, rx.ctypes.data, L)
        nb_oracle.nb_orc_demodulate(h, sigma, rx.ctypes.data, lch[f].ctypes.data)
    o_out = np.zeros((F, N), np.int32); o_it = np.zeros(F, np.int32); o_ok = np.zeros(F, np.int32)
    nb_oracle.nb_orc_decode_batch(h, m.ALGO_EMS, 1, lch.ctypes.data, F, 20, nm, nc, o_out.ctypes.data,
                                  o_it.ctypes.data, o_ok.ctypes.data)
    out, it, ok = code.decode(lch, 20, algo=m.ALGO_EMS, ems_nm=nm, ems_nc=nc)
    assert (ok == o_ok).all() and (it == o_it).all()
    assert (out.astype(np.int32) == o_out).all()
    assert o_ok.sum() > 0


def test_log_qspa_full_configuration_set(nb_oracle, gf_dir, meta):
    """decoder_method 2 of the reference = Decoding_EMS(..., GFQ, maxdc - 1, ...) (NB/src/Simulation.cpp:64-67):
    every one of the q^(dc-1) configurations.  The kernel's (max,+) forward recursion must reproduce the
    oracle's explicit enumeration bit for bit (3 iterations of 2 frames: the enumeration is 5e7 leaves per
    frame and iteration)."""
    cfg = meta["configs"]["BDS"]
    mt, gf, cs = paths(cfg, gf_dir)
    h = orc_load(nb_oracle, cfg, gf_dir, 0)
    code = m.NbLdpcCode(mt, None, cs, coef_is_exponent=False)
    N, q, p = code.N, code.q, code.p
    F, L = 2, N * p
    sym = np.zeros(N, np.int32)
    tx = np.zeros(2 * L, np.float32)
    nb_oracle.nb_orc_modulate(h, sym.ctypes.data, tx.ctypes.data)
    sigma = nb_oracle.nb_orc_sigma(h, 0, 1.5)
    seed = np.array([173, 173, 173], np.int32)
    lch = np.zeros((F, N * (q - 1)), np.float32)
    rx = np.zeros(2 * L, np.float32)
    for f in range(F):
        nb_oracle.nb_orc_awgn(seed.ctypes.data, sigma, tx.ctypes.data, rx.ctypes.data, L)
        nb_oracle.nb_orc_demodulate(h, sigma, rx.ctypes.data, lch[f].ctypes.data)
    dc_max = 4
    o_out = np.zeros((F, N), np.int32); o_it = np.zeros(F, np.int32); o_ok = np.zeros(F, np.int32)
    nb_oracle.nb_orc_decode_batch(h, m.ALGO_EMS, 1, lch.ctypes.data, F, 3, q, dc_max - 1, o_out.ctypes.data,
                                  o_it.ctypes.data, o_ok.ctypes.data)
    out, it, ok = code.decode(lch, 3, algo=m.ALGO_EMS, ems_nm=q, ems_nc=dc_max - 1)
    assert (ok == o_ok).all() and (it == o_it).all()
    assert (out.astype(np.int32) == o_out).all()
    # and it decodes: at a comfortable SNR the full set converges at least as often as EMS(2,2)
    sigma2 = nb_oracle.nb_orc_sigma(h, 0, 2.5)
    rxs = np.zeros((32, L), np.float32)
    for f in range(32):
        nb_oracle.nb_orc_awgn(seed.ctypes.data, sigma2, tx.ctypes.data, rx.ctypes.data, L)
        rxs[f] = rx[0::2]
    _, _, ok_full = code.decode(rxs, 20, algo=m.ALGO_EMS, in_kind=m.IN_BPSK, sigma=sigma2, ems_nm=q, ems_nc=dc_max - 1)
    _, _, ok_22 = code.decode(rxs, 20, algo=m.ALGO_EMS, in_kind=m.IN_BPSK, sigma=sigma2)
    assert ok_full.sum() >= ok_22.sum() - 1 and ok_full.sum() >= 16


@pytest.mark.parametrize("name,exp,snr,algo", [("C4", 1, 13.0, "ALGO_EMS"), ("C4", 1, 13.0, "ALGO_LAYERED_TMM"),
                                               ("C5", 1, 7.0, "ALGO_TMM"), ("C5", 1, 7.0, "ALGO_FFT_BP")])
def test_large_batch_properties(meta, name, exp, snr, algo):
    """Bench-sized batches (device path, Philox channel): at a comfortable SNR every frame of the all-zero word
    converges to the all-zero word, the result does not depend on how the batch is cut into calls, and a second
    decode of the same samples is identical (no state leaks between frames or calls)."""
    import torch
    cfg = meta["configs"][name]
    code = m.NbLdpcCode(os.path.join(NB_DATA, cfg["matrix"]), None,
                        os.path.join(NB_DATA, cfg["constellation"].replace("./", "")), coef_is_exponent=bool(exp))
    F = 8192 if name == "C4" else 2048
    kind = m.IN_BPSK if cfg["n_qam"] == 2 else m.IN_QAM
    per = code.in_elems(kind)
    sigma = m.lib.nb_ldpc_sigma(code._h, 0, snr, 0)
    x = torch.empty(F * per, dtype=torch.float32, device="cuda")
    rc = m.lib.nb_ldpc_modulate_awgn(code._h, x.data_ptr(), F, sigma, 7, 0, None, torch.cuda.current_stream().cuda_stream)
    assert rc >= 0
    x = x.view(F, per)
    a = getattr(m, algo)
    out, it, ok = code.decode(x, 20, algo=a, in_kind=kind, sigma=sigma)
    assert int(ok.sum()) == F and int(out.abs().sum()) == 0
    out2, it2, ok2 = code.decode(x, 20, algo=a, in_kind=kind, sigma=sigma)
    assert torch.equal(out, out2) and torch.equal(it, it2)
    h = F // 2 + 37  # ragged split
    o_a, i_a, _ = code.decode(x[:h].contiguous(), 20, algo=a, in_kind=kind, sigma=sigma)
    o_b, i_b, _ = code.decode(x[h:].contiguous(), 20, algo=a, in_kind=kind, sigma=sigma)
    assert torch.equal(torch.cat([o_a, o_b]), out) and torch.equal(torch.cat([i_a, i_b]), it)


@pytest.mark.parametrize("name,exp,snr,F", [("C4", 1, 12.5, 24), ("C5", 1, 5.0, 16)])
def test_encoded_codewords_through_channel_and_decoder(nb_oracle, gf_dir, meta, name, exp, snr, F):
    """nb_ldpc_encode -> oracle Modulate + AWGNChannel_CPU (NB/src/LDPC_Encoder.cpp:18-68) -> fused demapper + TMM /
    layered TMM on the GPU == the oracle's decoder on the same samples, for a RANDOM codeword (the reference can only
    send zeros for these matrices); converged frames return exactly the transmitted codeword; and the device
    transmitter (nb_ldpc_modulate_awgn with a codeword) + nb_ldpc_statistic count zero errors at high SNR."""
    import torch
    cfg = meta["configs"][name]
    mt, gf, cs = paths(cfg, gf_dir)
    h = orc_load(nb_oracle, cfg, gf_dir, exp)
    code = m.NbLdpcCode(mt, None, cs, coef_is_exponent=bool(exp))
    N, q, p = code.N, code.q, code.p
    pos = code.info_positions()
    cw = code.encode(np.random.default_rng(5).integers(0, q, len(pos)).astype(np.uint16))
    sym = cw.astype(np.int32)
    assert nb_oracle.nb_orc_syndrome_ok(h, sym.ctypes.data) == 1
    L = N * p if cfg["n_qam"] == 2 else N
    tx = np.zeros(2 * L, np.float32)
    nb_oracle.nb_orc_modulate(h, sym.ctypes.data, tx.ctypes.data)
    sigma = nb_oracle.nb_orc_sigma(h, 0, snr)
    seed = np.array([173, 173, 173], np.int32)
    lch = np.zeros((F, N * (q - 1)), np.float32)
    rxs = np.zeros((F, 2 * L), np.float32)
    for f in range(F):
        nb_oracle.nb_orc_awgn(seed.ctypes.data, sigma, tx.ctypes.data, rxs[f].ctypes.data, L)
        nb_oracle.nb_orc_demodulate(h, sigma, rxs[f].ctypes.data, lch[f].ctypes.data)
    inp, kind = channel_input(cfg, rxs, N, p)
    for a in (m.ALGO_TMM, m.ALGO_LAYERED_TMM):
        o_out = np.zeros((F, N), np.int32); o_it = np.zeros(F, np.int32); o_ok = np.zeros(F, np.int32)
        nb_oracle.nb_orc_decode_batch(h, a, 0, lch.ctypes.data, F, 20, 2, 2, o_out.ctypes.data, o_it.ctypes.data,
                                      o_ok.ctypes.data)
        out, it, ok = code.decode(inp, 20, algo=a, in_kind=kind, sigma=sigma)
        assert (ok == o_ok).all() and (it == o_it).all() and (out.astype(np.int32) == o_out).all(), (name, a)
        assert o_ok.sum() >= F // 2 and (out[ok == 1] == cw[None, :]).all()
    # device transmitter with the codeword, far above the waterfall: every frame decodes to it, the counters say so
    cwd = torch.as_tensor(cw.astype(np.int16), device="cuda")
    hi_sigma = code.sigma(0, snr + 6.0)
    x = code.modulate_awgn(2048, hi_sigma, seed=3, codeword=cwd)
    out, it, ok = code.decode(x.view(2048, -1), 20, algo=m.ALGO_LAYERED_TMM, in_kind=kind, sigma=hi_sigma)
    torch.cuda.synchronize()
    assert (out.view(torch.int16) == cwd[None, :]).all() and ok.all()
    cnt = torch.zeros(6, dtype=torch.int64, device="cuda")
    assert m.lib.nb_ldpc_statistic(code._h, out.data_ptr(), it.data_ptr(), ok.data_ptr(), 2048, cwd.data_ptr(),
                                   cnt.data_ptr(), torch.cuda.current_stream().cuda_stream) >= 0
    c = cnt.cpu().numpy()
    assert c[0] == 2048 and c[1] == 0 and c[2] == 0
