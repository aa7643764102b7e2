"""GPU parity tests of the binary decode path, through the C-ABI (ctypes), against the CPU oracle
and the golden vectors generated from the reference's own sources."""
import os

import numpy as np
import pytest

from conftest import DATA, GOLDEN, OracleCode, ip, fp

import cuda_ldpc_b200 as m

pytestmark = pytest.mark.gpu
BL = os.path.join(DATA, "bldpc")

CODES = {
    "C1": ("J4_L24_Z96_BlockH.txt", 4, 24, 96),
    "C2": ("J15_L30_Z1280_BlockH.txt", 15, 30, 1280),
    "C3": ("PON_LDPC.txt", 12, 69, 256),
    "J32": ("J32_L64_Z64_BlockH.txt", 32, 64, 64),
    "J10": ("J10_L60_Z160_BlockH.txt", 10, 60, 160),
    "J6": ("J6_L24_Z96_BlockH.txt", 6, 24, 96),
}


def load(oracle, key):
    f, J, L, Z = CODES[key]
    path = os.path.join(BL, f)
    return m.LdpcCode(path, J, L, Z), OracleCode(oracle, path, J, L, Z, literal=0)


def noisy(oracle, N, F, snr_db, seed=173, rate=1.0, snrtype=1):
    s = np.array([seed, seed, seed], np.int32)
    y = np.zeros(N * F, np.float32)
    oracle.orc_awgn(ip(s), oracle.orc_sigma(snrtype, snr_db, rate), None, fp(y), N, F)
    return y.reshape(N, F)


def orc_flood(oracle, oc, y, maxit, mode):
    N, F = y.shape
    D = np.zeros((N + 1) * F, np.int32)
    it = np.zeros(F, np.int32)
    rq = np.zeros(oc.M * int(oc.Wc[oc.J]) * F, np.float32)
    rc = oracle.orc_flooding_fp32(oc.J, oc.L, oc.Z, ip(oc.H), ip(oc.Wc), ip(oc.Wv), ip(oc.addr),
                                  fp(np.ascontiguousarray(y)), F, maxit, mode, oc.K, ip(D), ip(it), rq.ctypes.data)
    assert rc == 0
    return D.reshape(N + 1, F), it, rq


def orc_i8(oracle, oc, y, maxit, mode, scale=8.0, amax=127, bnum=0, bshift=0):
    N, F = y.shape
    D = np.zeros((N + 1) * F, np.int32)
    it = np.zeros(F, np.int32)
    app = np.zeros(N * F, np.int8)
    rec = np.zeros(oc.M * 4 * F, np.uint32)
    rc = oracle.orc_layered_i8(oc.J, oc.L, oc.Z, ip(oc.H), fp(np.ascontiguousarray(y)), F, maxit, scale, amax, bnum,
                               bshift, mode, ip(D), ip(it), app.ctypes.data, rec.ctypes.data)
    assert rc == 0
    return D.reshape(N + 1, F), it, app.reshape(N, F), expand_messages(rec, oc, F)


def expand_messages(rec, oc, F):
    """The oracle keeps one record {min1, min2, idx, sign bits} per check; the message of edge k is
    (sign bit k ? -1 : +1) * (k == idx ? min2 : min1).  Returns int8 [M * dc_max * F] in the layout of
    ldpc_decode_opts_t.debug_msgs (0 for absent edges)."""
    M, dcm = oc.M, int(oc.Wc[oc.J])
    r = rec.reshape(M, 4, F).astype(np.int64)
    dc_row = np.repeat((oc.H.reshape(oc.J, oc.L) >= 0).sum(1), oc.Z)          # [M]
    k = np.arange(dcm)[None, :, None]
    mag = np.where(k == r[:, 2][:, None, :], r[:, 1][:, None, :], r[:, 0][:, None, :])
    neg = (r[:, 3][:, None, :] >> k) & 1
    msg = np.where(neg == 1, -mag, mag) * (k < dc_row[:, None, None])
    return msg.astype(np.int8).reshape(-1)


# ------------------------------------------------------------------ flooding fp32 (strict parity)

def test_flooding_matches_reference_golden_batch(oracle):
    """The reference's own LDPC_Decoder_GPU output (one-line-fixed Transform_H) on 16 frames."""
    batch = np.load(os.path.join(GOLDEN, "binary_C1_batch16.npz"))
    code, _ = load(oracle, "C1")
    y = np.ascontiguousarray(batch["y"])
    r = code.decode(y, 10, early_exit=m.EXIT_GENIE)
    want = np.unpackbits(batch["D_fixed"], axis=0)[: code.N].astype(np.int32)
    assert (r.D[: code.N] == want).all()
    assert (r.D[code.N] == batch["flag_fixed"]).all()
    assert (r.iters == int(batch["iters_fixed"])).all()
    assert r.launches >= 20


@pytest.mark.parametrize("key,F,snr", [("C1", 64, 3.0), ("C1", 37, 2.5), ("C3", 16, 3.5), ("C2", 8, -0.5),
                                       ("J32", 24, 1.0), ("J10", 12, 4.0)])
@pytest.mark.parametrize("mode", [m.EXIT_NONE, m.EXIT_GENIE, m.EXIT_SYNDROME])
def test_flooding_bit_exact_vs_oracle(oracle, key, F, snr, mode):
    code, oc = load(oracle, key)
    y = noisy(oracle, code.N, F, snr)
    r = code.decode(y, 10, early_exit=mode, debug=True)
    D, it, rq = orc_flood(oracle, oc, y, 10, mode)
    assert (r.D == D).all()
    assert (r.iters == it).all()
    if mode != m.EXIT_SYNDROME or it.max() == 10:
        # every message, bit for bit (after an early syndrome exit the oracle stops at the same pass)
        assert (r.msgs.view(np.uint32) == rq.view(np.uint32)).all()


def test_flooding_input_variants(oracle):
    """fp16 / int8 inputs and the [F][N] layout give the result of the converted fp32 [N][F] input."""
    code, oc = load(oracle, "C1")
    F = 20
    y = noisy(oracle, code.N, F, 3.0)
    y16 = y.astype(np.float16)
    base = code.decode(y16.astype(np.float32), 8)
    r16 = code.decode(y16, 8)
    assert (r16.D == base.D).all()
    rfn = code.decode(np.ascontiguousarray(y16.astype(np.float32).T), 8, layout=m.LAYOUT_FN, out_format=m.OUT_U8)
    assert (rfn.D.T == base.D[: code.N]).all()
    q = np.clip(np.rint(y * 8), -127, 127).astype(np.int8)
    r8 = code.decode(q, 8, out_format=m.OUT_BITPACK)
    b8 = code.decode(q.astype(np.float32), 8)
    bits = ((r8.D[:, :, None] >> np.arange(32)[None, None, :]) & 1).reshape(F, -1)[:, : code.N]
    assert (bits.T == b8.D[: code.N]).all()


# ------------------------------------------------------------------ layered int8 (throughput mode)

@pytest.mark.parametrize("key,F,snr,it", [("C1", 64, 3.0, 10), ("C1", 13, 2.0, 5), ("C2", 16, -0.5, 10),
                                          ("C2", 6, 0.5, 3), ("C3", 12, 3.0, 10), ("J32", 20, 0.5, 6),
                                          ("J10", 8, 4.0, 6), ("J6", 32, 2.0, 8)])
@pytest.mark.parametrize("mode", [m.EXIT_NONE, m.EXIT_SYNDROME])
def test_layered_i8_bit_exact_vs_oracle(oracle, key, F, snr, it, mode):
    code, oc = load(oracle, key)
    y = noisy(oracle, code.N, F, snr)
    r = code.decode(y, it, schedule=m.SCHED_LAYERED, early_exit=mode, debug=True)
    D, its, app, rec = orc_i8(oracle, oc, y, it, mode)
    assert (r.iters == its).all()
    assert (r.ok == D[code.N]).all()
    assert (r.app == app).all(), "APP values differ"
    assert (r.msgs == rec).all(), "check-to-variable messages differ"
    assert (r.D == D).all()


@pytest.mark.parametrize("bnum,bshift,amax,scale", [(1, 2, 127, 8.0), (1, 3, 63, 4.0), (3, 4, 127, 16.0),
                                                    (5, 5, 31, 8.0), (0, 0, 15, 2.0)])
def test_layered_i8_scaling_and_clipping_variants(oracle, bnum, bshift, amax, scale):
    code, oc = load(oracle, "C1")
    y = noisy(oracle, code.N, 32, 2.5)
    r = code.decode(y, 7, schedule=m.SCHED_LAYERED, debug=True, beta_num=bnum, beta_shift=bshift, msg_max=amax,
                    llr_scale=scale)
    D, its, app, rec = orc_i8(oracle, oc, y, 7, 0, scale, amax, bnum, bshift)
    assert (r.app == app).all() and (r.msgs == rec).all() and (r.D == D).all()


@pytest.mark.parametrize("seed", range(6))
def test_layered_i8_randomised_settings(oracle, seed):
    """Seeded sweep over code, quantiser scale, message clip, beta and SNR (heavy saturation included): APP values,
    check records (min1, min2, first index, signs), hard bits and iteration counts against the oracle."""
    rng = np.random.default_rng(1000 + seed)
    key = ["C1", "C3", "J32", "J6", "J10", "C2"][seed % 6]
    code, oc = load(oracle, key)
    for _ in range(3):
        bshift = int(rng.integers(1, 6))
        bnum = int(rng.integers(0, min(8, (1 << bshift) - 1) + 1))
        amax = int(rng.choice([1, 3, 7, 15, 31, 63, 100, 127]))
        scale = float(rng.choice([0.5, 2.0, 8.0, 16.0, 64.0]))  # 64: almost every LLR saturates
        snr = float(rng.uniform(-1.0, 4.0))
        it = int(rng.integers(1, 7))
        mode = int(rng.choice([m.EXIT_NONE, m.EXIT_SYNDROME]))
        F = int(rng.choice([3, 8, 17])) if key != "C2" else 5
        y = noisy(oracle, code.N, F, snr)
        r = code.decode(y, it, schedule=m.SCHED_LAYERED, early_exit=mode, debug=True, beta_num=bnum, beta_shift=bshift,
                        msg_max=amax, llr_scale=scale)
        D, its, app, rec = orc_i8(oracle, oc, y, it, mode, scale, amax, bnum, bshift)
        tag = (key, bnum, bshift, amax, scale, snr, it, mode, F)
        assert (r.iters == its).all(), tag
        assert (r.app == app).all(), tag
        assert (r.msgs == rec).all(), tag
        assert (r.D == D).all(), tag


def test_layered_i8_io_variants(oracle):
    code, oc = load(oracle, "C3")
    F = 10
    y = noisy(oracle, code.N, F, 3.0)
    D, its, app, rec = orc_i8(oracle, oc, y, 6, 2)
    r = code.decode(np.ascontiguousarray(y.T), 6, schedule=m.SCHED_LAYERED, layout=m.LAYOUT_FN, early_exit=2,
                    out_format=m.OUT_U8)
    assert (r.D.T == D[: code.N]).all() and (r.iters == its).all()
    r = code.decode(y, 6, schedule=m.SCHED_LAYERED, early_exit=2, out_format=m.OUT_BITPACK)
    bits = ((r.D[:, :, None] >> np.arange(32)[None, None, :]) & 1).reshape(F, -1)[:, : code.N]
    assert (bits.T == D[: code.N]).all()
    q = np.clip(np.rint(y * 8), -127, 127).astype(np.int8)
    r = code.decode(q, 6, schedule=m.SCHED_LAYERED, early_exit=2, out_format=m.OUT_U8)
    assert (r.D == D[: code.N]).all()


@pytest.mark.parametrize("F", [12, 13])
def test_layered_i8_fp16_and_int8_inputs_match_fp32(oracle, F):
    """fp16 and int8 channel values in the [N][F] layout (vector fast paths when F is a multiple of 4, scalar
    path otherwise), host and device buffers: same decode as the fp32 input holding the same values."""
    import torch
    code, oc = load(oracle, "C1")
    y16 = noisy(oracle, code.N, F, 2.8).astype(np.float16)
    kw = dict(schedule=m.SCHED_LAYERED, early_exit=m.EXIT_SYNDROME, out_format=m.OUT_U8, msg_max=31, beta_num=1, beta_shift=3)
    want = code.decode(y16.astype(np.float32), 8, **kw)
    for got in (code.decode(y16, 8, **kw), code.decode(torch.as_tensor(y16, device="cuda"), 8, **kw)):
        D = got.D if isinstance(got.D, np.ndarray) else got.D.cpu().numpy()
        it = got.iters if isinstance(got.iters, np.ndarray) else got.iters.cpu().numpy()
        assert (D == want.D).all() and (it == want.iters).all()
    q = np.clip(np.rint(y16.astype(np.float32) * 8), -128, 127).astype(np.int8)  # -128 must behave like -127
    q[0, 0] = -128
    want8 = code.decode(np.maximum(q, -127).astype(np.float32) / 8, 8, **kw)
    for got in (code.decode(q, 8, **kw), code.decode(torch.as_tensor(q, device="cuda"), 8, **kw)):
        D = got.D if isinstance(got.D, np.ndarray) else got.D.cpu().numpy()
        assert (D == want8.D).all()


def test_device_path_and_large_batch_properties(oracle):
    """Device-resident I/O (torch tensors, no host copies) at a batch that fills the GPU; size-
    independent checks: decoded words satisfy H x = 0 when flagged ok, equal frames decode equally,
    and a slice re-decoded alone gives the same bits (no cross-frame leakage)."""
    import torch
    code, oc = load(oracle, "C2")
    F = 1200
    g = torch.Generator(device="cuda").manual_seed(7)
    sigma = m.sigma_from_snr(0, 2.2, code.rate)
    y = 1.0 + sigma * torch.randn(code.N, F, device="cuda", generator=g)
    y[:, 5] = y[:, 3]
    r = code.decode(y, 10, schedule=m.SCHED_LAYERED, early_exit=m.EXIT_SYNDROME, out_format=m.OUT_U8, msg_max=31)
    torch.cuda.synchronize()
    D = r.D.cpu().numpy().astype(np.int32)
    ok = r.ok.cpu().numpy()
    assert ok.mean() > 0.5
    assert (D[:, 5] == D[:, 3]).all()
    chk = np.zeros(F, np.int32)
    Dfull = np.concatenate([D, ok[None, :]], 0).astype(np.int32)
    oracle.orc_syndrome_ok(oc.J, oc.L, oc.Z, ip(oc.H), ip(np.ascontiguousarray(Dfull)), F, ip(chk))
    assert (chk == ok).all()
    assert (D[:, ok == 1] == 0).all()  # all-zero codeword was sent
    sub = code.decode(y[:, 40:52].contiguous(), 10, schedule=m.SCHED_LAYERED, early_exit=m.EXIT_SYNDROME,
                      out_format=m.OUT_U8, msg_max=31)
    assert (sub.D.cpu().numpy() == D[:, 40:52]).all()
    # oracle on a few frames of the big batch
    yo = y[:, :8].cpu().numpy()
    Do, its, _, _ = orc_i8(oracle, oc, yo, 10, 2, amax=31)
    assert (Do[: code.N] == D[:, :8]).all() and (its == r.iters.cpu().numpy()[:8]).all()


@pytest.mark.parametrize("layout,fmt", [(m.LAYOUT_NF, m.OUT_BITPACK), (m.LAYOUT_NF, m.OUT_INT32_REF),
                                        (m.LAYOUT_FN, m.OUT_U8), (m.LAYOUT_NF, m.OUT_U8)])
def test_chunked_host_path_equals_device_path(oracle, layout, fmt):
    """Large host batches are decoded in chunks on two internal streams (copy/compute overlap): same
    bits, iteration counts and flags as one device-resident call, for every layout / output format."""
    import torch
    code, _ = load(oracle, "C1")
    F = 3001  # not a multiple of the chunk (4*SMs*2) nor of 4
    y = noisy(oracle, code.N, 64, 2.8)
    y = np.tile(y, (1, 47))[:, :F] * (1.0 + 0.01 * np.arange(F, dtype=np.float32)[None, :] / F)
    yy = np.ascontiguousarray(y if layout == m.LAYOUT_NF else y.T)
    kw = dict(schedule=m.SCHED_LAYERED, early_exit=m.EXIT_SYNDROME, out_format=fmt, layout=layout, msg_max=31,
              beta_num=1, beta_shift=3)
    rh = code.decode(yy, 10, **kw)
    rd = code.decode(torch.as_tensor(yy, device="cuda"), 10, **kw)
    torch.cuda.synchronize()
    got = rd.D.cpu().numpy()
    assert (rh.D == (got.view(np.uint32) if fmt == m.OUT_BITPACK else got)).all()
    assert (rh.iters == rd.iters.cpu().numpy()).all() and (rh.ok == rd.ok.cpu().numpy()).all()
    assert rh.launches >= 3 and rd.launches == 1  # 3 equal chunks, or 4 in the hybrid pack / copy feed
    # host pack: the chunks quantised to int8 on host threads before the copy — same result bit for bit
    rp = code.decode(yy, 10, host_pack_threads=3, **kw)
    assert (rp.D == rh.D).all() and (rp.iters == rh.iters).all() and (rp.ok == rh.ok).all() and rp.launches >= 3


def test_host_pack_extreme_values(oracle):
    """Quantisation on the host must follow the kernel's rule everywhere: ties to even, saturation at +-127,
    huge and tiny values, and a batch that is a multiple of 4 per chunk edge case."""
    import torch
    code, _ = load(oracle, "C1")
    F = 2 * 4 * 148 * 2 + 8  # chunked path with a short last chunk
    rng = np.random.default_rng(5)
    y = (rng.standard_normal((code.N, F)) * 3.0).astype(np.float32)
    special = np.array([0.0625, -0.0625, 0.1875, -0.1875, 0.3125, 15.875, 15.9375, -15.9375, 1e9, -1e9, 1e-30, -1e-30,
                        15.8125, -15.8125, 0.0, -0.0], np.float32)
    y[:special.size, 0] = special
    y[1, : special.size] = special
    kw = dict(schedule=m.SCHED_LAYERED, out_format=m.OUT_BITPACK, msg_max=127, llr_scale=8.0, early_exit=m.EXIT_NONE)
    ra = code.decode(y, 3, **kw)
    rb = code.decode(y, 3, host_pack_threads=4, debug=False, **kw)
    assert (ra.D == rb.D).all() and (ra.ok == rb.ok).all()
    rd = code.decode(torch.as_tensor(y, device="cuda"), 3, **kw)
    assert (rd.D.cpu().numpy().view(np.uint32) == rb.D).all()


ALL_FILES = sorted(f for f in os.listdir(BL) if f.endswith(".txt"))


@pytest.mark.parametrize("fname", ALL_FILES)
def test_every_shipped_code_both_schedules(oracle, fname):
    """All 18 H files of the reference (dc 4..24, dv 2..15, Z 64..1280): flooding fp32 and layered int8
    bit-exact against the oracle on a small seeded batch (exercises every degree bucket, the exact-degree
    instances and the out-of-line generic path)."""
    import re
    geo = (12, 69, 256) if fname == "PON_LDPC.txt" else tuple(int(x) for x in re.match(r"J(\d+)_L(\d+)_Z(\d+)", fname).groups())
    path = os.path.join(BL, fname)
    code, oc = m.LdpcCode(path, *geo), OracleCode(oracle, path, *geo)
    F = 8 if code.N > 20000 else 12
    snr = 10 * np.log10(1.0 / (2 * code.rate)) + 3.0 + (1.5 if code.rate > 0.7 else 0.0)  # near the waterfall (Es/N0)
    y = noisy(oracle, code.N, F, snr)
    r = code.decode(y, 6, schedule=m.SCHED_LAYERED, early_exit=m.EXIT_SYNDROME, debug=True, msg_max=31, beta_num=1,
                    beta_shift=3)
    D, its, app, rec = orc_i8(oracle, oc, y, 6, 2, amax=31, bnum=1, bshift=3)
    assert (r.D == D).all() and (r.iters == its).all() and (r.app == app).all() and (r.msgs == rec).all(), fname
    r = code.decode(y, 4, debug=True)
    D, it, rq = orc_flood(oracle, oc, y, 4, 0)
    assert (r.D == D).all() and (r.msgs.view(np.uint32) == rq.view(np.uint32)).all(), fname


def orc_l32(oracle, oc, y, maxit, mode, alpha=1.0):
    N, F = y.shape
    D = np.zeros((N + 1) * F, np.int32)
    it = np.zeros(F, np.int32)
    app = np.zeros(N * F, np.float32)
    assert oracle.orc_layered_fp32(oc.J, oc.L, oc.Z, ip(oc.H), fp(np.ascontiguousarray(y)), F, maxit, alpha, mode,
                                   ip(D), ip(it), app.ctypes.data) == 0
    return D.reshape(N + 1, F), it, app.reshape(N, F)


@pytest.mark.parametrize("key,F,snr,alpha,mode", [("C1", 32, 2.8, 1.0, 0), ("C1", 21, 2.8, 0.8125, 2),
                                                  ("C3", 8, 3.2, 0.875, 2), ("C2", 8, 0.0, 1.0, 2), ("J10", 8, 4.0, 0.75, 0)])
def test_layered_fp32_bit_exact_vs_oracle(oracle, key, F, snr, alpha, mode):
    """The float twin of the throughput mode: every APP value (bit pattern), hard bit, iteration count
    and flag equals the oracle's layered fp32 (same operation order, no FMA contraction)."""
    code, oc = load(oracle, key)
    y = noisy(oracle, code.N, F, snr)
    r = code.decode(y, 8, schedule=m.SCHED_LAYERED, msg_dtype=m.DTYPE_FP32, early_exit=mode, alpha=alpha, debug=True)
    D, it, app = orc_l32(oracle, oc, y, 8, mode, alpha)
    assert (r.iters == it).all() and (r.D == D).all()
    assert (r.app.view(np.uint32) == app.view(np.uint32)).all()


def test_full_size_roundtrip_properties(oracle):
    """BASELINE.json's full size (J15_L30_Z1280, 9472 frames per GPU per step, as in bench.py), checked
    through size-independent properties: encode -> AWGN -> decode returns the transmitted codeword for
    every frame flagged ok; the flag equals an independent syndrome computation; permuting the frames
    permutes the results; the fused-channel path and the buffer path agree on the whole batch."""
    import torch
    code, oc = load(oracle, "C2")
    F = 148 * 4 * 16
    u = np.random.default_rng(11).integers(0, 2, code.K).astype(np.uint8)
    cw = np.zeros(code.N, np.uint8)
    assert m.lib.ldpc_encode(code.handle, u.ctypes.data, cw.ctypes.data) == 0
    cwd = torch.as_tensor(cw, device="cuda")
    sigma = m.sigma_from_snr(0, 2.4, code.rate)
    y = torch.empty(code.N * F, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert m.lib.ldpc_awgn_bpsk(code.handle, y.data_ptr(), F, m.LAYOUT_NF, sigma, 5, 0, cwd.data_ptr(), st) >= 0
    y = y.view(code.N, F)
    kw = dict(early_exit=m.EXIT_SYNDROME, out_format=m.OUT_U8, msg_max=31, beta_num=1, beta_shift=3)
    r = code.decode(y, 10, schedule=m.SCHED_LAYERED, **kw)
    ok = r.ok.bool()
    assert ok.float().mean() > 0.99
    assert (r.D[:, ok] == cwd[:, None]).all()                       # decoded = transmitted
    # syndrome of every decoded word, computed independently on the GPU with torch
    H = oc.H.reshape(oc.J, oc.L)
    D = r.D.view(oc.L, oc.Z, F)
    syn_bad = torch.zeros(F, dtype=torch.bool, device="cuda")
    for rr in range(oc.J):
        acc = torch.zeros(oc.Z, F, dtype=torch.uint8, device="cuda")
        for c in range(oc.L):
            if H[rr, c] >= 0:
                acc ^= torch.roll(D[c], -int(H[rr, c]), 0)         # row i reads column (i + s) mod Z
        syn_bad |= (acc != 0).any(0)
    assert (syn_bad == ~ok).all()
    perm = torch.randperm(F, device="cuda")
    rp = code.decode(y[:, perm].contiguous(), 10, schedule=m.SCHED_LAYERED, **kw)
    assert (rp.D == r.D[:, perm]).all() and (rp.iters == r.iters[perm]).all()
    rf = code.decode_channel(F, 10, sigma, seed=5, first_frame=0, codeword=cwd, **kw)
    assert (rf.D == r.D).all() and (rf.iters == r.iters).all() and (rf.ok == r.ok).all()


@pytest.mark.parametrize("mode", [m.EXIT_GENIE, m.EXIT_SYNDROME])
def test_flooding_device_path_exits_on_the_device(oracle, mode):
    """With device buffers ldpc_decode_batch only enqueues: the genie / syndrome stop is evaluated on the device
    (remaining iterations return at their first instruction).  Same bits, flags, iteration counts and messages as the
    host-buffer call (which may stop enqueuing early) and as the oracle."""
    import torch
    code, oc = load(oracle, "C1")
    y = noisy(oracle, code.N, 48, 4.2)   # every frame converges well before maxIT = 30
    rh = code.decode(y, 30, early_exit=mode, debug=True)
    rd = code.decode(torch.as_tensor(y, device="cuda"), 30, early_exit=mode, debug=True)
    torch.cuda.synchronize()
    D, it, rq = orc_flood(oracle, oc, y, 30, mode)
    assert it.max() < 30
    assert (rd.D.cpu().numpy() == D).all() and (rd.iters.cpu().numpy() == it).all()
    assert (rh.D == D).all() and (rh.iters == it).all()
    if mode == m.EXIT_GENIE:  # the whole batch freezes at the stop (syndrome mode freezes the decisions frame by frame,
        # the check-node pass of already latched frames keeps running in both paths)
        assert (rd.msgs.cpu().numpy().view(np.uint32) == rh.msgs.view(np.uint32)).all()
        assert (rd.msgs.cpu().numpy().view(np.uint32) == rq.view(np.uint32)).all()
    assert rd.launches > rh.launches   # the device path enqueued all 30 iterations, the host path stopped early
    for alpha in (1.0,):
        lh = code.decode(y, 30, schedule=m.SCHED_LAYERED, msg_dtype=m.DTYPE_FP32, early_exit=m.EXIT_SYNDROME, debug=True)
        ld = code.decode(torch.as_tensor(y, device="cuda"), 30, schedule=m.SCHED_LAYERED, msg_dtype=m.DTYPE_FP32,
                         early_exit=m.EXIT_SYNDROME, debug=True)
        torch.cuda.synchronize()
        assert (ld.D.cpu().numpy() == lh.D).all() and (ld.iters.cpu().numpy() == lh.iters).all()
        assert (ld.app.cpu().numpy().view(np.uint32) == lh.app.view(np.uint32)).all()


def test_one_handle_from_two_threads_and_streams(oracle):
    """include/ldpc_b200.h: calls on one handle are safe from any host thread / stream and serialise on its scratch
    arena.  Two threads hammer the same handle on their own streams; every result equals the single-threaded one."""
    import threading
    import torch
    code, _ = load(oracle, "C3")
    ys = [torch.as_tensor(noisy(oracle, code.N, 64, 3.2, seed=100 + i), device="cuda") for i in range(2)]
    kw = dict(schedule=m.SCHED_LAYERED, early_exit=m.EXIT_SYNDROME, out_format=m.OUT_U8, msg_max=31)
    want = []
    for y in ys:
        r = code.decode(y, 10, **kw)
        torch.cuda.synchronize()
        want.append((r.D.clone(), r.iters.clone()))
    errs = []

    def worker(i):
        try:
            st = torch.cuda.Stream()
            for _ in range(40):
                r = code.decode(ys[i], 10, stream=st.cuda_stream, **kw)
                st.synchronize()
                if not ((r.D == want[i][0]).all().item() and (r.iters == want[i][1]).all().item()):
                    errs.append(f"thread {i}: result differs")
                    return
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
