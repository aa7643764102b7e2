"""GPU parity tests of the fp16 message mode (schedule LAYERED, msg_dtype FP16; bldpc_layered_f16.cu) against the
oracle's fp16 rule (oracle/bldpc_oracle.c "fp16 layered rules"): hard bits, flags, iteration counts, every final APP
value and every final check-to-variable message, compared as binary16 bit patterns."""
import numpy as np
import pytest

import os
import re

from conftest import OracleCode, ip, fp
from test_binary_gpu import ALL_FILES, BL, load, noisy

import cuda_ldpc_b200 as m

pytestmark = pytest.mark.gpu


def orc_f16(oracle, oc, y, maxit, mode, scale=8.0, amax=127, bnum=0, bshift=0):
    N, F = y.shape
    dcm = int(oc.Wc[oc.J])
    D = np.zeros((N + 1) * F, np.int32)
    it = np.zeros(F, np.int32)
    app = np.zeros(N * F, np.uint16)
    msg = np.zeros(oc.M * dcm * F, np.uint16)
    rc = oracle.orc_layered_f16(oc.J, oc.L, oc.Z, ip(oc.H), fp(np.ascontiguousarray(y, np.float32)), F, maxit, scale,
                                amax, bnum, bshift, mode, ip(D), ip(it), app.ctypes.data, msg.ctypes.data)
    assert rc == 0
    return D.reshape(N + 1, F), it, app.reshape(N, F), msg


def dec(code, y, it, mode, debug=True, **kw):
    return code.decode(y, it, schedule=m.SCHED_LAYERED, msg_dtype=m.DTYPE_FP16, early_exit=mode, debug=debug, **kw)


@pytest.mark.parametrize("key,F,snr,it", [("C1", 64, 2.5, 10), ("C1", 37, 3.0, 7), ("C3", 16, 3.2, 8), ("C2", 6, 0.5, 5),
                                          ("J32", 24, 1.0, 6), ("J10", 11, 4.0, 10), ("J6", 20, 2.0, 12)])
@pytest.mark.parametrize("mode", [m.EXIT_NONE, m.EXIT_SYNDROME])
def test_layered_f16_bit_exact_vs_oracle(oracle, key, F, snr, it, mode):
    code, oc = load(oracle, key)
    y = noisy(oracle, code.N, F, snr)
    r = dec(code, y, it, mode, msg_max=31, beta_num=1, beta_shift=3)
    D, its, app, msg = orc_f16(oracle, oc, y, it, mode, amax=31, bnum=1, bshift=3)
    assert (r.D == D).all()
    assert (r.iters == its).all()
    assert (r.ok == D[code.N]).all()
    assert (r.app.view(np.uint16) == app).all()
    assert (r.msgs.view(np.uint16) == msg).all()
    assert r.launches == 1


@pytest.mark.parametrize("bnum,bshift,amax,scale", [(0, 0, 127, 8.0), (1, 2, 20, 4.0), (3, 4, 127, 16.0), (1, 3, 1, 8.0),
                                                    (5, 7, 64, 3.3)])
def test_layered_f16_scaling_and_clipping_variants(oracle, bnum, bshift, amax, scale):
    code, oc = load(oracle, "C1")
    y = noisy(oracle, code.N, 24, 2.0)
    r = dec(code, y, 9, m.EXIT_SYNDROME, msg_max=amax, beta_num=bnum, beta_shift=bshift, llr_scale=scale)
    D, its, app, msg = orc_f16(oracle, oc, y, 9, m.EXIT_SYNDROME, scale=scale, amax=amax, bnum=bnum, bshift=bshift)
    assert (r.D == D).all() and (r.iters == its).all()
    assert (r.app.view(np.uint16) == app).all()
    assert (r.msgs.view(np.uint16) == msg).all()


def test_layered_f16_io_variants_and_device_path(oracle):
    """fp16 / int8 inputs, [F][N] layout, the three output formats and the device path give the same decisions."""
    import torch
    code, oc = load(oracle, "C3")
    F = 19
    y = noisy(oracle, code.N, F, 3.0)
    y16 = y.astype(np.float16)
    kw = dict(msg_max=31, beta_num=1, beta_shift=3)
    D, its, app, _ = orc_f16(oracle, oc, y16.astype(np.float32), 8, m.EXIT_SYNDROME, amax=31, bnum=1, bshift=3)
    base = dec(code, y16.astype(np.float32), 8, m.EXIT_SYNDROME, **kw)
    assert (base.D == D).all() and (base.iters == its).all()
    r16 = dec(code, y16, 8, m.EXIT_SYNDROME, **kw)
    assert (r16.D == D).all() and (r16.app.view(np.uint16) == app).all()
    rfn = dec(code, np.ascontiguousarray(y16.T), 8, m.EXIT_SYNDROME, layout=m.LAYOUT_FN, out_format=m.OUT_U8, **kw)
    assert (rfn.D.T == D[: code.N]).all()
    ru8 = dec(code, y16, 8, m.EXIT_SYNDROME, out_format=m.OUT_U8, **kw)
    assert (ru8.D == D[: code.N]).all()
    rbp = dec(code, y16, 8, m.EXIT_SYNDROME, out_format=m.OUT_BITPACK, **kw)
    bits = ((rbp.D[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(F, -1)[:, : code.N]
    assert (bits.T == D[: code.N]).all()
    q = np.clip(np.rint(y * 8), -127, 127).astype(np.int8)
    r8 = dec(code, q, 8, m.EXIT_SYNDROME, llr_scale=1.0, **kw)
    D8, its8, _, _ = orc_f16(oracle, oc, q.astype(np.float32), 8, m.EXIT_SYNDROME, scale=1.0, amax=31, bnum=1, bshift=3)
    assert (r8.D == D8).all() and (r8.iters == its8).all()
    # device path
    yd = torch.from_numpy(y16.astype(np.float32)).cuda()
    rd = dec(code, yd, 8, m.EXIT_SYNDROME, **kw)
    torch.cuda.synchronize()
    assert (rd.D.cpu().numpy() == D).all()
    assert (rd.app.cpu().numpy().view(np.uint16) == app).all()


def test_layered_f16_extreme_inputs(oracle):
    """+-inf, NaN, +-0, values below the binary16 subnormal range and beyond the clamp follow the oracle's rule."""
    code, oc = load(oracle, "C1")
    F = 8
    y = noisy(oracle, code.N, F, 3.0)
    rng = np.random.default_rng(5)
    special = np.array([np.inf, -np.inf, np.nan, 0.0, -0.0, 1e-9, -1e-9, 1e30, -1e30, 15.875, -15.9, 3e-8, -6e-8], np.float32)
    idx = rng.integers(0, y.size, 400)
    y.reshape(-1)[idx] = special[rng.integers(0, special.size, 400)]
    r = dec(code, y, 6, m.EXIT_NONE, msg_max=31, beta_num=1, beta_shift=3)
    D, its, app, msg = orc_f16(oracle, oc, y, 6, m.EXIT_NONE, amax=31, bnum=1, bshift=3)
    assert (r.D == D).all()
    assert (r.app.view(np.uint16) == app).all()
    assert (r.msgs.view(np.uint16) == msg).all()


def test_layered_f16_fused_channel_and_large_batch(oracle):
    """Fused Philox channel = channel kernel + decode; a batch larger than one wave of CTAs, odd F; repeated runs
    are identical (dynamic group scheduling must not change results)."""
    import torch
    code, oc = load(oracle, "C1")
    F = 2 * 148 * 10 + 5
    sigma = m.sigma_from_snr(1, 2.5, code.rate)
    y = torch.empty(code.N * F, dtype=torch.float32, device="cuda")
    assert m.lib.ldpc_awgn_bpsk(code.handle, y.data_ptr(), F, m.LAYOUT_NF, sigma, 99, 0, None,
                                torch.cuda.current_stream().cuda_stream) >= 0
    y = y.view(code.N, F)
    kw = dict(msg_max=31, beta_num=1, beta_shift=3)
    a = dec(code, y, 10, m.EXIT_SYNDROME, debug=False, out_format=m.OUT_BITPACK, **kw)
    b = code.decode_channel(F, 10, sigma, seed=99, msg_dtype=m.DTYPE_FP16, early_exit=m.EXIT_SYNDROME, **kw)
    torch.cuda.synchronize()
    assert torch.equal(a.D, b.D) and torch.equal(a.iters, b.iters) and torch.equal(a.ok, b.ok)
    for _ in range(3):
        c = dec(code, y, 10, m.EXIT_SYNDROME, debug=False, out_format=m.OUT_BITPACK, **kw)
        torch.cuda.synchronize()
        assert torch.equal(a.D, c.D) and torch.equal(a.iters, c.iters)
    # a sample of frames against the oracle
    sel = [0, 1, 2, 3, 777, F - 2, F - 1]
    ys = y[:, sel].cpu().numpy()
    D, its, _, _ = orc_f16(oracle, oc, ys, 10, m.EXIT_SYNDROME, amax=31, bnum=1, bshift=3)
    W = a.D.cpu().numpy().view(np.uint32)[sel]
    bits = ((W[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(len(sel), -1)[:, : code.N]
    assert (bits.T == D[: code.N]).all()
    assert (a.iters.cpu().numpy()[sel] == its).all()
    # decoding works: most frames converge at this SNR, and no worse than the int8 mode by more than a few frames
    i8 = code.decode(y, 10, schedule=m.SCHED_LAYERED, early_exit=m.EXIT_SYNDROME, out_format=m.OUT_BITPACK, **kw)
    torch.cuda.synchronize()
    assert int(a.ok.sum()) >= int(i8.ok.sum()) - F // 50


@pytest.mark.parametrize("fname", ALL_FILES)
def test_layered_f16_every_shipped_code(oracle, fname):
    """All 18 H files of the reference (dc 4..24, dv 2..15, Z 64..1280): every degree bucket, the exact-degree
    instances and the out-of-line generic path of the fp16 kernel, bit-exact against the oracle (F odd: the last
    group holds one codeword)."""
    geo = (12, 69, 256) if fname == "PON_LDPC.txt" else tuple(int(x) for x in re.match(r"J(\d+)_L(\d+)_Z(\d+)", fname).groups())
    path = os.path.join(BL, fname)
    code, oc = m.LdpcCode(path, *geo), OracleCode(oracle, path, *geo)
    F = 7 if code.N > 20000 else 11
    snr = 10 * np.log10(1.0 / (2 * code.rate)) + 3.0 + (1.5 if code.rate > 0.7 else 0.0)
    y = noisy(oracle, code.N, F, snr)
    r = dec(code, y, 6, m.EXIT_SYNDROME, msg_max=31, beta_num=1, beta_shift=3)
    D, its, app, msg = orc_f16(oracle, oc, y, 6, m.EXIT_SYNDROME, amax=31, bnum=1, bshift=3)
    assert (r.D == D).all() and (r.iters == its).all(), fname
    assert (r.app.view(np.uint16) == app).all() and (r.msgs.view(np.uint16) == msg).all(), fname


@pytest.mark.parametrize("it", [1, 2])
@pytest.mark.parametrize("mode", [m.EXIT_NONE, m.EXIT_SYNDROME])
def test_layered_f16_one_and_two_iterations(oracle, it, mode):
    """The first sweep has its own code path (no records to read) and the last fixed sweep skips the record stores."""
    code, oc = load(oracle, "J10")
    y = noisy(oracle, code.N, 9, 6.0)
    for debug in (True, False):
        r = dec(code, y, it, mode, debug=debug, msg_max=31, beta_num=1, beta_shift=3)
        D, its, app, msg = orc_f16(oracle, oc, y, it, mode, amax=31, bnum=1, bshift=3)
        assert (r.D == D).all() and (r.iters == its).all()
        if debug:
            assert (r.app.view(np.uint16) == app).all() and (r.msgs.view(np.uint16) == msg).all()


@pytest.mark.parametrize("layout,fmt", [(m.LAYOUT_NF, m.OUT_BITPACK), (m.LAYOUT_NF, m.OUT_INT32_REF), (m.LAYOUT_FN, m.OUT_U8)])
def test_layered_f16_chunked_host_path_equals_device_path(oracle, layout, fmt):
    """Large host batches are decoded in chunks on two internal streams (copy / compute overlap, fp32 chunks only —
    the host quantiser is an int8 feature): same bits, iteration counts and flags as one device-resident call."""
    import torch
    code, _ = load(oracle, "C1")
    F = 3001  # not a multiple of the chunk (4 * SMs * 2) nor of 2
    y = noisy(oracle, code.N, 64, 2.8)
    y = np.tile(y, (1, 47))[:, :F] * (1.0 + 0.01 * np.arange(F, dtype=np.float32)[None, :] / F)
    yy = np.ascontiguousarray(y if layout == m.LAYOUT_NF else y.T)
    kw = dict(out_format=fmt, layout=layout, msg_max=31, beta_num=1, beta_shift=3, debug=False)
    rh = dec(code, yy, 10, m.EXIT_SYNDROME, **kw)
    rd = dec(code, torch.as_tensor(yy, device="cuda"), 10, m.EXIT_SYNDROME, **kw)
    torch.cuda.synchronize()
    got = rd.D.cpu().numpy()
    assert (rh.D == (got.view(np.uint32) if fmt == m.OUT_BITPACK else got)).all()
    assert (rh.iters == rd.iters.cpu().numpy()).all() and (rh.ok == rd.ok.cpu().numpy()).all()
    assert rh.launches >= 3 and rd.launches == 1
    rp = dec(code, yy, 10, m.EXIT_SYNDROME, host_pack_threads=3, **kw)  # ignored in this mode, not an error
    assert (rp.D == rh.D).all() and (rp.iters == rh.iters).all()
