"""Edge cases of ldpc_decode_batch through the C-ABI: tiny and ragged batches, degenerate inputs (all zero: every
comparison a tie; all saturated; mixed infinities), and the error contract (negative codes, nothing printed, nothing
decoded).  The reference has none of this: it exits on every error (B/LDPC_Decoder.cu:39-44) and fixes F at compile
time (B/define.cuh:60)."""
import ctypes as C

import numpy as np
import pytest

from test_binary_gpu import load, noisy, orc_flood, orc_i8

import cuda_ldpc_b200 as m

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("F", [1, 2, 3, 5])
def test_tiny_and_ragged_batches(oracle, F):
    """F below one group of 4, host and device entry, both schedules"""
    import torch
    code, oc = load(oracle, "C1")
    y = noisy(oracle, code.N, F, 2.8)
    D, its, app, msgs = orc_i8(oracle, oc, y, 8, m.EXIT_SYNDROME, amax=31, bnum=1, bshift=3)
    kw = dict(schedule=m.SCHED_LAYERED, early_exit=m.EXIT_SYNDROME, msg_max=31, beta_num=1, beta_shift=3, debug=True)
    rh = code.decode(y, 8, **kw)
    rd = code.decode(torch.as_tensor(y, device="cuda"), 8, **kw)
    torch.cuda.synchronize()
    for r, D_, it_, app_, msg_ in ((rh, rh.D, rh.iters, rh.app, rh.msgs),
                                   (rd, rd.D.cpu().numpy(), rd.iters.cpu().numpy(), rd.app.cpu().numpy(), rd.msgs.cpu().numpy())):
        assert (D_ == D).all() and (it_ == its).all() and (app_ == app).all() and (msg_ == msgs).all()
    Df, itf, rq = orc_flood(oracle, oc, y, 6, m.EXIT_NONE)
    rf = code.decode(y, 6, debug=True)
    assert (rf.D == Df).all() and (rf.msgs.view(np.uint32) == rq.view(np.uint32)).all()


@pytest.mark.parametrize("kind", ["zeros", "saturated", "infinite", "alternating"])
def test_degenerate_inputs_bit_exact(oracle, kind):
    code, oc = load(oracle, "J6")
    F = 9
    rng = np.random.default_rng(3)
    if kind == "zeros":            # every |t| equal: first-index ties in every check, sign(0) = +
        y = np.zeros((code.N, F), np.float32)
    elif kind == "saturated":      # every LLR clips at +-127
        y = np.where(rng.random((code.N, F)) < 0.5, -1e3, 1e3).astype(np.float32)
    elif kind == "infinite":
        y = noisy(oracle, code.N, F, 3.0)
        y[rng.random((code.N, F)) < 0.01] = np.inf
        y[rng.random((code.N, F)) < 0.01] = -np.inf
    else:                          # +-0.0625 = exactly half a quantiser step at scale 8: ties-to-even -> 0
        y = (0.0625 * (1 - 2 * (np.add.outer(np.arange(code.N), np.arange(F)) & 1))).astype(np.float32)
    for mode in (m.EXIT_NONE, m.EXIT_SYNDROME):
        r = code.decode(y, 5, schedule=m.SCHED_LAYERED, early_exit=mode, debug=True)
        D, its, app, msgs = orc_i8(oracle, oc, y, 5, mode)
        assert (r.D == D).all() and (r.iters == its).all() and (r.app == app).all() and (r.msgs == msgs).all(), (kind, mode)


def test_error_contract_returns_codes_and_decodes_nothing(oracle):
    code, _ = load(oracle, "C1")
    y = noisy(oracle, code.N, 8, 3.0)
    out = np.full(code.out_bytes(8, m.OUT_INT32_REF), 0x5A, np.uint8)

    def call(iters=5, llr=y, **kw):
        o = code.make_opts(8, **dict(dict(mem_space=m.binary.MEM_HOST), **kw))
        return m.lib.ldpc_decode_batch(code.handle, llr.ctypes.data if llr is not None else None, out.ctypes.data,
                                       iters, C.byref(o))
    ERR_ARG, ERR_UNSUP = -3, -6
    assert call(iters=0) == ERR_ARG
    assert call(llr=None) == ERR_ARG
    assert call(batch=0) == ERR_ARG
    assert call(struct_size=12) == ERR_ARG
    assert call(layout=7) == ERR_ARG
    assert call(out_format=9) == ERR_ARG
    assert call(schedule=m.SCHED_LAYERED, msg_dtype=m.DTYPE_INT8, msg_max=0) == ERR_ARG
    assert call(schedule=m.SCHED_LAYERED, msg_dtype=m.DTYPE_INT8, msg_max=31, beta_num=4, beta_shift=2) == ERR_ARG
    assert call(schedule=m.SCHED_LAYERED, msg_dtype=m.DTYPE_INT8, early_exit=m.EXIT_GENIE) == ERR_UNSUP
    assert call(schedule=m.SCHED_FLOODING, msg_dtype=m.DTYPE_INT8) == ERR_UNSUP
    assert call(schedule=m.SCHED_FLOODING, llr_dtype=m.binary.DTYPE_CHANNEL, channel_sigma=0.5) == ERR_UNSUP
    assert (out == 0x5A).all(), "a rejected call must not touch the output buffer"
    assert m.lib.ldpc_strerror(ERR_UNSUP).decode() and m.lib.ldpc_strerror(-12345).decode()
    assert call() >= 1  # and the handle still works afterwards


def test_wave_frames_batches_decode_bit_exact(oracle):
    """ldpc_wave_frames = resident CTAs x codewords per group; batches of one wave and one wave + 1 frame decode
    bit for bit like the oracle (checked on the first and last 8 frames)"""
    import torch
    code, oc = load(oracle, "C1")
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    w8, w16 = code.wave_frames(), code.wave_frames(m.DTYPE_FP16)
    assert w8 > 0 and w8 % (4 * sms) == 0
    assert w16 > 0 and w16 % (2 * sms) == 0
    assert m.lib.ldpc_wave_frames(code.handle, m.DTYPE_FP32) == -6
    for F in (w8, w8 + 1):
        y = noisy(oracle, code.N, F, 2.8)
        r = code.decode(y, 5, schedule=m.SCHED_LAYERED, msg_max=31, beta_num=1, beta_shift=3)
        sl = np.r_[0:8, F - 8:F]
        D, its, _, _ = orc_i8(oracle, oc, np.ascontiguousarray(y[:, sl]), 5, m.EXIT_NONE, amax=31, bnum=1, bshift=3)
        assert (r.D[:, sl] == D).all() and (r.iters[sl] == its).all()
