"""Repeated-run determinism of the layered int8 path at batch sizes that fill the machine.

Round 1 shipped two shared-memory races in the syndrome pass (bldpc_layered.cu: one `s_fail` word read and
re-written without a barrier in between; latched frames' outputs read while the next sweep already rewrote the
APP words).  The bit-exact tests of test_binary_gpu.py use F <= 64 frames and could not see them; these run
F >= 4 * SMs * 8 frames (every SM busy, dynamic group scheduling under EXIT_SYNDROME), 20 times, through both the
host (chunked, two internal streams) and the device entry of ldpc_decode_batch, demand that every output byte
equals run 0, and pin a random 64-frame slice of run 0 to the CPU oracle.
Reference: the stop test this pass replaces is B/LDPC_Decoder.cu:134-153.
"""
import numpy as np
import pytest

from conftest import ip, fp  # noqa: F401
from test_binary_gpu import load, noisy, orc_i8

import cuda_ldpc_b200 as m

pytestmark = pytest.mark.gpu

REPEATS = 20


def big_batch(oracle, N, F, snr, base=64):
    """F frames from `base` oracle-RNG frames, each scaled a little differently so no two are equal."""
    y = noisy(oracle, N, base, snr)
    reps = (F + base - 1) // base
    y = np.tile(y, (1, reps))[:, :F] * (1.0 + 0.02 * np.arange(F, dtype=np.float32)[None, :] / F)
    return np.ascontiguousarray(y.astype(np.float32))


@pytest.mark.parametrize("key,snr,fmt", [("C1", 2.8, m.OUT_INT32_REF), ("C1", 2.8, m.OUT_BITPACK),
                                         ("C2", -0.2, m.OUT_U8), ("C3", 3.3, m.OUT_INT32_REF)])
@pytest.mark.parametrize("mode", [m.EXIT_SYNDROME, m.EXIT_NONE])
def test_repeated_runs_identical_and_equal_oracle(oracle, key, snr, fmt, mode):
    import torch
    code, oc = load(oracle, key)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    F = 4 * sms * 8 + (3 if key == "C1" else 0)     # C1: ragged tail (last group partly valid, slow load path)
    y = big_batch(oracle, code.N, F, snr)
    kw = dict(schedule=m.SCHED_LAYERED, early_exit=mode, out_format=fmt, msg_max=31, beta_num=1, beta_shift=3)
    yd = torch.as_tensor(y, device="cuda")
    ref = None
    for rep in range(REPEATS):
        rd = code.decode(yd, 10, **kw)
        torch.cuda.synchronize()
        got = (rd.D.cpu().numpy().copy(), rd.iters.cpu().numpy().copy(), rd.ok.cpu().numpy().copy())
        if fmt == m.OUT_BITPACK:
            got = (got[0].view(np.uint32),) + got[1:]
        if ref is None:
            ref = got
            if mode == m.EXIT_SYNDROME:
                its = ref[1]
                assert its.min() < its.max(), "operating point must mix early and late frames inside the groups"
        for a, b, what in zip(got, ref, ("D", "iters", "ok")):
            assert (a == b).all(), f"device path, repeat {rep}: {what} differs from run 0"
        if rep % 4 == 0:  # the host path moves 0.7 GB per call for C2: every 4th repeat
            rh = code.decode(y, 10, **kw)
            for a, b, what in zip((rh.D, rh.iters, rh.ok), ref, ("D", "iters", "ok")):
                assert (a == b).all(), f"host path, repeat {rep}: {what} differs from the device result"
    # pin run 0 to the oracle on a random 64-frame slice (whole groups and not: any frames)
    sel = np.sort(np.random.default_rng(7).choice(F, 64, replace=False))
    D, its, app, rec = orc_i8(oracle, oc, np.ascontiguousarray(y[:, sel]), 10, mode, amax=31, bnum=1, bshift=3)
    if fmt == m.OUT_BITPACK:
        bits = ((ref[0][sel][:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(64, -1)[:, : code.N].T
        okrow = ref[2][sel]
    elif fmt == m.OUT_U8:
        bits, okrow = ref[0][:, sel], ref[2][sel]
    else:
        bits, okrow = ref[0][: code.N][:, sel], ref[0][code.N][sel]
    assert (bits == D[: code.N]).all() and (okrow == D[code.N]).all() and (ref[1][sel] == its).all()


def test_fused_channel_repeated_runs_identical(oracle):
    """Same for the fused-channel entry (Philox inside the kernel's load phase)."""
    import torch
    code, _ = load(oracle, "C3")
    F = 4 * torch.cuda.get_device_properties(0).multi_processor_count * 8
    sigma = m.sigma_from_snr(1, 3.3, code.rate)
    ref = None
    for rep in range(REPEATS):
        r = code.decode_channel(F, 12, sigma, seed=9, first_frame=1000, msg_max=31, beta_num=1, beta_shift=3)
        torch.cuda.synchronize()
        got = (r.D.cpu().numpy().copy(), r.iters.cpu().numpy().copy(), r.ok.cpu().numpy().copy())
        if ref is None:
            ref = got
            assert 0 < got[2].mean() < 1 or got[1].min() < got[1].max()
        assert all((a == b).all() for a, b in zip(got, ref)), f"repeat {rep}"
