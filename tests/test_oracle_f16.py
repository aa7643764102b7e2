"""CPU checks of the fp16 oracle (oracle/bldpc_oracle.c "fp16 layered rules"): its binary16 conversions against numpy
for every pattern, and the whole rule against an independent numpy.float16 restatement (numpy's half arithmetic is
correctly rounded), bit for bit."""
import os

import numpy as np

from conftest import DATA, OracleCode, ip, fp

BL = os.path.join(DATA, "bldpc")


def test_binary16_conversions_match_numpy(oracle):
    for h in range(0, 65536):
        want = np.array([h], np.uint16).view(np.float16)[0]
        got = oracle.orc_f16_to_f32(h)
        if np.isnan(want):
            assert np.isnan(got)
        else:
            assert np.float32(want) == got and np.signbit(want) == np.signbit(np.float32(got))
            assert oracle.orc_f16_from_f32(float(np.float32(want))) == h
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.standard_normal(3000).astype(np.float32) * s for s in (1e-8, 1e-7, 1e-5, 1e-3, 1, 100, 30000)])
    x = np.concatenate([x, np.array([65519.9, 65520, 65504, -0.0, 0.0, 6.1e-5, 6.103515625e-05, 5.96e-8, 2.98e-8,
                                     2.9802322e-8, 8.9e-8, np.inf, -np.inf], np.float32)])
    with np.errstate(over="ignore"):
        ref = x.astype(np.float16).view(np.uint16)
    got = np.array([oracle.orc_f16_from_f32(float(v)) for v in x], np.uint16)
    assert (got == ref).all()


def np_layered_f16(oc, y, maxit, scale, amax, bnum, bshift):
    """numpy.float16 restatement of the rule for one frame; returns (APP patterns, iterations, ok)."""
    H = oc.H.reshape(oc.J, oc.L)
    Z = oc.Z
    with np.errstate(invalid="ignore"):
        a = np.clip(y.astype(np.float32) * np.float32(scale), -127, 127)
    a = np.where(np.isnan(a), np.float32(-127), a).astype(np.float16)
    dcm = int((H >= 0).sum(1).max())
    msg = np.zeros((oc.M, dcm), np.float16)
    c = np.float16(1.0 - bnum / (1 << bshift))
    i = np.arange(Z)
    for it in range(1, maxit + 1):
        for r in range(oc.J):
            cols = [cc for cc in range(oc.L) if H[r, cc] >= 0]
            v = np.stack([cc * Z + (i + H[r, cc]) % Z for cc in cols], 1)          # [Z][dc]
            rows = r * Z + i
            t = a[v] - msg[rows][:, : len(cols)]                                    # float16 - float16 -> float16
            neg = np.signbit(t)
            ak = np.abs(t)
            srt = np.sort(ak, 1)
            min1 = srt[:, 0]
            min2 = srt[:, 1] if len(cols) > 1 else np.full(Z, np.inf, np.float16)
            P = np.logical_xor.reduce(neg, 1)
            m1 = np.minimum(min1, np.float16(amax))
            m2 = np.minimum(min2, np.float16(amax))
            if bnum:
                m1, m2 = m1 * c, m2 * c
            mag = np.where(ak == min1[:, None], m2[:, None], m1[:, None])
            s = np.logical_xor(P[:, None], neg)
            nw = np.where(s, -mag, mag).astype(np.float16)
            a[v] = (t + nw).astype(np.float16)
            msg[rows[:, None], np.arange(len(cols))[None, :]] = nw
        ok = True
        for r in range(oc.J):
            cols = [cc for cc in range(oc.L) if H[r, cc] >= 0]
            v = np.stack([cc * Z + (i + H[r, cc]) % Z for cc in cols], 1)
            if np.logical_xor.reduce(np.signbit(a[v]), 1).any():
                ok = False
                break
        if ok:
            return a.view(np.uint16), it, 1
    return a.view(np.uint16), maxit, 0


def test_fp16_rule_against_numpy_restatement(oracle):
    path = os.path.join(BL, "J4_L24_Z96_BlockH.txt")
    oc = OracleCode(oracle, path, 4, 24, 96, literal=0)
    N, F = oc.N, 6
    s = np.array([173, 173, 173], np.int32)
    y = np.zeros(N * F, np.float32)
    oracle.orc_awgn(ip(s), oracle.orc_sigma(1, 3.5, 1.0), None, fp(y), N, F)
    y = y.reshape(N, F)
    y[5, 0], y[6, 0], y[7, 1], y[8, 1], y[9, 2] = np.inf, -1e-9, np.nan, -0.0, -1e30
    fastest = 8
    for (bnum, bshift, amax, scale) in [(1, 3, 31, 8.0), (0, 0, 127, 8.0), (3, 4, 20, 3.3)]:
        D = np.zeros((N + 1) * F, np.int32)
        it = np.zeros(F, np.int32)
        app = np.zeros(N * F, np.uint16)
        assert oracle.orc_layered_f16(oc.J, oc.L, oc.Z, ip(oc.H), fp(np.ascontiguousarray(y)), F, 8, scale, amax, bnum,
                                      bshift, 2, ip(D), ip(it), app.ctypes.data, None) == 0
        app = app.reshape(N, F)
        D = D.reshape(N + 1, F)
        for f in range(F):
            a, its, ok = np_layered_f16(oc, y[:, f], 8, scale, amax, bnum, bshift)
            assert (a == app[:, f]).all(), (bnum, f)
            assert its == it[f] and ok == D[N, f]
            assert ((a >> 15) == D[:N, f]).all()
        fastest = min(fastest, int(it.min()))
    assert fastest < 8  # some frames converge: the early exit path is covered
