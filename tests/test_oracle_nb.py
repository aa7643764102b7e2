"""CPU tests: the non-binary oracle against golden vectors recorded from the reference's own CPU
decoder (tests/golden/gen_nb_golden.py; myNBLDPC/src compiled unchanged under oracle/_ref)."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, NB_DATA


@pytest.fixture(scope="module")
def meta():
    with open(os.path.join(GOLDEN, "nb_ref.json")) as f:
        return json.load(f)


def load(nb_oracle, gf_dir, cfg, exp=None):
    const = os.path.join(NB_DATA, cfg["constellation"].replace("./", ""))
    h = nb_oracle.nb_orc_load(os.path.join(NB_DATA, cfg["matrix"]).encode(),
                              os.path.join(gf_dir, f"Arith.Table.GF.{cfg['q']}.txt").encode(), const.encode(),
                              cfg["n_qam"], cfg["coef_is_exponent"] if exp is None else exp)
    assert h, cfg
    return C.c_void_p(h)


def test_generated_gf_tables_match_reference_files(meta):
    """cuda_ldpc_b200/gf.py regenerates Arith.Table.GF.<q>.txt: same multiply / inverse tables."""
    from cuda_ldpc_b200.gf import gf_tables
    for q, want in meta["gf_sha256"].items():
        mul, add, inv = gf_tables(int(q))
        assert hashlib.sha256(mul.astype(np.int32).tobytes()).hexdigest() == want["mul"], q
        assert hashlib.sha256(inv.astype(np.int32).tobytes()).hexdigest() == want["inv"], q
        a = np.arange(1, int(q))
        assert (mul[a, inv[a]] == 1).all() and (add == (np.arange(int(q))[:, None] ^ np.arange(int(q))[None, :])).all()
        x, order = 2, 1
        while x != 1:
            x, order = int(mul[x, 2]), order + 1
        assert order == int(q) - 1  # alpha = 2 is primitive


def test_codeword_test_vector_is_a_codeword(nb_oracle, gf_dir, meta):
    """CodeWord_sym_test (myNBLDPC/include/codeword_test.h): the only KAT in the reference tree."""
    h = load(nb_oracle, gf_dir, meta["configs"]["BDS"])
    sym = np.array(meta["CodeWord_sym_test"], np.int32)
    assert nb_oracle.nb_orc_syndrome_ok(h, sym.ctypes.data) == 1
    sym[7] ^= 5
    assert nb_oracle.nb_orc_syndrome_ok(h, sym.ctypes.data) == 0
    # it is NOT a codeword of the other GF(64) matrix
    h2 = load(nb_oracle, gf_dir, meta["configs"]["C4"], exp=1)
    assert nb_oracle.nb_orc_syndrome_ok(h2, np.array(meta["CodeWord_sym_test"], np.int32).ctypes.data) == 0


@pytest.mark.parametrize("name", ["BDS", "C4", "C5"])
def test_channel_and_decoders_match_reference(nb_oracle, gf_dir, meta, name):
    """Modulate + AWGNChannel_CPU + Demodulate bit-exact, and Decoding_EMS / Decoding_TMM /
    Decoding_layered_TMM frame-exact (return value, iter_number, every decided symbol)."""
    cfg = meta["configs"][name]
    g = np.load(os.path.join(GOLDEN, f"nb_{name}.npz"))
    h = load(nb_oracle, gf_dir, cfg)
    info = np.zeros(7, np.int32)
    nb_oracle.nb_orc_info(h, info.ctypes.data)
    N, M, q, p = map(int, info[:4])
    sigma = float(g["sigma"])
    assert nb_oracle.nb_orc_sigma(h, 0, cfg["snr_db"]) == pytest.approx(sigma, abs=0)
    sym = np.ascontiguousarray(g["sym"].astype(np.int32))
    L = N * p if cfg["n_qam"] == 2 else N
    tx = np.zeros(2 * L, np.float32)
    assert nb_oracle.nb_orc_modulate(h, sym.ctypes.data, tx.ctypes.data) == L
    seed = np.array([173, 173, 173], np.int32)
    for f in range(cfg["frames"]):
        rx = np.zeros(2 * L, np.float32)
        nb_oracle.nb_orc_awgn(seed.ctypes.data, sigma, tx.ctypes.data, rx.ctypes.data, L)
        assert (rx.view(np.uint32) == g["rx"][f].view(np.uint32)).all(), f"frame {f}: channel samples differ"
        lch = np.zeros(N * (q - 1), np.float32)
        nb_oracle.nb_orc_demodulate(h, sigma, rx.ctypes.data, lch.ctypes.data)
        if f < 2:
            assert (lch.view(np.uint32) == g["L_ch_first2"][f].view(np.uint32)).all()
        for a in cfg["algos"]:
            out = np.zeros(N, np.int32)
            it = C.c_int(0)
            r = nb_oracle.nb_orc_decode(h, a, 0, lch.ctypes.data, cfg["maxit"], cfg["ems_nm"], cfg["ems_nc"],
                                        out.ctypes.data, C.byref(it))
            assert r == g[f"ret_{a}"][f] and it.value == g[f"iter_{a}"][f], (name, a, f)
            assert (out == g[f"out_{a}"][f]).all(), (name, a, f)
    assert seed.tolist() == g["seed_after"].tolist()


def test_fresh_sum_agrees_with_literal_on_converged_frames(nb_oracle, gf_dir, meta):
    """the `fresh` EMS summation (the CUDA kernel's specification) gives the same
    iteration count and decisions as the literal running sum whenever the frame converges."""
    cfg = meta["configs"]["BDS"]
    g = np.load(os.path.join(GOLDEN, "nb_BDS.npz"))
    h = load(nb_oracle, gf_dir, cfg)
    N, q, p = 96, 64, 6
    agree = 0
    rx_all = np.ascontiguousarray(g["rx"])  # keep the buffer alive: a temporary's .ctypes.data dangles once it is freed
    for f in range(cfg["frames"]):
        lch = np.zeros(N * (q - 1), np.float32)
        nb_oracle.nb_orc_demodulate(h, float(g["sigma"]), rx_all[f].ctypes.data, lch.ctypes.data)
        out = np.zeros(N, np.int32)
        it = C.c_int(0)
        r = nb_oracle.nb_orc_decode(h, 0, 1, lch.ctypes.data, 20, 2, 2, out.ctypes.data, C.byref(it))
        same = r == g["ret_0"][f] and it.value == g["iter_0"][f] and (out == g["out_0"][f]).all()
        if g["ret_0"][f] == 1:
            assert same, f
        agree += same
    assert agree >= cfg["frames"] - 4  # non-converged frames amplify rounding noise
