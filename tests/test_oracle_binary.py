"""CPU tests: the binary oracle against the golden vectors generated from the reference's own
sources (tests/golden/gen_binary_golden.py, CPU shim build under oracle/_ref)."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, DATA, OracleCode, ip, fp

BL = os.path.join(DATA, "bldpc")


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "binary_ref.json")) as f:
        meta = json.load(f)
    batch = np.load(os.path.join(GOLDEN, "binary_C1_batch16.npz"))
    return meta, batch


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_tables_match_reference(oracle, golden):
    """Get_H / Transform_H (B/Simulation.cu:292-387): H, weights and both address tables."""
    meta, _ = golden
    for hfile, t in meta["tables"].items():
        for literal, key in ((1, "addr_sha256_literal"), (0, "addr_sha256_fixed")):
            code = OracleCode(oracle, os.path.join(BL, hfile), t["J"], t["L"], t["Z"], literal)
            assert code.H.tolist() == t["H"], hfile
            assert code.Wc.tolist() == t["Wc"] and code.Wv.tolist() == t["Wv"], hfile
            assert sha(code.addr) == t[key], (hfile, key)


def test_literal_table_is_miswired_circulant_is_a_permutation(oracle):
    """the literal table has duplicate slots; the circulant table is a bijection
    onto the used slots and agrees with row = (col - s) mod Z."""
    lit = OracleCode(oracle, os.path.join(BL, "J4_L24_Z96_BlockH.txt"), 4, 24, 96, 1)
    cir = OracleCode(oracle, os.path.join(BL, "J4_L24_Z96_BlockH.txt"), 4, 24, 96, 0)
    used_l = lit.addr[lit.addr >= 0]
    used_c = cir.addr[cir.addr >= 0]
    assert len(used_c) == 7680 and len(np.unique(used_c)) == 7680
    assert len(used_l) == 7680 and len(np.unique(used_l)) == 7680 - 1542
    assert int((lit.addr == cir.addr).sum()) - int((cir.addr < 0).sum()) == 4394
    Z, L, Wcm, Wvm = 96, 24, int(cir.Wc[4]), int(cir.Wv[24])
    H = cir.H.reshape(4, 24)
    for c in range(L):
        rows = [r for r in range(4) if H[r, c] != -1]
        for k, r in enumerate(rows):
            pos = int((H[r, :c] != -1).sum())
            i = np.arange(Z)
            want = (r * Z + (i - H[r, c]) % Z) * Wcm + pos
            assert (cir.addr.reshape(-1, Wvm)[c * Z + i, k] == want).all()


def test_rng_and_awgn_bit_exact(oracle, golden):
    """RandomModule + AWGNChannel_CPU (B/LDPC_Encoder.cu:25-56), seeds 173/173/173."""
    _, batch = golden
    y = batch["y"]
    N, F = y.shape
    seed = np.array([173, 173, 173], np.int32)
    out = np.zeros(N * F, np.float32)
    oracle.orc_awgn(ip(seed), float(batch["sigma"]), None, fp(out), N, F)
    assert (out.view(np.uint32) == y.reshape(-1).view(np.uint32)).all()
    assert seed.tolist() == batch["seed_after"].tolist()
    assert abs(float(batch["sigma"]) - oracle.orc_sigma(1, 3.0, 1.0)) < 1e-7


@pytest.mark.parametrize("variant,literal", [("literal", 1), ("fixed", 0)])
def test_flooding_decode_matches_reference_batch(oracle, golden, variant, literal):
    """LDPC_Decoder_GPU (B/LDPC_Decoder.cu:23-398) on one 16-frame batch: every hard bit, the
    per-frame flag row and the batch iteration count."""
    _, batch = golden
    y = np.ascontiguousarray(batch["y"])
    N, F = y.shape
    code = OracleCode(oracle, os.path.join(BL, "J4_L24_Z96_BlockH.txt"), 4, 24, 96, literal)
    D = np.zeros((N + 1) * F, np.int32)
    iters = np.zeros(F, np.int32)
    rc = oracle.orc_flooding_fp32(4, 24, 96, ip(code.H), ip(code.Wc), ip(code.Wv), ip(code.addr), fp(y), F, 10, 1,
                                  code.K, ip(D), ip(iters), None)
    assert rc == 0
    want = np.unpackbits(batch[f"D_{variant}"], axis=0)[:N].astype(np.int32)
    assert (D[: N * F].reshape(N, F) == want).all()
    assert (D[N * F:] == batch[f"flag_{variant}"]).all()
    assert (iters == int(batch[f"iters_{variant}"])).all()


def test_sim_points_match_reference_counters(oracle, golden):
    """Simulation_GPU + Statistic (B/Simulation.cu:12-285) with the reference RNG: frames, error
    frames, error bits, iterations, false and alarm counters are seed-exact."""
    meta, _ = golden
    for p in meta["sim_points"]:
        literal = 1 if p["variant"] == "literal" else 0
        code = OracleCode(oracle, os.path.join(BL, p["code"]), 4, 24, 96, literal)
        c = np.zeros(6, np.int64)
        oracle.orc_sim_point(4, 24, 96, ip(code.H), ip(code.Wc), ip(code.Wv), ip(code.addr), p["snrtype"],
                             p["snr_db"], p["F"], p["maxit"], code.K, 1, 50, p["leastTestFrames"], 0, 173,
                             c.ctypes.data_as(C.POINTER(C.c_long)))
        assert c.tolist() == p["counters"], p


def test_layered_modes_decode_clean_and_noisy(oracle):
    """The oracle-defined modes (layered fp32 / int8): all-zero codeword at high SNR decodes with
    zero errors in one iteration, and at 3 dB Es/N0 layered int8 tracks layered fp32."""
    code = OracleCode(oracle, os.path.join(BL, "J4_L24_Z96_BlockH.txt"), 4, 24, 96, 0)
    N, F = code.N, 64
    seed = np.array([173, 173, 173], np.int32)
    y = np.zeros(N * F, np.float32)
    oracle.orc_awgn(ip(seed), oracle.orc_sigma(1, 3.0, 1.0), None, fp(y), N, F)
    D32 = np.zeros((N + 1) * F, np.int32); it32 = np.zeros(F, np.int32)
    D8 = np.zeros((N + 1) * F, np.int32); it8 = np.zeros(F, np.int32)
    assert oracle.orc_layered_fp32(4, 24, 96, ip(code.H), fp(y), F, 10, 1.0, 2, ip(D32), ip(it32), None) == 0
    assert oracle.orc_layered_i8(4, 24, 96, ip(code.H), fp(y), F, 10, 8.0, 127, 0, 0, 2, ip(D8), ip(it8), None,
                                 None) == 0
    ok32, ok8 = D32[N * F:], D8[N * F:]
    assert ok32.sum() >= 60 and ok8.sum() >= 58
    # converged frames are the all-zero codeword
    assert (D32[: N * F].reshape(N, F)[:, ok32 == 1] == 0).all()
    assert (D8[: N * F].reshape(N, F)[:, ok8 == 1] == 0).all()
    # syndrome flag agrees with the independent syndrome routine
    chk = np.zeros(F, np.int32)
    oracle.orc_syndrome_ok(4, 24, 96, ip(code.H), ip(D8), F, ip(chk))
    assert (chk == ok8).all()
