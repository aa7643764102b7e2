"""CPU tests of the C-ABI library: it loads, exports every symbol include/ldpc_b200.h declares,
its loader agrees with the oracle's Get_H / Transform_H on all shipped H files, and decoding
without a GPU fails loudly instead of falling back."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

from conftest import DATA, GOLDEN, ROOT, OracleCode

import cuda_ldpc_b200 as m

BL = os.path.join(DATA, "bldpc")


def declared_functions():
    src = open(os.path.join(ROOT, "include", "ldpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:nb_)?ldpc_[a-z0-9_]+|nb_decode_opts_default)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 20
    L = C.CDLL(m.lib_path)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ldpc_b200.h but not exported"
    assert b"sm_100a" in m.lib.ldpc_version()


def geometry(path):
    base = os.path.basename(path)
    if base == "PON_LDPC.txt":
        return 12, 69, 256
    return tuple(int(x) for x in re.match(r"J(\d+)_L(\d+)_Z(\d+)", base).groups())


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(BL, "*.txt"))), ids=os.path.basename)
def test_loader_matches_oracle_tables(oracle, path):
    J, L, Z = geometry(path)
    code = m.LdpcCode(path, *(geometry(path) if "PON" in path else (0, 0, 0)))
    ref = OracleCode(oracle, path, J, L, Z, literal=0)
    assert (code.J, code.L, code.Z, code.N, code.K, code.M) == (J, L, Z, L * Z, (L - J) * Z, J * Z)
    H, Wc, Wv, addr = code.tables()
    assert (H == ref.H).all() and (Wc == ref.Wc).all() and (Wv == ref.Wv).all()
    assert (addr == ref.addr).all()
    assert code.E == int((ref.H >= 0).sum()) * Z
    assert code.dc_max == ref.Wc[J] and code.dv_max == ref.Wv[L]


def test_loader_errors(tmp_path):
    h = C.c_void_p()
    assert m.lib.ldpc_load_code(b"/nonexistent/J4_L24_Z96_BlockH.txt", 0, 0, 0, C.byref(h)) == -1
    assert m.lib.ldpc_load_code(os.path.join(BL, "PON_LDPC.txt").encode(), 0, 0, 0, C.byref(h)) == -3
    # wrong geometry: too few / too many integers, shift >= Z
    assert m.lib.ldpc_load_code(os.path.join(BL, "J4_L24_Z96_BlockH.txt").encode(), 5, 24, 96, C.byref(h)) == -2
    assert m.lib.ldpc_load_code(os.path.join(BL, "J4_L24_Z96_BlockH.txt").encode(), 3, 24, 96, C.byref(h)) == -2
    assert m.lib.ldpc_load_code(os.path.join(BL, "J4_L24_Z96_BlockH.txt").encode(), 4, 24, 64, C.byref(h)) == -2
    # CRLF, tabs, leading blanks and no trailing newline parse like the reference's fscanf
    p = tmp_path / "J2_L3_Z5_BlockH.txt"
    p.write_bytes(b" 1\t-1\t 4\r\n\t0 2 -1")
    c = m.LdpcCode(str(p))
    assert c.tables()[0].tolist() == [1, -1, 4, 0, 2, -1] and (c.N, c.K, c.E) == (15, 5, 20)
    assert m.lib.ldpc_strerror(-7).decode().startswith("no CUDA device")


def test_decode_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    c = m.LdpcCode(os.path.join(BL, "J4_L24_Z96_BlockH.txt"))
    with pytest.raises(m.LdpcError) as e:
        c.decode(np.ones((c.N, 4), np.float32), 2)
    assert e.value.code == -7


def test_wave_frames_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    c = m.LdpcCode(os.path.join(BL, "J4_L24_Z96_BlockH.txt"))
    assert m.lib.ldpc_wave_frames(c.handle, m.DTYPE_INT8) == -7
    assert m.lib.ldpc_wave_frames(c.handle, m.DTYPE_FP32) == -6  # only the layered int8 / fp16 kernels have waves
    assert m.lib.ldpc_wave_frames(None, m.DTYPE_INT8) == -3


def test_no_product_code_touches_the_oracle():
    """The product path must never import, link or execute anything under oracle/."""
    bad = []
    for root, _, files in os.walk(os.path.join(ROOT, "cuda_ldpc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", "Makefile")):
                txt = open(os.path.join(root, f), errors="replace").read()
                # comments may cite the oracle as the specification; code may not use it
                if re.search(r"liboracle|#\s*include\s*[<\"][^>\"]*oracle|^\s*(import|from)\s+\S*oracle|dlopen|"
                             r"CDLL\([^)]*oracle", txt, flags=re.M):
                    bad.append(os.path.join(root, f))
    assert not bad, bad


def test_host_quantiser_follows_the_kernel_rule():
    """ldpcb_host_pack_nf (csrc/host_pack.cc, AVX-512 with non-temporal stores or scalar): q = sat127(rint(y * scale)),
    ties to even, NaN -> -127 — the layered kernel's own load rule (bldpc_layered.cu `quant`), for ragged chunk widths,
    unaligned outputs and any thread count."""
    import ctypes as C
    import cuda_ldpc_b200 as m
    L = C.CDLL(m.lib_path)
    L.ldpcb_host_pack_nf.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int]
    L.ldpcb_host_pack_nf.restype = None
    rng = np.random.default_rng(3)
    N, F = 37, 1000
    y = (rng.standard_normal((N, F)) * 6).astype(np.float32)
    special = np.array([0.0625, -0.0625, 0.1875, -0.1875, 0.3125, 15.875, 15.9375, -15.9375, 1e9, -1e9, 1e-30, -1e-30,
                        15.8125, -15.8125, 0.0, -0.0, np.nan, np.inf, -np.inf, 3.4e38, -3.4e38], np.float32)
    y[0, : special.size] = special
    y[5, 100: 100 + special.size] = special
    for scale in (8.0, 1.0, 0.37):
        with np.errstate(invalid="ignore", over="ignore"):
            want = np.clip(np.rint(y * np.float32(scale)), -127, 127)
        want = np.where(np.isnan(want), -127, want).astype(np.int8)
        for f0, fc, threads in ((0, F, 1), (3, 333, 2), (64, 640, 5), (1, 65, 3), (500, 7, 1)):
            buf = np.full(N * fc + 64, 77, np.int8)
            for shift in (0, 1, 13):  # output alignment relative to 64 bytes
                out = buf[shift: shift + N * fc]
                L.ldpcb_host_pack_nf(y.ctypes.data, F, N, f0, fc, scale, out.ctypes.data, threads)
                assert (out.reshape(N, fc) == want[:, f0: f0 + fc]).all(), (scale, f0, fc, threads, shift)


@pytest.mark.parametrize("matrix,q,exp", [("BDS.576.288.GF.64.txt", 64, 0), ("LDPC_N576_K288_GF64_d1_exp.txt", 64, 1),
                                          ("LDPC_N576_K480_GF256_exp.txt", 256, 1), ("LDPC_N96_K48_GF256_d1_exp.txt", 256, 1),
                                          ("Tanner_74_9_Z128_GF16.txt", 16, 0)])
def test_nb_encoder_produces_codewords(nb_oracle, gf_dir, matrix, q, exp):
    """nb_ldpc_encode (GF(q) Gauss-Jordan, csrc/nb_encoder.cpp): random information symbols -> H c = 0 by the ORACLE's
    syndrome routine; the information symbols sit where nb_ldpc_encode_info says; the map is GF-linear."""
    import ctypes as C
    import cuda_ldpc_b200 as m
    mat = os.path.join(DATA, "nbldpc", matrix)
    gf = os.path.join(gf_dir, f"Arith.Table.GF.{q}.txt")
    const = os.path.join(DATA, "nbldpc", "Constellation", "BPSK.txt")
    code = m.NbLdpcCode(mat, gf, const, bool(exp))
    h = C.c_void_p(nb_oracle.nb_orc_load(mat.encode(), gf.encode(), const.encode(), 2, exp))
    pos = code.info_positions()
    assert len(pos) >= code.N - code.M and (np.diff(pos) > 0).all()
    rng = np.random.default_rng(q)
    words = []
    for _ in range(4):
        u = rng.integers(0, q, len(pos)).astype(np.uint16)
        cw = code.encode(u)
        assert (cw[pos] == u).all()
        assert nb_oracle.nb_orc_syndrome_ok(h, cw.astype(np.int32).ctypes.data) == 1
        words.append((u, cw))
    # linearity over GF(2^p): addition is XOR
    (u0, c0), (u1, c1) = words[0], words[1]
    assert (code.encode(u0 ^ u1) == (c0 ^ c1)).all()
    assert (code.encode(np.zeros(len(pos), np.uint16)) == 0).all()


def test_nb_encoder_rederives_the_reference_codeword(nb_oracle, gf_dir):
    """CodeWord_sym_test (myNBLDPC/include/codeword_test.h:1), the only codeword the reference ships: encoding its own
    information symbols gives it back (H has full rank, so the codeword is determined by them)."""
    import json
    import cuda_ldpc_b200 as m
    with open(os.path.join(GOLDEN, "nb_ref.json")) as f:
        cwt = np.array(json.load(f)["CodeWord_sym_test"], np.uint16)
    code = m.NbLdpcCode(os.path.join(DATA, "nbldpc", "BDS.576.288.GF.64.txt"),
                        os.path.join(gf_dir, "Arith.Table.GF.64.txt"),
                        os.path.join(DATA, "nbldpc", "Constellation", "BPSK.txt"), False)
    pos = code.info_positions()
    assert len(pos) == code.N - code.M == 48
    assert (code.encode(cwt[pos]) == cwt).all()
