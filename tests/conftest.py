import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "cuda_ldpc_b200", "data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the product library is built in-tree (sm_100a cross-compiles without a GPU); build it if a fresh
    # checkout has not run __graft_entry__.build() yet
    lib = os.path.join(ROOT, "cuda_ldpc_b200", "libldpc_b200.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "cuda_ldpc_b200", "csrc")])


def ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/liboracle.so) — the checker, never the product path."""
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = C.CDLL(so)
    lib.orc_sim_point.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] * 4 + [
        C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_long, C.c_long, C.c_long, C.c_int,
        C.POINTER(C.c_long)]
    lib.orc_sigma.restype = C.c_float
    lib.orc_sigma.argtypes = [C.c_int, C.c_float, C.c_float]
    lib.orc_random_module.restype = C.c_float
    lib.orc_awgn.argtypes = [C.POINTER(C.c_int), C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int,
                             C.c_int]
    lib.orc_layered_fp32.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int, C.c_int,
                                                      C.c_float, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                                      C.c_void_p]
    lib.orc_layered_i8.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int, C.c_int,
                                                    C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                                    C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
    lib.orc_layered_f16.argtypes = lib.orc_layered_i8.argtypes
    lib.orc_f16_from_f32.restype = C.c_uint16
    lib.orc_f16_from_f32.argtypes = [C.c_float]
    lib.orc_f16_to_f32.restype = C.c_float
    lib.orc_f16_to_f32.argtypes = [C.c_uint16]
    lib.orc_flooding_fp32.argtypes = [C.c_int] * 3 + [C.POINTER(C.c_int)] * 4 + [
        C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
        C.c_void_p]
    return lib


class OracleCode:
    """Get_H + Transform_H through the oracle for one BlockH file."""

    def __init__(self, lib, path, J, L, Z, literal=0):
        self.J, self.L, self.Z = J, L, Z
        self.N, self.M, self.K = L * Z, J * Z, (L - J) * Z
        self.H = np.zeros(J * L, np.int32)
        self.Wc = np.zeros(J + 1, np.int32)
        self.Wv = np.zeros(L + 1, np.int32)
        rc = lib.orc_get_h(path.encode(), J, L, ip(self.H), ip(self.Wc), ip(self.Wv))
        assert rc == 0, f"orc_get_h({path}) -> {rc}"
        self.addr = np.full(self.N * int(self.Wv[L]), -1, np.int32)
        lib.orc_transform_h(ip(self.H), J, L, Z, ip(self.Wc), ip(self.Wv), ip(self.addr), literal)


@pytest.fixture(scope="session")
def data_dir():
    return DATA


NB_DATA = os.path.join(DATA, "nbldpc")


@pytest.fixture(scope="session")
def gf_dir(tmp_path_factory):
    """GF table files regenerated from the primitive polynomials (cuda_ldpc_b200/gf.py)."""
    from cuda_ldpc_b200.gf import write_table_file
    d = tmp_path_factory.mktemp("gf")
    for q in (16, 64, 256):
        write_table_file(q, str(d / f"Arith.Table.GF.{q}.txt"))
    return str(d)


@pytest.fixture(scope="session")
def nb_oracle(oracle):
    lib = oracle
    lib.nb_orc_load.restype = C.c_void_p
    lib.nb_orc_load.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    lib.nb_orc_free.argtypes = [C.c_void_p]
    lib.nb_orc_info.argtypes = [C.c_void_p, C.c_void_p]
    lib.nb_orc_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    lib.nb_orc_constellation.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.nb_orc_syndrome_ok.argtypes = [C.c_void_p, C.c_void_p]
    lib.nb_orc_sigma.restype = C.c_float
    lib.nb_orc_sigma.argtypes = [C.c_void_p, C.c_int, C.c_float]
    lib.nb_orc_awgn.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int]
    lib.nb_orc_modulate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.nb_orc_demodulate.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
    lib.nb_orc_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p]
    lib.nb_orc_decode_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
    return lib
