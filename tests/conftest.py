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
