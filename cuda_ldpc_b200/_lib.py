"""ctypes binding of libldpc_b200.so.  Fails loudly if the library is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LDPC_B200_LIB: file name of an alternative build of the SAME library next to this file (the race-hunting builds
# libldpc_b200_stress*.so of tests/test_stress_gpu.py); never a different implementation
lib_path = os.path.join(_HERE, os.path.basename(os.environ.get("LDPC_B200_LIB", "libldpc_b200.so")))


class LdpcError(RuntimeError):
    def __init__(self, code, where=""):
        self.code = code
        msg = lib.ldpc_strerror(code).decode()
        cuda = lib.ldpc_last_cuda_error().decode()
        super().__init__(f"{where}: {msg} ({code})" + (f" [{cuda}]" if cuda and code == -4 else ""))


class DecodeOpts(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int), ("batch", C.c_int), ("layout", C.c_int), ("llr_dtype", C.c_int),
        ("mem_space", C.c_int), ("schedule", C.c_int), ("msg_dtype", C.c_int), ("early_exit", C.c_int),
        ("out_format", C.c_int), ("alpha", C.c_float), ("llr_scale", C.c_float), ("msg_max", C.c_int),
        ("beta_num", C.c_int), ("beta_shift", C.c_int), ("iters_out", C.c_void_p), ("ok_out", C.c_void_p),
        ("stream", C.c_void_p), ("debug_app", C.c_void_p), ("debug_msgs", C.c_void_p),
        ("channel_sigma", C.c_float), ("channel_seed", C.c_uint64), ("channel_first_frame", C.c_uint64),
        ("channel_codeword", C.c_void_p), ("host_pack_threads", C.c_int),
    ]


class CodeInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("J", "L", "Z", "N", "K", "M", "E", "dc_max", "dv_max", "dc_min", "dv_min")]


class SimCounters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("num_Frames", "num_Error_Frames", "num_Error_Bits", "Total_Iteration",
                                         "num_False_Frames", "num_Alarm_Frames")]


class NbCodeInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("N", "M", "q", "p", "dv_max", "dc_max", "n_const")]


class NbDecodeOpts(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int), ("batch", C.c_int), ("algo", C.c_int), ("in_kind", C.c_int),
        ("mem_space", C.c_int), ("ems_nm", C.c_int), ("ems_nc", C.c_int), ("sigma", C.c_float),
        ("iters_out", C.c_void_p), ("ok_out", C.c_void_p), ("stream", C.c_void_p),
    ]


def _load():
    if not os.path.exists(lib_path):
        raise ImportError(
            f"{lib_path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C cuda_ldpc_b200/csrc).  There is no CPU fallback.")
    L = C.CDLL(lib_path)
    L.ldpc_strerror.restype = C.c_char_p
    L.ldpc_strerror.argtypes = [C.c_int]
    L.ldpc_last_cuda_error.restype = C.c_char_p
    L.ldpc_version.restype = C.c_char_p
    L.ldpc_load_code.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.ldpc_free_code.argtypes = [C.c_void_p]
    L.ldpc_free_code.restype = None
    L.ldpc_code_info.argtypes = [C.c_void_p, C.POINTER(CodeInfo)]
    L.ldpc_code_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    L.ldpc_decode_opts_default.argtypes = [C.POINTER(DecodeOpts)]
    L.ldpc_decode_opts_default.restype = None
    L.ldpc_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(DecodeOpts)]
    L.ldpc_wave_frames.restype = C.c_int
    L.ldpc_wave_frames.argtypes = [C.c_void_p, C.c_int]
    L.ldpc_last_h2d_bytes.restype = C.c_size_t
    L.ldpc_last_h2d_bytes.argtypes = [C.c_void_p]
    L.ldpc_out_bytes.restype = C.c_size_t
    L.ldpc_out_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ldpc_awgn_bpsk.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_uint64, C.c_uint64,
                                 C.c_void_p, C.c_void_p]
    L.ldpc_statistic.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p]
    L.ldpc_philox4x32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ldpc_philox4x32.restype = None
    L.ldpc_sigma.restype = C.c_float
    L.ldpc_sigma.argtypes = [C.c_int, C.c_float, C.c_float]
    L.ldpc_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.nb_ldpc_load_code.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    L.nb_ldpc_free_code.argtypes = [C.c_void_p]
    L.nb_ldpc_free_code.restype = None
    L.nb_ldpc_code_info.argtypes = [C.c_void_p, C.POINTER(NbCodeInfo)]
    L.nb_ldpc_code_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    L.nb_decode_opts_default.argtypes = [C.POINTER(NbDecodeOpts)]
    L.nb_decode_opts_default.restype = None
    L.nb_ldpc_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(NbDecodeOpts)]
    L.nb_ldpc_sigma.restype = C.c_float
    L.nb_ldpc_sigma.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int]
    L.nb_ldpc_modulate_awgn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_uint64,
                                        C.c_void_p, C.c_void_p]
    L.nb_ldpc_statistic.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p]
    L.nb_ldpc_encode_info.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_void_p]
    L.nb_ldpc_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return L


lib = _load()
