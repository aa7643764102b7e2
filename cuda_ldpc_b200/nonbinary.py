"""Non-binary GF(q) LDPC decode path: Python mirror of the reference's interface.

Reference call sites this replaces (gsw4869/CUDA_LDPC, myNBLDPC/):
  Get_H + GFInitial + Get_CONSTELLATION + table flattening (src/main.cu:44-188) -> NbLdpcCode(...)
  Demodulate + Decoding_EMS / Decoding_TMM / Decoding_layered_TMM (src/Simulation.cpp:50-70,
  128-144)                                                         -> NbLdpcCode.decode(...)
decode() returns (DecodeOutput[F][N], iter_number[F], ok[F]) with the reference's meaning: ok = the
decoder's return value (1 = syndrome satisfied), iter_number = iterations-1 on success.
"""
import ctypes as C

import numpy as np

from ._lib import LdpcError, NbCodeInfo, NbDecodeOpts, lib

ALGO_EMS, ALGO_TMM, ALGO_LAYERED_TMM, ALGO_FFT_BP = 0, 1, 3, 4
IN_SYMBOL_LLR, IN_BPSK, IN_QAM = 0, 1, 2


class NbLdpcCode:
    def __init__(self, matrix, gf_table=None, constellation=None, coef_is_exponent=False):
        h = C.c_void_p()
        rc = lib.nb_ldpc_load_code(str(matrix).encode(), str(gf_table).encode() if gf_table else None,
                                   str(constellation).encode() if constellation else None,
                                   1 if coef_is_exponent else 0, C.byref(h))
        if rc != 0:
            raise LdpcError(rc, f"nb_ldpc_load_code({matrix})")
        self._h = h
        info = NbCodeInfo()
        lib.nb_ldpc_code_info(h, C.byref(info))
        for n, _ in NbCodeInfo._fields_:
            setattr(self, n, getattr(info, n))
        self.K_bits = (self.N - self.M) * self.p

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.nb_ldpc_free_code(h)

    def tables(self):
        mul = np.zeros(self.q * self.q, np.uint16)
        inv = np.zeros(self.q, np.uint16)
        vn = np.zeros(self.M * self.dc_max, np.int32)
        coef = np.zeros(self.M * self.dc_max, np.int32)
        w = np.zeros(self.M, np.int32)
        rc = lib.nb_ldpc_code_tables(self._h, mul.ctypes.data, inv.ctypes.data, vn.ctypes.data, coef.ctypes.data,
                                     w.ctypes.data)
        if rc != 0:
            raise LdpcError(rc, "nb_ldpc_code_tables")
        return mul.reshape(self.q, self.q), inv, vn.reshape(self.M, self.dc_max), coef.reshape(self.M, self.dc_max), w

    def in_elems(self, in_kind):
        return {IN_SYMBOL_LLR: self.N * (self.q - 1), IN_BPSK: self.N * self.p, IN_QAM: self.N * 2}[in_kind]

    def info_positions(self):
        """positions of the K = N - rank(H) information symbols inside a codeword (nb_ldpc_encode_info)"""
        k = C.c_int(0)
        rc = lib.nb_ldpc_encode_info(self._h, C.byref(k), None)
        if rc != 0:
            raise LdpcError(rc, "nb_ldpc_encode_info")
        pos = np.zeros(k.value, np.int32)
        lib.nb_ldpc_encode_info(self._h, C.byref(k), pos.ctypes.data)
        return pos

    def encode(self, info_syms):
        """info symbols [K] -> codeword [N] (uint16) with H c = 0; the reference has no encoder (codeword_test.h:1)."""
        u = np.ascontiguousarray(info_syms, dtype=np.uint16)
        cw = np.zeros(self.N, np.uint16)
        rc = lib.nb_ldpc_encode(self._h, u.ctypes.data, cw.ctypes.data)
        if rc != 0:
            raise LdpcError(rc, "nb_ldpc_encode")
        return cw

    def sigma(self, snrtype, snr_db, n_qam=0):
        """main.cu:221-228 (n_qam <= 0: the loaded constellation's size)."""
        return float(lib.nb_ldpc_sigma(self._h, int(snrtype), float(snr_db), int(n_qam)))

    def modulate_awgn(self, F, sigma, *, seed=173, first_frame=0, codeword=None, stream=None):
        """BitToSym / Modulate + complex AWGNChannel_CPU (LDPC_Encoder.cpp:6-68) on the device: returns a CUDA
        float32 tensor [F, N*p] (BPSK) or [F, N, 2] (QAM).  codeword: CUDA int16/uint16 tensor [N] or None."""
        import torch
        bpsk = self.n_const == 2
        x = torch.empty((F, self.N * self.p) if bpsk else (F, self.N, 2), dtype=torch.float32, device="cuda")
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        rc = lib.nb_ldpc_modulate_awgn(self._h, x.data_ptr(), int(F), float(sigma), int(seed), int(first_frame),
                                       codeword.data_ptr() if codeword is not None else None, st)
        if rc < 0:
            raise LdpcError(rc, "nb_ldpc_modulate_awgn")
        return x

    def decode(self, inp, iters=20, *, algo=ALGO_EMS, in_kind=IN_SYMBOL_LLR, sigma=1.0, ems_nm=2, ems_nc=2, stream=None):
        """inp: float32 [F, in_elems] numpy array (host path) or torch CUDA tensor (device path)."""
        per = self.in_elems(in_kind)
        is_torch = type(inp).__module__.startswith("torch")
        F = int(np.prod(inp.shape)) // per
        if F * per != int(np.prod(inp.shape)):
            raise ValueError(f"input of {tuple(inp.shape)} is not a multiple of {per} values per frame")
        o = NbDecodeOpts()
        lib.nb_decode_opts_default(C.byref(o))
        o.batch, o.algo, o.in_kind, o.ems_nm, o.ems_nc, o.sigma = F, algo, in_kind, ems_nm, ems_nc, float(sigma)
        if is_torch:
            import torch
            if not inp.is_cuda or not inp.is_contiguous() or inp.dtype != torch.float32:
                raise ValueError("device path wants a contiguous float32 CUDA tensor")
            out = torch.empty((F, self.N), dtype=torch.int16, device=inp.device)
            it = torch.empty(F, dtype=torch.int32, device=inp.device)
            ok = torch.empty(F, dtype=torch.int32, device=inp.device)
            o.mem_space = 1
            o.iters_out, o.ok_out = it.data_ptr(), ok.data_ptr()
            o.stream = stream if stream is not None else torch.cuda.current_stream(inp.device).cuda_stream
            rc = lib.nb_ldpc_decode_batch(self._h, inp.data_ptr(), out.data_ptr(), int(iters), C.byref(o))
        else:
            a = np.ascontiguousarray(inp, dtype=np.float32)
            out = np.zeros((F, self.N), np.uint16)
            it = np.zeros(F, np.int32)
            ok = np.zeros(F, np.int32)
            o.mem_space = 0
            o.iters_out, o.ok_out = it.ctypes.data, ok.ctypes.data
            o.stream = stream
            rc = lib.nb_ldpc_decode_batch(self._h, a.ctypes.data, out.ctypes.data, int(iters), C.byref(o))
        if rc < 0:
            raise LdpcError(rc, "nb_ldpc_decode_batch")
        return out, it, ok
