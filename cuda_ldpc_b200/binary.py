"""Binary QC-LDPC decode path: Python mirror of the reference's interface.

Reference call sites this replaces (gsw4869/CUDA_LDPC, bldpc_实习/):
  Get_H + Transform_H (Simulation.cu:292-387)      -> LdpcCode(path, J, L, Z)
  LDPC_Decoder_GPU(D, Channel_Out, ...) (LDPC_Decoder.cu:23)  -> LdpcCode.decode(Channel_Out, maxIT)
The arrays keep the reference's meaning: Channel_Out is [N][F] fp32 with the frame index
fastest, D is int32 [(N+1)][F] whose last row is the per-frame flag.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import CodeInfo, DecodeOpts, LdpcError, lib

LAYOUT_NF, LAYOUT_FN = 0, 1
DTYPE_FP32, DTYPE_FP16, DTYPE_INT8, DTYPE_CHANNEL = 0, 1, 2, 3
MEM_HOST, MEM_DEVICE = 0, 1
SCHED_FLOODING, SCHED_LAYERED = 0, 1
EXIT_NONE, EXIT_GENIE, EXIT_SYNDROME = 0, 1, 2
OUT_INT32_REF, OUT_U8, OUT_BITPACK = 0, 1, 2

_NP_DTYPES = {np.dtype(np.float32): DTYPE_FP32, np.dtype(np.float16): DTYPE_FP16, np.dtype(np.int8): DTYPE_INT8}


def sigma_from_snr(snrtype, snr_db, rate):
    """main.cu:120-127."""
    return float(lib.ldpc_sigma(int(snrtype), float(snr_db), float(rate)))


@dataclass
class DecodeResult:
    D: object          # hard decisions in the requested format
    iters: object      # [F] iterations per frame
    ok: object         # [F] flag per frame
    launches: int      # kernel launches enqueued by the call
    app: object = None
    msgs: object = None


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class LdpcCode:
    def __init__(self, path, J=0, L=0, Z=0):
        h = C.c_void_p()
        rc = lib.ldpc_load_code(str(path).encode(), int(J), int(L), int(Z), C.byref(h))
        if rc != 0:
            raise LdpcError(rc, f"ldpc_load_code({path})")
        self._h = h
        info = CodeInfo()
        lib.ldpc_code_info(h, C.byref(info))
        for n, _ in CodeInfo._fields_:
            setattr(self, n, getattr(info, n))
        self.rate = self.K / self.N

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.ldpc_free_code(h)

    @property
    def handle(self):
        return self._h

    def tables(self):
        """H[J*L], Wc[J+1], Wv[L+1], Address_Variablenode[N*Wv_max] (fixed Transform_H)."""
        H = np.zeros(self.J * self.L, np.int32)
        Wc = np.zeros(self.J + 1, np.int32)
        Wv = np.zeros(self.L + 1, np.int32)
        addr = np.zeros(self.N * self.dv_max, np.int32)
        rc = lib.ldpc_code_tables(self._h, H.ctypes.data, Wc.ctypes.data, Wv.ctypes.data, addr.ctypes.data)
        if rc != 0:
            raise LdpcError(rc, "ldpc_code_tables")
        return H, Wc, Wv, addr

    def out_bytes(self, F, out_format):
        return int(lib.ldpc_out_bytes(self._h, int(F), int(out_format)))

    def wave_frames(self, msg_dtype=None):
        """frames one full wave of the layered kernel's resident CTAs decodes (ldpc_wave_frames): batches that are a
        multiple of it leave no CTA idle in the last round of a fixed-iteration decode"""
        n = int(lib.ldpc_wave_frames(self._h, int(DTYPE_INT8 if msg_dtype is None else msg_dtype)))
        if n < 0:
            raise LdpcError(n, "ldpc_wave_frames")
        return n

    def make_opts(self, F, **kw):
        o = DecodeOpts()
        lib.ldpc_decode_opts_default(C.byref(o))
        o.batch = int(F)
        for k, v in kw.items():
            if not hasattr(o, k):
                raise TypeError(f"unknown decode option {k}")
            setattr(o, k, v)
        return o

    def decode_channel(self, F, iters, sigma, *, seed=173, first_frame=0, codeword=None, early_exit=EXIT_SYNDROME,
                       out_format=OUT_BITPACK, llr_scale=8.0, msg_max=31, beta_num=0, beta_shift=0, stream=None, out=None,
                       iters_out=None, ok_out=None, device=None, msg_dtype=DTYPE_INT8):
        """Layered int8 (or fp16: msg_dtype) decode of F frames whose channel values are generated INSIDE the kernel
        (y = 1 - 2c + sigma*n, the Philox stream of ldpc_awgn_bpsk keyed by the global frame index):
        the fused form of AWGNChannel_CPU + cudaMemcpy + LDPC_Decoder_GPU (Simulation.cu:137-143).
        codeword: CUDA uint8 tensor [N] or None (all-zero).  Device path only."""
        import torch
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        if out is None:
            out = torch.empty(self.out_bytes(F, out_format), dtype=torch.uint8, device=dev)
        if iters_out is None:
            iters_out = torch.empty(F, dtype=torch.int32, device=dev)
        if ok_out is None:
            ok_out = torch.empty(F, dtype=torch.int32, device=dev)
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        o = self.make_opts(F, layout=LAYOUT_NF, llr_dtype=DTYPE_CHANNEL, mem_space=MEM_DEVICE, schedule=SCHED_LAYERED,
                           msg_dtype=msg_dtype, early_exit=early_exit, out_format=out_format, llr_scale=llr_scale,
                           msg_max=msg_max, beta_num=beta_num, beta_shift=beta_shift, iters_out=iters_out.data_ptr(),
                           ok_out=ok_out.data_ptr(), stream=stream, channel_sigma=float(sigma), channel_seed=int(seed),
                           channel_first_frame=int(first_frame),
                           channel_codeword=codeword.data_ptr() if codeword is not None else None)
        rc = lib.ldpc_decode_batch(self._h, None, out.data_ptr(), int(iters), C.byref(o))
        if rc < 0:
            raise LdpcError(rc, "ldpc_decode_batch")
        if out_format == OUT_INT32_REF:
            D = out.view(torch.int32).view(self.N + 1, F)
        elif out_format == OUT_U8:
            D = out.view(self.N, F)
        else:
            D = out.view(torch.int32).view(F, (self.N + 31) // 32)
        return DecodeResult(D, iters_out, ok_out, rc)

    def decode(self, llr, iters, *, schedule=SCHED_FLOODING, msg_dtype=None, layout=LAYOUT_NF, early_exit=EXIT_NONE,
               out_format=OUT_INT32_REF, alpha=1.0, llr_scale=8.0, msg_max=127, beta_num=0, beta_shift=0,
               stream=None, debug=False, out=None, iters_out=None, ok_out=None, host_pack_threads=0):
        """llr: numpy array (host path: copies in/out, synchronises) or torch CUDA tensor (device
        path: enqueues on `stream` / the current torch stream).  Shape [N, F] (LAYOUT_NF) or
        [F, N] (LAYOUT_FN).  host_pack_threads: see ldpc_decode_opts_t (host path, layered int8)."""
        if msg_dtype is None:
            msg_dtype = DTYPE_FP32 if schedule == SCHED_FLOODING else DTYPE_INT8
        N = self.N
        shape = tuple(llr.shape)
        if len(shape) != 2 or (shape[0] != N if layout == LAYOUT_NF else shape[1] != N):
            raise ValueError(f"llr shape {shape} does not match N={N} in layout {layout}")
        F = shape[1] if layout == LAYOUT_NF else shape[0]
        if _is_torch(llr):
            import torch
            if not llr.is_cuda or not llr.is_contiguous():
                raise ValueError("device path wants a contiguous CUDA tensor")
            dt = {torch.float32: DTYPE_FP32, torch.float16: DTYPE_FP16, torch.int8: DTYPE_INT8}[llr.dtype]
            dev = llr.device
            nbytes = self.out_bytes(F, out_format)
            if out is None:
                out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            if iters_out is None:
                iters_out = torch.empty(F, dtype=torch.int32, device=dev)
            if ok_out is None:
                ok_out = torch.empty(F, dtype=torch.int32, device=dev)
            app = msgs = None
            if debug:
                if schedule == SCHED_LAYERED and msg_dtype == DTYPE_INT8:
                    app = torch.zeros(N * F, dtype=torch.uint8, device=dev)
                    msgs = torch.zeros(self.M * self.dc_max * F, dtype=torch.int8, device=dev)
                elif schedule == SCHED_LAYERED and msg_dtype == DTYPE_FP16:
                    app = torch.zeros(N * F, dtype=torch.float16, device=dev)
                    msgs = torch.zeros(self.M * self.dc_max * F, dtype=torch.float16, device=dev)
                elif schedule == SCHED_LAYERED:
                    app = torch.zeros(N * F, dtype=torch.float32, device=dev)
                    msgs = torch.zeros(self.M * self.dc_max * F, dtype=torch.float32, device=dev)
                else:
                    msgs = torch.zeros(self.M * self.dc_max * F, dtype=torch.float32, device=dev)
            if stream is None:
                stream = torch.cuda.current_stream(dev).cuda_stream
            o = self.make_opts(F, layout=layout, llr_dtype=dt, mem_space=MEM_DEVICE, schedule=schedule,
                               msg_dtype=msg_dtype, early_exit=early_exit, out_format=out_format, alpha=alpha,
                               llr_scale=llr_scale, msg_max=msg_max, beta_num=beta_num, beta_shift=beta_shift,
                               iters_out=iters_out.data_ptr(), ok_out=ok_out.data_ptr(), stream=stream,
                               debug_app=app.data_ptr() if app is not None else None,
                               debug_msgs=msgs.data_ptr() if msgs is not None else None)
            rc = lib.ldpc_decode_batch(self._h, llr.data_ptr(), out.data_ptr(), int(iters), C.byref(o))
            if rc < 0:
                raise LdpcError(rc, "ldpc_decode_batch")
            if out_format == OUT_INT32_REF:
                D = out.view(torch.int32).view(N + 1, F)
            elif out_format == OUT_U8:
                D = out.view(N, F) if layout == LAYOUT_NF else out.view(F, N)
            else:
                D = out.view(torch.int32).view(F, (N + 31) // 32)
            if app is not None:
                app = app.view(torch.int8).view(N, F) if msg_dtype == DTYPE_INT8 else app.view(N, F)
            return DecodeResult(D, iters_out, ok_out, rc, app, msgs)
        # ---- host path
        a = np.ascontiguousarray(llr)
        if a.dtype not in _NP_DTYPES:
            a = a.astype(np.float32)
        dt = _NP_DTYPES[a.dtype]
        nbytes = self.out_bytes(F, out_format)
        # caller-provided (e.g. pinned) result buffers avoid a pageable D2H copy per call
        outb = np.zeros(nbytes, np.uint8) if out is None else out.view(np.uint8).reshape(-1)
        it = np.zeros(F, np.int32) if iters_out is None else iters_out
        ok = np.zeros(F, np.int32) if ok_out is None else ok_out
        if outb.size != nbytes or it.size != F or ok.size != F:
            raise ValueError("result buffers have the wrong size")
        app = msgs = None
        if debug:
            if schedule == SCHED_LAYERED and msg_dtype == DTYPE_INT8:
                app = np.zeros(N * F, np.int8)
                msgs = np.zeros(self.M * self.dc_max * F, np.int8)
            elif schedule == SCHED_LAYERED and msg_dtype == DTYPE_FP16:
                app = np.zeros(N * F, np.float16)
                msgs = np.zeros(self.M * self.dc_max * F, np.float16)
            elif schedule == SCHED_LAYERED:
                app = np.zeros(N * F, np.float32)
                msgs = np.zeros(self.M * self.dc_max * F, np.float32)
            else:
                msgs = np.zeros(self.M * self.dc_max * F, np.float32)
        o = self.make_opts(F, layout=layout, llr_dtype=dt, mem_space=MEM_HOST, schedule=schedule, msg_dtype=msg_dtype,
                           early_exit=early_exit, out_format=out_format, alpha=alpha, llr_scale=llr_scale,
                           msg_max=msg_max, beta_num=beta_num, beta_shift=beta_shift, iters_out=it.ctypes.data,
                           ok_out=ok.ctypes.data, stream=stream, host_pack_threads=int(host_pack_threads),
                           debug_app=app.ctypes.data if app is not None else None,
                           debug_msgs=msgs.ctypes.data if msgs is not None else None)
        rc = lib.ldpc_decode_batch(self._h, a.ctypes.data, outb.ctypes.data, int(iters), C.byref(o))
        if rc < 0:
            raise LdpcError(rc, "ldpc_decode_batch")
        if out_format == OUT_INT32_REF:
            D = outb.view(np.int32).reshape(N + 1, F)
        elif out_format == OUT_U8:
            D = outb.reshape(N, F) if layout == LAYOUT_NF else outb.reshape(F, N)
        else:
            D = outb.view(np.uint32).reshape(F, (N + 31) // 32)
        if app is not None:
            app = app.reshape(N, F)
        return DecodeResult(D, it, ok, rc, app, msgs)
