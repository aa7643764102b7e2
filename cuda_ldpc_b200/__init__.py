"""cuda_ldpc_b200 — host-side mirror (ctypes) of the C-ABI in include/ldpc_b200.h.

The product is libldpc_b200.so (hand-written CUDA for sm_100a behind a C ABI); this package
only loads it and marshals numpy / torch buffers.  There is no CPU fallback: importing works
without a GPU (loaders, table access), decoding raises LdpcError(LDPC_ERR_NO_DEVICE).
"""
from ._lib import LdpcError, lib, lib_path  # noqa: F401
from .binary import (  # noqa: F401
    LdpcCode, DecodeResult, sigma_from_snr,
    LAYOUT_NF, LAYOUT_FN, DTYPE_FP32, DTYPE_FP16, DTYPE_INT8, SCHED_FLOODING, SCHED_LAYERED,
    EXIT_NONE, EXIT_GENIE, EXIT_SYNDROME, OUT_INT32_REF, OUT_U8, OUT_BITPACK,
)

from .nonbinary import NbLdpcCode, ALGO_EMS, ALGO_TMM, ALGO_LAYERED_TMM, ALGO_FFT_BP, IN_SYMBOL_LLR, IN_BPSK, IN_QAM  # noqa: F401,E402

DATA_DIR = __import__("os").path.join(__import__("os").path.dirname(__file__), "data")
