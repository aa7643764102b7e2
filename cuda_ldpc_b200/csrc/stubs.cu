// Entry points declared in include/ldpc_b200.h that are not implemented yet return
// LDPC_ERR_UNSUPPORTED (they never fall back to the CPU).
#include "common.h"
extern "C" int ldpc_awgn_bpsk(const ldpc_code_t *, float *, int, int, float, uint64_t, uint64_t, const uint8_t *, void *) { return LDPC_ERR_UNSUPPORTED; }
extern "C" int ldpc_statistic(const ldpc_code_t *, const int *, const int *, int, int, const uint8_t *, int64_t *, void *) { return LDPC_ERR_UNSUPPORTED; }
extern "C" int ldpc_encode(ldpc_code_t *, const uint8_t *, uint8_t *) { return LDPC_ERR_UNSUPPORTED; }

