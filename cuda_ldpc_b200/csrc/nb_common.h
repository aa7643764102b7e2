// Internal definitions of the non-binary GF(q) decode path.  NB/ = myNBLDPC/ (gsw4869/CUDA_LDPC).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <vector>

#include "../../include/ldpc_b200.h"

struct nb_ldpc_code {
    int N, M, q, p, dv_max, dc_max, E;  // E = number of edges
    int n_const;
    float rate;
    // host tables
    std::vector<uint16_t> mul, inv;     // [q*q], [q]
    std::vector<int> vw, cw;            // weights
    std::vector<int> v_cn, v_gf, v_pos; // [N*dv_max]: check, coefficient, position inside that check
    std::vector<int> c_vn, c_gf, c_pos; // [M*dc_max]: variable, coefficient, position inside that variable
    std::vector<float> cre, cim;        // constellation
    // device tables (uploaded lazily by the first decode call)
    int device;
    int num_sms;
    uint16_t *d_mul, *d_inv;
    int *d_vw, *d_cw, *d_v_cn, *d_v_pos, *d_c_vn, *d_c_gf, *d_c_pos;
    float *d_cre, *d_cim;
    void *scratch;
    size_t scratch_bytes;
    // calls on one handle serialise on its scratch arena: `call_mu` for host threads, `last_use` (recorded on the
    // call's stream) is waited for by the next call's stream — same contract as the binary handle (common.h)
    std::mutex call_mu;
    cudaEvent_t last_use;
    bool last_use_valid;
    // encoder cache (host), built lazily by nb_ldpc_encode_info: information positions, pivot column per reduced row,
    // and the dense map pivot symbol i = sum_j enc_P[i][j] * info symbol j
    int enc_state;
    std::vector<int> enc_info_pos, enc_piv_col;
    std::vector<uint16_t> enc_P;
};
