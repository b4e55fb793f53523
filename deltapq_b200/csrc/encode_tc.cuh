// Tensor-core PQ encoder (encode_tc.cu): arguments and launcher shared with secondary.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace dpq {

struct EncTcArgs {
    const float* x;   // [n][D] device
    int64_t n;
    int D, M, K;
    const float* cw;  // [M][K][Ds] device, exact centroids (the re-score reads these)
    const unsigned char* bsplit;  // [M][256 x 128 B] bf16 operand blocks (encode_tc_prep_kernel)
    const float* cmax;            // [M] upper bound of max_k ||c_k||
    uint8_t* codes;   // [n][M] device
    uint32_t* error;  // set to 1 when an MMA completion never arrived
    bool vec_ok;      // rows may be read with 16-byte loads
};

bool encode_tc_supported(int M, int K, int Ds);
size_t encode_tc_scratch_bytes(int M);
// d_scratch: encode_tc_scratch_bytes(M) bytes, 1024-byte aligned (cudaMalloc); *d_error must be 0 on entry
cudaError_t launch_encode_tc(const float* d_cw, int M, int K, int Ds, const float* d_x, int64_t n, int D, uint8_t* d_codes,
                             unsigned char* d_scratch, uint32_t* d_error, int n_sms, cudaStream_t st);

}  // namespace dpq
