// Device-side helpers of the code-array engine (dpq_internal.h "v2"): a tree that was just built
// (dpq_index_open_tree) or decoded on the GPU keeps its codes by DFS position on the device and
// only needs them padded to the scan's word stride.
#include <cuda_runtime.h>

#include "dpq_internal.h"
#include "kernels.cuh"

namespace dpq {

// codes [n][M] -> out [n][stride] (stride 8 or 16), pad bytes 0; one thread per output 32-bit word
__global__ void pad_codes_kernel(const uint8_t* __restrict__ codes, int64_t n, int M, int stride,
                                 uint32_t* __restrict__ out) {
    const int wpn = stride / 4;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * wpn) return;
    const int64_t node = i / wpn;
    const int b0 = (int)(i % wpn) * 4;
    const uint8_t* c = codes + (size_t)node * M;
    uint32_t w = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
        if (b0 + b < M) w |= (uint32_t)c[b0 + b] << (8 * b);
    out[i] = w;
}

cudaError_t launch_pad_codes(const uint8_t* codes, int64_t n, int M, int stride, uint8_t* out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int64_t words = n * (stride / 4);
    const unsigned blocks = (unsigned)((words + 255) / 256);
    pad_codes_kernel<<<blocks, 256, 0, st>>>(codes, n, M, stride, reinterpret_cast<uint32_t*>(out));
    return cudaGetLastError();
}

}  // namespace dpq
