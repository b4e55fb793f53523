// Device-side tree compiler: the fixed-record scan program (dpq_internal.h "v2") straight from the
// layout arrays of a tree that was just built (codes by position, parent position, depth), one
// thread per 64-node chunk.  Same records as program.cpp's Emitter2 produces from the byte stream
// -- that path decodes a strictly sequential code on one host core (8.9 s per 125M nodes); chunks
// are independent here.  tests/test_gpu_parity.py compares searches through both.
#include <cuda_runtime.h>

#include "dpq_internal.h"
#include "kernels.cuh"

namespace dpq {

__global__ void build_recs_kernel(const uint8_t* __restrict__ codes_p, const uint32_t* __restrict__ parent_pos,
                                  const uint8_t* __restrict__ depth, int64_t n, int M, int K, int nf, int lpg,
                                  int chunk_nodes, uint32_t pos_shift, uint32_t* __restrict__ recs,
                                  ChunkDesc2* __restrict__ chunks, unsigned long long* __restrict__ n_delta) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t first = c * chunk_nodes;
    if (first >= n) return;
    const int cnt = (int)(n - first < chunk_nodes ? n - first : chunk_nodes);
    const int half = nf / 2, rw = nf / 2;
    ChunkDesc2 cd;
    cd.rec_begin = (uint32_t)first;
    cd.n_nodes = (uint32_t)cnt;
    cd.first_pos = (uint32_t)first + pos_shift;
    cd.pad = 0;
    chunks[c] = cd;
    int prev_depth = 0, reg_owner_depth = -1;
    unsigned long long deltas = 0;
    for (int i = 0; i < cnt; ++i) {
        const int64_t p = first + i;
        const int d = depth[p];
        const uint8_t* cur = codes_p + (size_t)p * M;
        const uint8_t* par = p == 0 ? cur : codes_p + (size_t)parent_pos[p] * M;
        if (i > 0 && d == prev_depth + 1) {  // the previous node is this node's parent
            recs[(size_t)(p - 1) * rw] |= V2_CHILD;
            reg_owner_depth = prev_depth;
        } else if (reg_owner_depth >= d) {
            reg_owner_depth = -1;  // the register's node left the path
        }
        uint32_t f[16];
        int nd = 0;
        for (int m = 0; m < M; ++m) nd += par[m] != cur[m];
        const bool abs = !(reg_owner_depth >= 0 && reg_owner_depth == d - 1 && nd <= half);
        if (!abs) {
            for (int j = 0; j < nf; ++j) f[j] = 0;
            int j = 0;
            for (int m = 0; m < M; ++m)
                if (par[m] != cur[m]) {
                    f[j] = (uint32_t)(m * K + cur[m]) * (uint32_t)lpg;         // plus: new centroid
                    f[half + j] = (uint32_t)(m * K + par[m]) * (uint32_t)lpg;  // minus: old centroid
                    ++j;
                }
            ++deltas;
        } else {
            const uint32_t zero_row = (uint32_t)(M * K) * (uint32_t)lpg;
            for (int j = 0; j < nf; ++j) f[j] = j < M ? (uint32_t)(j * K + cur[j]) * (uint32_t)lpg : zero_row;
        }
        uint32_t* r = recs + (size_t)p * rw;
        for (int w = 0; w < rw; ++w) r[w] = f[2 * w] | (f[2 * w + 1] << 16);
        if (abs) r[0] |= V2_ABS;
        prev_depth = d;
    }
    if (deltas) atomicAdd(n_delta, deltas);
}

cudaError_t launch_build_recs(const uint8_t* codes_p, const uint32_t* parent_pos, const uint8_t* depth, int64_t n, int M,
                              int K, const V2Shape& sh, int chunk_nodes, uint32_t pos_shift, uint32_t* recs,
                              ChunkDesc2* chunks, unsigned long long* n_delta, cudaStream_t st) {
    const int64_t n_chunks = (n + chunk_nodes - 1) / chunk_nodes;
    build_recs_kernel<<<(unsigned)((n_chunks + 127) / 128), 128, 0, st>>>(codes_p, parent_pos, depth, n, M, K, sh.nf, sh.lpg,
                                                                          chunk_nodes, pos_shift, recs, chunks, n_delta);
    return cudaGetLastError();
}

}  // namespace dpq
