// Tensor-core ground-truth filter (gt_tc.cu): arguments and launchers shared with secondary.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace dpq {

struct GtTcArgs {
    const float* base;      // [n][D] device
    const float* queries;   // [Q][D] device
    int64_t n;
    int Q, D;
    const float* x_nlo;     // [n] lower bound of ||x||^2
    const float* x_len;     // [n] upper bound of ||x||
    const float* thr;       // [Q] cap_q - ||q||^2 (upper bound)
    const float* qerr;      // [Q] c_err * ||q||
    uint32_t* cand;         // [Q][cand_cap] candidate indices into base
    uint32_t* cand_cnt;     // [Q] (may exceed cand_cap: overflow)
    int cand_cap;
    uint32_t* error;        // set to 1 when an MMA completion never arrived
    const unsigned char* q_split;  // pipelined form: bf16 hi / lo operand blocks (gt_split_kernel)
    const unsigned char* x_split;
};

cudaError_t launch_gt_prep(const float* x, int64_t n, int D, float* nlo, float* nhi, float* len, cudaStream_t st);
cudaError_t launch_gt_thr(const unsigned long long* state, int topk, int Q, const float* q_nlo, const float* q_nhi,
                          const float* q_len, float c_err, float* thr, float* qerr, cudaStream_t st);
cudaError_t launch_gt_tc_filter(const GtTcArgs& a, int n_sms, cudaStream_t st);
cudaError_t launch_gt_tc_filter2(const GtTcArgs& a, int n_sms, cudaStream_t st);
cudaError_t launch_gt_split(const float* src, int64_t n_rows, int D, bool queries, unsigned char* dst, cudaStream_t st);
size_t gt_split_bytes(int64_t n_rows, int D, bool queries);
cudaError_t launch_gt_rescore(const GtTcArgs& a, int64_t id0, int topk, unsigned long long* state, uint32_t* flagged,
                              uint32_t* n_flagged, cudaStream_t st);

}  // namespace dpq
