// Kernel-side declarations (launch wrappers implemented in kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "dpq_internal.h"

namespace dpq {

constexpr int kMaxSmem = 232448;  // 227 KB opt-in per CTA on sm_100
constexpr uint32_t kInf31 = 0x7FFFFFFFu;

struct ScanGeom {
    int M, K, rb;      // rows = 1 << rb words per query-lane
    int pack;          // 1: one 31-bit query per lane; 2: two 15-bit queries per lane
    int qgl;           // query lanes per group (<= 32)
    int qpg;           // queries per group = qgl * pack
    int n_groups;      // ceil(Q / qpg)
    int n_slices;      // tree slices (chunk ranges) per group
    int n_warps;       // warps per CTA
    int kp;            // candidates kept per (warp, query): topk + slack, <= 256
    int kps;           // CTA-shared top-distance array length per query (kp, or 0 = disabled)
    int bcap;          // candidate buffer capacity per (warp, query): multiple of 32, >= 2*kp, <= 512
    int levels;        // depth stack levels
    size_t smem_bytes;
};

struct ScanArgs {
    ScanGeom g;
    const uint4* ops;
    const ChunkDesc* chunks;
    const uint8_t* anc;
    int n_chunks;
    const uint32_t* qlut;   // [n_groups][qgl][rows]   (swizzled, see lut kernel)
    uint64_t* cand;         // [n_items][n_warps][32*pack][bcap]
    uint32_t* gthr;         // [n_groups*qpg] exclusive distance bounds shared across slices
    uint32_t* cand_cnt;     // [n_items][n_warps][32*pack]
    int Q;
};

// ADC tables: exact float table [Q][M*K] + per-query scale + quantised swizzled table.
void launch_lut(const float* d_cw, int M, int K, int Ds, const float* d_queries, int Q,
                float* d_lutf, double* d_scale, uint32_t* d_qlut, uint32_t* d_gthr,
                const ScanGeom& g, cudaStream_t st);
// plain float tables only (dpq_adc_tables)
void launch_lut_plain(const float* d_cw, int M, int K, int Ds, const float* d_queries, int Q,
                      float* d_lutf, cudaStream_t st);

cudaError_t launch_scan(const ScanArgs& a, cudaStream_t st);

// ---- second-generation scan (scan2.cu) --------------------------------------------------
struct Scan2Args {
    V2Shape shape;
    const uint4* recs;            // one 16- or 32-byte record per node
    const ChunkDesc2* chunks;
    int n_chunks, chunk_nodes, rec_stride;
    const uint16_t* qlut;         // [n_groups][2048][56] fixed-point tables
    uint64_t* cand;               // [n_items][56][bcap] candidate keys (dist << 32 | pos)
    uint32_t* cand_cnt;           // [n_items][56]
    uint32_t* gthr;               // [n_groups*56] exclusive bounds shared across tree slices
    uint32_t* ovf;                // [n_groups*56] a candidate buffer overflowed: query needs the exact fallback
    int Q, n_groups, n_slices, n_warps;
    int kp, bcap, trigger, epoch, ramp;
};
void launch_lut2(const float* d_cw, int M, int K, int Ds, const float* d_queries, int Q, float* d_lutf,
                 double* d_scale, uint16_t* d_qlut, uint32_t* d_gthr, uint32_t* d_ovf, int n_groups,
                 const V2Shape& sh, uint32_t bound0, cudaStream_t st);
cudaError_t launch_scan2(const Scan2Args& a, cudaStream_t st);

// One ADC table entry in the reference's arithmetic (DCAT.h:3754-3757): float accumulator,
// each term the double square of the float difference, rounded back to float after every add.
__device__ __forceinline__ float adc_entry(const float* __restrict__ c, const float* q, int Ds) {
    float acc = 0.0f;
    for (int d = 0; d < Ds; ++d) {
        float diff = __fsub_rn(c[d], q[d]);
        double t = __dmul_rn((double)diff, (double)diff);
        acc = (float)__dadd_rn((double)acc, t);
    }
    return acc;
}

struct SelectArgs {
    ScanGeom g;
    int v2;                  // 1: lists are per (slice) with 56 queries per group (scan2.cu)
    const uint32_t* ovf;     // v2: per-query overflow flags
    const uint64_t* cand;
    const uint32_t* cand_cnt;
    const float* lutf;       // [Q][M*K]
    const double* scale;     // [Q]
    const uint8_t* codes;    // [n_local][M]
    int64_t base_pos;
    int64_t n_local;
    int Q, topk;
    uint64_t* out_key;       // [Q][topk]
    uint32_t* flagged;       // [max_flagged] compacted ids of queries needing the fallback
    uint32_t* n_flagged;     // [1]
    int max_flagged;
    float* bound;            // [Q]: exact distance bound for the fallback
    int force_fallback;
};
void launch_select(const SelectArgs& a, cudaStream_t st);

// Exact fallback for flagged queries: collect every node with exact distance <= bound.
struct FallbackArgs {
    const uint32_t* flagged;    // [max_flagged] query ids
    const uint32_t* n_flagged;  // [1] device-side count
    int max_flagged;
    const float* lutf;
    const float* bound;
    const uint8_t* codes;
    int64_t base_pos, n_local;
    int M, K, topk;
    uint64_t* buf;            // [max_flagged][cap]
    uint32_t* buf_cnt;        // [max_flagged]
    int cap;
    uint64_t* out_key;        // [Q][topk]
    uint32_t* overflow;       // [1]
};
void launch_fallback(const FallbackArgs& a, cudaStream_t st);

void launch_merge(const uint64_t* d_keys, int n_lists, int Q, int topk, uint64_t* d_out,
                  cudaStream_t st);

}  // namespace dpq
