// DeltaTree build, stage 1: edge search ("find_edge").
//
// Reference: driver find_edges_by_diff_approx (DCAT.h:1207-1332), one round
// partition_linear_opt_approx_with_constraint (DCAT.h:445-627, method 1) and its "WOH"
// variant (DCAT.h:629-792, method 2); kept-subspace combinations nchoosek (CT.h:75-90).
// Canonical tie rule = every sort stable: the reference's unstable
// sort makes its own output thread-count dependent.
//
// Device formulation.  All pool state lives in HBM for the whole build: the codes, heights,
// the round's id pool and its merged flags.  One
// kept-subspace combination is a fixed sequence of launches, no host logic in between:
//
//   mask_key_kernel    key = packed code with the dropped subspaces blanked   (DCAT.h:495-511)
//   cub radix sort     stable (key, pool slot) sort                           (DCAT.h:524-528)
//   head_kernel        run heads: key differs from the predecessor (XOR != 0)
//   cub inclusive sum  run index per element
//   run_best_kernel    per run: first member with the largest height          (DCAT.h:547-558)
//   run_second_kernel  per run: largest height among the other members        (DCAT.h:561-568)
//   run_decide_kernel  height bump, finalist decision, edge count per run     (DCAT.h:569-576)
//   cub exclusive sum  edge / finalist output offsets in run order
//   star_emit_kernel   parent -> member edges in run order, merged flags      (DCAT.h:577-598)
//   cub select         the still-unmerged pool slots, ascending               (DCAT.h:486-490)
//
// Runs are independent within one combination (every element is in exactly one run and a
// run only touches the heights of its own members), so the reference's serial walk and this
// data-parallel form emit identical edges in identical order.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/cub.cuh>
#include <string>
#include <vector>

#include "../../include/dpq.h"

namespace dpq {
int api_fail(int code, const std::string& msg);
int api_check_device();
int api_device();
int check_code_range(const uint8_t* codes, int64_t n, int M, int K);  // layout.cu
}  // namespace dpq

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return dpq::api_fail(DPQ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

namespace {

struct Buf {
    void* p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
    template <class T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
    ~Buf() {
        if (p) cudaFree(p);
    }
};

// key of one code under a kept-subspace selector: OR of the kept centroid ids shifted by
// LOG_K * m (DCAT.h:495-511), as two 64-bit halves.  Computed from the code bytes exactly as
// the reference does, so K that is not a power of two behaves the same.
__device__ __forceinline__ void masked_key(const uint8_t* __restrict__ codes, uint32_t id, int M, int log_k,
                                           uint32_t sel, uint64_t& lo, uint64_t& hi) {
    lo = 0;
    hi = 0;
    if (M == 8 && log_k == 8) {  // the packed code itself, dropped bytes blanked
        const uint64_t c = *reinterpret_cast<const uint64_t*>(codes + (size_t)id * 8);
        uint64_t mask = 0;
#pragma unroll
        for (int m = 0; m < 8; ++m) mask |= ((sel >> m) & 1u) ? (0xFFull << (8 * m)) : 0ull;
        lo = c & mask;
        return;
    }
    for (int m = 0; m < M; ++m) {
        if (!((sel >> m) & 1u)) continue;
        const uint64_t c = codes[(size_t)id * M + m];
        const int sh = log_k * m;
        if (sh < 64) {
            lo |= c << sh;
            if (sh > 56) hi |= c >> (64 - sh);
        } else if (sh < 128) {
            hi |= c << (sh - 64);
        }
    }
}

__global__ void iota_kernel(uint32_t* a, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (uint32_t)i;
}

// compact != 0: the kept centroid ids packed next to each other (subspace order kept).  Two blanked
// 128-bit keys compare exactly like their compacted forms -- the dropped positions are zero in both
// -- so sorting by the compact key gives the reference's order with kept * log_k key bits instead of
// 128 (one radix sort, fewer digit passes).
__global__ void mask_key_kernel(const uint32_t* __restrict__ live, int64_t n_live,
                                const uint32_t* __restrict__ ids, const uint8_t* __restrict__ codes, int M,
                                int log_k, uint32_t sel, int compact, uint64_t* __restrict__ key_lo,
                                uint32_t* __restrict__ slot) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_live) return;
    const uint32_t s = live[l];
    uint64_t lo, hi;
    if (compact) {
        const uint8_t* c = codes + (size_t)ids[s] * M;
        lo = 0;
        int sh = 0;
        for (int m = 0; m < M; ++m)
            if ((sel >> m) & 1u) {
                lo |= (uint64_t)c[m] << sh;
                sh += log_k;
            }
    } else {
        masked_key(codes, ids[s], M, log_k, sel, lo, hi);
    }
    key_lo[l] = lo;
    slot[l] = s;
}

// after the first sort: the high halves in sorted order (pass 2 of the 128-bit LSD sort)
__global__ void gather_hi_kernel(const uint32_t* __restrict__ slot, int64_t n, const uint32_t* __restrict__ ids,
                                 const uint8_t* __restrict__ codes, int M, int log_k, uint32_t sel,
                                 uint64_t* __restrict__ key_hi) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t lo, hi;
    masked_key(codes, ids[slot[i]], M, log_k, sel, lo, hi);
    key_hi[i] = hi;
}

// run heads over the sorted order: XOR of neighbouring keys != 0
__global__ void head_kernel(const uint32_t* __restrict__ slot, int64_t n, const uint32_t* __restrict__ ids,
                            const uint8_t* __restrict__ codes, int M, int log_k, uint32_t sel,
                            uint32_t* __restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = 1;
    if (i > 0) {
        uint64_t alo, ahi, blo, bhi;
        masked_key(codes, ids[slot[i]], M, log_k, sel, alo, ahi);
        masked_key(codes, ids[slot[i - 1]], M, log_k, sel, blo, bhi);
        h = ((alo ^ blo) | (ahi ^ bhi)) != 0;
    }
    head[i] = h;
}

struct RunState {
    unsigned long long* best;  // [runs] (height + 1) << 32 | ~index : max = first member with the largest height
    uint32_t* second;          // [runs] largest height among the non-parent members
    uint32_t* start;           // [runs + 1] first sorted index of each run
    unsigned long long* cnt;   // [runs] finalist << 40 | edges of the run (exclusive-summed in place)
};

__global__ void run_best_kernel(const uint32_t* __restrict__ slot, int64_t n, const uint32_t* __restrict__ ids,
                                const uint8_t* __restrict__ heights, const uint32_t* __restrict__ head,
                                const uint32_t* __restrict__ run_of, RunState rs, int method) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;  // every lane takes part in the shuffles
    uint32_t r = 0xFFFFFFFFu;
    unsigned long long v = 0;
    if (valid) {
        r = run_of[i] - 1;
        if (head[i]) rs.start[r] = (uint32_t)i;
        if (i == n - 1) rs.start[r + 1] = (uint32_t)n;
        v = ((unsigned long long)(heights[ids[slot[i]]] + 1u) << 32) | (0xFFFFFFFFu - (uint32_t)i);
    }
    if (method == 2) return;  // WOH: the parent is the run's first member (DCAT.h:731)
    // members of one run are adjacent: the first lane of each same-run group of the warp
    // carries the group's maximum, so a long run costs one atomic per warp, not per member
    const int lane = threadIdx.x & 31;
    unsigned long long m = v;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_down_sync(0xffffffffu, m, o);
        const uint32_t tr = __shfl_down_sync(0xffffffffu, r, o);
        if (lane + o < 32 && tr == r && t > m) m = t;
    }
    const uint32_t pr = __shfl_up_sync(0xffffffffu, r, 1);
    if (valid && (lane == 0 || pr != r)) atomicMax(&rs.best[r], m);
}

__global__ void run_second_kernel(const uint32_t* __restrict__ slot, int64_t n, const uint32_t* __restrict__ ids,
                                  const uint8_t* __restrict__ heights, const uint32_t* __restrict__ run_of,
                                  RunState rs, int method) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = run_of[i] - 1;
    const uint32_t parent_i = method == 2 ? rs.start[r] : 0xFFFFFFFFu - (uint32_t)rs.best[r];
    if ((uint32_t)i == parent_i) return;
    const uint32_t h = heights[ids[slot[i]]];
    if (h) atomicMax(&rs.second[r], h);  // "second" starts from 0 (DCAT.h:561)
}

__global__ void run_decide_kernel(int64_t n_runs, const uint32_t* __restrict__ slot,
                                  const uint32_t* __restrict__ ids, uint8_t* __restrict__ heights, RunState rs,
                                  int method, int max_height) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_runs) return;
    const uint32_t len = rs.start[r + 1] - rs.start[r];
    unsigned long long out = 0;
    if (len >= 2) {
        const uint32_t parent_i = method == 2 ? rs.start[r] : 0xFFFFFFFFu - (uint32_t)rs.best[r];
        const uint32_t pc = ids[slot[parent_i]];
        const int second = (int)rs.second[r];
        bool finalist;
        if (method == 2) {  // DCAT.h:731-743
            int h = heights[pc];
            if (second + 1 > h) {
                h = second + 1;
                heights[pc] = (uint8_t)h;
            }
            finalist = h >= max_height - 2;
        } else {  // DCAT.h:545-575
            const int top = (int)(rs.best[r] >> 32) - 1;
            if (second == top) heights[pc] = (uint8_t)(heights[pc] + 1);
            finalist = top + 1 >= max_height - 2;
        }
        out = ((unsigned long long)finalist << 40) | (unsigned long long)(len - 1);
    }
    rs.cnt[r] = out;
}

__global__ void star_emit_kernel(const uint32_t* __restrict__ slot, int64_t n, const uint32_t* __restrict__ ids,
                                 const uint32_t* __restrict__ run_of, RunState rs,
                                 const unsigned long long* __restrict__ off, int method,
                                 uint8_t* __restrict__ is_merged, uint32_t* __restrict__ edges, int64_t edge_base,
                                 uint32_t* __restrict__ finalists, int64_t fin_base) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = run_of[i] - 1;
    const uint32_t s0 = rs.start[r];
    if (rs.start[r + 1] - s0 < 2) return;
    const uint32_t parent_i = method == 2 ? s0 : 0xFFFFFFFFu - (uint32_t)rs.best[r];
    const unsigned long long o = off[r], c = rs.cnt[r];
    const uint32_t my_slot = slot[i];
    const uint32_t pc = ids[slot[parent_i]];
    if ((uint32_t)i == parent_i) {
        if (c >> 40) {  // finalist: leaves the pool but stays a tree root candidate
            finalists[fin_base + (int64_t)(o >> 40)] = pc;
            is_merged[my_slot] = 1;
        }
        return;
    }
    const int64_t e = edge_base + (int64_t)(o & ((1ull << 40) - 1)) + ((int64_t)i - s0) - ((uint32_t)i > parent_i);
    edges[2 * e] = pc;
    edges[2 * e + 1] = ids[my_slot];
    is_merged[my_slot] = 1;
}

struct NotMerged {
    const uint8_t* is_merged;
    __device__ bool operator()(uint32_t s) const { return !is_merged[s]; }
};

__global__ void gather_ids_kernel(const uint32_t* __restrict__ live, int64_t n, const uint32_t* __restrict__ ids,
                                  uint32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ids[live[i]];
}

inline unsigned blocks(int64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

// ---------------------------------------------------------------------------------------------
// Futile-pass prefilter (M >= 12: thousands of kept-subspace combinations).  A pass with kept set S
// merges something only if two pool codes agree on all of S, i.e. their changed-subspace mask is a
// subset of the dropped set D.  The pool only ever shrinks, so the masks of all pairs of ORIGINAL
// codes bound every later pass from above.  Every pair that differs in at most M - 2 subspaces
// shares at least one pair of subspaces: bucketing the codes by each of the M (M - 1) / 2
// two-subspace keys enumerates all those pairs.  active[m] = some pair has changed-subspace mask m;
// the host then skips a pass when no active mask is a subset of its dropped set (exactly the passes
// that would have found all keys distinct and changed nothing).
__global__ void pair_key_kernel(const uint8_t* __restrict__ codes, int64_t n, int M, int sa, int sb,
                                uint32_t* __restrict__ key, uint32_t* __restrict__ id) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = (uint32_t)codes[(size_t)i * M + sa] | ((uint32_t)codes[(size_t)i * M + sb] << 8);
    id[i] = (uint32_t)i;
}

// bucket work = sum over elements of the number of later elements in the same bucket
__global__ void pair_work_kernel(const uint32_t* __restrict__ key, int64_t n, unsigned long long* __restrict__ work) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long w = 0;
    if (i < n && (i == 0 || key[i - 1] != key[i])) {  // bucket head: walk to its end
        int64_t j = i + 1;
        while (j < n && key[j] == key[i]) ++j;
        const unsigned long long len = (unsigned long long)(j - i);
        w = len * (len - 1) / 2;
    }
    for (int o = 16; o; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(work, w);
}

__global__ void pair_mask_kernel(const uint32_t* __restrict__ key, const uint32_t* __restrict__ id, int64_t n,
                                 const uint8_t* __restrict__ codes, int M, uint32_t* __restrict__ active) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = key[i];
    uint8_t x[16];
    const uint8_t* cx = codes + (size_t)id[i] * M;
    for (int m = 0; m < M; ++m) x[m] = cx[m];
    for (int64_t j = i + 1; j < n && key[j] == k; ++j) {
        const uint8_t* cy = codes + (size_t)id[j] * M;
        uint32_t mask = 0;
        for (int m = 0; m < M; ++m) mask |= (uint32_t)(x[m] != cy[m]) << m;
        const uint32_t bit = 1u << (mask & 31u);
        if (!(active[mask >> 5] & bit)) atomicOr(&active[mask >> 5], bit);
    }
}

// hit[D] != 0 iff some pair of codes has exactly the changed-subspace mask D.  Empty when the
// prefilter does not apply (few combinations, K > 256 keys, or buckets too large to enumerate).
int futile_pass_prefilter(const uint8_t* d_codes, int64_t n, int M, std::vector<uint8_t>* hit) {
    hit->clear();
    if (M < 12 || M > 16 || n < 2) return DPQ_OK;
    Buf d_key, d_key2, d_id, d_id2, d_active, d_work, d_tmp;
    CU(d_key.alloc((size_t)n * 4));
    CU(d_key2.alloc((size_t)n * 4));
    CU(d_id.alloc((size_t)n * 4));
    CU(d_id2.alloc((size_t)n * 4));
    CU(d_active.alloc(((size_t)1 << M) / 8));
    CU(d_work.alloc(8));
    size_t tmp_bytes = 0;
    {
        cub::DoubleBuffer<uint32_t> kb(d_key.as<uint32_t>(), d_key2.as<uint32_t>());
        cub::DoubleBuffer<uint32_t> vb(d_id.as<uint32_t>(), d_id2.as<uint32_t>());
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)n, 0, 16));
    }
    CU(d_tmp.alloc(tmp_bytes));
    CU(cudaMemset(d_active.p, 0, ((size_t)1 << M) / 8));
    // at most ~2 x 10^10 pair comparisons in total (about a second); denser data keeps every pass
    const unsigned long long budget = 20000000000ull;
    unsigned long long spent = 0;
    for (int sa = 0; sa < M; ++sa)
        for (int sb = sa + 1; sb < M; ++sb) {
            pair_key_kernel<<<blocks(n), 256>>>(d_codes, n, M, sa, sb, d_key.as<uint32_t>(), d_id.as<uint32_t>());
            cub::DoubleBuffer<uint32_t> kb(d_key.as<uint32_t>(), d_key2.as<uint32_t>());
            cub::DoubleBuffer<uint32_t> vb(d_id.as<uint32_t>(), d_id2.as<uint32_t>());
            size_t t = tmp_bytes;
            CU(cub::DeviceRadixSort::SortPairs(d_tmp.p, t, kb, vb, (int)n, 0, 16));
            CU(cudaMemsetAsync(d_work.p, 0, 8));
            pair_work_kernel<<<blocks(n), 256>>>(kb.Current(), n, d_work.as<unsigned long long>());
            unsigned long long w = 0;
            CU(cudaMemcpy(&w, d_work.p, 8, cudaMemcpyDeviceToHost));
            spent += w;
            if (spent > budget) return DPQ_OK;  // too dense: no prefilter
            pair_mask_kernel<<<blocks(n), 256>>>(kb.Current(), vb.Current(), n, d_codes, M, d_active.as<uint32_t>());
            CU(cudaGetLastError());
        }
    std::vector<uint32_t> active(((size_t)1 << M) / 32);
    CU(cudaMemcpy(active.data(), d_active.p, active.size() * 4, cudaMemcpyDeviceToHost));
    // A pair with changed-subspace mask m lands in one run in the pass whose dropped set IS m (round
    // |m|), and every run keeps at most its parent in the pool: afterwards no two pool codes have mask
    // m.  So when a pass with dropped set D runs, every surviving pair with mask inside D has mask
    // exactly D, and the pass can merge something only if active[D] (exact duplicates, mask 0, only
    // matter to the first pass).
    hit->assign((size_t)1 << M, 0);
    for (size_t m = 0; m < hit->size(); ++m) (*hit)[m] = (active[m >> 5] >> (m & 31)) & 1u;
    return DPQ_OK;
}

}  // namespace

// edges: host output (nullable when d_edges_out is given); d_edges_out: the device array is handed to the caller
static int find_edges_core(const uint8_t* codes, int64_t n_codes, int M, int K, int max_height_folds, int method,
                           uint32_t* edges, uint32_t** d_edges_out, uint32_t* root_id) {
    if (!codes || !root_id || (n_codes > 1 && !edges && !d_edges_out) || n_codes < 1 || M < 1 || M > 16 || K < 1 || K > 256 ||
        max_height_folds < 1 || (method != 1 && method != 2))
        return dpq::api_fail(DPQ_ERR_ARG, "dpq_find_edges: bad argument (1<=M<=16, K<=256, method 1|2)");
    if (n_codes >= 0x7FFFFFFFLL) return dpq::api_fail(DPQ_ERR_ARG, "dpq_find_edges: n_codes must be < 2^31-1 (DCAT.h:982)");
    int rc = dpq::api_check_device();
    if (rc) return rc;
    CU(cudaSetDevice(dpq::api_device()));
    if ((rc = dpq::check_code_range(codes, n_codes, M, K))) return rc;
    const int log_k = (int)std::lround(std::log2((double)K));  // DCAT.h:454
    const int max_height = M * max_height_folds;               // DCAT.h:1262
    const int key_bits = std::min(128, log_k * (M - 1) + 8);
    const bool wide = key_bits > 64;
    const int64_t n = n_codes;

    Buf d_codes, d_heights, d_ids, d_ids2, d_live, d_live2, d_merged, d_klo, d_klo2, d_khi, d_khi2,
        d_slot, d_slot2, d_head, d_runof, d_best, d_second, d_start, d_cnt, d_off, d_edges, d_fin, d_nsel, d_tmp;
    CU(d_codes.alloc((size_t)n * M + 8));
    CU(d_heights.alloc((size_t)n));
    CU(d_ids.alloc((size_t)n * 4));
    CU(d_ids2.alloc((size_t)n * 4));
    CU(d_live.alloc((size_t)n * 4));
    CU(d_live2.alloc((size_t)n * 4));
    CU(d_merged.alloc((size_t)n));
    CU(d_klo.alloc((size_t)n * 8));
    CU(d_klo2.alloc((size_t)n * 8));
    if (wide) {
        CU(d_khi.alloc((size_t)n * 8));
        CU(d_khi2.alloc((size_t)n * 8));
    }
    CU(d_slot.alloc((size_t)n * 4));
    CU(d_slot2.alloc((size_t)n * 4));
    CU(d_head.alloc((size_t)n * 4));
    CU(d_runof.alloc((size_t)n * 4));
    CU(d_best.alloc((size_t)n * 8));
    CU(d_second.alloc((size_t)n * 4));
    CU(d_start.alloc((size_t)(n + 1) * 4));
    CU(d_cnt.alloc((size_t)n * 8));
    CU(d_off.alloc((size_t)n * 8));
    CU(d_edges.alloc((size_t)std::max<int64_t>(n - 1, 1) * 8));
    CU(d_fin.alloc((size_t)n * 4));
    CU(d_nsel.alloc(16));

    // one temp buffer sized for the largest cub request at n items
    size_t tmp_bytes = 0, t = 0;
    {
        cub::DoubleBuffer<uint64_t> kb(d_klo.as<uint64_t>(), d_klo2.as<uint64_t>());
        cub::DoubleBuffer<uint32_t> vb(d_slot.as<uint32_t>(), d_slot2.as<uint32_t>());
        CU(cub::DeviceRadixSort::SortPairs(nullptr, t, kb, vb, (int)n, 0, 64));
        tmp_bytes = std::max(tmp_bytes, t);
        CU(cub::DeviceScan::InclusiveSum(nullptr, t, d_head.as<uint32_t>(), d_runof.as<uint32_t>(), (int)n));
        tmp_bytes = std::max(tmp_bytes, t);
        CU(cub::DeviceScan::ExclusiveSum(nullptr, t, d_cnt.as<unsigned long long>(), d_off.as<unsigned long long>(), (int)n));
        tmp_bytes = std::max(tmp_bytes, t);
        CU(cub::DeviceSelect::If(nullptr, t, d_live.as<uint32_t>(), d_live2.as<uint32_t>(), d_nsel.as<int>(), (int)n,
                                 NotMerged{d_merged.as<uint8_t>()}));
        tmp_bytes = std::max(tmp_bytes, t);
    }
    CU(d_tmp.alloc(tmp_bytes));

    CU(cudaMemcpy(d_codes.p, codes, (size_t)n * M, cudaMemcpyDefault));  // host or device source
    CU(cudaMemset(d_heights.p, 0, (size_t)n));
    iota_kernel<<<blocks(n), 256>>>(d_ids.as<uint32_t>(), n);
    CU(cudaGetLastError());

    // passes that cannot merge anything are skipped (exact: see futile_pass_prefilter)
    std::vector<uint8_t> hit;
    if (!getenv("DPQ_NO_PREFILTER") && K <= 256) {
        rc = futile_pass_prefilter(d_codes.as<uint8_t>(), n, M, &hit);
        if (rc) return rc;
    }
    const uint32_t all_bits = M >= 32 ? 0xFFFFFFFFu : ((1u << M) - 1u);
    int64_t skipped_passes = 0;

    uint32_t* ids = d_ids.as<uint32_t>();
    uint32_t* ids_alt = d_ids2.as<uint32_t>();
    uint32_t* live = d_live.as<uint32_t>();
    uint32_t* live_alt = d_live2.as<uint32_t>();
    const uint8_t* dc = d_codes.as<uint8_t>();
    uint8_t* heights = d_heights.as<uint8_t>();
    uint8_t* is_merged = d_merged.as<uint8_t>();
    RunState rs{d_best.as<unsigned long long>(), d_second.as<uint32_t>(), d_start.as<uint32_t>(),
                d_cnt.as<unsigned long long>()};
    int64_t n_ids = n, n_edges = 0, n_fin = 0;

    const bool stats = getenv("DPQ_EDGE_STATS") != nullptr;  // per-round pool size / passes / seconds on stderr
    for (int diff = 0; diff <= M && n_ids > 1; ++diff) {  // dmain:126 forces diff_argument = M
        const auto round_t0 = std::chrono::steady_clock::now();
        const int64_t round_n0 = n_ids;
        int64_t round_passes = 0, round_live_sum = 0;
        CU(cudaMemsetAsync(is_merged, 0, (size_t)n_ids));
        iota_kernel<<<blocks(n_ids), 256>>>(live, n_ids);
        int64_t n_live = n_ids;
        std::vector<char> sel((size_t)M);
        for (int m = 0; m < M; ++m) sel[(size_t)m] = m < M - diff;
        do {  // one kept-subspace combination
            if (n_live < 2) break;  // no run of two can form any more in this round
            uint32_t selbits = 0;
            for (int m = 0; m < M; ++m) selbits |= sel[(size_t)m] ? (1u << m) : 0u;
            // pairs that differ in more than M - 2 subspaces were not enumerated: those rounds always run
            if (!hit.empty() && diff <= M - 2 && !hit[(size_t)(~selbits & all_bits)]) {
                ++skipped_passes;
                continue;
            }
            ++round_passes;
            round_live_sum += n_live;
            // a centroid id needs log_k bits unless K is not a power of two (ids up to K - 1 < 2^log_k may
            // not hold then): the compact form is used only when every id fits its field
            const int kept = M - diff;
            const bool compact = kept * log_k <= 64 && (1 << log_k) >= K;
            mask_key_kernel<<<blocks(n_live), 256>>>(live, n_live, ids, dc, M, log_k, selbits, compact ? 1 : 0,
                                                     d_klo.as<uint64_t>(), d_slot.as<uint32_t>());
            cub::DoubleBuffer<uint64_t> kb(d_klo.as<uint64_t>(), d_klo2.as<uint64_t>());
            cub::DoubleBuffer<uint32_t> vb(d_slot.as<uint32_t>(), d_slot2.as<uint32_t>());
            t = tmp_bytes;
            CU(cub::DeviceRadixSort::SortPairs(d_tmp.p, t, kb, vb, (int)n_live, 0,
                                               compact ? std::max(1, kept * log_k) : std::min(64, key_bits)));
            if (wide && !compact) {  // LSD over the two halves: stable sort by the high half second
                gather_hi_kernel<<<blocks(n_live), 256>>>(vb.Current(), n_live, ids, dc, M, log_k, selbits, d_khi.as<uint64_t>());
                cub::DoubleBuffer<uint64_t> kh(d_khi.as<uint64_t>(), d_khi2.as<uint64_t>());
                t = tmp_bytes;
                CU(cub::DeviceRadixSort::SortPairs(d_tmp.p, t, kh, vb, (int)n_live, 0, key_bits - 64));
            }
            const uint32_t* slot = vb.Current();
            head_kernel<<<blocks(n_live), 256>>>(slot, n_live, ids, dc, M, log_k, selbits, d_head.as<uint32_t>());
            t = tmp_bytes;
            CU(cub::DeviceScan::InclusiveSum(d_tmp.p, t, d_head.as<uint32_t>(), d_runof.as<uint32_t>(), (int)n_live));
            CU(cudaMemsetAsync(rs.best, 0, (size_t)n_live * 8));
            CU(cudaMemsetAsync(rs.second, 0, (size_t)n_live * 4));
            run_best_kernel<<<blocks(n_live), 256>>>(slot, n_live, ids, heights, d_head.as<uint32_t>(),
                                                     d_runof.as<uint32_t>(), rs, method);
            run_second_kernel<<<blocks(n_live), 256>>>(slot, n_live, ids, heights, d_runof.as<uint32_t>(), rs, method);
            uint32_t n_runs_u = 0;
            CU(cudaMemcpy(&n_runs_u, d_runof.as<uint32_t>() + (n_live - 1), 4, cudaMemcpyDeviceToHost));
            const int64_t n_runs = n_runs_u;
            if (n_runs == n_live) continue;  // all keys distinct: nothing merges
            run_decide_kernel<<<blocks(n_runs), 256>>>(n_runs, slot, ids, heights, rs, method, max_height);
            t = tmp_bytes;
            CU(cub::DeviceScan::ExclusiveSum(d_tmp.p, t, rs.cnt, d_off.as<unsigned long long>(), (int)n_runs));
            star_emit_kernel<<<blocks(n_live), 256>>>(slot, n_live, ids, d_runof.as<uint32_t>(), rs,
                                                      d_off.as<unsigned long long>(), method, is_merged,
                                                      d_edges.as<uint32_t>(), n_edges, d_fin.as<uint32_t>(), n_fin);
            unsigned long long last_off = 0, last_cnt = 0;
            CU(cudaMemcpy(&last_off, d_off.as<unsigned long long>() + (n_runs - 1), 8, cudaMemcpyDeviceToHost));
            CU(cudaMemcpy(&last_cnt, rs.cnt + (n_runs - 1), 8, cudaMemcpyDeviceToHost));
            const unsigned long long tot = last_off + last_cnt;
            n_edges += (int64_t)(tot & ((1ull << 40) - 1));
            n_fin += (int64_t)(tot >> 40);
            t = tmp_bytes;
            CU(cub::DeviceSelect::If(d_tmp.p, t, live, live_alt, d_nsel.as<int>(), (int)n_live, NotMerged{is_merged}));
            int n_sel = 0;
            CU(cudaMemcpy(&n_sel, d_nsel.p, 4, cudaMemcpyDeviceToHost));
            std::swap(live, live_alt);
            n_live = n_sel;
        } while (std::prev_permutation(sel.begin(), sel.end()));
        // next round's pool: the unmerged ids in pool order (DCAT.h:611-615, 1283-1286)
        if (n_live > 0) gather_ids_kernel<<<blocks(n_live), 256>>>(live, n_live, ids, ids_alt);
        std::swap(ids, ids_alt);
        n_ids = n_live;
        CU(cudaGetLastError());
        if (stats) {
            CU(cudaDeviceSynchronize());
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - round_t0).count();
            fprintf(stderr, "dpq_find_edges: diff %d pool %lld -> %lld, %lld passes (%lld skipped so far), mean live %lld, %.3f s\n",
                    diff, (long long)round_n0, (long long)n_ids, (long long)round_passes, (long long)skipped_passes,
                    (long long)(round_passes ? round_live_sum / round_passes : 0), sec);
        }
    }
    if (n_edges + n_fin + (n_ids > 0 ? 1 : 0) != n_codes)
        return dpq::api_fail(DPQ_ERR_CUDA, "dpq_find_edges: internal edge count mismatch");
    std::vector<uint32_t> fin((size_t)n_fin + 1);
    if (n_fin) CU(cudaMemcpy(fin.data(), d_fin.p, (size_t)n_fin * 4, cudaMemcpyDeviceToHost));
    if (n_ids > 0) {  // the last survivor joins the finalists (DCAT.h:1292-1294)
        CU(cudaMemcpy(&fin[(size_t)n_fin], ids, 4, cudaMemcpyDeviceToHost));
        ++n_fin;
    }
    *root_id = fin[0];  // DCAT.h:1297-1313: a star under the first finalist
    std::vector<uint32_t> star((size_t)std::max<int64_t>(2 * (n_fin - 1), 0));
    for (int64_t i = 1; i < n_fin; ++i) {
        star[(size_t)(2 * (i - 1))] = fin[0];
        star[(size_t)(2 * (i - 1) + 1)] = fin[(size_t)i];
    }
    if (edges) {
        if (n_edges) CU(cudaMemcpy(edges, d_edges.p, (size_t)n_edges * 8, cudaMemcpyDeviceToHost));
        if (!star.empty()) memcpy(edges + 2 * n_edges, star.data(), star.size() * 4);
    }
    if (d_edges_out) {  // the star goes behind the device copy, which the caller takes over
        if (!star.empty())
            CU(cudaMemcpy(d_edges.as<uint32_t>() + 2 * n_edges, star.data(), star.size() * 4, cudaMemcpyHostToDevice));
        *d_edges_out = d_edges.as<uint32_t>();
        d_edges.p = nullptr;
    }
    return DPQ_OK;
}

extern "C" int dpq_find_edges(const uint8_t* codes, int64_t n_codes, int M, int K, int max_height_folds,
                              int method, uint32_t* edges, uint32_t* root_id) {
    return find_edges_core(codes, n_codes, M, K, max_height_folds, method, edges, nullptr, root_id);
}

namespace dpq {
int find_edges_device(const uint8_t* codes, int64_t n, int M, int K, int max_height_folds, int method,
                      uint32_t** d_edges_out, uint32_t* root_id) {
    return find_edges_core(codes, n, M, K, max_height_folds, method, nullptr, d_edges_out, root_id);
}
}  // namespace dpq
